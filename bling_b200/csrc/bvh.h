// bvh.h -- accelerator layout and the per-ray traversal bodies.
// Replaces scIntersect/occluded -> kdTreePrimitive.inter/inters + traverse/traverse' (Scene.hs:45-51,
// KdTree.hs:210-246, Primitive.hs:29-43). The reference walks a SAH kd-tree; the result of that walk is the
// globally nearest hit over all primitives (leaf primitives are tested against the full ray range and
// `near` shrinks rayMax), so any conservative accelerator returns the same (t, prim) as long as
// the primitive tests are the same arithmetic. The library builds its own BVH (bvh_build.cpp).
//
// Node = 64 B = 4 x 16-byte loads, holding BOTH children's boxes (one fetch decides both):
//   n0 = c0.lo.x c0.hi.x c0.lo.y c0.hi.y     n1 = c1.lo.x c1.hi.x c1.lo.y c1.hi.y
//   n2 = c0.lo.z c0.hi.z c1.lo.z c1.hi.z     n3 = child0 child1 (int bits) - -
// child >= 0: node index. child < 0: leaf, ~child = (first_item << 4) | count.
// Leaf item = 48 B = 3 x 16-byte loads:
//   triangle: (p1.xyz, prim_id) (e1.xyz, 0) (e2.xyz, -)       shape: (-, -, -, prim_id) (-, -, -, 1 + shape index) -
// Boxes are inflated by the builder, so the slab test needs no epsilon (see bvh_build.cpp).
#pragma once
#include "geom.h"

namespace bl {

struct Bvh {
   const F4 *nodes;
   const F4 *items;
   const blingcu_shape *shapes;
   int root;          // node index (there is always at least one node unless the scene is empty: root = -1 and n_nodes = 0)
   int n_nodes;
};

struct HitRec { float t; int prim; float b1, b2; };

#define BL_STACK 64

HD bool leafItemNearest(const Bvh &bvh, int item, Ray &r, HitRec &h) {
   F4 q0 = ld4(bvh.items + 3 * item), q1 = ld4(bvh.items + 3 * item + 1), q2 = ld4(bvh.items + 3 * item + 2);   // one 48-byte record, three independent LDG.128
   int tag = f2i(q1.w);
   if (tag == 0) {
      float t, b1, b2;
      if (!triHit(mk3(q0.x, q0.y, q0.z), mk3(q1.x, q1.y, q1.z), mk3(q2.x, q2.y, q2.z), r, t, b1, b2)) return false;
      r.tmax = t; h.t = t; h.prim = f2i(q0.w); h.b1 = b1; h.b2 = b2;
      return true;
   }
   const blingcu_shape &s = bvh.shapes[tag - 1];
   float t; DG dg;
   if (!shapeIntersect<false>(s, transRay(s.w2o, r), t, dg)) return false;   // Geometry.hs:33-36: direction not renormalised, t preserved
   r.tmax = t; h.t = t; h.prim = f2i(q0.w); h.b1 = 0; h.b2 = 0;
   return true;
}
HD bool leafItemAny(const Bvh &bvh, int item, const Ray &r) {
   F4 q0 = ld4(bvh.items + 3 * item), q1 = ld4(bvh.items + 3 * item + 1), q2 = ld4(bvh.items + 3 * item + 2);   // one 48-byte record, three independent LDG.128
   int tag = f2i(q1.w);
   if (tag == 0) {
      float t, b1, b2;
      return triHit(mk3(q0.x, q0.y, q0.z), mk3(q1.x, q1.y, q1.z), mk3(q2.x, q2.y, q2.z), r, t, b1, b2);   // TriangleMesh.hs:140-158 == same predicate
   }
   const blingcu_shape &s = bvh.shapes[tag - 1];
   return shapeIntersects(s, transRay(s.w2o, r));
}

// slab test of both children of one node against [r.tmin, r.tmax]
HD void nodeTest(const F4 &n0, const F4 &n1, const F4 &n2, const Ray &r, V3 idir, float &tn0, float &tn1, bool &h0, bool &h1) {
   float c0lx = (n0.x - r.o.x) * idir.x, c0hx = (n0.y - r.o.x) * idir.x;
   float c0ly = (n0.z - r.o.y) * idir.y, c0hy = (n0.w - r.o.y) * idir.y;
   float c0lz = (n2.x - r.o.z) * idir.z, c0hz = (n2.y - r.o.z) * idir.z;
   float c1lx = (n1.x - r.o.x) * idir.x, c1hx = (n1.y - r.o.x) * idir.x;
   float c1ly = (n1.z - r.o.y) * idir.y, c1hy = (n1.w - r.o.y) * idir.y;
   float c1lz = (n2.z - r.o.z) * idir.z, c1hz = (n2.w - r.o.z) * idir.z;
   tn0 = fmaxf(fmaxf(fminf(c0lx, c0hx), fminf(c0ly, c0hy)), fmaxf(fminf(c0lz, c0hz), r.tmin));
   float tf0 = fminf(fminf(fmaxf(c0lx, c0hx), fmaxf(c0ly, c0hy)), fminf(fmaxf(c0lz, c0hz), r.tmax));
   tn1 = fmaxf(fmaxf(fminf(c1lx, c1hx), fminf(c1ly, c1hy)), fmaxf(fminf(c1lz, c1hz), r.tmin));
   float tf1 = fminf(fminf(fmaxf(c1lx, c1hx), fmaxf(c1ly, c1hy)), fminf(fmaxf(c1lz, c1hz), r.tmax));
   h0 = tn0 <= tf0; h1 = tn1 <= tf1;
}

// nearest hit (Primitive.intersect). STATS counts node fetches / primitive tests like dbgTraverse (KdTree.hs:260-281).
template <bool STATS>
HD HitRec traceNearest(const Bvh &bvh, Ray r, uint32_t *nNodes, uint32_t *nPrims) {
   HitRec h; h.t = 0; h.prim = -1; h.b1 = 0; h.b2 = 0;
   if (bvh.root < 0) return h;
   V3 idir = mk3(1.0f / r.d.x, 1.0f / r.d.y, 1.0f / r.d.z);
   int stack[BL_STACK]; int sp = 0;
   int cur = bvh.root;
   for (;;) {
      if (cur >= 0) {
         const F4 *np = bvh.nodes + 4 * (size_t)cur;
         F4 n0 = ld4(np), n1 = ld4(np + 1), n2 = ld4(np + 2), n3 = ld4(np + 3);
         if (STATS) (*nNodes)++;
         float tn0, tn1; bool h0, h1;
         nodeTest(n0, n1, n2, r, idir, tn0, tn1, h0, h1);
         int c0 = f2i(n3.x), c1 = f2i(n3.y);
         if (h0 && h1) {
            if (tn1 < tn0) { int t = c0; c0 = c1; c1 = t; }
            if (sp < BL_STACK) stack[sp++] = c1;
            cur = c0;
            continue;
         }
         if (h0) { cur = c0; continue; }
         if (h1) { cur = c1; continue; }
      } else {
         int enc = ~cur; int first = enc >> 4, cnt = enc & 15;
         for (int i = 0; i < cnt; ++i) { if (STATS) (*nPrims)++; leafItemNearest(bvh, first + i, r, h); }
      }
      if (sp == 0) break;
      cur = stack[--sp];
   }
   return h;
}

// any hit (Primitive.intersects)
HD bool traceAny(const Bvh &bvh, const Ray &r) {
   if (bvh.root < 0) return false;
   V3 idir = mk3(1.0f / r.d.x, 1.0f / r.d.y, 1.0f / r.d.z);
   int stack[BL_STACK]; int sp = 0;
   int cur = bvh.root;
   for (;;) {
      if (cur >= 0) {
         const F4 *np = bvh.nodes + 4 * (size_t)cur;
         F4 n0 = ld4(np), n1 = ld4(np + 1), n2 = ld4(np + 2), n3 = ld4(np + 3);
         float tn0, tn1; bool h0, h1;
         nodeTest(n0, n1, n2, r, idir, tn0, tn1, h0, h1);
         int c0 = f2i(n3.x), c1 = f2i(n3.y);
         if (h0 && h1) { if (sp < BL_STACK) stack[sp++] = c1; cur = c0; continue; }
         if (h0) { cur = c0; continue; }
         if (h1) { cur = c1; continue; }
      } else {
         int enc = ~cur; int first = enc >> 4, cnt = enc & 15;
         for (int i = 0; i < cnt; ++i) if (leafItemAny(bvh, first + i, r)) return true;
      }
      if (sp == 0) break;
      cur = stack[--sp];
   }
   return false;
}

// ---- host-side builder (bvh_build.cpp)
struct BvhBuildInput {
   size_t n;             // number of leaf items
   const float *lo;      // n*3 item bounds (NOT inflated)
   const float *hi;      // n*3
   int max_leaf;         // 1..15
   int threads;
};
struct BvhBuildOutput {
   F4 *nodes;            // malloc'ed, 4*n_nodes
   uint32_t *order;      // malloc'ed, n: order[k] = input item stored at leaf position k
   int n_nodes;
   int root;
   float scene_lo[3], scene_hi[3];
};
int bvhBuild(const BvhBuildInput &in, BvhBuildOutput &out);

}  // namespace bl
