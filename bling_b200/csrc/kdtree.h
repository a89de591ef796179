// kdtree.h -- SURVEY.md §8(f)3: the REFERENCE's own accelerator, flattened by the host, as an alternative input. The product
// path traverses the library's BVH (bvh.h); this one walks the host's SAH kd-tree (Primitive/KdTree.hs:29-33) exactly as
// `traverse` does (KdTree.hs:223-242, entered through intersectAABB, AABB.hs:79-94) and counts what `dbgTraverse` counts
// (KdTree.hs:252-281), so that a host can hold the GPU's answers AND its TraversalStats against its own, node for node.
#pragma once
#include "shading.h"

namespace bl {

struct KdTreeDev {
   const blingcu_kdnode *nodes;
   const uint32_t *leaf;        // primitive ids (position in mkScene's list), in leaf order
   const int32_t *primRef;      // primitive id -> hit reference (bvh.h::mkRef): where the primitive's geometry lives
   int root;
   float lo[3], hi[3];          // the tree's bounds (KdTree b _)
};
#define BL_KD_STACK 96          // maximum depth is round (8 + 3 ln n) <= 64 for n <= 2^26 (KdTree.hs:107-112)

// intersectAABB (AABB.hs:79-94) with Haskell's max / min on NaN (hd.h)
HD bool kdBoundsHit(const KdTreeDev &kd, const Ray &r, float &tn, float &tf) {
   float nearT = r.tmin, farT = r.tmax;
   for (int dim = 0; dim < 3; ++dim) {
      if (nearT > farT) return false;
      const float oc = comp(r.o, dim), dInv = 1.0f / comp(r.d, dim);
      const float tFar = (kd.hi[dim] - oc) * dInv, tNear = (kd.lo[dim] - oc) * dInv;
      float n2, f2;
      if (tNear > tFar) { n2 = tFar; f2 = tNear; } else { n2 = tNear; f2 = tFar; }
      nearT = hmaxf(nearT, n2); farT = hminf(farT, f2);
   }
   if (nearT > farT) return false;
   tn = nearT; tf = farT;
   return true;
}

// one primitive of a leaf: `near` of Primitive.hs:29-43 (accepts t == rayMax, so the later primitive wins a tie)
HD void kdPrim(const DScene &S, int ref, Ray &r, HitRec &h) {
   if (!refIsShape(ref)) {
      const uint32_t i = refIndex(ref);
      const F4 a = S.tri_p[BL_TRI_F4 * (size_t)i], b = S.tri_p[BL_TRI_F4 * (size_t)i + 1], c = S.tri_p[BL_TRI_F4 * (size_t)i + 2];
      const V3 p1 = mk3(a.x, a.y, a.z);
      float t, b1, b2;
      if (!triHit(p1, mk3(b.x, b.y, b.z) - p1, mk3(c.x, c.y, c.z) - p1, r, t, b1, b2)) return;   // e1, e2 as TriangleMesh.hs:169
      r.tmax = t; h.t = t; h.prim = ref; h.b1 = b1; h.b2 = b2;
      return;
   }
   const blingcu_shape &s = S.shapes[refIndex(ref)];
   float t; DG dg;
   if (!shapeIntersect<false>(s, transRay(s.w2o, r), t, dg)) return;
   r.tmax = t; h.t = t; h.prim = ref; h.b1 = 0; h.b2 = 0;
}

// traverse (KdTree.hs:223-234) without the recursion: the second child waits on a stack with its (tmin, tmax) while the first
// one is walked, and sees the ray as the first one left it. nt / ni count like dbgTraverse' (:268-281): a leaf is one node
// and all its primitives, an interior node counts only where a single child is entered.
HD HitRec kdTraceNearest(const DScene &S, const KdTreeDev &kd, Ray r, uint32_t &nt, uint32_t &ni) {
   HitRec h; h.t = 0; h.prim = -1; h.b1 = 0; h.b2 = 0;
   nt = 0; ni = 0;
   float tmin, tmax;
   if (kd.root < 0 || !kdBoundsHit(kd, r, tmin, tmax)) return h;
   const V3 inv = mk3(1.0f / r.d.x, 1.0f / r.d.y, 1.0f / r.d.z);
   int sNode[BL_KD_STACK]; float sMin[BL_KD_STACK], sMax[BL_KD_STACK]; int sp = 0;
   int node = kd.root;
   for (;;) {
      const blingcu_kdnode n = kd.nodes[node];
      bool done = false;
      if (n.left < 0) {   // Leaf ps -> nearest' ps ri
         nt += 1; ni += n.count;
         for (uint32_t i = 0; i < n.count; ++i) kdPrim(S, kd.primRef[kd.leaf[n.first + i]], r, h);
         done = true;
      } else if (r.tmax < tmin) done = true;   // :226
      else {
         const float oa = comp(r.o, n.axis), da = comp(r.d, n.axis);
         const float tp = (n.split - oa) * comp(inv, n.axis);
         const bool lf = (oa < n.split) || (oa == n.split && da <= 0.0f);
         const int fc = lf ? n.left : n.right, sc = lf ? n.right : n.left;
         if (tp > tmax || tp <= 0.0f) { nt += 1; node = fc; }
         else if (tp < tmin) { nt += 1; node = sc; }
         else {
            if (sp < BL_KD_STACK) { sNode[sp] = sc; sMin[sp] = tp; sMax[sp] = tmax; sp++; }
            node = fc; tmax = tp;
         }
      }
      if (done) {
         if (sp == 0) break;
         --sp; node = sNode[sp]; tmin = sMin[sp]; tmax = sMax[sp];
      }
   }
   return h;
}

struct TraceKdBody {
   const DScene *sc; KdTreeDev kd; const F4 *o, *d; F4 *hit; uint32_t *nodes, *prims;
   HD void operator()(uint32_t i) const {
      uint32_t nt, ni;
      HitRec h = kdTraceNearest(*sc, kd, loadRay(o, d, i), nt, ni);
      F4 v; v.x = h.t; v.y = h.b1; v.z = h.b2; v.w = i2f(h.prim); hit[i] = v;
      nodes[i] = nt; prims[i] = ni;
   }
};

}  // namespace bl
