// bvh.h -- accelerator layout and the per-ray traversal bodies.
// Replaces scIntersect/occluded -> kdTreePrimitive.inter/inters + traverse/traverse' (Scene.hs:45-51,
// KdTree.hs:210-246, Primitive.hs:29-43). The reference walks a SAH kd-tree; the result of that walk is the
// globally nearest hit over all primitives (leaf primitives are tested against the full ray range and
// `near` shrinks rayMax), so any conservative accelerator returns the same (t, prim) as long as
// the primitive tests are the same arithmetic. The library builds its own BVH (bvh_build.cpp).
//
// Node = 64 B = a 4-wide BVH node with child boxes quantised to 8 bits on a node-local grid (two 32-byte loads):
//   F4[0] = (P.x P.y P.z  E)      grid origin (f32) and the three grid exponents packed as biased float exponents
//                                  (E = ex | ey << 8 | ez << 16, cell size 2^(e-127) per axis)
//   F4[1] = child refs [0..3]      ref >= 0: node index. ref < 0: leaf, ~ref = (first_item << 4) | count
//   F4[2] = (qlo.x qlo.y qlo.z qhi.x), F4[3] = (qhi.y qhi.z 0x47000000 -)   each q word packs the byte of the four children
// Dequantised plane = P + cell * q. The ray/plane distance is ONE fma per plane: q_as_float * (cell/d) + (P/d - o/d),
// with q_as_float built by a byte permute into the mantissa of 2^15 (bits 0x47000000 | q << 8 = 32768 + q exactly) and the
// 2^15 folded into the addend. The folded addend is rounded at the magnitude of 2^15 cells, i.e. to 1/512 of a cell, so the
// builder moves every bound out by 1/64 of a cell before it rounds to the grid. (Until the end of round 2 the byte went into
// the LOW mantissa byte of 2^23; that addend is rounded to HALF a cell and every bound carried one extra cell per side, which
// cost 3-5 % node visits and 5 % primitive tests on the cfg-5 ray streams: tools/travsim.cpp, QPAD.) Unused child slots
// carry qlo = 255 > qhi = 0, which the direction-sign based near/far selection rejects on its own. Why quantise: ncu showed the LSU data pipe as the
// limiter with 128-byte nodes (one L1 wavefront per lane per 16 bytes); 64-byte nodes halve it and halve the
// L2/DRAM bytes per visit. The builder (bvh_build.cpp) builds a binned-SAH binary tree, collapses it into 4-wide
// nodes -- SAH-optimally by dynamic programming for scenes of more than 4096 items, else by repeatedly opening the child with the
// largest area -- then quantises.
// Leaf item = 64 B = 2 x 32-byte loads (BL_ITEM_F4 = 4; the fourth float4 is padding):
//   triangle: (p1.xyz, ref) (e1.xyz, 0) (e2.xyz, -) -       shape: (-, -, -, ref) (-, -, -, 1 + shape index) - -      (ref: see HitRec)
// Why padded: ncu (profiles/r02_trace_warpq.md) shows the L1 data pipe as the busiest unit of the traversal kernels, and for
// scattered accesses it moves one 32-byte sector per load instruction per lane: a 48-byte record read as three LDG.128 costs three
// wavefronts, a 64-byte record read as two LDG.256 costs two. DRAM and L2 (8-25 % / 33-44 % busy) have room for the extra 16 bytes.
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
   int max_stack;     // worst-case traversal stack entries (from the builder)
};

// Hit record. `prim` is the leaf item's REFERENCE, not the reference implementation's primitive id:
//   bit 31 = analytic shape, bits 26..30 = shade kind (BLINGCU_MAT_*, + 16 when the material's textures compute),
//   bits 0..25 = triangle / shape index;  -1 = miss.
// The wavefront classifies by material kind straight from this word and finds the geometry without an indirection;
// the C ABI converts it to the primitive id of `mkScene`'s list on the way out (HitToAbiBody).
struct HitRec { float t; int prim; float b1, b2; };
#define BL_REF_MISS (-1)
HD bool refIsShape(int ref) { return ((uint32_t)ref >> 31) != 0; }
HD int refKind(int ref) { return (int)(((uint32_t)ref >> 26) & 31u); }
HD uint32_t refIndex(int ref) { return (uint32_t)ref & 0x03ffffffu; }
HD int mkRef(bool shape, int kind, uint32_t index) { return (int)((shape ? 0x80000000u : 0u) | ((uint32_t)kind << 26) | index); }


#ifndef BL_ITEM_F4
#define BL_ITEM_F4 4     // F4 per leaf item (64 B); 3 = the packed 48-byte records of round 1 (A/B builds only)
#endif
#ifndef BL_L2_POLICY
#define BL_L2_POLICY 0   // A/B: bit 0 = leaf items with L2 evict_first (640 MB, little reuse), bit 1 = nodes with L2 evict_last
#endif
#ifndef BL_L1_POLICY
#define BL_L1_POLICY 0   // A/B: bit 0 = leaf items bypass L1 (no_allocate: 640 MB, read once per test), bit 1 = nodes with L1 evict_last
#endif
#if defined(__CUDACC__)
// the same 32-byte load with an L1 policy: leaf items are never re-read by the SM that fetched them, the top of the tree always is
__device__ __forceinline__ void ld8NoAlloc(const F4 *p, F4 &a, F4 &b) {
   asm volatile("ld.global.nc.L1::no_allocate.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p));
}
__device__ __forceinline__ void ld8KeepL1(const F4 *p, F4 &a, F4 &b) {
   asm volatile("ld.global.nc.L1::evict_last.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p));
}
__device__ __forceinline__ void ld8(const F4 *p, F4 &a, F4 &b) {   // one 32-byte load (LDG.E.256 on sm_100)
   asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p));
}
// the same with an L2 eviction policy (the policy is a pure value: the compiler hoists its creation out of the loop)
template <bool LAST>
__device__ __forceinline__ void ld8Policy(const F4 *p, F4 &a, F4 &b) {
   unsigned long long pol;
   if (LAST) asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
   else asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
   asm volatile("ld.global.nc.L2::cache_hint.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8], %9;"
                : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p), "l"(pol));
}
#endif
HD void ldItem(const F4 *items, int item, F4 &q0, F4 &q1, F4 &q2) {
   const F4 *p = items + (size_t)BL_ITEM_F4 * (size_t)item;
#if defined(__CUDA_ARCH__) && BL_ITEM_F4 == 4
   F4 q3;
   if (BL_L2_POLICY & 1) { ld8Policy<false>(p, q0, q1); ld8Policy<false>(p + 2, q2, q3); }
   else if (BL_L1_POLICY & 1) { ld8NoAlloc(p, q0, q1); ld8NoAlloc(p + 2, q2, q3); }
   else { ld8(p, q0, q1); ld8(p + 2, q2, q3); }
#else
   q0 = ld4(p); q1 = ld4(p + 1); q2 = ld4(p + 2);
#endif
}

HD bool leafItemNearest(const Bvh &bvh, int item, Ray &r, HitRec &h) {
   F4 q0, q1, q2; ldItem(bvh.items, item, q0, q1, q2);
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
   F4 q0, q1, q2; ldItem(bvh.items, item, q0, q1, q2);
   int tag = f2i(q1.w);
   if (tag == 0) {
      float t, b1, b2;
      return triHit(mk3(q0.x, q0.y, q0.z), mk3(q1.x, q1.y, q1.z), mk3(q2.x, q2.y, q2.z), r, t, b1, b2);   // TriangleMesh.hs:140-158 == same predicate
   }
   const blingcu_shape &s = bvh.shapes[tag - 1];
   return shapeIntersects(s, transRay(s.w2o, r));
}

#define BL_NODE_F4 4     // F4 per node (64 B)
#define BL_TRI_F4 4      // F4 per triangle of the shading geometry (shading.h::DScene::tri_p)
#define BL_STACK 192     // worst case 3 pushes per level of a 56-level binary tree; real scenes use < 48

#if defined(__CUDA_ARCH__)
#define BL_FMA(a, b, c) __fmaf_rn(a, b, c)
#else
#define BL_FMA(a, b, c) fmaf(a, b, c)
#endif
struct RayPre { V3 idir; };   // 1/d: plane distance = (plane - o) * idir; the node-constant part (P - o) * idir is folded into the FMA's addend
// |1/d| is clamped to 1e18 so that the distances never evaluate inf - inf: an axis-parallel ray then sees
// (-huge, +huge) when its origin is inside the slab and two same-signed huge values (a miss) when it is outside.
// (Round 1 also kept o/d per ray; that cost three registers the 64-register traversal kernels do not have -- ncu showed them
// spilled and re-read from local memory at every node step -- for no fewer instructions: (P - o) * idir is one FADD + one FMUL.)
HD float safeInv(float d) { return fminf(fmaxf(1.0f / d, -1e18f), 1e18f); }
HD RayPre rayPre(const Ray &r) { RayPre p; p.idir = mk3(safeInv(r.d.x), safeInv(r.d.y), safeInv(r.d.z)); return p; }

HD uint32_t f2u(float f) { return (uint32_t)f2i(f); }
HD float u2f(uint32_t u) { return i2f((int)u); }
// float with bits 0x47000000 | (byte k of w) << 8  ==  2^15 + q   (one PRMT on the device)
#define BL_QMAGIC 0x47000000u     // 32768.0f
#define BL_QBIAS 32768.0f
#if defined(__CUDA_ARCH__)
// the 2^15 pattern comes from the node itself (word F4[3].z, written by the builder) so that it sits in a register and
// the byte selector can be the PRMT's immediate; as a literal it would take the immediate slot and every PRMT would
// need a MOV of its selector first
#define BL_QF(w, k) __uint_as_float(__byte_perm((w), bl_k23, 0x7604u + ((k) << 4)))
#else
#define BL_QF(w, k) u2f(BL_QMAGIC | ((((w) >> (8 * (k))) & 0xffu) << 8))
#endif

// slab test of the four children of one quantised node against [r.tmin, r.tmax]: tn[k] = entry distance of child k,
// +inf when the ray misses it. OVERLAP: tn[k] = entry - exit distance instead, i.e. minus the length of the ray inside child k:
// <= 0 for a hit, > 0 (or NaN: a ray parallel to a slab it lies outside of, with tmax = inf) for a miss. The any-hit kernel
// enters the child the ray stays in longest (trace_warpq.cuh).
template <bool OVERLAP = false>
HD void node4Near(const F4 &n0, const F4 &n2, const F4 &n3, const Ray &r, const RayPre &p, float tnear[4]) {
#if defined(__CUDA_ARCH__)
   const uint32_t bl_k23 = f2u(n3.z);   // BL_QMAGIC
#endif
   const uint32_t E = f2u(n0.w);
   const float ax = u2f((E & 0xffu) << 23) * p.idir.x, ay = u2f(((E >> 8) & 0xffu) << 23) * p.idir.y, az = u2f(((E >> 16) & 0xffu) << 23) * p.idir.z;
   // addend: (P - o)/d - 2^15 * cell/d
   const float bx = BL_FMA(-BL_QBIAS, ax, (n0.x - r.o.x) * p.idir.x);
   const float by = BL_FMA(-BL_QBIAS, ay, (n0.y - r.o.y) * p.idir.y);
   const float bz = BL_FMA(-BL_QBIAS, az, (n0.z - r.o.z) * p.idir.z);
   const uint32_t qlx = f2u(n2.x), qly = f2u(n2.y), qlz = f2u(n2.z), qhx = f2u(n2.w), qhy = f2u(n3.x), qhz = f2u(n3.y);
   const bool px = p.idir.x >= 0.0f, py = p.idir.y >= 0.0f, pz = p.idir.z >= 0.0f;
   const uint32_t nx = px ? qlx : qhx, fx = px ? qhx : qlx, ny = py ? qly : qhy, fy = py ? qhy : qly, nz = pz ? qlz : qhz, fz = pz ? qhz : qlz;
   BL_UNROLL for (int k = 0; k < 4; ++k) {
      float tn = fmaxf(fmaxf(BL_FMA(BL_QF(nx, k), ax, bx), BL_FMA(BL_QF(ny, k), ay, by)), fmaxf(BL_FMA(BL_QF(nz, k), az, bz), r.tmin));
      float tf = fminf(fminf(BL_FMA(BL_QF(fx, k), ax, bx), BL_FMA(BL_QF(fy, k), ay, by)), fminf(BL_FMA(BL_QF(fz, k), az, bz), r.tmax));
      if (OVERLAP) tnear[k] = tn - tf;
      else tnear[k] = (tn <= tf) ? tn : BL_INF;
   }
}
// sort the four (entry distance, child reference) pairs by distance: 5 compare-exchanges, each one compare + four
// selects (no key packing, no indexed pick afterwards). Misses carry +inf and end up last. On equal distances the
// lower slot stays first; every traversal variant uses this same network, so they all visit nodes in the same order.
HD void cswap(float &ta, int &ra, float &tb, int &rb) {
   const bool p = tb < ta;
   const float t0 = p ? tb : ta, t1 = p ? ta : tb; const int r0 = p ? rb : ra, r1 = p ? ra : rb;
   ta = t0; tb = t1; ra = r0; rb = r1;
}
HD void sort4(float t[4], int c[4]) {
   cswap(t[0], c[0], t[1], c[1]); cswap(t[2], c[2], t[3], c[3]);
   cswap(t[0], c[0], t[2], c[2]); cswap(t[1], c[1], t[3], c[3]);
   cswap(t[1], c[1], t[2], c[2]);
}

// nearest hit (Primitive.intersect). STATS counts node fetches / primitive tests like dbgTraverse (KdTree.hs:260-281).
// Children are entered nearest-first; every variant of the traversal kernel follows this same order.
template <bool STATS>
HD HitRec traceNearest(const Bvh &bvh, Ray r, uint32_t *nNodes, uint32_t *nPrims) {
   HitRec h; h.t = 0; h.prim = -1; h.b1 = 0; h.b2 = 0;
   if (bvh.root < 0) return h;
   RayPre pre = rayPre(r);
   int stack[BL_STACK]; int sp = 0;
   int cur = bvh.root;
   for (;;) {
      if (cur >= 0) {
         const F4 *np = bvh.nodes + BL_NODE_F4 * (size_t)cur;
         if (STATS) (*nNodes)++;
         float tn[4];
         F4 n0 = ld4(np), n1 = ld4(np + 1), n2 = ld4(np + 2), n3 = ld4(np + 3);
         node4Near(n0, n2, n3, r, pre, tn);
         int n6[4] = {f2i(n1.x), f2i(n1.y), f2i(n1.z), f2i(n1.w)};
         sort4(tn, n6);
         if (tn[0] < BL_INF) {
            for (int j = 3; j >= 1; --j) if (tn[j] < BL_INF && sp < BL_STACK) stack[sp++] = n6[j];
            cur = n6[0];
            continue;
         }
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
   RayPre pre = rayPre(r);
   int stack[BL_STACK]; int sp = 0;
   int cur = bvh.root;
   for (;;) {
      if (cur >= 0) {
         const F4 *np = bvh.nodes + BL_NODE_F4 * (size_t)cur;
         float tn[4];
         F4 n0 = ld4(np), n1 = ld4(np + 1), n2 = ld4(np + 2), n3 = ld4(np + 3);
         node4Near(n0, n2, n3, r, pre, tn);
         const int n6[4] = {f2i(n1.x), f2i(n1.y), f2i(n1.z), f2i(n1.w)};
         int next = 0; bool have = false;
         for (int k = 0; k < 4; ++k) if (tn[k] < BL_INF) { int c = n6[k]; if (!have) { next = c; have = true; } else if (sp < BL_STACK) stack[sp++] = c; }
         if (have) { cur = next; continue; }
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
   float trav_cost = 0;  // cost of visiting a node in units of one primitive test (SAH termination for ranges of <= 15 items); 0 = split whenever it lowers the test count
   int force_leaf = 1;   // 1: a range of <= max_leaf items always becomes a leaf (round 1); 0: the SAH decides there too
   float collapse_cp = 0; // > 0: SAH-optimal collapse (bvh_build.cpp::Collapse) with a primitive test costing collapse_cp node visits; 0: greedy collapse over leaves of <= max_leaf items
};
struct BvhBuildOutput {
   F4 *nodes;            // malloc'ed, BL_NODE_F4*n_nodes (4-wide nodes)
   uint32_t *order;      // malloc'ed, n: order[k] = input item stored at leaf position k
   int n_nodes;
   int root;
   int max_stack;        // worst-case number of stack entries a traversal of this tree can hold
   float scene_lo[3], scene_hi[3];
};
int bvhBuild(const BvhBuildInput &in, BvhBuildOutput &out);

}  // namespace bl
