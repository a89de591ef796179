// trace_kernels.cuh -- K2 trace_nearest / K3 trace_any: BVH traversal kernels for sm_100a.
// Replaces the per-ray recursion of KdTree.hs:210-246 (a4, a5 in SURVEY.md §8a).
//
// variant 0  one ray per thread, grid-stride, traversal stack in local memory (the reference point)
// variant 1  persistent warps, one ray per lane, majority-vote stepping, LDG.256 node fetch, shared-memory stack
//            (kTracePersistent below; the product path)
// B200 has no RT cores and traversal is not a contraction: no tensor cores here. Measured limiters (DESIGN.md §5,
// profiles/): the ALU pipe and issue slots of the slab/sort/push arithmetic, then the LSU data pipe; HBM is at 9-17 %.
#pragma once
#include "bodies.h"
#include <cuda_runtime.h>

namespace bl {

struct TraceConfig {
   int sms = 148;
   int variant = 3;                   // 0 reference point, 1 majority-vote stepping (round 1), 2 warp-level leaf queue, 3 = 2 with unsorted any-hit (the product)
   int blocksPerSm = 8;
   int maxStack = 64;                 // worst-case stack entries of the uploaded tree (Bvh::max_stack)
   uint32_t *workCounter = nullptr;   // device, one uint32 per launch slot
   bool countStats = false;           // option "traversal_stats": nearest-hit launches count node fetches / primitive tests
   unsigned long long *travCounters = nullptr;   // device: nodes, prims, rays of the nearest-hit launches, then the same three of the any-hit launches
   Bvh bvh{};                         // host copy of the uploaded accelerator's (device) pointers: kernel parameter of variant 2 / 3
};

// ---------------------------------------------------------------------------------------------- variant 0
__global__ void __launch_bounds__(128) kTraceNearestSimple(const uint32_t *__restrict__ q, const uint32_t *__restrict__ cnt, uint32_t n,
                                                          const DScene *__restrict__ sc, const F4 *__restrict__ O, const F4 *__restrict__ D, F4 *__restrict__ hit) {
   uint32_t total = cnt ? *cnt : n;
   Bvh bvh = sc->bvh;
   for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < total; k += gridDim.x * blockDim.x) {
      uint32_t i = q ? q[k] : k;
      HitRec h = traceNearest<false>(bvh, loadRay(O, D, i), nullptr, nullptr);
      F4 v; v.x = h.t; v.y = h.b1; v.z = h.b2; v.w = i2f(h.prim); hit[i] = v;
   }
}
__global__ void __launch_bounds__(128) kTraceAnySimple(const uint32_t *__restrict__ q, const uint32_t *__restrict__ cnt, uint32_t n,
                                                      const DScene *__restrict__ sc, const F4 *__restrict__ O, const F4 *__restrict__ D, uint8_t *__restrict__ occl) {
   uint32_t total = cnt ? *cnt : n;
   Bvh bvh = sc->bvh;
   for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < total; k += gridDim.x * blockDim.x) {
      uint32_t i = q ? q[k] : k;
      occl[i] = traceAny(bvh, loadRay(O, D, i)) ? 1 : 0;
   }
}
__global__ void __launch_bounds__(128) kTraceStats(uint32_t n, const DScene *__restrict__ sc, const F4 *__restrict__ O, const F4 *__restrict__ D,
                                                  F4 *__restrict__ hit, uint32_t *__restrict__ nodes, uint32_t *__restrict__ prims) {
   Bvh bvh = sc->bvh;
   for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
      uint32_t nn = 0, np = 0;
      HitRec h = traceNearest<true>(bvh, loadRay(O, D, i), &nn, &np);
      F4 v; v.x = h.t; v.y = h.b1; v.z = h.b2; v.w = i2f(h.prim); hit[i] = v;
      nodes[i] = nn; prims[i] = np;
   }
}

// instrumented nearest-hit kernel (same traversal order as every variant): totals for the roofline's n_nodes / n_prims
__global__ void __launch_bounds__(128) kTraceNearestCount(const uint32_t *__restrict__ q, const uint32_t *__restrict__ cnt, uint32_t n,
                                                         const DScene *__restrict__ sc, const F4 *__restrict__ O, const F4 *__restrict__ D, F4 *__restrict__ hit,
                                                         unsigned long long *__restrict__ totals) {
   uint32_t total = cnt ? *cnt : n;
   Bvh bvh = sc->bvh;
   unsigned long long nn = 0, np = 0, nr = 0;
   for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < total; k += gridDim.x * blockDim.x) {
      uint32_t i = q ? q[k] : k;
      uint32_t a = 0, b = 0;
      HitRec h = traceNearest<true>(bvh, loadRay(O, D, i), &a, &b);
      F4 v; v.x = h.t; v.y = h.b1; v.z = h.b2; v.w = i2f(h.prim); hit[i] = v;
      nn += a; np += b; nr += 1;
   }
   for (int o = 16; o > 0; o >>= 1) { nn += __shfl_down_sync(0xffffffffu, nn, o); np += __shfl_down_sync(0xffffffffu, np, o); nr += __shfl_down_sync(0xffffffffu, nr, o); }
   if ((threadIdx.x & 31) == 0 && nr) { atomicAdd(totals, nn); atomicAdd(totals + 1, np); atomicAdd(totals + 2, nr); }
}

// ---------------------------------------------------------------------------------------------- variant 1
// Persistent warps, one ray per lane, MAJORITY-VOTE stepping (the product path).
//
// History (profiles/r01_trace_experiments.md): v1 let every lane run its own while loop -> 3.5 of 32 threads active
// per instruction. v2 keeps the warp converged: every trip the lanes vote (ballot) whether more of them stand at an
// inner node or inside a leaf, and the whole warp executes ONLY that step, predicated per lane (2.1x). v3 went to
// 4-wide nodes; ncu then showed the LSU data pipe at 81 % because a node cost 7 LDG.128 per lane, each one L1
// wavefront per lane. v4 (one ray per quad) cut the wavefronts 4x but doubled the instructions per ray. This
// version keeps one ray per lane, fetches a (now 64-byte, quantised: bvh.h) node as 2 x LDG.256, keeps the first
// TR_SS levels of the traversal stack in shared memory (PTX-predicated pushes/pops, no compiler-made branches) with a
// local-memory tail for deeper trees, writes accepted hits straight to the output record, and refills idle lanes
// from the global queue with a warp-aggregated atomic once enough lanes have retired.
#define TR_THREADS 128
#ifndef TR_REFILL
#define TR_REFILL 6       // refill once this many lanes are idle (a refill stalls the whole warp for three dependent loads; cfg 5:
                          // 4 -> 589 / 628 ms nearest / any per 5 steps, 6 -> 585 / 622, 8 -> 589 / 624, 12 -> 610 / 631; tools/gpu_r02_q.sh)
#endif
#ifndef TR_MINBLOCKS
#define TR_MINBLOCKS 8    // resident CTAs per SM of the nearest-hit kernel (64 registers); the any-hit kernel carries no hit
#endif                  // record and fits 9 (56 registers, no spills): +5 % on shadow / MIS rays
#ifndef TR_SS
#define TR_SS 16        // stack levels per thread kept in shared memory; deeper levels live in local memory (see below)
#endif
#ifndef TR_LEAF_W
#define TR_LEAF_W 6       // vote weights (quarters): leaf step when TR_LEAF_W * #leaf lanes > 4 * #node lanes
#endif


// ---- small PTX helpers: keep the hot loop free of compiler-made branches and generic->shared address conversions
__device__ __forceinline__ void stsIf(uint32_t addr, int v, bool p) {   // predicated st.shared.b32
   asm volatile("{ .reg .pred q; setp.ne.b32 q, %2, 0; @q st.shared.b32 [%0], %1; }" ::"r"(addr), "r"(v), "r"((int)p) : "memory");
}
__device__ __forceinline__ int ldsIf(uint32_t addr, int old, bool p) {   // predicated ld.shared.b32, keeps `old` otherwise
   int v = old;
   asm volatile("{ .reg .pred q; setp.ne.b32 q, %2, 0; @q ld.shared.b32 %0, [%1]; }" : "+r"(v) : "r"(addr), "r"((int)p) : "memory");
   return v;
}
// Fused NEE resolve (any-hit instantiation, fuseL != null): an UNOCCLUDED shadow ray adds its pending contribution to the path's
// radiance when the kernel retires the ray (`L += pending`, one float4 quarter at a time; the retire path is cold and the
// traversal registers are dead there: still 56 registers), instead of writing an occlusion flag for a separate resolve launch.
// Same arithmetic in the same order as ResolveShadowBody: films are bit-identical; +1.4 .. 2.9 % on the named scenes
// (profiles/r01_trace_experiments.md).
template <bool ANY>
__global__ void __launch_bounds__(TR_THREADS, ANY ? TR_MINBLOCKS + 1 : TR_MINBLOCKS) kTracePersistent(const uint32_t *__restrict__ q, const uint32_t *__restrict__ cnt, uint32_t n,
                                                                 const DScene *__restrict__ sc, const F4 *__restrict__ O, const F4 *__restrict__ D,
                                                                 F4 *__restrict__ hit, uint8_t *__restrict__ occl, uint32_t *__restrict__ work, F4 *__restrict__ fuseL, const F4 *__restrict__ fuseP, uint32_t fuseCap) {
   extern __shared__ int sstack[];            // [level][TR_THREADS]: conflict-free, one column per thread
   const unsigned FULL = 0xffffffffu;
   const uint32_t total = cnt ? *cnt : n;
   const Bvh bvh = sc->bvh;
   const unsigned lane = threadIdx.x & 31u;
   const uint32_t stackBase = (uint32_t)__cvta_generic_to_shared(sstack + threadIdx.x);   // 32-bit shared address of my column
   const uint32_t LV = TR_THREADS * (uint32_t)sizeof(int);                                // bytes per stack level
   // Only the first TR_SS levels live in shared memory (8 KB per CTA whatever the tree depth, so 8 CTAs per SM always
   // fit and most of the 256 KB L1/shared array stays L1); deeper levels (rare) spill to a local-memory tail.
   int tail[BL_STACK - TR_SS];
   const int EMPTY = (int)0x80000000;   // lane holds no ray; otherwise cur = child reference (>= 0 node, < 0 encoded leaf)
   int cur = EMPTY, li = 0;
   int sp = 0;                          // stack entries in use
   uint32_t slot = 0;
   Ray r; RayPre pre;
   bool exhausted = false;
   r.o = mk3(0, 0, 0); r.d = mk3(0, 0, 1); r.tmin = 0; r.tmax = 0; pre.idir = mk3(0, 0, 0);

   for (;;) {
      // ---- where does every lane stand?
      bool atNode = cur >= 0, atLeaf = cur < 0 && cur != EMPTY;
      unsigned mN = __ballot_sync(FULL, atNode), mL = __ballot_sync(FULL, atLeaf);
      // ---- refill idle lanes (warp-uniform decision)
      if (!exhausted && __popc(mN | mL) <= 32 - TR_REFILL) {
         const unsigned idle = ~(mN | mL);
         uint32_t base = 0;
         const int leader = __ffs(idle) - 1;
         if ((int)lane == leader) base = atomicAdd(work, (uint32_t)__popc(idle));
         base = __shfl_sync(FULL, base, leader);
         if (cur == EMPTY) {
            const uint32_t k = base + __popc(idle & ((1u << lane) - 1u));
            if (k < total) {
               slot = q ? q[k] : k;
               r = loadRay(O, D, slot);
               pre = rayPre(r);
               if (!ANY) { F4 v_; v_.x = 0; v_.y = 0; v_.z = 0; v_.w = i2f(-1); hit[slot] = v_; }   // a miss until a leaf step says otherwise
               sp = 0;
               cur = (bvh.root >= 0) ? bvh.root : ~0;   // empty scene: a leaf with zero items
               li = 0;
            }
         }
         if (base + (uint32_t)__popc(idle) >= total) exhausted = true;   // warp-uniform: the queue is drained
         atNode = cur >= 0; atLeaf = cur < 0 && cur != EMPTY;
         mN = __ballot_sync(FULL, atNode); mL = __ballot_sync(FULL, atLeaf);
      }
      if ((mN | mL) == 0) { if (exhausted) break; continue; }
      bool pop = false;
      // ---- vote: node step or leaf step
      if (4 * __popc(mN) >= TR_LEAF_W * __popc(mL)) {
         if (atNode) {
            const F4 *np = bvh.nodes + BL_NODE_F4 * (size_t)cur;
            F4 n0, n1, n2, n3;
            ld8(np, n0, n1); ld8(np + 2, n2, n3);
            float tn[4];
            node4Near(n0, n2, n3, r, pre, tn);
            int c[4] = {f2i(n1.x), f2i(n1.y), f2i(n1.z), f2i(n1.w)};
            sort4(tn, c);   // nearest first, misses (+inf) last
            // branch-free push of the far hits (nearest on top), then enter the nearest
            const bool h0 = tn[0] < BL_INF, h1 = tn[1] < BL_INF, h2 = tn[2] < BL_INF, h3 = tn[3] < BL_INF;
            const int nh = (int)h0 + (int)h1 + (int)h2 + (int)h3;
            const int r0 = c[0], r1 = c[1], r2 = c[2], r3 = c[3];
            // far hits go to levels sp .. sp+nh-2, the nearest of them on top
            const int l1 = sp + nh - 2, l2 = l1 - 1, l3 = l1 - 2;
            if (l1 < TR_SS) {   // common case: everything fits in the shared-memory part (l3 <= l2 <= l1)
               const uint32_t top = stackBase + (uint32_t)l1 * LV;
               stsIf(top, r1, h1); stsIf(top - LV, r2, h2); stsIf(top - 2 * LV, r3, h3);
            } else {
               if (h1) { if (l1 < TR_SS) stsIf(stackBase + (uint32_t)l1 * LV, r1, true); else tail[l1 - TR_SS] = r1; }
               if (h2) { if (l2 < TR_SS) stsIf(stackBase + (uint32_t)l2 * LV, r2, true); else tail[l2 - TR_SS] = r2; }
               if (h3) { if (l3 < TR_SS) stsIf(stackBase + (uint32_t)l3 * LV, r3, true); else tail[l3 - TR_SS] = r3; }
            }
            sp += (nh > 0) ? nh - 1 : 0;
            cur = r0; li = 0;
            pop = !h0;
         }
      } else {
         if (atLeaf) {
            const int enc = ~cur; const int first = enc >> 4, cntl = enc & 15;
            bool found = false;
            if (li < cntl) {
               if (ANY) found = leafItemAny(bvh, first + li, r);
               else {   // an accepted hit shrinks r.tmax and goes straight to the output record (a few stores per ray) instead
                  // of riding in three registers for the whole traversal
                  HitRec hh; hh.t = 0; hh.prim = -1; hh.b1 = hh.b2 = 0;
                  if (leafItemNearest(bvh, first + li, r, hh)) { F4 v_; v_.x = hh.t; v_.y = hh.b1; v_.z = hh.b2; v_.w = i2f(hh.prim); hit[slot] = v_; }
               }
               li++;
            }
            if (ANY && found) { occl[slot] = 1; cur = EMPTY; }
            else pop = li >= cntl;
         }
      }
      // ---- pop (predicated load) or, with an empty stack, finish the ray (rare: once per ray)
      const bool more = sp != 0;
      sp -= (pop && more) ? 1 : 0;
      if (pop && more && sp >= TR_SS) cur = tail[sp - TR_SS];
      else cur = ldsIf(stackBase + (uint32_t)sp * LV, cur, pop && more);
      li = pop ? 0 : li;
      if (pop && !more) {
         if (ANY) {
            if (fuseL) {   // L += pending, one quarter (float4) at a time
               for (int qq = 0; qq < 4; ++qq) {
                  const size_t at = spec4At(fuseCap, slot, qq);
                  F4 l = fuseL[at]; const F4 p_ = fuseP[at];
                  l.x += p_.x; l.y += p_.y; l.z += p_.z; l.w += p_.w;
                  fuseL[at] = l;
               }
            } else occl[slot] = 0;
         }
         cur = EMPTY;
      }
   }
}

// levels of shared-memory stack per thread: the builder's worst case (a pop precedes every push burst of <= 3)
static inline size_t traceSmemBytes(int maxStack) { int lv = maxStack < TR_SS ? maxStack : TR_SS; if (lv < 1) lv = 1; return (size_t)lv * TR_THREADS * sizeof(int); }

}  // namespace bl
#include "trace_warpq.cuh"
namespace bl {

static inline uint32_t traceGrid(const TraceConfig &cfg, uint32_t n, uint32_t raysPerBlock) {
   uint32_t need = (n + raysPerBlock - 1) / raysPerBlock;
   uint32_t full = (uint32_t)cfg.sms * (uint32_t)cfg.blocksPerSm;
   return need < full ? (need ? need : 1) : full;
}
static inline void launchTraceNearest(TraceConfig &cfg, cudaStream_t st, const uint32_t *q, const uint32_t *cnt, uint32_t n, const DScene *sc,
                                      const F4 *O, const F4 *D, F4 *hit) {
   if (cfg.countStats && cfg.travCounters && cfg.variant < 2) { kTraceNearestCount<<<traceGrid(cfg, n, 128), 128, 0, st>>>(q, cnt, n, sc, O, D, hit, cfg.travCounters); return; }
   if (cfg.variant == 0) { kTraceNearestSimple<<<traceGrid(cfg, n, 128), 128, 0, st>>>(q, cnt, n, sc, O, D, hit); return; }
   if (!cfg.workCounter) cudaMalloc(&cfg.workCounter, sizeof(uint32_t));
   cudaMemsetAsync(cfg.workCounter, 0, sizeof(uint32_t), st);
   if (cfg.variant >= 2) {
      TraceConfig cn = cfg; cn.blocksPerSm = cfg.blocksPerSm - (TR_MINBLOCKS - TQ_NEAR_BLOCKS);
      // option "traversal_stats": the SAME kernel with counters (node visits / primitive tests of the product's own schedule)
      if (cfg.countStats && cfg.travCounters) kTraceWarpQ<false, true, true, false><<<traceGrid(cn, n, TR_THREADS), TR_THREADS, traceWarpQSmemBytes(cfg.maxStack), st>>>(q, cnt, n, sc, O, D, hit, nullptr, cfg.workCounter, nullptr, nullptr, 0u, cfg.bvh, cfg.travCounters);
      else kTraceWarpQ<false, true, false, false><<<traceGrid(cn, n, TR_THREADS), TR_THREADS, traceWarpQSmemBytes(cfg.maxStack), st>>>(q, cnt, n, sc, O, D, hit, nullptr, cfg.workCounter, nullptr, nullptr, 0u, cfg.bvh, nullptr);
      return;
   }
   kTracePersistent<false><<<traceGrid(cfg, n, TR_THREADS), TR_THREADS, traceSmemBytes(cfg.maxStack), st>>>(q, cnt, n, sc, O, D, hit, nullptr, cfg.workCounter, nullptr, nullptr, 0u);

}
static inline void launchTraceAny(TraceConfig &cfg, cudaStream_t st, const uint32_t *q, const uint32_t *cnt, uint32_t n, const DScene *sc,
                                  const F4 *O, const F4 *D, uint8_t *occl, F4 *fuseL = nullptr, const F4 *fuseP = nullptr, uint32_t fuseCap = 0) {
   if (cfg.variant == 0) { kTraceAnySimple<<<traceGrid(cfg, n, 128), 128, 0, st>>>(q, cnt, n, sc, O, D, occl); return; }
   if (!cfg.workCounter) cudaMalloc(&cfg.workCounter, sizeof(uint32_t));
   cudaMemsetAsync(cfg.workCounter, 0, sizeof(uint32_t), st);
   TraceConfig c9 = cfg; c9.blocksPerSm = cfg.blocksPerSm + 1;
   if (cfg.variant == 2) {
      if (fuseL) kTraceWarpQ<true, true, false, true><<<traceGrid(cfg, n, TR_THREADS), TR_THREADS, traceWarpQSmemBytes(cfg.maxStack), st>>>(q, cnt, n, sc, O, D, nullptr, occl, cfg.workCounter, fuseL, fuseP, fuseCap, cfg.bvh, nullptr);
      else kTraceWarpQ<true, true, false, false><<<traceGrid(c9, n, TR_THREADS), TR_THREADS, traceWarpQSmemBytes(cfg.maxStack), st>>>(q, cnt, n, sc, O, D, nullptr, occl, cfg.workCounter, nullptr, nullptr, 0u, cfg.bvh, nullptr);
      return;
   }
   if (cfg.variant >= 3) {
      const uint32_t g = traceGrid(fuseL ? cfg : c9, n, TR_THREADS); const size_t sm = traceWarpQSmemBytes(cfg.maxStack);   // the fused instantiation takes 64 registers: 8 CTAs per SM
      if (cfg.countStats && cfg.travCounters) {
         if (fuseL) kTraceWarpQ<true, false, true, true><<<g, TR_THREADS, sm, st>>>(q, cnt, n, sc, O, D, nullptr, occl, cfg.workCounter, fuseL, fuseP, fuseCap, cfg.bvh, cfg.travCounters + 3);
         else kTraceWarpQ<true, false, true, false><<<g, TR_THREADS, sm, st>>>(q, cnt, n, sc, O, D, nullptr, occl, cfg.workCounter, nullptr, nullptr, 0u, cfg.bvh, cfg.travCounters + 3);
      } else if (fuseL) kTraceWarpQ<true, false, false, true><<<g, TR_THREADS, sm, st>>>(q, cnt, n, sc, O, D, nullptr, occl, cfg.workCounter, fuseL, fuseP, fuseCap, cfg.bvh, nullptr);
      else kTraceWarpQ<true, false, false, false><<<g, TR_THREADS, sm, st>>>(q, cnt, n, sc, O, D, nullptr, occl, cfg.workCounter, nullptr, nullptr, 0u, cfg.bvh, nullptr);
      return;
   }
   kTracePersistent<true><<<traceGrid(c9, n, TR_THREADS), TR_THREADS, traceSmemBytes(cfg.maxStack), st>>>(q, cnt, n, sc, O, D, nullptr, occl, cfg.workCounter, fuseL, fuseP, fuseCap);

}
static inline void launchTraceStats(TraceConfig &cfg, cudaStream_t st, uint32_t n, const DScene *sc, const F4 *O, const F4 *D, F4 *hit, uint32_t *nodes, uint32_t *prims) {
   uint32_t need = (n + 127) / 128;
   uint32_t full = (uint32_t)cfg.sms * 8u;
   kTraceStats<<<need < full ? (need ? need : 1) : full, 128, 0, st>>>(n, sc, O, D, hit, nodes, prims);
}

}  // namespace bl
