// trace_kernels.cuh -- K2 trace_nearest / K3 trace_any: BVH traversal kernels for sm_100a.
// Replaces the per-ray recursion of KdTree.hs:210-246 (a4, a5 in SURVEY.md §8a).
//
// variant 0  one ray per thread, grid-stride, traversal stack in local memory (the reference point)
// variant 1  persistent threads: grid = SMs x blocksPerSm, rays pulled from a global work counter with a
//            warp-aggregated atomic; lanes that finish refill themselves once the warp's live-lane count
//            drops below a threshold (ballot compaction); traversal stack lives in shared memory
//            ([level][thread], conflict-free) with a local-memory tail; nodes are fetched as 4 x LDG.128,
//            leaf items as 3 x LDG.128 through the read-only path.
// B200 has no RT cores and traversal is not a contraction: no tensor cores here. The bound is L2/HBM latency
// and bandwidth on the node/triangle fetches (DESIGN.md "Roofline").
#pragma once
#include "bodies.h"
#include <cuda_runtime.h>

namespace bl {

struct TraceConfig {
   int sms = 148;
   int variant = 1;
   int blocksPerSm = 8;
   uint32_t *workCounter = nullptr;   // device, one uint32 per launch slot
   bool countStats = false;           // option "traversal_stats": nearest-hit launches count node fetches / primitive tests
   unsigned long long *travCounters = nullptr;   // device: nodes, prims, rays
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
// Persistent warps with MAJORITY-VOTE stepping. The first version of this kernel let every lane run its own
// data-dependent while loop; ncu (profiles/r01_trace_v1_divergent.md) showed 3.5 of 32 threads active per issued
// instruction. Here the warp stays converged: every trip the lanes vote (ballot) whether more of them stand at an
// inner node or inside a leaf, and the whole warp executes ONLY that step, predicated per lane. At least half of
// the live lanes are active in every trip; idle lanes are refilled from the global queue (warp-aggregated atomic)
// once enough of them have retired.
#define TR_THREADS 128
#define TR_SSTACK 24      // shared-memory stack levels per thread
#define TR_LSTACK 40      // local-memory tail (tree depth is bounded by the builder: 32 SAH + 24 median levels)
#define TR_REFILL 8       // refill once this many lanes are idle

template <bool ANY>
__global__ void __launch_bounds__(TR_THREADS) kTracePersistent(const uint32_t *__restrict__ q, const uint32_t *__restrict__ cnt, uint32_t n,
                                                              const DScene *__restrict__ sc, const F4 *__restrict__ O, const F4 *__restrict__ D,
                                                              F4 *__restrict__ hit, uint8_t *__restrict__ occl, uint32_t *__restrict__ work) {
   __shared__ int sstack[TR_SSTACK][TR_THREADS];
   int lstack[TR_LSTACK];
   const unsigned FULL = 0xffffffffu;
   const uint32_t total = cnt ? *cnt : n;
   const Bvh bvh = sc->bvh;
   const unsigned lane = threadIdx.x & 31u;
   const int tid = threadIdx.x;
   const int EMPTY = (int)0x80000000;   // lane holds no ray (negative: never mistaken for a node index)
   const int LEAF = -1;             // lane iterates the items [li, le) of a leaf
   int cur = EMPTY, sp = 0, li = 0, le = 0;
   uint32_t slot = 0;
   Ray r; V3 idir; HitRec h;
   bool exhausted = false;
   r.o = mk3(0, 0, 0); r.d = mk3(0, 0, 1); r.tmin = 0; r.tmax = 0; idir = mk3(0, 0, 0); h.t = 0; h.prim = -1; h.b1 = h.b2 = 0;

   // enter child reference c (node index or encoded leaf)
#define TR_ENTER(c) do { int c_ = (c); if (c_ >= 0) cur = c_; else { int enc_ = ~c_; li = enc_ >> 4; le = li + (enc_ & 15); cur = LEAF; } } while (0)
   // ray finished: write the result, free the lane
#define TR_FINISH(found_) do { if (ANY) occl[slot] = (found_) ? 1 : 0; else { F4 v_; v_.x = h.t; v_.y = h.b1; v_.z = h.b2; v_.w = i2f(h.prim); hit[slot] = v_; } cur = EMPTY; } while (0)
#define TR_POP() do { if (sp == 0) TR_FINISH(false); else { sp--; int c2_ = (sp < TR_SSTACK) ? sstack[sp][tid] : lstack[sp - TR_SSTACK]; TR_ENTER(c2_); } } while (0)

   for (;;) {
      // ---- refill idle lanes (warp-uniform decision)
      unsigned idle = __ballot_sync(FULL, cur == EMPTY);
      if (!exhausted && __popc(idle) >= TR_REFILL) {
         uint32_t base = 0;
         int leader = __ffs(idle) - 1;
         if ((int)lane == leader) base = atomicAdd(work, (uint32_t)__popc(idle));
         base = __shfl_sync(FULL, base, leader);
         if (cur == EMPTY) {
            uint32_t k = base + __popc(idle & ((1u << lane) - 1u));
            if (k < total) {
               slot = q ? q[k] : k;
               r = loadRay(O, D, slot);
               idir = mk3(1.0f / r.d.x, 1.0f / r.d.y, 1.0f / r.d.z);
               h.t = 0; h.prim = -1; h.b1 = 0; h.b2 = 0;
               sp = 0;
               if (bvh.root >= 0) cur = bvh.root; else { li = le = 0; cur = LEAF; }   // empty scene: a leaf with zero items
            }
         }
         if (base + (uint32_t)__popc(idle) >= total) exhausted = true;   // warp-uniform: the queue is drained
         idle = __ballot_sync(FULL, cur == EMPTY);
      }
      if (idle == FULL) { if (exhausted) break; continue; }
      // ---- vote: node step or leaf step
      const bool atNode = cur >= 0, atLeaf = cur == LEAF;
      const unsigned mN = __ballot_sync(FULL, atNode), mL = __ballot_sync(FULL, atLeaf);
      if (__popc(mN) >= __popc(mL)) {
         if (atNode) {
            const F4 *np = bvh.nodes + 4 * (size_t)cur;
            F4 n0 = ld4(np), n1 = ld4(np + 1), n2 = ld4(np + 2), n3 = ld4(np + 3);
            float tn0, tn1; bool h0, h1;
            nodeTest(n0, n1, n2, r, idir, tn0, tn1, h0, h1);
            int c0 = f2i(n3.x), c1 = f2i(n3.y);
            if (h0 && h1) {
               if (!ANY && tn1 < tn0) { int t = c0; c0 = c1; c1 = t; }
               if (sp < TR_SSTACK) sstack[sp][tid] = c1; else if (sp < TR_SSTACK + TR_LSTACK) lstack[sp - TR_SSTACK] = c1;
               sp++;
               TR_ENTER(c0);
            } else if (h0) TR_ENTER(c0);
            else if (h1) TR_ENTER(c1);
            else TR_POP();
         }
      } else {
         if (atLeaf) {
            bool found = false;
            if (li < le) {
               if (ANY) found = leafItemAny(bvh, li, r);
               else leafItemNearest(bvh, li, r, h);
               li++;
            }
            if (ANY && found) TR_FINISH(true);
            else if (li >= le) TR_POP();
         }
      }
   }
#undef TR_ENTER
#undef TR_FINISH
#undef TR_POP
}

static inline void launchTraceNearest(TraceConfig &cfg, cudaStream_t st, const uint32_t *q, const uint32_t *cnt, uint32_t n, const DScene *sc,
                                      const F4 *O, const F4 *D, F4 *hit) {
   uint32_t need = (n + TR_THREADS - 1) / TR_THREADS;
   uint32_t full = (uint32_t)cfg.sms * (uint32_t)cfg.blocksPerSm;
   uint32_t grid = need < full ? (need ? need : 1) : full;
   if (cfg.countStats && cfg.travCounters) { kTraceNearestCount<<<grid, 128, 0, st>>>(q, cnt, n, sc, O, D, hit, cfg.travCounters); return; }
   if (cfg.variant == 0) { kTraceNearestSimple<<<grid, 128, 0, st>>>(q, cnt, n, sc, O, D, hit); return; }
   if (!cfg.workCounter) cudaMalloc(&cfg.workCounter, sizeof(uint32_t));
   cudaMemsetAsync(cfg.workCounter, 0, sizeof(uint32_t), st);
   kTracePersistent<false><<<grid, TR_THREADS, 0, st>>>(q, cnt, n, sc, O, D, hit, nullptr, cfg.workCounter);
}
static inline void launchTraceAny(TraceConfig &cfg, cudaStream_t st, const uint32_t *q, const uint32_t *cnt, uint32_t n, const DScene *sc,
                                  const F4 *O, const F4 *D, uint8_t *occl) {
   uint32_t need = (n + TR_THREADS - 1) / TR_THREADS;
   uint32_t full = (uint32_t)cfg.sms * (uint32_t)cfg.blocksPerSm;
   uint32_t grid = need < full ? (need ? need : 1) : full;
   if (cfg.variant == 0) { kTraceAnySimple<<<grid, 128, 0, st>>>(q, cnt, n, sc, O, D, occl); return; }
   if (!cfg.workCounter) cudaMalloc(&cfg.workCounter, sizeof(uint32_t));
   cudaMemsetAsync(cfg.workCounter, 0, sizeof(uint32_t), st);
   kTracePersistent<true><<<grid, TR_THREADS, 0, st>>>(q, cnt, n, sc, O, D, nullptr, occl, cfg.workCounter);
}
static inline void launchTraceStats(TraceConfig &cfg, cudaStream_t st, uint32_t n, const DScene *sc, const F4 *O, const F4 *D, F4 *hit, uint32_t *nodes, uint32_t *prims) {
   uint32_t need = (n + 127) / 128;
   uint32_t full = (uint32_t)cfg.sms * 8u;
   kTraceStats<<<need < full ? (need ? need : 1) : full, 128, 0, st>>>(n, sc, O, D, hit, nodes, prims);
}

}  // namespace bl
