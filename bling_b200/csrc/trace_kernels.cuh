// trace_kernels.cuh -- K2 trace_nearest / K3 trace_any: BVH traversal kernels for sm_100a.
// Replaces the per-ray recursion of KdTree.hs:210-246 (a4, a5 in SURVEY.md §8a).
//
// variant 0  one ray per thread, grid-stride, traversal stack in local memory (the reference point)
// variant 1  persistent warps, one ray per QUAD, majority-vote stepping (kTraceQuad below; the product path)
// B200 has no RT cores and traversal is not a contraction: no tensor cores here. The bound is L2/HBM latency
// and bandwidth on the node/triangle fetches (DESIGN.md "Roofline").
#pragma once
#include "bodies.h"
#include <cuda_runtime.h>

namespace bl {

struct TraceConfig {
   int sms = 148;
   int variant = 1;
   int blocksPerSm = 10;
   int maxStack = 64;                 // worst-case stack entries of the uploaded tree (Bvh::max_stack)
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
// Persistent warps, ONE RAY PER QUAD (4 lanes), majority-vote stepping.
//
// History (profiles/): v1 let every lane run its own while loop -> 3.5 of 32 threads active per instruction. v2 kept
// the warp converged by voting between "node step" and "leaf step" -> 2.1x, but with one ray per lane every node
// fetch costs one L1 wavefront per lane per 16-byte load (7 LDG.128 x lanes): ncu showed the LSU data pipe at 81 %
// (l1tex__data_pipe_lsu_wavefronts) with DRAM at 12 %. Here the four lanes of a quad own one ray and each lane
// tests ONE child of the 4-wide node: the quad's four 32-byte loads fall into one 128-byte line = one wavefront per
// node visit, the box test costs a quarter of the instructions per lane, and leaf items are tested four at a time.
//   - the quad's traversal stack lives in shared memory ([level][quad], sized from the builder's worst case);
//     a node step pushes every hit child, nearest on top (rank by three quad shuffles of the sort key), then pops;
//   - every trip the warp votes (ballot) whether more quads stand at a node or at a leaf and executes only that
//     step; idle quads refill from the global queue with a warp-aggregated atomic once enough quads have retired.
#define TR_THREADS 128
#define TR_QUADS (TR_THREADS / 4)
#define TR_REFILL_Q 2     // refill once this many quads of the warp are idle

__device__ __forceinline__ void ld8(const F4 *p, F4 &a, F4 &b) {   // one 32-byte load (LDG.E.256 on sm_100)
   asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p));
}

template <bool ANY>
__global__ void __launch_bounds__(TR_THREADS, 10) kTraceQuad(const uint32_t *__restrict__ q, const uint32_t *__restrict__ cnt, uint32_t n,
                                                           const DScene *__restrict__ sc, const F4 *__restrict__ O, const F4 *__restrict__ D,
                                                           F4 *__restrict__ hit, uint8_t *__restrict__ occl, uint32_t *__restrict__ work) {
   extern __shared__ int qstack[];            // [level][TR_QUADS]
   const unsigned FULL = 0xffffffffu;
   const uint32_t total = cnt ? *cnt : n;
   const Bvh bvh = sc->bvh;
   const unsigned lane = threadIdx.x & 31u;
   const unsigned sub = lane & 3u;            // which child / leaf item this lane tests
   const unsigned qshift = lane & ~3u;        // first lane of my quad
   const unsigned qmask = 0xfu << qshift;
   int *const myStack = qstack + (threadIdx.x >> 2);
   const int EMPTY = (int)0x80000000;         // quad holds no ray
   int cur = EMPTY, sp = 0;
   uint32_t slot = 0;
   Ray r; RayPre pre; int hPrim = -1; float hB1 = 0, hB2 = 0;
   bool exhausted = false;
   r.o = mk3(0, 0, 0); r.d = mk3(0, 0, 1); r.tmin = 0; r.tmax = 0; pre.idir = mk3(0, 0, 0); pre.ood = mk3(0, 0, 0);

#define TQ_FINISH(found_) do { if (sub == 0) { if (ANY) occl[slot] = (found_) ? 1 : 0; else { F4 v_; v_.x = hPrim >= 0 ? r.tmax : 0.0f; v_.y = hB1; v_.z = hB2; v_.w = i2f(hPrim); hit[slot] = v_; } } cur = EMPTY; } while (0)
#define TQ_POP() do { if (sp == 0) TQ_FINISH(false); else { sp--; cur = myStack[sp * TR_QUADS]; } } while (0)

   for (;;) {
      // ---- refill idle quads (warp-uniform decision)
      unsigned idle = __ballot_sync(FULL, cur == EMPTY) & 0x11111111u;   // one bit per quad
      if (!exhausted && __popc(idle) >= TR_REFILL_Q) {
         uint32_t base = 0;
         int leader = __ffs(idle) - 1;
         if ((int)lane == leader) base = atomicAdd(work, (uint32_t)__popc(idle));
         base = __shfl_sync(FULL, base, leader);
         if (cur == EMPTY) {
            uint32_t k = base + __popc(idle & ((1u << qshift) - 1u));
            if (k < total) {
               slot = q ? q[k] : k;
               r = loadRay(O, D, slot);
               pre = rayPre(r);
               hPrim = -1; hB1 = 0; hB2 = 0;
               sp = 0;
               cur = bvh.root;
               if (bvh.root < 0) TQ_FINISH(false);   // empty scene
            }
         }
         if (base + (uint32_t)__popc(idle) >= total) exhausted = true;   // warp-uniform: the queue is drained
         idle = __ballot_sync(FULL, cur == EMPTY) & 0x11111111u;
      }
      if (idle == 0x11111111u) { if (exhausted) break; continue; }
      // ---- vote: node step or leaf step
      const bool atNode = cur >= 0, atLeaf = cur < 0 && cur != EMPTY;
      const unsigned mN = __ballot_sync(FULL, atNode), mL = __ballot_sync(FULL, atLeaf);
      if (__popc(mN) >= __popc(mL)) {
         if (atNode) {
            F4 a, b;
            ld8(bvh.nodes + BL_NODE_F4 * (size_t)cur + 2 * sub, a, b);
            const uint32_t key = childKey(a, b, r, pre, sub);
            const bool hitc = key != 0xffffffffu;
            const uint32_t k1 = __shfl_xor_sync(qmask, key, 1), k2 = __shfl_xor_sync(qmask, key, 2), k3 = __shfl_xor_sync(qmask, key, 3);
            const int nh = (int)hitc + (int)(k1 != 0xffffffffu) + (int)(k2 != 0xffffffffu) + (int)(k3 != 0xffffffffu);
            if (hitc) {
               const int rank = (int)(k1 < key) + (int)(k2 < key) + (int)(k3 < key);   // 0 = nearest (keys are distinct)
               myStack[(sp + nh - 1 - rank) * TR_QUADS] = f2i(b.z);
            }
            sp += nh;
            __syncwarp(qmask);
            TQ_POP();
         }
      } else {
         if (atLeaf) {
            const int enc = ~cur; const int first = enc >> 4, cntl = enc & 15;
            bool found = false;
            for (int base = 0; base < cntl; base += 4) {   // quad-uniform trip count (leaves hold <= 4 items by default)
               const int it = base + (int)sub;
               bool hitp = false; float t = BL_INF, b1 = 0, b2 = 0; int prim = -1;
               if (it < cntl) {
                  if (ANY) hitp = leafItemAny(bvh, first + it, r);
                  else { HitRec hh; hh.t = 0; hh.prim = -1; hh.b1 = hh.b2 = 0; Ray rr = r; hitp = leafItemNearest(bvh, first + it, rr, hh); if (hitp) { t = hh.t; b1 = hh.b1; b2 = hh.b2; prim = hh.prim; } }
               }
               const unsigned hb = __ballot_sync(qmask, hitp) & qmask;
               if (hb) {
                  if (ANY) { found = true; break; }
                  // nearest of the quad's hits; on equal t the later item wins, as in the sequential fold (Primitive.hs:29-43)
                  float tt = t; unsigned w = sub;
                  { float o = __shfl_xor_sync(qmask, tt, 1); unsigned ow = sub ^ 1u; if (o < tt || (o == tt && ow > w)) { tt = o; w = ow; } }
                  { float o = __shfl_xor_sync(qmask, tt, 2); unsigned ow = __shfl_xor_sync(qmask, w, 2); if (o < tt || (o == tt && ow > w)) { tt = o; w = ow; } }
                  const int src = (int)(qshift + w);
                  r.tmax = tt;
                  hPrim = __shfl_sync(qmask, prim, src); hB1 = __shfl_sync(qmask, b1, src); hB2 = __shfl_sync(qmask, b2, src);
               }
            }
            if (ANY && found) TQ_FINISH(true);
            else TQ_POP();
         }
      }
   }
#undef TQ_FINISH
#undef TQ_POP
}

static inline size_t traceSmemBytes(int maxStack) { int lv = maxStack < 16 ? 16 : maxStack; return (size_t)lv * TR_QUADS * sizeof(int); }

static inline uint32_t traceGrid(const TraceConfig &cfg, uint32_t n, uint32_t raysPerBlock) {
   uint32_t need = (n + raysPerBlock - 1) / raysPerBlock;
   uint32_t full = (uint32_t)cfg.sms * (uint32_t)cfg.blocksPerSm;
   return need < full ? (need ? need : 1) : full;
}
static inline void launchTraceNearest(TraceConfig &cfg, cudaStream_t st, const uint32_t *q, const uint32_t *cnt, uint32_t n, const DScene *sc,
                                      const F4 *O, const F4 *D, F4 *hit) {
   if (cfg.countStats && cfg.travCounters) { kTraceNearestCount<<<traceGrid(cfg, n, 128), 128, 0, st>>>(q, cnt, n, sc, O, D, hit, cfg.travCounters); return; }
   if (cfg.variant == 0) { kTraceNearestSimple<<<traceGrid(cfg, n, 128), 128, 0, st>>>(q, cnt, n, sc, O, D, hit); return; }
   if (!cfg.workCounter) cudaMalloc(&cfg.workCounter, sizeof(uint32_t));
   cudaMemsetAsync(cfg.workCounter, 0, sizeof(uint32_t), st);
   kTraceQuad<false><<<traceGrid(cfg, n, TR_QUADS), TR_THREADS, traceSmemBytes(cfg.maxStack), st>>>(q, cnt, n, sc, O, D, hit, nullptr, cfg.workCounter);
}
static inline void launchTraceAny(TraceConfig &cfg, cudaStream_t st, const uint32_t *q, const uint32_t *cnt, uint32_t n, const DScene *sc,
                                  const F4 *O, const F4 *D, uint8_t *occl) {
   if (cfg.variant == 0) { kTraceAnySimple<<<traceGrid(cfg, n, 128), 128, 0, st>>>(q, cnt, n, sc, O, D, occl); return; }
   if (!cfg.workCounter) cudaMalloc(&cfg.workCounter, sizeof(uint32_t));
   cudaMemsetAsync(cfg.workCounter, 0, sizeof(uint32_t), st);
   kTraceQuad<true><<<traceGrid(cfg, n, TR_QUADS), TR_THREADS, traceSmemBytes(cfg.maxStack), st>>>(q, cnt, n, sc, O, D, nullptr, occl, cfg.workCounter);
}
static inline void launchTraceStats(TraceConfig &cfg, cudaStream_t st, uint32_t n, const DScene *sc, const F4 *O, const F4 *D, F4 *hit, uint32_t *nodes, uint32_t *prims) {
   uint32_t need = (n + 127) / 128;
   uint32_t full = (uint32_t)cfg.sms * 8u;
   kTraceStats<<<need < full ? (need ? need : 1) : full, 128, 0, st>>>(n, sc, O, D, hit, nodes, prims);
}

}  // namespace bl
