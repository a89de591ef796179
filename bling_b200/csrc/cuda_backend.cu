// cuda_backend.cu -- the product: sm_100a kernels + the C ABI (libblingcu.so).
//
// Kernels:
//   K1 raygen, K4 classify, K5 shade (miss + one launch per material kind present), K6 resolve, K7 film,
//   finalize, advance                    -> kRun / kRunQueue over the functors of bodies.h
//   K2 trace_nearest, K3 trace_any       -> trace_kernels.cuh (persistent-thread BVH traversal)
// All queue-driven kernels read their item count from device memory, so the whole bounce loop is enqueued
// without a host round trip. Grids are sized in multiples of the SM count.
#include "api_impl.h"
#include "trace_kernels.cuh"
#include "comm.h"
#include <cuda_runtime.h>
#include <algorithm>
#include <cstring>
#include <thread>
#include <utility>
#include <vector>

namespace bl {

struct CudaError { cudaError_t e; const char *what; };
#define CU(x) do { cudaError_t _e = (x); if (_e != cudaSuccess) throw CudaError{_e, #x}; } while (0)


template <class B> __global__ void __launch_bounds__(256) kRun(B b, uint32_t n) {
   for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) b(i);
}
template <class B> __global__ void __launch_bounds__(256) kRunQueue(B b, const uint32_t *__restrict__ q, const uint32_t *__restrict__ cnt) {
   uint32_t n = *cnt;
   for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) b(q[i]);
}
// the shade kernel carries 16-band spectra in registers: smaller blocks, one item per thread per trip
#ifndef SH_MINBLOCKS
#define SH_MINBLOCKS 3
#endif
// Resident CTAs per SM the heavy kernels are compiled for: SH_MINBLOCKS (3: up to 168 registers), or 4 (128 registers) for the
// hit-shading kernels of the material kinds in SH_KINDS4 (bit k = kind k): given 168 registers ptxas takes 160-162 for the plastic
// and translucent-matte bodies, given 128 it needs 127 / 125 and still does not spill (cuobjdump -res-usage).
#ifndef SH_KINDS4
#define SH_KINDS4 0x88    // plastic, translucent matte: measured on the B200 (tools/gpu_r02_z10.sh) ducky +2.4 %, environment +2.8 %, sun-sky +11.2 %, films bit-identical
#endif
#ifndef SH_KINDS5
#define SH_KINDS5 0       // A/B: kinds compiled for FIVE CTAs per SM (96 registers). Plastic alone (0x08) is faster in isolation (ducky +3.1 %,
                          // environment +1.9 %, tools/gpu_r02_z12.sh) but spills: its stack frame grows from 280 to 392 bytes, the largest of any
                          // kernel in a pass, and inside the full bench.py run (the cfg-5 context with its 33 GB of path state alive beside it)
                          // sun-sky then came back at 660 instead of 935 Msamples/s, twice (tools/gpu_r02_final3.sh; cause not established --
                          // the round's GPU time ended there; with plastic at four CTAs the full run measures 952). Matte at five: cfg 5 +0.4 %, cornell-box -2.3 % (0x1F, _z11.sh). Off.
#endif
template <class B> struct HeavyBlocks { static const int v = SH_MINBLOCKS; };
template <int MK> struct HeavyBlocks<ShadeHitBody<MK>> {
   static const int v = (MK >= 0 && MK < 16 && ((SH_KINDS5 >> MK) & 1)) ? 5 : ((MK >= 0 && MK < 16 && ((SH_KINDS4 >> MK) & 1)) ? 4 : SH_MINBLOCKS);
};
template <class B> __global__ void __launch_bounds__(128, HeavyBlocks<B>::v) kRunQueueHeavy(B b, const uint32_t *__restrict__ q, const uint32_t *__restrict__ cnt) {
   uint32_t n = *cnt;
   for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) b(q[i]);
}
#ifndef SH_PREFETCH
#define SH_PREFETCH 0     // 1: the hit-shading kernels prefetch their next slot (bodies.h::ShadeHitBody::prefetchSlot); 0: plain loop.
                          // Measured (tools/gpu_r02_z3.sh): the pipelined loop is SLOWER -- shade 35.0 -> 42.0 ms per cfg-5 step, the named
                          // scenes -1 .. -7 % -- although it hides two of the three load levels: the kernels sit on the request rate
                          // of their scattered accesses (profiles/r01_shade_experiments.md), and eight prefetches plus two early loads
                          // per item are eight more requests. Kept as an A/B build only.
#endif
// The same loop, software-pipelined over the items of one thread: at the top of iteration k the thread issues the loads that
// ADDRESS iteration k + 1 (its hit reference, and the queue entry after it) and prefetches that slot's records; after the body,
// when the hit reference has arrived, it prefetches the geometry behind it.
template <int MK> __global__ void __launch_bounds__(128, SH_MINBLOCKS) kRunQueueShade(ShadeHitBody<MK> b, const uint32_t *__restrict__ q, const uint32_t *__restrict__ cnt) {
   const uint32_t n = *cnt, stride = gridDim.x * blockDim.x;
   uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= n) return;
   uint32_t slot = q[i];
   uint32_t next = (i + stride < n) ? q[i + stride] : slot;
   for (;;) {
      const bool more = i + stride < n;
      int href = BL_REF_MISS; uint32_t next2 = next;
      if (more) {
         href = b.peekHit(next);
         if (i + 2 * stride < n) next2 = q[i + 2 * stride];
         b.prefetchSlot(next);
      }
      b(slot);
      if (!more) break;
      b.prefetchSurface(href);
      i += stride; slot = next; next = next2;
   }
}

// K7 film, tiled: one CTA per 16x16 film-pixel tile. For every sample index of the batch the samples of all sample pixels
// that can reach the tile ((16 + 2R + 1)^2 of them) are staged in shared memory once, then every thread gathers its own
// (2R + 1)^2 neighbourhood from there: each sample is read from L2 once per tile instead of once per reached pixel.
// Atomic-free; same per-sample arithmetic as FilmBody (bodies.h), summed sample-index-major.
#define FT_TILE 16
#define FT_MAXC 24   // staged cells per axis: 16 + floor(f + 0.5) + ceil(f + 0.5) + 1 <= 24  <=>  filter radius <= 3.5
__global__ void __launch_bounds__(FT_TILE * FT_TILE) kFilmTile(const DScene *__restrict__ sc, PathState ps, F4 *__restrict__ film, uint32_t k, uint32_t npix) {
   __shared__ F4 sXyz[FT_MAXC * FT_MAXC];
   __shared__ F2 sPos[FT_MAXC * FT_MAXC];
   const DScene &S = *sc;
   const FilmGeom g = filmGeom(S);
   const int tilesX = (S.W + FT_TILE - 1) / FT_TILE;
   const int X0 = (int)(blockIdx.x % (unsigned)tilesX) * FT_TILE, Y0 = (int)(blockIdx.x / (unsigned)tilesX) * FT_TILE;
   const int tx = (int)threadIdx.x % FT_TILE, ty = (int)threadIdx.x / FT_TILE;
   const int x = X0 + tx, y = Y0 + ty;
   const bool inside = x < S.W && y < S.H;
   // cells staged for the whole tile, and this thread's own neighbourhood inside them
   const int cx0 = filmCellLo(X0, g.fw), cy0 = filmCellLo(Y0, g.fh);
   const int ncx = filmCellHi(X0 + FT_TILE - 1, g.fw) - cx0 + 1, ncy = filmCellHi(Y0 + FT_TILE - 1, g.fh) - cy0 + 1;
   const int ixlo = imax(S.ex0, filmCellLo(x, g.fw)), ixhi = imin(S.ex1, filmCellHi(x, g.fw));
   const int iylo = imax(S.ey0, filmCellLo(y, g.fh)), iyhi = imin(S.ey1, filmCellHi(y, g.fh));
   float aw = 0, ax = 0, ay = 0, az = 0;
   for (uint32_t sl = 0; sl < k; ++sl) {
      for (int c = (int)threadIdx.x; c < ncx * ncy; c += FT_TILE * FT_TILE) {
         const int ix = cx0 + c % ncx, iy = cy0 + c / ncx;
         F4 v; v.x = v.y = v.z = v.w = 0; F2 p; p.x = p.y = 0;
         if (ix >= S.ex0 && ix <= S.ex1 && iy >= S.ey0 && iy <= S.ey1) {
            const uint32_t slot = sl * npix + (uint32_t)(iy - S.ey0) * (uint32_t)S.EW + (uint32_t)(ix - S.ex0);
            v = ps.xyz[slot]; p = ps.spos[slot];
         }
         sXyz[c] = v; sPos[c] = p;
      }
      __syncthreads();
      if (inside) {
         for (int iy = iylo; iy <= iyhi; ++iy) {
            int toy, tymax; filmTileSpan(iy, S.ey0, S.ey1, g.exty, toy, tymax);
            if (y < toy || y > tymax) continue;
            for (int ix = ixlo; ix <= ixhi; ++ix) {
               int tox, txmax; filmTileSpan(ix, S.ex0, S.ex1, g.extx, tox, txmax);
               if (x < tox || x > txmax) continue;
               const int c = (iy - cy0) * ncx + (ix - cx0);
               filmAddSample(S, g, x, y, tox, txmax, toy, tymax, sXyz[c], sPos[c], aw, ax, ay, az);
            }
         }
      }
      __syncthreads();
   }
   if (inside) {
      const size_t fp = (size_t)y * S.W + x;
      F4 f = film[fp];
      f.x = f.x + aw; f.y = f.y + ax; f.z = f.z + ay; f.w = f.w + az;
      film[fp] = f;
   }
}

struct CudaBackend {
   int device = -1, sms = 148;
   cudaStream_t stream = nullptr, ownStream = nullptr;
   bool timerPending = false;
   cudaEvent_t ev0 = nullptr, ev1 = nullptr;
   TraceConfig tcfg;
   // optional per-class kernel timing (option "profile_kernels"): one event pair per launch on the launching stream
   bool profile = false; int curClass = BLINGCU_KC_OTHER;
   struct Rec { cudaEvent_t a, b; int cls; };
   std::vector<Rec> recs; std::vector<cudaEvent_t> pool;
   double clsMs[BLINGCU_KC_COUNT] = {}; uint64_t clsLaunches[BLINGCU_KC_COUNT] = {};

   void tag(int c) { curClass = c; }
   cudaEvent_t getEvent() { if (!pool.empty()) { cudaEvent_t e = pool.back(); pool.pop_back(); return e; } cudaEvent_t e; CU(cudaEventCreate(&e)); return e; }
   struct Scope {
      CudaBackend *b; Rec r; bool on;
      Scope(CudaBackend *be) : b(be), on(be->profile) { if (on) { r.a = b->getEvent(); r.b = b->getEvent(); r.cls = b->curClass; cudaEventRecord(r.a, b->stream); } else b->clsLaunches[b->curClass]++; }
      ~Scope() { if (on) { cudaEventRecord(r.b, b->stream); b->recs.push_back(r); if (b->recs.size() > 4096) b->collect(); } }
   };
   void collect() {
      if (recs.empty()) return;
      cudaStreamSynchronize(stream);
      for (Rec &r : recs) { float ms = 0; if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) { clsMs[r.cls] += ms; clsLaunches[r.cls]++; } pool.push_back(r.a); pool.push_back(r.b); }
      recs.clear();
   }
   void kernelTimes(double *ms, uint64_t *l, int n) { collect(); for (int i = 0; i < n; ++i) { ms[i] = i < BLINGCU_KC_COUNT ? clsMs[i] : 0; l[i] = i < BLINGCU_KC_COUNT ? clsLaunches[i] : 0; } }
   void resetProfile() { collect(); for (int i = 0; i < BLINGCU_KC_COUNT; ++i) { clsMs[i] = 0; clsLaunches[i] = 0; } if (tcfg.travCounters) cudaMemsetAsync(tcfg.travCounters, 0, 6 * sizeof(unsigned long long), stream); }
   void traversalTotals(uint64_t *six) {   // nearest-hit nodes, prims, rays; any-hit nodes, prims, rays
      for (int i = 0; i < 6; ++i) six[i] = 0;
      if (!tcfg.travCounters) return;
      unsigned long long h[6]; download(h, tcfg.travCounters, sizeof(h)); for (int i = 0; i < 6; ++i) six[i] = h[i];
   }

   int init(int dev, std::string &err) {
      int n = 0;
      cudaError_t e = cudaGetDeviceCount(&n);
      if (e != cudaSuccess || n == 0) { err = std::string("no CUDA device visible (") + cudaGetErrorString(e) + "); this library has no CPU fallback"; return BLINGCU_ENOGPU; }
      if (dev < 0 || dev >= n) { err = "device index out of range"; return BLINGCU_EINVAL; }
      if ((e = cudaSetDevice(dev)) != cudaSuccess) { err = cudaGetErrorString(e); return BLINGCU_ECUDA; }
      cudaDeviceProp p;
      if ((e = cudaGetDeviceProperties(&p, dev)) != cudaSuccess) { err = cudaGetErrorString(e); return BLINGCU_ECUDA; }
      if (p.major < 10) { err = "built for sm_100a (B200); found sm_" + std::to_string(p.major * 10 + p.minor); return BLINGCU_ENOGPU; }
      device = dev; sms = p.multiProcessorCount;
      if (cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking) != cudaSuccess || cudaEventCreate(&ev0) != cudaSuccess || cudaEventCreate(&ev1) != cudaSuccess) { err = "stream/event creation failed"; return BLINGCU_ECUDA; }
      ownStream = stream; tcfg.sms = sms;
      // the traversal stack lives in dynamic shared memory: allow the builder's worst case (BL_STACK levels)
      cudaFuncSetAttribute(kTracePersistent<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)traceSmemBytes(BL_STACK));
      cudaFuncSetAttribute(kTracePersistent<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)traceSmemBytes(BL_STACK));
      const int wq = (int)traceWarpQSmemBytes(BL_STACK);
      cudaFuncSetAttribute(kTraceWarpQ<false, true, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, wq);
      cudaFuncSetAttribute(kTraceWarpQ<false, true, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, wq);
      cudaFuncSetAttribute(kTraceWarpQ<true, true, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, wq);
      cudaFuncSetAttribute(kTraceWarpQ<true, true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, wq);
      cudaFuncSetAttribute(kTraceWarpQ<true, false, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, wq);
      cudaFuncSetAttribute(kTraceWarpQ<true, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, wq);
      cudaFuncSetAttribute(kTraceWarpQ<true, false, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, wq);
      cudaFuncSetAttribute(kTraceWarpQ<true, false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, wq);
      return 0;
   }
   // run on a caller-owned stream (e.g. torch's current stream, so an NCCL all-reduce of the film orders after the
   // pass without a host sync); nullptr restores the context's own stream
   void setStream(void *s) { cudaStreamSynchronize(stream); collect(); stream = s ? (cudaStream_t)s : ownStream; }
   void shutdown() {
      if (device >= 0) cudaSetDevice(device);
      if (stream) cudaStreamSynchronize(stream);
      commDestroy();
      if (fc.stream) { cudaStreamDestroy(fc.stream); fc.stream = nullptr; }
      if (fc.rendered) { cudaEventDestroy(fc.rendered); fc.rendered = nullptr; }
      if (fc.reduced) { cudaEventDestroy(fc.reduced); fc.reduced = nullptr; }
      freeTracePipe();
      if (sIn) { cudaStreamDestroy(sIn); sIn = nullptr; } if (sOut) { cudaStreamDestroy(sOut); sOut = nullptr; }
      if (tcfg.workCounter) { cudaFree(tcfg.workCounter); tcfg.workCounter = nullptr; }
      if (tcfg.travCounters) { cudaFree(tcfg.travCounters); tcfg.travCounters = nullptr; }
      if (ownStream) { cudaStreamDestroy(ownStream); ownStream = nullptr; }
      stream = nullptr;
      collect(); for (cudaEvent_t e : pool) cudaEventDestroy(e); pool.clear();
      if (ev0) cudaEventDestroy(ev0);
      if (ev1) cudaEventDestroy(ev1);
      ev0 = ev1 = nullptr;
   }
   template <class F> int guard(std::string &err, F f) {
      try {
         CU(cudaSetDevice(device));
         int rc = f();
         cudaError_t e = cudaGetLastError();
         if (e != cudaSuccess) { err = std::string("CUDA: ") + cudaGetErrorString(e); return BLINGCU_ECUDA; }
         return rc;
      } catch (const CudaError &c) { err = std::string("CUDA: ") + cudaGetErrorString(c.e) + " in " + c.what; return BLINGCU_ECUDA; }
      catch (const std::bad_alloc &) { err = "host out of memory"; return BLINGCU_EINVAL; }
   }
   bool setOption(const std::string &k, double v) {
      if (k == "trace_variant") { tcfg.variant = (int)v; return true; }
      if (k == "profile_kernels") { collect(); profile = v != 0; return true; }
      if (k == "traversal_stats") {
         tcfg.countStats = v != 0;
         if (tcfg.countStats && !tcfg.travCounters) { tcfg.travCounters = (unsigned long long *)alloc(6 * sizeof(unsigned long long)); zero(tcfg.travCounters, 6 * sizeof(unsigned long long)); }
         return true;
      }
      if (k == "trace_blocks_per_sm") { tcfg.blocksPerSm = (int)v; return true; }
      if (k == "trace_chunk") { if (!(v >= 1024) || v > 2147483647.0) return false; traceChunk = (size_t)v; return true; }
      if (k == "copy_threads") { if (!(v >= 1) || v > 64) return false; copyThreads = (int)v; return true; }
      return false;
   }
   void setMaxStack(int m) { tcfg.maxStack = m; }
   void setBvh(const Bvh &b) { tcfg.bvh = b; }
   // ---- film reduction over ranks (comm.h). Per-process mode: comm_unique_id on one rank, comm_init on every rank. In-process
   // mode (one host thread driving several contexts, e.g. a Haskell host): comm_init_all / reduce_film_group wrap the same
   // calls in ncclGroupStart/End.
   FilmComm fc;
   int commEnsure(std::string &err) {
      if (fc.stream) return 0;
      if (cudaStreamCreateWithFlags(&fc.stream, cudaStreamNonBlocking) != cudaSuccess || cudaEventCreateWithFlags(&fc.rendered, cudaEventDisableTiming) != cudaSuccess ||
          cudaEventCreateWithFlags(&fc.reduced, cudaEventDisableTiming) != cudaSuccess) { err = "comm stream/event creation failed"; return BLINGCU_ECUDA; }
      return 0;
   }
   static int commUniqueId(uint8_t *id, std::string &err) {
      NcclApi &N = ncclApi();
      if (!N.load()) { err = N.err; return BLINGCU_EUNSUPPORTED; }
      NcclUniqueId u; int rc = N.GetUniqueId(&u);
      if (rc != kNcclSuccess) { err = N.what(rc); return BLINGCU_ECUDA; }
      std::memcpy(id, u.internal, sizeof(u.internal));
      return 0;
   }
   static int groupStart(std::string &err) { NcclApi &N = ncclApi(); if (!N.load()) { err = N.err; return BLINGCU_EUNSUPPORTED; } int rc = N.GroupStart(); if (rc != kNcclSuccess) { err = N.what(rc); return BLINGCU_ECUDA; } return 0; }
   static int groupEnd(std::string &err) { NcclApi &N = ncclApi(); int rc = N.GroupEnd(); if (rc != kNcclSuccess) { err = N.what(rc); return BLINGCU_ECUDA; } return 0; }
   int commInit(int rank, int nranks, const uint8_t *id, std::string &err) {
      if (nranks < 1 || rank < 0 || rank >= nranks) { err = "comm_init: rank out of range"; return BLINGCU_EINVAL; }
      commDestroy();
      int rc = commEnsure(err); if (rc) return rc;
      fc.rank = rank; fc.nranks = nranks;
      if (nranks == 1) return 0;   // nothing to talk to: reduce_film is a device copy
      NcclApi &N = ncclApi();
      if (!N.load()) { err = N.err; return BLINGCU_EUNSUPPORTED; }
      NcclUniqueId u; std::memcpy(u.internal, id, sizeof(u.internal));
      CU(cudaSetDevice(device));
      int nr = N.CommInitRank(&fc.comm, nranks, u, rank);
      if (nr != kNcclSuccess) { fc.comm = nullptr; fc.nranks = 1; fc.rank = 0; err = N.what(nr); return BLINGCU_ECUDA; }
      return 0;
   }
   void commDestroy() {
      if (fc.stream) cudaStreamSynchronize(fc.stream);
      if (fc.comm) { ncclApi().CommDestroy(fc.comm); fc.comm = nullptr; }
      fc.rank = 0; fc.nranks = 1; fc.pending = false;
   }
   // film_sum = sum over ranks of film (root < 0: on every rank; else on `root` only), enqueued behind everything rendered so
   // far, on the reduction's own stream
   int reduceFilm(const float *film, float *filmSum, size_t nFloats, int root, std::string &err) {
      int rc = commEnsure(err); if (rc) return rc;
      if (root >= fc.nranks) { err = "reduce_film: root out of range"; return BLINGCU_EINVAL; }
      waitReduced();   // one reduction in flight at a time (film_sum is a single buffer)
      CU(cudaEventRecord(fc.rendered, stream));
      CU(cudaStreamWaitEvent(fc.stream, fc.rendered, 0));
      if (fc.comm) {
         NcclApi &N = ncclApi();
         int nr = root < 0 ? N.AllReduce(film, filmSum, nFloats, kNcclFloat32, kNcclSum, fc.comm, fc.stream)
                           : N.Reduce(film, filmSum, nFloats, kNcclFloat32, kNcclSum, root, fc.comm, fc.stream);
         if (nr != kNcclSuccess) { err = N.what(nr); return BLINGCU_ECUDA; }
      } else CU(cudaMemcpyAsync(filmSum, film, nFloats * sizeof(float), cudaMemcpyDeviceToDevice, fc.stream));
      CU(cudaEventRecord(fc.reduced, fc.stream));
      fc.pending = true; fc.reductions++; fc.bytes += nFloats * sizeof(float);
      return 0;
   }
   // the compute stream waits (on the device) for the reduction in flight: called before anything writes `film` again
   void waitReduced() { if (fc.pending) { cudaStreamWaitEvent(stream, fc.reduced, 0); fc.pending = false; } }
   void syncComm() { if (fc.stream) cudaStreamSynchronize(fc.stream); }
   void downloadOnComm(void *d, const void *s, size_t n) { CU(cudaMemcpyAsync(d, s, n, cudaMemcpyDeviceToHost, fc.stream ? fc.stream : stream)); CU(cudaStreamSynchronize(fc.stream ? fc.stream : stream)); }

   // ---- explicit ray batches on HOST buffers (blingcu_trace_nearest / _occluded), pipelined: the batch is cut into chunks and
   // three of them are in flight at any time -- chunk k+1 goes host -> device on the copy-in stream while chunk k is traced on the
   // context's stream and chunk k-1 goes device -> host on the copy-out stream (PCIe is full duplex). Buffers that are already
   // page-locked (blingcu_host_alloc, cudaHostRegister, torch pin_memory) are used as they are; pageable ones are staged through
   // the context's own pinned ring by a few helper threads (one memcpy thread moves ~10 GB/s, the link 50).
   enum { TB_SLOTS = 3 };
   struct TraceSlot { void *hIn = nullptr, *hOut = nullptr; F4 *dR = nullptr, *dH = nullptr; uint8_t *dC = nullptr; cudaEvent_t in = nullptr, done = nullptr, out = nullptr; };
   TraceSlot tslot[TB_SLOTS]; size_t tchunkCap = 0; cudaStream_t sIn = nullptr, sOut = nullptr;
   size_t traceChunk = 1u << 20;   // rays per chunk (option "trace_chunk")
   int copyThreads = 4;            // helper threads for pageable host buffers (option "copy_threads")
   void freeTracePipe() {
      for (TraceSlot &t : tslot) {
         if (t.hIn) cudaFreeHost(t.hIn); if (t.hOut) cudaFreeHost(t.hOut);
         cudaFree(t.dR); cudaFree(t.dH); cudaFree(t.dC);
         if (t.in) cudaEventDestroy(t.in); if (t.done) cudaEventDestroy(t.done); if (t.out) cudaEventDestroy(t.out);
         t = TraceSlot{};
      }
      tchunkCap = 0;
   }
   void ensureTracePipe(size_t chunk) {
      if (!sIn) { CU(cudaStreamCreateWithFlags(&sIn, cudaStreamNonBlocking)); CU(cudaStreamCreateWithFlags(&sOut, cudaStreamNonBlocking)); }
      if (chunk <= tchunkCap) return;
      freeTracePipe();
      for (TraceSlot &t : tslot) {
         CU(cudaMallocHost(&t.hIn, chunk * sizeof(blingcu_ray))); CU(cudaMallocHost(&t.hOut, chunk * sizeof(blingcu_hit)));
         CU(cudaMalloc(&t.dR, chunk * 2 * sizeof(F4)));
         CU(cudaMalloc(&t.dH, chunk * sizeof(F4))); CU(cudaMalloc(&t.dC, chunk));
         CU(cudaEventCreateWithFlags(&t.in, cudaEventDisableTiming)); CU(cudaEventCreateWithFlags(&t.done, cudaEventDisableTiming)); CU(cudaEventCreateWithFlags(&t.out, cudaEventDisableTiming));
      }
      tchunkCap = chunk;
   }
   static bool isPinned(const void *p) {
      cudaPointerAttributes a;
      if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
      return a.type == cudaMemoryTypeHost;
   }
   void parallelCopy(void *dst, const void *src, size_t n) {
      const int nt = (n < (4u << 20) || copyThreads <= 1) ? 1 : copyThreads;
      if (nt == 1) { std::memcpy(dst, src, n); return; }
      std::vector<std::thread> th;
      const size_t per = (n / nt + 4095) & ~(size_t)4095;
      for (int i = 1; i < nt; ++i) { const size_t b = std::min(n, per * i), e = std::min(n, per * (i + 1)); if (e > b) th.emplace_back([=]() { std::memcpy((char *)dst + b, (const char *)src + b, e - b); }); }
      std::memcpy(dst, src, std::min(n, per));
      for (std::thread &t : th) t.join();
   }
   // returns false when the backend does not take the batch (never here); hits are converted to ABI form on the device
   bool traceHostBatch(const blingcu_ray *rays, size_t n, blingcu_hit *outHit, uint8_t *outOccl, const DScene *dscene) {
      const size_t chunk = std::min(n, traceChunk);
      ensureTracePipe(chunk);
      const bool pinIn = isPinned(rays), pinOut = outHit ? isPinned(outHit) : isPinned(outOccl);
      const size_t nChunks = (n + chunk - 1) / chunk;
      const size_t outStride = outHit ? sizeof(blingcu_hit) : 1;
      auto drain = [&](size_t k) {   // chunk k has landed in its slot's staging buffer: hand it to the caller's (pageable) buffer
         TraceSlot &t = tslot[k % TB_SLOTS];
         CU(cudaEventSynchronize(t.out));
         if (!pinOut) { const size_t b = k * chunk, m = std::min(chunk, n - b); parallelCopy((char *)(outHit ? (void *)outHit : (void *)outOccl) + b * outStride, t.hOut, m * outStride); }
      };
      for (size_t k = 0; k < nChunks; ++k) {
         TraceSlot &t = tslot[k % TB_SLOTS];
         if (k >= TB_SLOTS) drain(k - TB_SLOTS);   // the slot is free again (its device buffers and staging too)
         const size_t b = k * chunk, m = std::min(chunk, n - b);
         const void *src = rays + b;
         if (!pinIn) { parallelCopy(t.hIn, src, m * sizeof(blingcu_ray)); src = t.hIn; }
         CU(cudaMemcpyAsync(t.dR, src, m * sizeof(blingcu_ray), cudaMemcpyHostToDevice, sIn));
         CU(cudaEventRecord(t.in, sIn));
         CU(cudaStreamWaitEvent(stream, t.in, 0));
         void *dsrc;
         if (outHit) {
            tag(BLINGCU_KC_TRACE_NEAREST); traceNearest(nullptr, nullptr, (uint32_t)m, dscene, t.dR, t.dR + 1, t.dH);
            tag(BLINGCU_KC_OTHER); run(HitToAbiBody{dscene, t.dH}, (uint32_t)m);
            dsrc = t.dH;
         } else { tag(BLINGCU_KC_TRACE_ANY); traceAny(nullptr, nullptr, (uint32_t)m, dscene, t.dR, t.dR + 1, t.dC); dsrc = t.dC; }
         CU(cudaEventRecord(t.done, stream));
         CU(cudaStreamWaitEvent(sOut, t.done, 0));
         void *dst = pinOut ? (void *)((char *)(outHit ? (void *)outHit : (void *)outOccl) + b * outStride) : t.hOut;
         CU(cudaMemcpyAsync(dst, dsrc, m * outStride, cudaMemcpyDeviceToHost, sOut));
         CU(cudaEventRecord(t.out, sOut));
      }
      for (size_t k = (nChunks > TB_SLOTS ? nChunks - TB_SLOTS : 0); k < nChunks; ++k) drain(k);
      return true;
   }
   void *hostAlloc(size_t n) { void *p = nullptr; CU(cudaMallocHost(&p, n ? n : 1)); return p; }
   void hostFree(void *p) { if (p) cudaFreeHost(p); }

   void *alloc(size_t n) { void *p = nullptr; CU(cudaMalloc(&p, n ? n : 1)); return p; }
   void free(void *p) { if (p) cudaFree(p); }
   void upload(void *d, const void *s, size_t n) { CU(cudaMemcpyAsync(d, s, n, cudaMemcpyHostToDevice, stream)); CU(cudaStreamSynchronize(stream)); }
   void download(void *d, const void *s, size_t n) { CU(cudaMemcpyAsync(d, s, n, cudaMemcpyDeviceToHost, stream)); CU(cudaStreamSynchronize(stream)); }
   void zero(void *d, size_t n) { CU(cudaMemsetAsync(d, 0, n, stream)); }
   void sync() { CU(cudaStreamSynchronize(stream)); }
   // device time of a render call: events recorded on the launching stream, read back lazily (no host sync per pass)
   int timerStart() { CU(cudaEventRecord(ev0, stream)); return 0; }
   double timerStop(int) { CU(cudaEventRecord(ev1, stream)); timerPending = true; return -1.0; }
   double timerRead(double last) { if (!timerPending) return last; timerPending = false; CU(cudaEventSynchronize(ev1)); float ms = 0; CU(cudaEventElapsedTime(&ms, ev0, ev1)); return ms; }

   uint32_t gridFor(uint32_t n, uint32_t block, uint32_t perSm) const {
      uint32_t need = (n + block - 1) / block;
      uint32_t full = (uint32_t)sms * perSm;
      if (need >= full) return full;
      return need ? need : 1;
   }
   // CTAs of `kernel` that fit on one SM (registers / shared memory), cached per kernel: grid-stride kernels are
   // launched with exactly sms x resident CTAs so that every SM holds the same amount of work (no partial last wave)
   std::vector<std::pair<const void *, int>> residentCache;
   template <class K> uint32_t resident(K kernel, int block) {
      const void *key = (const void *)kernel;
      for (auto &e : residentCache) if (e.first == key) return (uint32_t)e.second;
      int nb = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, block, 0);
      residentCache.emplace_back(key, nb > 0 ? nb : 1);
      return (uint32_t)residentCache.back().second;
   }
   template <class B> void run(const B &b, uint32_t n) {
      if (n == 0) return;
      Scope sc_(this);
      kRun<B><<<gridFor(n, 256, resident(kRun<B>, 256)), 256, 0, stream>>>(b, n);
   }
   template <class B> void runQueue(const B &b, const uint32_t *q, const uint32_t *cnt, uint32_t bound) {
      if (bound == 0) return;
      Scope sc_(this);
      kRunQueue<B><<<gridFor(bound, 256, resident(kRunQueue<B>, 256)), 256, 0, stream>>>(b, q, cnt);
   }
   // film: tiled shared-memory kernel whenever the filter fits its staging area, else the per-pixel gather
   void run(const FilmBody &b, uint32_t n) {
      if (n == 0) return;
      Scope sc_(this);
      if (filmTiled) {
         const uint32_t tiles = (uint32_t)((filmW + FT_TILE - 1) / FT_TILE) * (uint32_t)((filmH + FT_TILE - 1) / FT_TILE);
         kFilmTile<<<tiles, FT_TILE * FT_TILE, 0, stream>>>(b.sc, b.ps, b.film, b.k, b.npix);
      } else kRun<FilmBody><<<gridFor(n, 256, resident(kRun<FilmBody>, 256)), 256, 0, stream>>>(b, n);
   }
   bool filmTiled = false; int filmW = 0, filmH = 0;
   void setFilm(int w, int h, float fw, float fh) {
      filmW = w; filmH = h;
      auto cells = [](float f) { return FT_TILE + (int)floorf(f + 0.5f) + (int)ceilf(f + 0.5f) + 1; };
      filmTiled = cells(fw) <= FT_MAXC && cells(fh) <= FT_MAXC;
   }
   template <int MK> void runQueue(const ShadeHitBody<MK> &b, const uint32_t *q, const uint32_t *cnt, uint32_t bound) {
      if (bound == 0) return;
      Scope sc_(this);
#if SH_PREFETCH
      kRunQueueShade<MK><<<gridFor(bound, 128, resident(kRunQueueShade<MK>, 128)), 128, 0, stream>>>(b, q, cnt);
#else
      kRunQueueHeavy<ShadeHitBody<MK>><<<gridFor(bound, 128, resident(kRunQueueHeavy<ShadeHitBody<MK>>, 128)), 128, 0, stream>>>(b, q, cnt);
#endif
   }
   template <bool DL> void runQueue(const ResolveMisBodyT<DL> &b, const uint32_t *q, const uint32_t *cnt, uint32_t bound) {
      if (bound == 0) return;
      Scope sc_(this);
      kRunQueueHeavy<ResolveMisBodyT<DL>><<<gridFor(bound, 128, resident(kRunQueueHeavy<ResolveMisBodyT<DL>>, 128)), 128, 0, stream>>>(b, q, cnt);
   }
   void runQueue(const DlShadeBody &b, const uint32_t *q, const uint32_t *cnt, uint32_t bound) {
      if (bound == 0) return;
      Scope sc_(this);
      kRunQueueHeavy<DlShadeBody><<<gridFor(bound, 128, resident(kRunQueueHeavy<DlShadeBody>, 128)), 128, 0, stream>>>(b, q, cnt);
   }
   void traceNearest(const uint32_t *q, const uint32_t *cnt, uint32_t n, const DScene *sc, const F4 *o, const F4 *d, F4 *hit) {
      if (!n) return;
      Scope sc_(this);
      launchTraceNearest(tcfg, stream, q, cnt, n, sc, o, d, hit);
   }
   void traceAny(const uint32_t *q, const uint32_t *cnt, uint32_t n, const DScene *sc, const F4 *o, const F4 *d, uint8_t *occl) {
      if (!n) return;
      Scope sc_(this);
      launchTraceAny(tcfg, stream, q, cnt, n, sc, o, d, occl);
   }
   bool fusesResolve() const { return tcfg.variant != 0; }   // the reference-point kernels of variant 0 only write occlusion flags
   // any-hit query that also resolves: L[slot] += P[slot] for every unoccluded ray (trace_kernels.cuh)
   void traceAnyFused(const uint32_t *q, const uint32_t *cnt, uint32_t n, const DScene *sc, const F4 *o, const F4 *d, uint8_t *occl, F4 *L, const F4 *P, uint32_t cap) {
      if (!n) return;
      Scope sc_(this);
      launchTraceAny(tcfg, stream, q, cnt, n, sc, o, d, occl, L, P, cap);
   }
   void traceStats(uint32_t n, const DScene *sc, const F4 *o, const F4 *d, F4 *hit, uint32_t *nodes, uint32_t *prims) {
      if (n) launchTraceStats(tcfg, stream, n, sc, o, d, hit, nodes, prims);
   }
};

}  // namespace bl

BL_DEFINE_API(blingcu, bl::CudaBackend)
