// emu.cpp -- CPU kernel-body EMULATOR. TEST INFRASTRUCTURE ONLY (tests -m "not gpu").
// It drives the exact kernel bodies of bling_b200/csrc/bodies.h one item at a time on the host so that the
// no-GPU CI can check the wavefront logic against the oracle. It is NOT a fallback: libblingcu.so does not
// contain it, bling_b200/ never loads it, and blingcu_create() fails with BLINGCU_ENOGPU without a device.
#include "../../bling_b200/csrc/api_impl.h"
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <vector>

struct EmuBackend {
   int init(int, std::string &) { return 0; }
   void shutdown() {}
   template <class F> int guard(std::string &, F f) { return f(); }
   bool setOption(const std::string &, double) { return false; }
   void *alloc(size_t n) { return std::malloc(n ? n : 1); }
   void free(void *p) { std::free(p); }
   void upload(void *d, const void *s, size_t n) { std::memcpy(d, s, n); }
   void download(void *d, const void *s, size_t n) { std::memcpy(d, s, n); }
   void zero(void *d, size_t n) { std::memset(d, 0, n); }
   void sync() {}
   void setStream(void *) {}
   void setMaxStack(int) {}
   bool traceHostBatch(const blingcu_ray *, size_t, blingcu_hit *, uint8_t *, const bl::DScene *) { return false; }
   void *hostAlloc(size_t n) { return std::malloc(n ? n : 1); }
   void hostFree(void *p) { std::free(p); }
   void setBvh(const bl::Bvh &) {}
   void setFilm(int, int, float, float) {}
   double timerRead(double last) { return last; }
   void tag(int) {}
   void kernelTimes(double *ms, uint64_t *l, int n) { for (int i = 0; i < n; ++i) { ms[i] = 0; l[i] = 0; } }
   void traversalTotals(uint64_t *six) { for (int i = 0; i < 6; ++i) six[i] = 0; }
   void resetProfile() {}
   std::chrono::steady_clock::time_point timerStart() { return std::chrono::steady_clock::now(); }
   double timerStop(std::chrono::steady_clock::time_point t0) { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); }
   template <class B> void run(const B &b, uint32_t n) { for (uint32_t i = 0; i < n; ++i) b(i); }
   template <class B> void runQueue(const B &b, const uint32_t *q, const uint32_t *cnt, uint32_t) { uint32_t n = *cnt; for (uint32_t i = 0; i < n; ++i) b(q[i]); }
   void traceNearest(const uint32_t *q, const uint32_t *cnt, uint32_t n, const bl::DScene *sc, const bl::F4 *o, const bl::F4 *d, bl::F4 *hit) {
      bl::TraceNearestBody b{sc, o, d, hit};
      if (q) runQueue(b, q, cnt, n); else run(b, n);
   }
   void traceAny(const uint32_t *q, const uint32_t *cnt, uint32_t n, const bl::DScene *sc, const bl::F4 *o, const bl::F4 *d, uint8_t *occl) {
      bl::TraceAnyBody b{sc, o, d, occl};
      if (q) runQueue(b, q, cnt, n); else run(b, n);
   }
   bool fusesResolve() const { return true; }
   // film reduction (comm.h in the product): the emulator sums the films of the contexts of ONE process, which is all the
   // no-GPU CI needs to check the group semantics (film_sum = sum, private films untouched)
   int commRank = 0, commRanks = 1; std::vector<EmuBackend *> *peers = nullptr; const float *myFilm = nullptr; float *mySum = nullptr; size_t myN = 0; int myRoot = -1;
   static std::vector<EmuBackend *> &groupBuf() { static std::vector<EmuBackend *> g; return g; }
   static bool &inGroup() { static bool b = false; return b; }
   static int commUniqueId(uint8_t *id, std::string &) { std::memset(id, 0x42, BLINGCU_COMM_ID_BYTES); return 0; }
   static int groupStart(std::string &) { inGroup() = true; groupBuf().clear(); return 0; }
   static int groupEnd(std::string &) {
      inGroup() = false;
      auto &g = groupBuf();
      for (EmuBackend *b : g) {
         if (!b->mySum) continue;
         if (b->myRoot >= 0 && b->commRank != b->myRoot) continue;
         for (size_t i = 0; i < b->myN; ++i) { float acc = 0; for (EmuBackend *o : g) if (o->myFilm) acc += o->myFilm[i]; b->mySum[i] = acc; }
      }
      for (EmuBackend *b : g) { b->myFilm = nullptr; b->mySum = nullptr; }
      g.clear();
      return 0;
   }
   int commInit(int rank, int nranks, const uint8_t *, std::string &err) {
      if (nranks < 1 || rank < 0 || rank >= nranks) { err = "comm_init: rank out of range"; return BLINGCU_EINVAL; }
      if (nranks > 1 && !inGroup()) { err = "emulator: no inter-process communicator (use comm_init_all)"; return BLINGCU_EUNSUPPORTED; }
      commRank = rank; commRanks = nranks; return 0;
   }
   void commDestroy() { commRank = 0; commRanks = 1; }
   int reduceFilm(const float *film, float *filmSum, size_t nFloats, int root, std::string &err) {
      if (root >= commRanks) { err = "reduce_film: root out of range"; return BLINGCU_EINVAL; }
      if (commRanks == 1) { std::memcpy(filmSum, film, nFloats * sizeof(float)); return 0; }
      if (!inGroup()) { err = "emulator: reduce_film of a multi-context communicator needs reduce_film_group"; return BLINGCU_EUNSUPPORTED; }
      myFilm = film; mySum = filmSum; myN = nFloats; myRoot = root; groupBuf().push_back(this);
      return 0;
   }
   void waitReduced() {}
   void syncComm() {}
   void downloadOnComm(void *d, const void *s, size_t n) { std::memcpy(d, s, n); }
   void traceAnyFused(const uint32_t *q, const uint32_t *cnt, uint32_t n, const bl::DScene *sc, const bl::F4 *o, const bl::F4 *d, uint8_t *occl, bl::F4 *L, const bl::F4 *P, uint32_t cap) {
      traceAny(q, cnt, n, sc, o, d, occl);
      bl::PathState ps{}; ps.cap = cap; ps.L = L; ps.PS = const_cast<bl::F4 *>(P); ps.occl = occl;
      bl::ResolveShadowBody b{ps};
      if (q) runQueue(b, q, cnt, n); else run(b, n);
   }
   void traceStats(uint32_t n, const bl::DScene *sc, const bl::F4 *o, const bl::F4 *d, bl::F4 *hit, uint32_t *nodes, uint32_t *prims) {
      for (uint32_t i = 0; i < n; ++i) {
         nodes[i] = 0; prims[i] = 0;
         bl::HitRec h = bl::traceNearest<true>(sc->bvh, bl::loadRay(o, d, i), nodes + i, prims + i);
         hit[i] = bl::F4{h.t, h.b1, h.b2, bl::i2f(h.prim)};
      }
   }
};

BL_DEFINE_API(blingemu, EmuBackend)
