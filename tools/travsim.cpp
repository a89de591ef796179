// travsim.cpp -- CPU model of the persistent traversal kernels (trace_kernels.cuh) for design decisions that would
// otherwise each cost a GPU call: it replays the warp-level schedule (32 lanes, vote stepping, refill) of a traversal
// policy over REAL ray streams of the cfg-5 soup (camera rays of a strip of the 3842-wide sample extent, then the
// extension / BSDF-MIS / shadow rays of the following bounces, in queue order) and reports per ray: node visits,
// primitive tests, warp trips, lanes active per step and a warp-instruction estimate from per-step costs read off the
// SASS of the shipped kernel. Test/tool code only (not linked into the product).
//   g++ -O2 -fopenmp -std=c++17 -I bling_b200/csrc tools/travsim.cpp -o /tmp/travsim && /tmp/travsim [ntris] [rows]
// Environment knobs: POLS="|name|name|" (policies to run beside the base), MAXLEAF, TRAVCOST, FORCELEAF (binary builder), DPCOLLAPSE="cn,cp,pmax"
// (SAH-optimal collapse instead of the greedy one), QPAD (cells a child bound is moved out by: 1 = the rule until late in round 2, default 1/64),
// NOQUANT (exact child boxes), SORTED, RANDOM_RAYS. The PRODUCT tree of cfg 5 is `MAXLEAF=1 DPCOLLAPSE=1,0.5,3` (profiles/r02_tree_and_requests.md).
#include "../bling_b200/csrc/bvh_build.cpp"
#include <chrono>
#include <cstdio>
#include <random>

using namespace bl;

static uint64_t sm_state;
static inline uint64_t splitmix(uint64_t &k) { uint64_t z = (k += 0x9E3779B97F4A7C15ULL); z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL; z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL; return z ^ (z >> 31); }

struct Tri { V3 p1, e1, e2; };

// ------------------------------------------------------------------ generic W-wide tree built from the binary SAH tree
struct WNode {
   int nc;
   float lo[8][3], hi[8][3];
   int ref[8];   // >= 0 node, < 0 leaf ~((first << 4) | cnt)
};
struct WTree { std::vector<WNode> nodes; int root; int W; };

struct Bin { const F4 *n; };
static void childrenOf(const F4 *nodes, int idx, Box &lb, int &lr, Box &rb, int &rr) {
   const F4 *np = nodes + 4 * (size_t)idx;
   lb.lo[0] = np[0].x; lb.hi[0] = np[0].y; lb.lo[1] = np[0].z; lb.hi[1] = np[0].w; lb.lo[2] = np[2].x; lb.hi[2] = np[2].y;
   rb.lo[0] = np[1].x; rb.hi[0] = np[1].y; rb.lo[1] = np[1].z; rb.hi[1] = np[1].w; rb.lo[2] = np[2].z; rb.hi[2] = np[2].w;
   lr = f2i(np[3].x); rr = f2i(np[3].y);
}

// ---- SAH-optimal collapse by dynamic programming (Ylitie, Karras, Laine 2017, section 3.1) as an alternative to "open the largest child":
// C(n, i) = cheapest way to represent the binary subtree n with at most i slots of its wide parent. DPCOLLAPSE="cn,cp,pmax"
struct DPTab { std::vector<float> c1, c2, c3; std::vector<int> first, cnt; float cn = 1, cp = 1; int pmax = 2; int W = 4; const F4 *bin = nullptr; };
static DPTab gDP;
static inline float dpChildC(int r, const Box &b, int i) { if (r < 0) return b.area() * (float)((~r) & 15) * gDP.cp; return i == 1 ? gDP.c1[r] : (i == 2 ? gDP.c2[r] : gDP.c3[r]); }
static inline float dpDist(int lr, const Box &lb, int rr, const Box &rb, int j, int *kbest) {
   float best = BL_INF; int kb = 1;
   for (int k = 1; k < j; ++k) { float c = dpChildC(lr, lb, std::min(k, 3)) + dpChildC(rr, rb, std::min(j - k, 3)); if (c < best) { best = c; kb = k; } }
   if (kbest) *kbest = kb; return best;
}
static void dpRec(int n, Box &box) {
   Box lb, rb; int lr, rr; childrenOf(gDP.bin, n, lb, lr, rb, rr);
   Box tmp;
   if (lr >= 0) dpRec(lr, tmp);
   if (rr >= 0) dpRec(rr, tmp);
   box = lb; box.grow(rb);
   const int lf = lr < 0 ? ((~lr) >> 4) : gDP.first[lr], lc = lr < 0 ? ((~lr) & 15) : gDP.cnt[lr];
   const int rc = rr < 0 ? ((~rr) & 15) : gDP.cnt[rr];
   gDP.first[n] = lf; gDP.cnt[n] = lc + rc;
   const float A = box.area();
   const float cleaf = (lc + rc <= gDP.pmax) ? A * (float)(lc + rc) * gDP.cp : BL_INF;
   const float cint = dpDist(lr, lb, rr, rb, gDP.W, nullptr) + A * gDP.cn;
   gDP.c1[n] = std::min(cleaf, cint);
   gDP.c2[n] = std::min(dpDist(lr, lb, rr, rb, 2, nullptr), gDP.c1[n]);
   gDP.c3[n] = std::min(dpDist(lr, lb, rr, rb, 3, nullptr), gDP.c2[n]);
}
static void dpExpand(int r, const Box &b, int i, Box *cb, int *cr, int &nc) {
   if (r < 0) { cb[nc] = b; cr[nc] = r; nc++; return; }
   Box lb, rb; int lr, rr; childrenOf(gDP.bin, r, lb, lr, rb, rr);
   if (i > 1) {
      int k; const float cd = dpDist(lr, lb, rr, rb, i, &k);
      const float cprev = i == 2 ? gDP.c1[r] : gDP.c2[r];
      if (cd < cprev) { dpExpand(lr, lb, k, cb, cr, nc); dpExpand(rr, rb, i - k, cb, cr, nc); return; }
      dpExpand(r, b, i - 1, cb, cr, nc); return;
   }
   const float A = b.area();
   const float cleaf = (gDP.cnt[r] <= gDP.pmax) ? A * (float)gDP.cnt[r] * gDP.cp : BL_INF;
   if (cleaf <= gDP.c1[r]) { cb[nc] = b; cr[nc] = ~((gDP.first[r] << 4) | gDP.cnt[r]); nc++; }
   else { cb[nc] = b; cr[nc] = r; nc++; }
}

// collapse + 8-bit quantisation (same rounding rules as bvh_build.cpp); octantOrder: assign children to slots so that
// slot ^ octant gives an approximate front-to-back order (Ylitie et al. 2017, greedy instead of auction)
static WTree collapse(const F4 *bin, int root2, int W, bool quantise, bool octantOrder) {
   WTree T; T.W = W;
   struct Work { int n2, nw; };
   std::vector<Work> work;
   T.nodes.emplace_back(); T.root = 0;
   work.push_back({root2, 0});
   while (!work.empty()) {
      Work w = work.back(); work.pop_back();
      Box cb[8]; int cr[8]; int nc = 2;
      childrenOf(bin, w.n2, cb[0], cr[0], cb[1], cr[1]);
      if (gDP.bin && W == gDP.W) {
         Box lb = cb[0], rb = cb[1]; int lr = cr[0], rr = cr[1]; int k; dpDist(lr, lb, rr, rb, W, &k);
         nc = 0; dpExpand(lr, lb, std::min(k, 3), cb, cr, nc); dpExpand(rr, rb, std::min(W - k, 3), cb, cr, nc);
      } else
      while (nc < W) {
         int best = -1; float ba = -1;
         for (int k = 0; k < nc; ++k) if (cr[k] >= 0 && cb[k].area() > ba) { ba = cb[k].area(); best = k; }
         if (best < 0) break;
         Box l, r; int lr, rr; childrenOf(bin, cr[best], l, lr, r, rr);
         cb[best] = l; cr[best] = lr; cb[nc] = r; cr[nc] = rr; nc++;
      }
      // drop empty leaves
      int m = 0;
      for (int k = 0; k < nc; ++k) if (!(cr[k] < 0 && ((~cr[k]) & 15) == 0)) { cb[m] = cb[k]; cr[m] = cr[k]; m++; }
      nc = m;
      Box nb; nb.reset(); for (int k = 0; k < nc; ++k) nb.grow(cb[k]);
      int slotOf[8]; for (int k = 0; k < 8; ++k) slotOf[k] = k;
      if (octantOrder) {
         // greedy: cost(c, s) = dot(centroid_c - centroid_node, sign vector of slot s); assign the largest |cost| first
         float cen[3]; for (int a = 0; a < 3; ++a) cen[a] = 0.5f * (nb.lo[a] + nb.hi[a]);
         bool usedS[8] = {}, usedC[8] = {};
         for (int it = 0; it < nc; ++it) {
            float bestv = -BL_INF; int bc = -1, bs = -1;
            for (int c = 0; c < nc; ++c) if (!usedC[c]) for (int s = 0; s < W; ++s) if (!usedS[s]) {
               float v = 0;
               for (int a = 0; a < 3; ++a) { float d = 0.5f * (cb[c].lo[a] + cb[c].hi[a]) - cen[a]; v += ((s >> a) & 1) ? d : -d; }
               if (v > bestv) { bestv = v; bc = c; bs = s; }
            }
            usedC[bc] = true; usedS[bs] = true; slotOf[bc] = bs;
         }
      }
      WNode nd; nd.nc = W;
      for (int s = 0; s < 8; ++s) { nd.ref[s] = ~0; for (int a = 0; a < 3; ++a) { nd.lo[s][a] = BL_INF; nd.hi[s][a] = -BL_INF; } }
      for (int a = 0; a < 3; ++a) {
         double lo = nb.lo[a], ext = (double)nb.hi[a] - (double)nb.lo[a]; if (!(ext > 0)) ext = 0;
         static const double qpad0 = getenv("QPAD") ? atof(getenv("QPAD")) : 1.0 / 64;   // QPAD=1: the rule until the end of round 2 (one extra cell per side)
         const double span = qpad0 >= 1 ? 252.0 : 254.0;
         int e = -100; if (ext > 0) { e = (int)std::ceil(std::log2(ext / span)); while (std::ldexp(span, e) < ext) e++; }
         double cell = std::ldexp(1.0, e); float Pf = (float)(lo - (qpad0 >= 1 ? cell : cell / 32));
         for (int k = 0; k < nc; ++k) {
            int s = slotOf[k];
            if (quantise) {
               static const double qpad = getenv("QPAD") ? atof(getenv("QPAD")) : 1.0 / 64;   // cells every bound is moved out by after rounding (bvh_build.cpp)
               int ql = (int)std::floor(((double)cb[k].lo[a] - (double)Pf) / cell - qpad), qh = (int)std::ceil(((double)cb[k].hi[a] - (double)Pf) / cell + qpad);
               ql = std::max(0, std::min(255, ql)); qh = std::max(0, std::min(255, qh));
               nd.lo[s][a] = (float)((double)Pf + cell * ql); nd.hi[s][a] = (float)((double)Pf + cell * qh);
            } else { nd.lo[s][a] = cb[k].lo[a]; nd.hi[s][a] = cb[k].hi[a]; }
         }
      }
      for (int k = 0; k < nc; ++k) {
         int s = slotOf[k];
         if (cr[k] >= 0) { int id = (int)T.nodes.size(); T.nodes.emplace_back(); work.push_back({cr[k], id}); nd.ref[s] = id; }
         else nd.ref[s] = cr[k];
      }
      T.nodes[w.nw] = nd;
   }
   return T;
}

// ------------------------------------------------------------------ per-ray traversal state machine
struct Policy {
   const char *name;
   int W = 4;
   bool sortChildren = true;     // distance sort (else slot order, or octant order)
   bool octant = false;          // slot ^ octant order (needs an octant-ordered tree)
   bool nearestFirstOnly = false;  // enter the nearest child, push the others in slot order
   bool cullOnPop = false;       // stack keeps tnear; stale entries are dropped at pop time (nearest-hit only)
   int leafW = 6;                // vote weights
   bool ifif = false;            // node step then leaf step each trip
   bool deferLeaves = false;     // warp-level leaf queue: leaf items are queued and tested 32 at a time by all lanes
   int enqPerTrip = 1; int flushAt = 32; int refillAt = 4;           // leaves a lane may queue per trip
   int anyOrder = 0;             // any-hit child order: 0 = as sortChildren says, 1 = longest overlap first, 2 = leaves first
   float cLeafPass = 110, cEnq = 14;
   // cost model (warp instructions)
   float cTrip = 40, cNode = 171, cLeaf = 86, cPopCull = 6;
};

struct Lane {
   bool active = false;
   Ray r; V3 idir, ood;
   int cur = 0, li = 0;
   int sp = 0; int stk[256]; float stn[256];
   HitRec h; uint32_t slot;
   bool occluded;
};

struct Stats {
   double nsIdle = 0, nsWait = 0, nsLeaf = 0, distinctLines = 0; double rays = 0, nodes = 0, prims = 0, trips = 0, nodeSteps = 0, leafSteps = 0, nodeLanes = 0, leafLanes = 0, cost = 0, popCulls = 0, lines = 0;
   void add(const Stats &o) { rays += o.rays; nodes += o.nodes; prims += o.prims; trips += o.trips; nodeSteps += o.nodeSteps; leafSteps += o.leafSteps; nodeLanes += o.nodeLanes; leafLanes += o.leafLanes; cost += o.cost; popCulls += o.popCulls; lines += o.lines; }
};

static std::vector<Tri> gItems;   // in leaf order

static inline void nodeStep(const WTree &T, const Policy &P, Lane &L, bool ANY, Stats &S) {
   const WNode &nd = T.nodes[L.cur];
   S.nodes++;
   float tn[8], tfv[8], area[8]; int rf[8]; int n = 0;
   const int oct = (L.r.d.x < 0 ? 1 : 0) | (L.r.d.y < 0 ? 2 : 0) | (L.r.d.z < 0 ? 4 : 0);
   for (int kk = 0; kk < T.W; ++kk) {
      const int k = P.octant ? (kk ^ oct) & (T.W - 1) : kk;   // for W=4 only two axes' bits matter (approximation)
      if (nd.ref[k] == ~0) continue;
      float t0 = L.r.tmin, t1 = L.r.tmax;
      const float o[3] = {L.r.o.x, L.r.o.y, L.r.o.z}, id[3] = {L.idir.x, L.idir.y, L.idir.z};
      for (int a = 0; a < 3; ++a) {
         float ta = (nd.lo[k][a] - o[a]) * id[a], tb = (nd.hi[k][a] - o[a]) * id[a];
         if (id[a] < 0) std::swap(ta, tb);
         t0 = std::max(t0, ta); t1 = std::min(t1, tb);
      }
      if (t0 <= t1) {
         const float ex = nd.hi[k][0] - nd.lo[k][0], ey = nd.hi[k][1] - nd.lo[k][1], ez = nd.hi[k][2] - nd.lo[k][2];
         area[n] = ex * ey + ey * ez + ez * ex;
         tn[n] = t0; tfv[n] = t1; rf[n] = nd.ref[k]; n++;
      }
   }
   if (n == 0) { L.cur = 0x7fffffff; return; }   // pop
   if (ANY && P.anyOrder) {   // experimental any-hit orders: the key replaces the entry distance
      for (int i = 0; i < n; ++i) {
         float key = tn[i];
         if (P.anyOrder == 1) key = -(tfv[i] - tn[i]);                       // longest overlap first
         else if (P.anyOrder == 2) key = (rf[i] < 0 ? -1e30f : 0.0f) + tn[i];   // leaves first, then nearest
         else if (P.anyOrder == 3) key = -tn[i];                              // farthest first
         else if (P.anyOrder == 4) key = (rf[i] < 0 ? 1e30f : 0.0f) + tn[i];    // inner nodes first, leaves last
         else if (P.anyOrder == 7) key = -area[i];                            // static: largest box first (a builder-time slot order, free at run time)
         else if (P.anyOrder == 8) key = area[i];                             // static: smallest box first
         tn[i] = key;
      }
      if (P.anyOrder == 5 || P.anyOrder == 6) {   // only the FIRST child is chosen by the key (5: longest overlap, 6: same, the rest reversed), the rest stay in slot order
         for (int i = 0; i < n; ++i) tn[i] = -(tfv[i] - (tn[i]));
         int b = 0; for (int i = 1; i < n; ++i) if (tn[i] < tn[b]) b = i;
         float tb = tn[b]; int rb = rf[b];
         for (int i = b; i > 0; --i) { tn[i] = tn[i - 1]; rf[i] = rf[i - 1]; }
         tn[0] = tb; rf[0] = rb;
      } else
      for (int i = 1; i < n; ++i) { float t = tn[i]; int r = rf[i]; int j = i; while (j > 0 && tn[j - 1] > t) { tn[j] = tn[j - 1]; rf[j] = rf[j - 1]; --j; } tn[j] = t; rf[j] = r; }
   } else if (P.sortChildren) {
      for (int i = 1; i < n; ++i) { float t = tn[i]; int r = rf[i]; int j = i; while (j > 0 && tn[j - 1] > t) { tn[j] = tn[j - 1]; rf[j] = rf[j - 1]; --j; } tn[j] = t; rf[j] = r; }
   } else if (P.nearestFirstOnly) {
      int b = 0; for (int i = 1; i < n; ++i) if (tn[i] < tn[b]) b = i;
      std::swap(tn[0], tn[b]); std::swap(rf[0], rf[b]);
   }
   for (int j = n - 1; j >= 1; --j) { L.stk[L.sp] = rf[j]; L.stn[L.sp] = tn[j]; L.sp++; }
   L.cur = rf[0]; L.li = 0;
}

static inline bool triTest(const Tri &tr, Ray &r, HitRec &h, bool ANY) {
   float t, b1, b2;
   if (!triHit(tr.p1, tr.e1, tr.e2, r, t, b1, b2)) return false;
   if (!ANY) { r.tmax = t; h.t = t; h.b1 = b1; h.b2 = b2; }
   return true;
}

// simulate one "launch": rays in queue order, persistent warps pulling from the queue. Returns stats; hits (nearest) out.
static Stats simulate(const WTree &T, const Policy &P, const std::vector<Ray> &rays, bool ANY, std::vector<HitRec> *hitsOut, std::vector<uint8_t> *occlOut, int nWarps = 0) {
   const size_t n = rays.size();
   if (hitsOut) hitsOut->assign(n, HitRec{0, -1, 0, 0});
   if (occlOut) occlOut->assign(n, 0);
   // the GPU runs ~ 148*8*4 = 4736 warps concurrently; each takes lanes from the shared queue as it goes. Model: chunks of
   // the queue are processed by independent simulated warps; the queue head advances in warp-trip lockstep (round-robin).
   if (nWarps <= 0) nWarps = (int)std::min<size_t>(4736, std::max<size_t>(1, n / 2048));
   size_t head = 0;
   std::vector<std::vector<Lane>> W(nWarps, std::vector<Lane>(32));
   std::vector<char> exhausted(nWarps, 0), done(nWarps, 0);
   Stats S; S.rays = (double)n;
   int live = nWarps;
   const int POPPED = 0x7fffffff;
   while (live > 0) {
      for (int w = 0; w < nWarps; ++w) {
         if (done[w]) continue;
         auto &lanes = W[w];
         int nN = 0, nL = 0, nIdle = 0;
         for (auto &L : lanes) { if (!L.active) nIdle++; else if (L.cur >= 0) nN++; else nL++; }
         if (!exhausted[w] && nIdle >= 4) {
            for (auto &L : lanes) if (!L.active) {
               if (head < n) {
                  L.active = true; L.slot = (uint32_t)head; L.r = rays[head]; head++;
                  RayPre p = rayPre(L.r); L.idir = p.idir;
                  L.sp = 0; L.cur = T.root; L.li = 0; L.h = HitRec{0, -1, 0, 0}; L.occluded = false;
               }
            }
            if (head >= n) exhausted[w] = 1;
            nN = nL = 0; for (auto &L : lanes) { if (L.active) { if (L.cur >= 0) nN++; else nL++; } }
         }
         if (nN + nL == 0) { if (exhausted[w]) { done[w] = 1; live--; } continue; }
         S.trips++; S.cost += P.cTrip;
         const bool doNode = P.ifif ? nN > 0 : 4 * nN >= P.leafW * nL;
         const bool doLeaf = P.ifif ? nL > 0 : !doNode;
         if (doNode) {
            S.nodeSteps++; S.nodeLanes += nN; S.cost += P.cNode;
            for (auto &L : lanes) if (L.active && L.cur >= 0 && L.cur != POPPED) nodeStep(T, P, L, ANY, S);
         }
         if (doLeaf) {
            S.leafSteps++; S.leafLanes += nL; S.cost += P.cLeaf;
            for (auto &L : lanes) if (L.active && L.cur < 0) {
               const int enc = ~L.cur, first = enc >> 4, cnt = enc & 15;
               bool found = false;
               if (L.li < cnt) { S.prims++; found = triTest(gItems[first + L.li], L.r, L.h, ANY); if (found && !ANY) L.h.prim = first + L.li; L.li++; }
               if (ANY && found) { L.occluded = true; L.active = false; if (occlOut) (*occlOut)[L.slot] = 1; }
               else if (L.li >= cnt) L.cur = POPPED;
            }
         }
         // pops
         for (auto &L : lanes) if (L.active && L.cur == POPPED) {
            for (;;) {
               if (L.sp == 0) { L.active = false; if (hitsOut) (*hitsOut)[L.slot] = L.h; break; }
               L.sp--; L.cur = L.stk[L.sp]; L.li = 0;
               if (P.cullOnPop && !ANY && L.stn[L.sp] > L.r.tmax) { S.popCulls++; S.cost += P.cPopCull / 8.0; continue; }
               break;
            }
         }
      }
   }
   return S;
}

// warp-level leaf queue: lanes never wait at a leaf. A lane that reaches a leaf appends (lane, item) pairs to a per-warp
// queue and pops on; once 32 pairs are queued ALL lanes test one pair each (the ray comes from its owner by shuffle).
// A ray whose stack runs dry waits for its queued pairs before it retires.
static Stats simulateDeferred(const WTree &T, const Policy &P, const std::vector<Ray> &rays, bool ANY, std::vector<HitRec> *hitsOut, std::vector<uint8_t> *occlOut) {
   const size_t n = rays.size();
   if (hitsOut) hitsOut->assign(n, HitRec{0, -1, 0, 0});
   if (occlOut) occlOut->assign(n, 0);
   int nWarps = (int)std::min<size_t>(4736, std::max<size_t>(1, n / 2048));
   size_t head = 0;
   struct Ent { int lane, item; uint32_t slot; };
   struct Warp { std::vector<Lane> lanes; std::vector<Ent> q; std::vector<int> pending; bool exhausted = false, done = false; };
   std::vector<Warp> W(nWarps);
   for (auto &w : W) { w.lanes.assign(32, Lane()); w.pending.assign(32, 0); }
   Stats S; S.rays = (double)n;
   int live = nWarps;
   const int POPPED = 0x7fffffff, WAIT = 0x7ffffffe;
   auto popNext = [&](Lane &L) {
      for (;;) {
         if (L.sp == 0) { L.cur = WAIT; return; }
         L.sp--; L.cur = L.stk[L.sp]; L.li = 0;
         if (P.cullOnPop && !ANY && L.stn[L.sp] > L.r.tmax) { S.popCulls++; continue; }
         return;
      }
   };
   while (live > 0) {
      for (auto &w : W) {
         if (w.done) continue;
         auto &lanes = w.lanes;
         int nIdle = 0; for (auto &L : lanes) if (!L.active) nIdle++;
         if (!w.exhausted && nIdle >= P.refillAt) {
            for (auto &L : lanes) if (!L.active) {
               if (head < n) {
                  L.active = true; L.slot = (uint32_t)head; L.r = rays[head]; head++;
                  RayPre p = rayPre(L.r); L.idir = p.idir;
                  L.sp = 0; L.cur = T.root; L.li = 0; L.h = HitRec{0, -1, 0, 0}; L.occluded = false;
               }
            }
            if (head >= n) w.exhausted = true;
         }
         int nAct = 0; for (auto &L : lanes) if (L.active) nAct++;
         if (nAct == 0) { if (w.exhausted) { w.done = true; live--; } continue; }
         S.trips++; S.cost += P.cTrip;
         // enqueue phase
         for (int rep = 0; rep < P.enqPerTrip; ++rep) {
            bool anyLeaf = false;
            for (int li = 0; li < 32; ++li) { Lane &L = lanes[li]; if (L.active && L.cur < 0) {
               anyLeaf = true;
               const int enc = ~L.cur, first = enc >> 4, cnt = enc & 15;
               for (int i = 0; i < cnt; ++i) { w.q.push_back({li, first + i, L.slot}); w.pending[li]++; }
               popNext(L);
            } }
            if (anyLeaf) S.cost += P.cEnq; else break;
         }
         // leaf passes
         int nNode = 0; for (auto &L : lanes) if (L.active && L.cur >= 0 && L.cur < WAIT) nNode++;
         while ((int)w.q.size() >= P.flushAt || (nNode == 0 && !w.q.empty())) {
            const size_t m = std::min<size_t>(32, w.q.size());
            S.leafSteps++; S.cost += P.cLeafPass;
            float tsnap[32]; for (int i = 0; i < 32; ++i) tsnap[i] = lanes[i].r.tmax;
            for (size_t e = 0; e < m; ++e) {
               const Ent &en = w.q[e]; Lane &L = lanes[en.lane];
               w.pending[en.lane]--;
               if (!L.active || L.slot != en.slot) continue;   // owner already terminated (any-hit) : a wasted lane
               S.leafLanes++; S.prims++;
               Ray r = L.r; r.tmax = tsnap[en.lane];
               float t, b1, b2;
               if (triHit(gItems[en.item].p1, gItems[en.item].e1, gItems[en.item].e2, r, t, b1, b2)) {
                  if (ANY) { L.occluded = true; }
                  else if (t <= L.r.tmax) { L.r.tmax = t; L.h.t = t; L.h.b1 = b1; L.h.b2 = b2; L.h.prim = en.item; }
               }
            }
            w.q.erase(w.q.begin(), w.q.begin() + m);
            for (int li = 0; li < 32; ++li) { Lane &L = lanes[li]; if (!L.active) continue;
               if (ANY && L.occluded) { L.active = false; if (occlOut) (*occlOut)[L.slot] = 1; }
               else if (L.cur == WAIT && w.pending[li] == 0) { L.active = false; if (hitsOut) (*hitsOut)[L.slot] = L.h; }
            }
            nNode = 0; for (auto &L : lanes) if (L.active && L.cur >= 0 && L.cur < WAIT) nNode++;
         }
         // retire lanes that wait on nothing
         for (int li = 0; li < 32; ++li) { Lane &L = lanes[li]; if (L.active && L.cur == WAIT && w.pending[li] == 0) { L.active = false; if (hitsOut) (*hitsOut)[L.slot] = L.h; } }
         // node step
         if (nNode > 0) {
            S.nodeSteps++; S.nodeLanes += nNode; S.cost += P.cNode;
            { int ids[32], m = 0; for (auto &L : lanes) if (L.active && L.cur >= 0 && L.cur < WAIT) ids[m++] = L.cur; std::sort(ids, ids + m); int dn = 0, dl = 0; for (int i = 0; i < m; ++i) { if (i == 0 || ids[i] != ids[i - 1]) dn++; if (i == 0 || (ids[i] >> 1) != (ids[i - 1] >> 1)) dl++; } S.lines += dn; S.popCulls += 0; S.distinctLines += dl; }
            for (auto &L : lanes) { if (!L.active) S.nsIdle++; else if (L.cur == WAIT) S.nsWait++; else if (L.cur < 0) S.nsLeaf++; }
            for (auto &L : lanes) if (L.active && L.cur >= 0 && L.cur < WAIT) { nodeStep(T, P, L, ANY, S); if (L.cur == POPPED) popNext(L); }
         }
      }
   }
   return S;
}

// ------------------------------------------------------------------ ray populations
static inline float u01r(uint64_t &k) { return (float)(splitmix(k) >> 40) * (1.0f / 16777216.0f); }
static V3 cosineDir(V3 n, uint64_t &k) {
   float u = u01r(k), v = u01r(k);
   float r = sqrtf(u), ph = 2 * BL_PI * v;
   Frame f = coordinateSystem(n);
   float z = sqrtf(std::max(0.0f, 1 - u));
   return normalize3(localToWorld(f, mk3(r * cosf(ph), r * sinf(ph), z)));
}
static V3 uniformSphere(uint64_t &k) { float z = 1 - 2 * u01r(k), ph = 2 * BL_PI * u01r(k); float s = sqrtf(std::max(0.0f, 1 - z * z)); return mk3(s * cosf(ph), s * sinf(ph), z); }

static uint32_t morton3(uint32_t x, uint32_t y, uint32_t z) {
   auto part = [](uint32_t v) { v &= 1023; v = (v | (v << 16)) & 0x30000ff; v = (v | (v << 8)) & 0x300f00f; v = (v | (v << 4)) & 0x30c30c3; v = (v | (v << 2)) & 0x9249249; return v; };
   return part(x) | (part(y) << 1) | (part(z) << 2);
}
static std::vector<Ray> sortRays(const std::vector<Ray> &rays, float ext, int cellBits, bool useOct) {
   std::vector<std::pair<uint64_t, uint32_t>> key(rays.size());
   for (size_t i = 0; i < rays.size(); ++i) {
      const Ray &r = rays[i];
      auto q = [&](float v) { float f = (v + ext) / (2 * ext); f = std::min(0.999999f, std::max(0.0f, f)); return (uint32_t)(f * (1 << cellBits)); };
      uint64_t m = morton3(q(r.o.x), q(r.o.y), q(r.o.z));
      int oct = (r.d.x < 0 ? 1 : 0) | (r.d.y < 0 ? 2 : 0) | (r.d.z < 0 ? 4 : 0);
      key[i] = {useOct ? (m << 3 | oct) : m, (uint32_t)i};
   }
   std::sort(key.begin(), key.end());
   std::vector<Ray> out(rays.size());
   for (size_t i = 0; i < rays.size(); ++i) out[i] = rays[key[i].second];
   return out;
}

static void report(const char *pop, const Policy &P, const Stats &S) {
   printf("%-10s %-26s rays %8.0f | nodes/ray %6.2f prims/ray %6.2f | trips/ray %6.3f | lanes node %5.2f leaf %5.2f | popculls/ray %5.2f | winst/ray %7.1f\n", pop, P.name, S.rays, S.nodes / S.rays,
          S.prims / S.rays, S.trips / S.rays, S.nodeLanes / std::max(1.0, S.nodeSteps), S.leafLanes / std::max(1.0, S.leafSteps), S.popCulls / S.rays, S.cost / S.rays);
   if (S.nsIdle + S.nsWait + S.nsLeaf > 0) printf("      during node steps: idle %.2f wait %.2f at-leaf %.2f lanes; distinct nodes per step %.2f, distinct 128-byte lines %.2f (of %.2f lanes)\n", S.nsIdle / S.nodeSteps, S.nsWait / S.nodeSteps, S.nsLeaf / S.nodeSteps, S.lines / S.nodeSteps, S.distinctLines / S.nodeSteps, S.nodeLanes / S.nodeSteps);
   fflush(stdout);
}

int main(int argc, char **argv) {
   size_t ntris = argc > 1 ? (size_t)atof(argv[1]) : 10000000;
   int rows = argc > 2 ? atoi(argv[2]) : 16;
   const float ext = 100.0f * cbrtf((float)ntris / 1e7f);
   // soup (bling_b200/host/soup.py)
   std::vector<float> lo(3 * ntris), hi(3 * ntris);
   std::vector<Tri> tris(ntris);
   {
      uint64_t k = 0xB11D6;
      for (size_t i = 0; i < ntris; ++i) {
         float u[9]; for (int j = 0; j < 9; ++j) u[j] = (float)(splitmix(k) >> 40) * (1.0f / 16777216.0f);
         V3 c = mk3(u[0] * 2 * ext - ext, u[1] * 2 * ext - ext, u[2] * 2 * ext - ext);
         V3 e1 = mk3(u[3] * 1.2f - 0.6f, u[4] * 1.2f - 0.6f, u[5] * 1.2f - 0.6f), e2 = mk3(u[6] * 1.2f - 0.6f, u[7] * 1.2f - 0.6f, u[8] * 1.2f - 0.6f);
         V3 v0 = c - scl(1.0f / 3.0f, e1 + e2), v1 = v0 + e1, v2 = v0 + e2;
         tris[i].p1 = v0; tris[i].e1 = v1 - v0; tris[i].e2 = v2 - v0;
         const float xs[3][3] = {{v0.x, v1.x, v2.x}, {v0.y, v1.y, v2.y}, {v0.z, v1.z, v2.z}};
         for (int a = 0; a < 3; ++a) { lo[3 * i + a] = std::min(xs[a][0], std::min(xs[a][1], xs[a][2])); hi[3 * i + a] = std::max(xs[a][0], std::max(xs[a][1], xs[a][2])); }
      }
   }
   auto t0 = std::chrono::steady_clock::now();
   // binary tree with the product builder's internals
   Box scene; scene.reset(); for (size_t i = 0; i < ntris; ++i) scene.grow(&lo[3 * i], &hi[3 * i]);
   float e = 0; for (int k = 0; k < 3; ++k) { e = std::max(e, scene.hi[k] - scene.lo[k]); e = std::max(e, std::max(std::fabs(scene.lo[k]), std::fabs(scene.hi[k]))); }
   const float eps = 4e-6f * e + 1e-30f;
   Builder B; B.items.resize(ntris);
   for (size_t i = 0; i < ntris; ++i) { Item &it = B.items[i]; for (int k = 0; k < 3; ++k) { it.lo[k] = lo[3 * i + k] - eps; it.hi[k] = hi[3 * i + k] + eps; it.c[k] = 0.5f * (lo[3 * i + k] + hi[3 * i + k]); } it.id = (uint32_t)i; }
   const int maxLeaf = getenv("MAXLEAF") ? atoi(getenv("MAXLEAF")) : 2;
   B.maxLeaf = maxLeaf; B.parLevels = 4;
   if (getenv("TRAVCOST")) B.travCost = (float)atof(getenv("TRAVCOST"));      // SAH termination with a node-visit cost (bvh.h::BvhBuildInput)
   if (getenv("FORCELEAF")) B.forceLeaf = atoi(getenv("FORCELEAF")) != 0;
   B.nodes = (F4 *)std::malloc(sizeof(F4) * 4 * (ntris + 1));
   Box rb; int root2 = B.build(0, ntris, 0, rb);
   gItems.resize(ntris); for (size_t i = 0; i < ntris; ++i) gItems[i] = tris[B.items[i].id];
   printf("binary tree: %d nodes, %.1f s\n", B.nextNode.load(), std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
   if (const char *dps = getenv("DPCOLLAPSE")) {
      float cn = 1, cp = 1; int pm = maxLeaf; sscanf(dps, "%f,%f,%d", &cn, &cp, &pm);
      gDP.cn = cn; gDP.cp = cp; gDP.pmax = std::min(15, pm); gDP.W = 4; gDP.bin = B.nodes;
      const size_t nn = (size_t)B.nextNode.load();
      gDP.c1.assign(nn, 0); gDP.c2.assign(nn, 0); gDP.c3.assign(nn, 0); gDP.first.assign(nn, 0); gDP.cnt.assign(nn, 0);
      Box bx; if (root2 >= 0) dpRec(root2, bx);
      printf("DP collapse: cn %.2f cp %.2f pmax %d, SAH cost of the root %.4g\n", cn, cp, gDP.pmax, root2 >= 0 ? gDP.c1[root2] / bx.area() : 0.0);
   }
   WTree T4 = collapse(B.nodes, root2, 4, getenv("NOQUANT") == nullptr, false);   // NOQUANT: what the 8-bit child boxes cost in visits
   WTree T8 = collapse(B.nodes, root2, 8, true, false);
   WTree T8o = collapse(B.nodes, root2, 8, true, true);
   printf("BVH4 %zu nodes, BVH8 %zu nodes\n", T4.nodes.size(), T8.nodes.size());

   // ---- populations: camera rays of `rows` rows of the 3842 x 2162 sample extent, then bounces
   const float camZ = -320.0f * ext / 100.0f;
   const int EW = 3842, EH = 2162;
   const float tanH = tanf(20.0f * BL_PI / 180.0f);   // fov 40 = vertical? Camera.hs: fov across the larger... keep simple: vertical
   std::vector<Ray> cam;
   uint64_t rk = 12345;
   const int y0 = EH / 2 - rows / 2;
   for (int y = y0; y < y0 + rows; ++y) for (int x = 0; x < EW; ++x) {
      float px = ((x + u01r(rk)) / EW * 2 - 1) * tanH * ((float)EW / EH), py = (1 - (y + u01r(rk)) / EH * 2) * tanH;
      Ray r; r.o = mk3(0, 0, camZ); r.d = normalize3(mk3(px, py, 1)); r.tmin = 0; r.tmax = BL_INF; cam.push_back(r);
   }
   Policy base; base.name = "base(sort4,vote)";
   std::vector<Policy> pols;
   pols.push_back(base);
   { Policy p = base; p.name = "cull-on-pop"; p.cullOnPop = true; p.cNode = 177; pols.push_back(p); }
   { Policy p = base; p.name = "cull+nearest-first-only"; p.cullOnPop = true; p.sortChildren = false; p.nearestFirstOnly = true; p.cNode = 160; pols.push_back(p); }
   { Policy p = base; p.name = "slot-order(no sort)"; p.sortChildren = false; p.cNode = 140; pols.push_back(p); }
   { Policy p = base; p.name = "nearest-first-only"; p.sortChildren = false; p.nearestFirstOnly = true; p.cNode = 152; pols.push_back(p); }
   { Policy p = base; p.name = "if-if"; p.ifif = true; pols.push_back(p); }
   { Policy p = base; p.name = "bvh8 sorted+cull"; p.W = 8; p.cullOnPop = true; p.cNode = 300; pols.push_back(p); }
   { Policy p = base; p.name = "bvh8 octant+cull"; p.W = 8; p.octant = true; p.sortChildren = false; p.cullOnPop = true; p.cNode = 215; p.cTrip = 50; pols.push_back(p); }
   { Policy p = base; p.name = "bvh8 octant"; p.W = 8; p.octant = true; p.sortChildren = false; p.cNode = 210; p.cTrip = 50; pols.push_back(p); }
   { Policy p = base; p.name = "defer-leaves"; p.deferLeaves = true; pols.push_back(p); }
   { Policy p = base; p.name = "defer flush24"; p.deferLeaves = true; p.flushAt = 24; pols.push_back(p); }
   { Policy p = base; p.name = "defer flush16"; p.deferLeaves = true; p.flushAt = 16; pols.push_back(p); }
   { Policy p = base; p.name = "defer refill2"; p.deferLeaves = true; p.refillAt = 2; pols.push_back(p); }
   { Policy p = base; p.name = "defer refill1 flush24"; p.deferLeaves = true; p.refillAt = 1; p.flushAt = 24; pols.push_back(p); }
   { Policy p = base; p.name = "defer nosort(any)"; p.deferLeaves = true; p.sortChildren = false; p.nearestFirstOnly = true; p.cNode = 152; pols.push_back(p); }
   { Policy p = base; p.name = "defer bvh8 octant"; p.deferLeaves = true; p.W = 8; p.octant = true; p.sortChildren = false; p.cNode = 210; p.cTrip = 50; pols.push_back(p); }
   { Policy p = base; p.name = "defer bvh8 octant+cull"; p.deferLeaves = true; p.W = 8; p.octant = true; p.sortChildren = false; p.cullOnPop = true; p.cNode = 215; p.cTrip = 50; pols.push_back(p); }
   { Policy p = base; p.name = "defer cull"; p.deferLeaves = true; p.cullOnPop = true; p.cNode = 177; pols.push_back(p); }
   { Policy p = base; p.name = "defer any:longest"; p.deferLeaves = true; p.anyOrder = 1; pols.push_back(p); }
   { Policy p = base; p.name = "defer any:leaves-first"; p.deferLeaves = true; p.anyOrder = 2; pols.push_back(p); }
   { Policy p = base; p.name = "defer any:farthest"; p.deferLeaves = true; p.anyOrder = 3; pols.push_back(p); }
   { Policy p = base; p.name = "defer any:longest-first-only"; p.deferLeaves = true; p.anyOrder = 5; pols.push_back(p); }
   { Policy p = base; p.name = "defer any:leaves-last"; p.deferLeaves = true; p.anyOrder = 4; pols.push_back(p); }
   { Policy p = base; p.name = "defer any:static-largest"; p.deferLeaves = true; p.anyOrder = 7; p.cNode = 140; pols.push_back(p); }
   { Policy p = base; p.name = "defer any:static-smallest"; p.deferLeaves = true; p.anyOrder = 8; p.cNode = 140; pols.push_back(p); }

   if (const char *f = getenv("POLS")) { std::vector<Policy> keep; keep.push_back(pols[0]); std::string fs(f); for (size_t i = 1; i < pols.size(); ++i) if (fs.find(std::string("|") + pols[i].name + "|") != std::string::npos) keep.push_back(pols[i]); pols = keep; }
   std::vector<Ray> ext_ = cam;
   if (getenv("RANDOM_RAYS")) { ext_.clear(); for (int i = 0; i < 60000; ++i) { Ray r; r.o = mk3((u01r(rk) * 2 - 1) * ext, (u01r(rk) * 2 - 1) * ext, (u01r(rk) * 2 - 1) * ext); r.d = uniformSphere(rk); r.tmin = 1e-3f; r.tmax = BL_INF; ext_.push_back(r); } }
   for (int depth = 0; depth <= 2; ++depth) {
      // nearest-hit population `ext_` at this depth
      std::vector<HitRec> hits;
      char nm[32];
      for (int sorted = 0; sorted < (getenv("SORTED") ? 2 : 1); ++sorted) {
         std::vector<Ray> rs = sorted ? sortRays(ext_, ext * 1.01f + 1, 6, true) : ext_;
         if (depth == 0 && sorted) continue;
         snprintf(nm, sizeof nm, "ext%d%s", depth, sorted ? "-srt" : "");
         for (const Policy &P : pols) {
            const WTree &T = P.W == 8 ? (P.octant ? T8o : T8) : T4;
            std::vector<HitRec> h; Stats S = P.deferLeaves ? simulateDeferred(T, P, rs, false, &h, nullptr) : simulate(T, P, rs, false, &h, nullptr);
            report(nm, P, S);
            if (!sorted && &P == &pols[0]) hits = h;
            else if (!sorted) { size_t bad = 0; for (size_t i = 0; i < h.size(); ++i) if (h[i].prim != hits[i].prim || h[i].t != hits[i].t) bad++; if (bad) printf("   !! %zu hits differ from base\n", bad); }
         }
      }
      // spawn: extension (cosine), MIS any-hit (cosine), shadow any-hit (uniform sphere, same side only)
      std::vector<Ray> nxt, anyRays;
      for (size_t i = 0; i < ext_.size(); ++i) {
         if (hits[i].prim < 0) continue;
         const Tri &tr = gItems[hits[i].prim];
         V3 n = normalize3(cross3(tr.e1, tr.e2)); if (dot3(n, ext_[i].d) > 0) n = -n;
         V3 p = rayAt(ext_[i], hits[i].t);
         const float eps_ = 1e-3f * hits[i].t;
         V3 ws = uniformSphere(rk);
         if (dot3(ws, n) > 0) { Ray s; s.o = p; s.d = ws; s.tmin = eps_; s.tmax = BL_INF; anyRays.push_back(s); }
         { Ray m; m.o = p; m.d = cosineDir(n, rk); m.tmin = eps_; m.tmax = BL_INF; anyRays.push_back(m); }
         { Ray x; x.o = p; x.d = cosineDir(n, rk); x.tmin = eps_; x.tmax = BL_INF; nxt.push_back(x); }
      }
      for (int sorted = 0; sorted < (getenv("SORTED") ? 2 : 1); ++sorted) {
         std::vector<Ray> rs = sorted ? sortRays(anyRays, ext * 1.01f + 1, 6, true) : anyRays;
         snprintf(nm, sizeof nm, "any%d%s", depth, sorted ? "-srt" : "");
         size_t occ0 = 0;
         for (const Policy &P : pols) {
            if (P.cullOnPop) continue;
            if (hits.empty()) {}
            const WTree &T = P.W == 8 ? (P.octant ? T8o : T8) : T4;
            std::vector<uint8_t> oc; Stats S = P.deferLeaves ? simulateDeferred(T, P, rs, true, nullptr, &oc) : simulate(T, P, rs, true, nullptr, &oc);
            size_t no = 0; for (uint8_t b : oc) no += b;
            report(nm, P, S);
            if (&P == &pols[0]) { occ0 = no; printf("   occluded fraction %.3f\n", (double)no / rs.size()); } else if (no != occ0) printf("   !! occluded count differs %zu vs %zu\n", no, occ0);
         }
      }
      ext_ = nxt;
   }
   return 0;
}
