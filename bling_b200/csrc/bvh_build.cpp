// bvh_build.cpp -- host-side binned-SAH builder (binary tree, then collapsed to 4-wide nodes; setup, not per-ray). Plays the role of
// mkKdTree/buildTree (KdTree.hs:107-203) for the GPU path; the tree shape is free to differ because the
// traversal result (globally nearest hit) does not depend on it.
//
// Collapse (round 2, second half): the binary tree is built down to single items and a dynamic programme over it decides, per
// subtree, whether it becomes one leaf (<= max_leaf items, cost = area * items * collapse_cp), one 4-wide node (cost = area + its
// children) or is dissolved into 2 / 3 slots of its parent -- the SAH-optimal collapse of Ylitie, Karras, Laine 2017 (section 3.1)
// restated for 4 slots. tools/travsim.cpp on the cfg-5 ray streams: same node visits as "open the largest child" over 2-item leaves,
// a quarter to a third fewer primitive tests, fewer nodes in memory. collapse_cp = 0 keeps the greedy collapse of round 1.
//
// Item boxes are inflated by eps = 4e-6 * scene extent (+ 1e-30) on every side before they are merged, so the
// slab test in bvh.h needs no epsilon: its rounding error (~2e-7 * t) is below eps * |1/d| for any ray that
// starts within ~20 scene diameters.
#include "bvh.h"
#include <algorithm>
#include <cmath>
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <future>
#include <vector>

namespace bl {

namespace {
struct Box {
   float lo[3], hi[3];
   void reset() { for (int k = 0; k < 3; ++k) { lo[k] = BL_INF; hi[k] = -BL_INF; } }
   void grow(const float *l, const float *h) { for (int k = 0; k < 3; ++k) { lo[k] = std::min(lo[k], l[k]); hi[k] = std::max(hi[k], h[k]); } }
   void grow(const Box &b) { grow(b.lo, b.hi); }
   void growP(const float *p) { for (int k = 0; k < 3; ++k) { lo[k] = std::min(lo[k], p[k]); hi[k] = std::max(hi[k], p[k]); } }
   float area() const { float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2]; return (dx < 0) ? 0.0f : 2 * (dx * dy + dx * dz + dy * dz); }
};
struct Item { float lo[3], hi[3], c[3]; uint32_t id; };

#ifndef BL_NBINS
#define BL_NBINS 16
#endif
constexpr int NBINS = BL_NBINS;

struct Builder {
   std::vector<Item> items;
   F4 *nodes;
   std::atomic<int> nextNode{0};
   int maxLeaf;
   int parLevels;
   float travCost = 0;
   bool forceLeaf = true;

   int allocNode() { return nextNode.fetch_add(1); }

   // builds [b,e), returns the child reference and its box
   int build(size_t b, size_t e, int depth, Box &box) {
      size_t n = e - b;
      box.reset();
      Box cb; cb.reset();
      for (size_t i = b; i < e; ++i) { box.grow(items[i].lo, items[i].hi); cb.growP(items[i].c); }
      if ((int)n <= maxLeaf && (forceLeaf || n <= 1)) return ~(int)((b << 4) | n);
      size_t mid = 0;
      bool done = false;
      if (depth < 32) {
         float bestCost = BL_INF; int bestAxis = -1, bestBin = -1;
         for (int axis = 0; axis < 3; ++axis) {
            float cmin = cb.lo[axis], cext = cb.hi[axis] - cb.lo[axis];
            if (!(cext > 0)) continue;
            Box bb[NBINS]; int cnt[NBINS];
            for (int k = 0; k < NBINS; ++k) { bb[k].reset(); cnt[k] = 0; }
            float scale = NBINS / cext;
            for (size_t i = b; i < e; ++i) {
               int k = std::min(NBINS - 1, std::max(0, (int)((items[i].c[axis] - cmin) * scale)));
               bb[k].grow(items[i].lo, items[i].hi); cnt[k]++;
            }
            float rightArea[NBINS]; int rightCnt[NBINS];
            Box acc; acc.reset(); int c = 0;
            for (int k = NBINS - 1; k > 0; --k) { acc.grow(bb[k]); c += cnt[k]; rightArea[k] = acc.area(); rightCnt[k] = c; }
            acc.reset(); c = 0;
            for (int k = 0; k < NBINS - 1; ++k) {
               acc.grow(bb[k]); c += cnt[k];
               if (c == 0 || rightCnt[k + 1] == 0) continue;
               float cost = acc.area() * c + rightArea[k + 1] * rightCnt[k + 1];
               if (cost < bestCost) { bestCost = cost; bestAxis = axis; bestBin = k; }
            }
         }
         float leafCost = box.area() * (float)n;
         if (bestAxis >= 0 && (bestCost + travCost * box.area() < leafCost || (int)n > 15)) {
            float cmin = cb.lo[bestAxis], scale = NBINS / (cb.hi[bestAxis] - cb.lo[bestAxis]);
            auto it = std::partition(items.begin() + b, items.begin() + e, [&](const Item &x) {
               int k = std::min(NBINS - 1, std::max(0, (int)((x.c[bestAxis] - cmin) * scale)));
               return k <= bestBin;
            });
            mid = (size_t)(it - items.begin());
            done = mid > b && mid < e;
         } else if (bestAxis >= 0 && (int)n <= 15) {
            return ~(int)((b << 4) | n);   // SAH says a leaf is cheaper
         }
      }
      if (!done) {   // median split along the widest centroid axis (also the depth guard)
         int axis = 0; float ext = -1;
         for (int k = 0; k < 3; ++k) if (cb.hi[k] - cb.lo[k] > ext) { ext = cb.hi[k] - cb.lo[k]; axis = k; }
         if (!(ext > 0) && (int)n <= 15) return ~(int)((b << 4) | n);
         mid = b + n / 2;
         std::nth_element(items.begin() + b, items.begin() + mid, items.begin() + e,
                          [axis](const Item &x, const Item &y) { return x.c[axis] < y.c[axis]; });
      }
      int idx = allocNode();
      Box lb, rb; int lc, rc;
      if (depth < parLevels && n > 65536) {
         auto fut = std::async(std::launch::async, [&]() { lc = build(b, mid, depth + 1, lb); });
         rc = build(mid, e, depth + 1, rb);
         fut.get();
      } else {
         lc = build(b, mid, depth + 1, lb);
         rc = build(mid, e, depth + 1, rb);
      }
      F4 *np = nodes + 4 * (size_t)idx;
      np[0] = F4{lb.lo[0], lb.hi[0], lb.lo[1], lb.hi[1]};
      np[1] = F4{rb.lo[0], rb.hi[0], rb.lo[1], rb.hi[1]};
      np[2] = F4{lb.lo[2], lb.hi[2], rb.lo[2], rb.hi[2]};
      np[3] = F4{i2f(lc), i2f(rc), 0, 0};
      return idx;
   }
};
// ---- SAH-optimal collapse: C(n, i) = cheapest representation of the binary subtree n in at most i slots of a 4-wide parent
struct Collapse {
   const F4 *bin = nullptr;
   std::vector<float> c1, c2, c3;
   std::vector<int> first, cnt;
   float cp = 0; int pmax = 2;
   static void childrenOf(const F4 *nodes, int idx, Box &l, int &lr, Box &r, int &rr) {
      const F4 *np = nodes + 4 * (size_t)idx;
      l.lo[0] = np[0].x; l.hi[0] = np[0].y; l.lo[1] = np[0].z; l.hi[1] = np[0].w; l.lo[2] = np[2].x; l.hi[2] = np[2].y;
      r.lo[0] = np[1].x; r.hi[0] = np[1].y; r.lo[1] = np[1].z; r.hi[1] = np[1].w; r.lo[2] = np[2].z; r.hi[2] = np[2].w;
      lr = f2i(np[3].x); rr = f2i(np[3].y);
   }
   float childC(int r, const Box &b, int i) const {
      if (r < 0) return b.area() * (float)((~r) & 15) * cp;
      return i == 1 ? c1[(size_t)r] : (i == 2 ? c2[(size_t)r] : c3[(size_t)r]);
   }
   // cheapest split of j slots between the two children of a binary node (a child never needs more than 3)
   float dist(int lr, const Box &lb, int rr, const Box &rb, int j, int *kbest) const {
      float best = BL_INF; int kb = 1;
      for (int k = 1; k < j; ++k) {
         const float c = childC(lr, lb, std::min(k, 3)) + childC(rr, rb, std::min(j - k, 3));
         if (c < best) { best = c; kb = k; }
      }
      if (kbest) *kbest = kb;
      return best;
   }
   float leafCost(int n, const Box &b) const { return cnt[(size_t)n] <= pmax ? b.area() * (float)cnt[(size_t)n] * cp : BL_INF; }
   int parLevels = 0;   // the two subtrees of the top levels are solved concurrently (they write disjoint table entries)
   void solve(int n, Box &box, int depth = 0) {   // post-order; the binary tree is at most ~60 levels deep
      Box lb, rb, tmp; int lr, rr; childrenOf(bin, n, lb, lr, rb, rr);
      if (depth < parLevels && lr >= 0 && rr >= 0) {
         auto fut = std::async(std::launch::async, [&]() { Box t2; solve(lr, t2, depth + 1); });
         solve(rr, tmp, depth + 1);
         fut.get();
      } else {
         if (lr >= 0) solve(lr, tmp, depth + 1);
         if (rr >= 0) solve(rr, tmp, depth + 1);
      }
      box = lb; box.grow(rb);
      const int lc = lr < 0 ? ((~lr) & 15) : cnt[(size_t)lr], rc = rr < 0 ? ((~rr) & 15) : cnt[(size_t)rr];
      first[(size_t)n] = lr < 0 ? (int)((uint32_t)(~lr) >> 4) : first[(size_t)lr];   // the items of a subtree are contiguous, left before right
      cnt[(size_t)n] = std::min(lc + rc, 1 << 20);
      const float cint = dist(lr, lb, rr, rb, 4, nullptr) + box.area();
      c1[(size_t)n] = std::min(leafCost(n, box), cint);
      c2[(size_t)n] = std::min(dist(lr, lb, rr, rb, 2, nullptr), c1[(size_t)n]);
      c3[(size_t)n] = std::min(dist(lr, lb, rr, rb, 3, nullptr), c2[(size_t)n]);
   }
   // the children that subtree r contributes to its wide parent when it may use i slots
   template <class ChildT> void expand(int r, const Box &b, int i, ChildT *c, int &nc) const {
      if (r < 0) { c[nc].box = b; c[nc].ref = r; nc++; return; }
      if (i > 1) {
         Box lb, rb; int lr, rr; childrenOf(bin, r, lb, lr, rb, rr);
         int k; const float cd = dist(lr, lb, rr, rb, i, &k);
         if (cd < (i == 2 ? c1[(size_t)r] : c2[(size_t)r])) { expand(lr, lb, k, c, nc); expand(rr, rb, i - k, c, nc); }
         else expand(r, b, i - 1, c, nc);
         return;
      }
      c[nc].box = b;
      c[nc].ref = (leafCost(r, b) <= c1[(size_t)r]) ? ~(int)(((uint32_t)first[(size_t)r] << 4) | (uint32_t)cnt[(size_t)r]) : r;
      nc++;
   }
};
}  // namespace

int bvhBuild(const BvhBuildInput &in, BvhBuildOutput &out) {
   std::memset(&out, 0, sizeof(out));
   out.root = -1;
   if (in.n == 0) return 0;
   if (in.n >= (1u << 26)) return 1;   // 26 index bits in the hit reference (bvh.h)
   Box scene; scene.reset();
   for (size_t i = 0; i < in.n; ++i) scene.grow(in.lo + 3 * i, in.hi + 3 * i);
   float ext = 0;
   for (int k = 0; k < 3; ++k) {
      ext = std::max(ext, scene.hi[k] - scene.lo[k]);
      ext = std::max(ext, std::max(std::fabs(scene.lo[k]), std::fabs(scene.hi[k])));
      out.scene_lo[k] = scene.lo[k]; out.scene_hi[k] = scene.hi[k];
   }
   float eps = 4e-6f * ext + 1e-30f;
   Builder B;
   B.items.resize(in.n);
   for (size_t i = 0; i < in.n; ++i) {
      Item &it = B.items[i];
      for (int k = 0; k < 3; ++k) {
         it.lo[k] = in.lo[3 * i + k] - eps; it.hi[k] = in.hi[3 * i + k] + eps;
         it.c[k] = 0.5f * (in.lo[3 * i + k] + in.hi[3 * i + k]);
      }
      it.id = (uint32_t)i;
   }
   const bool optimal = in.collapse_cp > 0;
   B.maxLeaf = optimal ? 1 : std::min(15, std::max(1, in.max_leaf));   // optimal collapse: the DP forms the leaves
   B.travCost = in.trav_cost; B.forceLeaf = in.force_leaf != 0;
   int th = std::max(1, in.threads);
   B.parLevels = 0; while ((1 << B.parLevels) < th) B.parLevels++;
   B.parLevels += 1;
   size_t maxNodes = std::max<size_t>(1, in.n);   // a binary tree over n leaves has n-1 inner nodes
   B.nodes = (F4 *)std::malloc(sizeof(F4) * 4 * (maxNodes + 1));
   Box rootBox;
   int root = B.build(0, in.n, 0, rootBox);
   if (root < 0) {   // single leaf: wrap it so traversal always starts at a node
      int idx = B.allocNode();
      F4 *np = B.nodes + 4 * (size_t)idx;
      np[0] = F4{rootBox.lo[0], rootBox.hi[0], rootBox.lo[1], rootBox.hi[1]};
      np[1] = F4{BL_INF, -BL_INF, BL_INF, -BL_INF};
      np[2] = F4{rootBox.lo[2], rootBox.hi[2], BL_INF, -BL_INF};
      np[3] = F4{i2f(root), i2f(~0), 0, 0};
      root = idx;
   }
   // ---- collapse the binary tree into 4-wide nodes (bvh.h layout): open the inner child with the largest area
   // until the node has four children; emitted in depth-first order so a subtree is contiguous in memory
   int n2 = B.nextNode.load();
   F4 *n4 = (F4 *)std::malloc(sizeof(F4) * BL_NODE_F4 * (size_t)(n2 + 1));
   int next4 = 0;
   struct Child { Box box; int ref; };
   Collapse dp;
   if (optimal) {
      dp.bin = B.nodes; dp.cp = in.collapse_cp; dp.pmax = std::min(15, std::max(1, in.max_leaf));
      dp.c1.assign((size_t)n2, 0); dp.c2.assign((size_t)n2, 0); dp.c3.assign((size_t)n2, 0); dp.first.assign((size_t)n2, 0); dp.cnt.assign((size_t)n2, 0);
      dp.parLevels = in.n > 65536 ? B.parLevels : 0;
      Box bx; dp.solve(root, bx);
   }
   auto childrenOf = [&](int idx, Child &l, Child &r) {
      const F4 *np = B.nodes + 4 * (size_t)idx;
      l.box.lo[0] = np[0].x; l.box.hi[0] = np[0].y; l.box.lo[1] = np[0].z; l.box.hi[1] = np[0].w; l.box.lo[2] = np[2].x; l.box.hi[2] = np[2].y;
      r.box.lo[0] = np[1].x; r.box.hi[0] = np[1].y; r.box.lo[1] = np[1].z; r.box.hi[1] = np[1].w; r.box.lo[2] = np[2].z; r.box.hi[2] = np[2].w;
      l.ref = f2i(np[3].x); r.ref = f2i(np[3].y);
   };
   struct Work { int n2idx; int n4idx; };
   std::vector<Work> work;
   int root4 = next4++;
   work.push_back(Work{root, root4});
   while (!work.empty()) {
      Work w = work.back(); work.pop_back();
      Child c[4]; int nc = 2;
      childrenOf(w.n2idx, c[0], c[1]);
      if (optimal) {
         const Child l = c[0], r = c[1]; int k; dp.dist(l.ref, l.box, r.ref, r.box, 4, &k);
         nc = 0; dp.expand(l.ref, l.box, k, c, nc); dp.expand(r.ref, r.box, 4 - k, c, nc);
      } else
      while (nc < 4) {
         int best = -1; float bestArea = -1;
         for (int k = 0; k < nc; ++k) if (c[k].ref >= 0 && c[k].box.area() > bestArea) { bestArea = c[k].box.area(); best = k; }
         if (best < 0) break;
         Child l, r; childrenOf(c[best].ref, l, r);
         c[best] = l; c[nc++] = r;
      }
      F4 *np = n4 + BL_NODE_F4 * (size_t)w.n4idx;
      int refs[4];
      Box nb; nb.reset();
      for (int k = 0; k < 4; ++k) {
         bool used = k < nc && !(c[k].ref < 0 && ((~c[k].ref) & 15) == 0);   // drop empty leaves
         refs[k] = used ? c[k].ref : ~0;
         if (used) nb.grow(c[k].box);
      }
      // inner children get their 4-wide index now: siblings are contiguous in memory
      for (int k = nc - 1; k >= 0; --k) if (k < nc && refs[k] >= 0) { int id = next4++; work.push_back(Work{refs[k], id}); refs[k] = id; }
      // quantise the child boxes on a node-local grid of 2^e cells: the grid origin lies 1/32 of a cell below the node box, every
      // bound is moved out by 1/64 of a cell (the kernel's folded 2^15 is good to 1/512 of a cell, bvh.h) and then rounded
      // outwards to the grid; the exponent grows until byte 255 covers the high bounds
      float P[3]; uint32_t E = 0; uint32_t qlo[3] = {0, 0, 0}, qhi[3] = {0, 0, 0};
      const double QPAD = 1.0 / 64.0;
      for (int a = 0; a < 3; ++a) {
         double lo = nb.lo[a], ext = (double)nb.hi[a] - (double)nb.lo[a];
         if (!(ext > 0)) ext = 0;
         int e = -100;
         if (ext > 0) { e = (int)std::ceil(std::log2(ext / 254.0)); while (std::ldexp(254.0, e) < ext) e++; }
         e = std::max(-120, std::min(120, e));
         for (;; ++e) {
            const double cell = std::ldexp(1.0, e);
            float Pf = (float)(lo - cell / 32);
            while ((double)Pf + cell / 64 > lo) Pf = std::nextafterf(Pf, -BL_INF);   // byte 0 must lie at least 1/64 cell below the node's low bound
            bool fits = true;
            uint32_t wl = 0, wh = 0;
            for (int k = 0; k < 4; ++k) {
               int ql = 255, qh = 0;   // unused slot: inverted, never hit
               if (refs[k] != ~0) {
                  const double ul = ((double)c[k].box.lo[a] - (double)Pf) / cell - QPAD, uh = ((double)c[k].box.hi[a] - (double)Pf) / cell + QPAD;
                  ql = (int)std::floor(ul); qh = (int)std::ceil(uh);
                  if (qh > 255 && e < 120) fits = false;
                  ql = std::max(0, std::min(255, ql)); qh = std::max(0, std::min(255, qh));
               }
               wl |= (uint32_t)ql << (8 * k); wh |= (uint32_t)qh << (8 * k);
            }
            if (!fits) continue;
            P[a] = Pf; E |= (uint32_t)(e + 127) << (8 * a); qlo[a] = wl; qhi[a] = wh;
            break;
         }
      }
      np[0] = F4{P[0], P[1], P[2], i2f((int)E)};
      np[1] = F4{i2f(refs[0]), i2f(refs[1]), i2f(refs[2]), i2f(refs[3])};
      np[2] = F4{i2f((int)qlo[0]), i2f((int)qlo[1]), i2f((int)qlo[2]), i2f((int)qhi[0])};
      np[3] = F4{i2f((int)qhi[1]), i2f((int)qhi[2]), i2f((int)BL_QMAGIC), 0};   // .z: the 2^15 bit pattern the kernel permutes bytes into
   }
   // worst-case traversal stack (entries) of the push-all-then-pop scheme: children indices are always larger than
   // the parent's, so one reverse sweep suffices
   {
      std::vector<int> cap((size_t)next4, 0);
      for (int i = next4 - 1; i >= 0; --i) {
         const F4 *np = n4 + BL_NODE_F4 * (size_t)i;
         int nc = 0, deepest = 0;
         for (int k = 0; k < 4; ++k) {
            int ref = f2i(k == 0 ? np[1].x : (k == 1 ? np[1].y : (k == 2 ? np[1].z : np[1].w)));
            if (ref == ~0) continue;
            nc++;
            if (ref >= 0) deepest = std::max(deepest, cap[(size_t)ref]);
         }
         cap[(size_t)i] = std::max(nc, nc - 1 + deepest);
      }
      out.max_stack = cap[(size_t)root4];
   }
   std::free(B.nodes);
   out.nodes = n4;
   out.n_nodes = next4;
   out.root = root4;
   out.order = (uint32_t *)std::malloc(sizeof(uint32_t) * in.n);
   for (size_t i = 0; i < in.n; ++i) out.order[i] = B.items[i].id;
   return 0;
}

}  // namespace bl
