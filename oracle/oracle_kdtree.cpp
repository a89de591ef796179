// ORACLE (test infrastructure, NOT product code). See oracle_math.h header. PARITY UNPINNED.
// SAH kd-tree build and traversal, restating Primitive/KdTree.hs.
#include "oracle_scene.h"
#include <cmath>
#include <future>

namespace orc {

static const float cT = 1.0f;   // KdTree.hs:82-83
static const float cI = 80.0f;  // KdTree.hs:86-87

struct Edge { uint32_t prim; float t; bool start; };  // KdTree.hs:89-92

// KdTree.hs:186-203
static inline float sahCost(const AABB &b, int a0, float t, int nl, int nr) {
   V3 d = b.hi - b.lo;
   int a1 = (a0 + 1) % 3, a2 = (a0 + 2) % 3;
   float dl = t - b.lo[a0], dr = b.hi[a0] - t;
   float sal = 2 * (d[a1] * d[a2] + dl * (d[a1] + d[a2]));
   float sar = 2 * (d[a1] * d[a2] + dr * (d[a1] + d[a2]));
   float invTot = 1 / surfaceArea(b);
   float pl = sal * invTot, pr = sar * invTot;
   float eb = (nl == 0 || nr == 0) ? 0.5f : 1.0f;
   float pI = pl * (float)nl + pr * (float)nr;
   return cT + cI * eb * pI;
}

struct Built { std::vector<KdNode> nodes; std::vector<uint32_t> leaf; int root; };

struct Builder {
   const std::vector<Prim> &prims;
   int parDepth;

   // returns a subtree in its own arrays (merged by the caller) so the left subtree can be built
   // on another thread like `par left` (KdTree.hs:130)
   void build(const AABB &bounds, std::vector<uint32_t> ps, int depth, int level, Built &out) {
      auto mkLeaf = [&]() {
         KdNode n; n.left = -1; n.right = -1; n.sp = 0; n.axis = 0;
         n.first = (uint32_t)out.leaf.size(); n.count = (uint32_t)ps.size();
         out.leaf.insert(out.leaf.end(), ps.begin(), ps.end());
         out.nodes.push_back(n);
         out.root = (int)out.nodes.size() - 1;
      };
      if (depth == 0 || ps.size() <= 1) { mkLeaf(); return; }  // KdTree.hs:114-119
      // trySplit (KdTree.hs:121-140): axes starting at the maximum extent; first axis that beats oldCost wins
      int a0 = dominant(bounds.hi - bounds.lo);
      float oldCost = cI * (float)ps.size();
      for (int k = 0; k < 3; ++k) {
         int axis = (a0 + k) % 3;
         // edges (KdTree.hs:167-183)
         std::vector<Edge> es(2 * ps.size());
         for (size_t i = 0; i < ps.size(); ++i) {
            const AABB &wb = prims[ps[i]].wb;
            es[2 * i] = Edge{ps[i], wb.lo[axis], true};
            es[2 * i + 1] = Edge{ps[i], wb.hi[axis], false};
         }
         std::sort(es.begin(), es.end(), [](const Edge &a, const Edge &b) {  // Ord Edge, KdTree.hs:97-100
            if (a.t == b.t) return a.start && !b.start;
            return a.t < b.t;
         });
         // allSplits + filterSplits + bestSplit (KdTree.hs:142-165)
         float bmin = bounds.lo[axis], bmax = bounds.hi[axis];
         int l = 0, r = (int)ps.size();
         float bestC = kInf; int bestI = -1; float bestT = 0;
         for (size_t i = 0; i < es.size(); ++i) {
            int nl, nr;
            if (!es[i].start) { r -= 1; nl = l; nr = r; }
            else { nl = l; nr = r; }
            float t = es[i].t;
            if (t > bmin && t < bmax) {
               float c = sahCost(bounds, axis, t, nl, nr);
               if (c < bestC) { bestC = c; bestI = (int)i; bestT = t; }
            }
            if (es[i].start) l += 1;
         }
         if (bestI < 0) continue;            // null fs
         if (!(bestC < oldCost)) continue;   // otherwise = go o axs
         // partition (KdTree.hs:147-151)
         std::vector<uint32_t> lp, rp;
         for (int i = 0; i < bestI; ++i) if (es[i].start) lp.push_back(es[i].prim);
         for (size_t i = (size_t)bestI + 1; i < es.size(); ++i) if (!es[i].start) rp.push_back(es[i].prim);
         std::vector<Edge>().swap(es);
         std::vector<uint32_t>().swap(ps);
         AABB lb = bounds, rb = bounds;  // splitAABB, AABB.hs:63-67
         lb.hi = setc(axis, bestT, bounds.hi);
         rb.lo = setc(axis, bestT, bounds.lo);
         Built L, R;
         if (level < parDepth) {
            auto fut = std::async(std::launch::async, [&]() { build(lb, std::move(lp), depth - 1, level + 1, L); });
            build(rb, std::move(rp), depth - 1, level + 1, R);
            fut.get();
         } else {
            build(lb, std::move(lp), depth - 1, level + 1, L);
            build(rb, std::move(rp), depth - 1, level + 1, R);
         }
         // merge: [out | L | R | interior]
         int offL = (int)out.nodes.size(); uint32_t leafL = (uint32_t)out.leaf.size();
         for (KdNode n : L.nodes) {
            if (n.left >= 0) { n.left += offL; n.right += offL; } else n.first += leafL;
            out.nodes.push_back(n);
         }
         out.leaf.insert(out.leaf.end(), L.leaf.begin(), L.leaf.end());
         int offR = (int)out.nodes.size(); uint32_t leafR = (uint32_t)out.leaf.size();
         for (KdNode n : R.nodes) {
            if (n.left >= 0) { n.left += offR; n.right += offR; } else n.first += leafR;
            out.nodes.push_back(n);
         }
         out.leaf.insert(out.leaf.end(), R.leaf.begin(), R.leaf.end());
         KdNode in; in.left = L.root + offL; in.right = R.root + offR; in.sp = bestT; in.axis = axis; in.first = 0; in.count = 0;
         out.nodes.push_back(in);
         out.root = (int)out.nodes.size() - 1;
         return;
      }
      mkLeaf();
   }
};

// KdTree.hs:107-112
void Geometry::buildKd() {
   bounds = emptyBox();
   for (const Prim &p : prims) bounds = extendB(bounds, p.wb);
   std::vector<uint32_t> ps(prims.size());
   for (uint32_t i = 0; i < prims.size(); ++i) ps[i] = i;
   int md = (int)std::nearbyint(8 + 3 * std::log((float)prims.size()));
   if (prims.empty()) md = 0;
   Builder b{prims, 4};
   Built out; out.root = -1;
   b.build(bounds, std::move(ps), md, 0, out);
   nodes.swap(out.nodes);
   leafPrims.swap(out.leaf);
   root = out.root;
   kd_built = true;
}

// KdTree.hs:223-234 traverse
void Geometry::trav(Ray &r, Hit &best, V3 inv, int ni, float tmin, float tmax, uint64_t *nt, uint64_t *nint) const {
   const KdNode &n = nodes[ni];
   if (n.left < 0) {  // Leaf: nearest' ps ri
      if (nt) { *nt += 1; *nint += n.count; }
      for (uint32_t i = 0; i < n.count; ++i) {
         Hit h;
         if (primIntersect(leafPrims[n.first + i], r, h)) { r.tmax = h.t; best = h; }
      }
      return;
   }
   if (r.tmax < tmin) return;
   float oa = r.o[n.axis], da = r.d[n.axis];
   float tp = (n.sp - oa) * inv[n.axis];
   bool lf = (oa < n.sp) || (oa == n.sp && da <= 0);
   int fc = lf ? n.left : n.right, sc = lf ? n.right : n.left;
   if (tp > tmax || tp <= 0) { if (nt) *nt += 1; trav(r, best, inv, fc, tmin, tmax, nt, nint); }
   else if (tp < tmin) { if (nt) *nt += 1; trav(r, best, inv, sc, tmin, tmax, nt, nint); }
   else {
      trav(r, best, inv, fc, tmin, tp, nt, nint);
      trav(r, best, inv, sc, tp, tmax, nt, nint);
   }
}

// KdTree.hs:240-242
Hit Geometry::kdNearest(const Ray &r0, uint64_t *nt, uint64_t *nint) const {
   Hit best; best.valid = false; best.prim = -1; best.t = 0; best.eps = 0;
   float tn, tf;
   if (root < 0 || !intersectAABB(bounds, r0, tn, tf)) return best;
   Ray r = r0;
   V3 inv = mk(1 / r.d.x, 1 / r.d.y, 1 / r.d.z);
   trav(r, best, inv, root, tn, tf, nt, nint);
   return best;
}

// KdTree.hs:210-220 traverse'
bool Geometry::travAny(const Ray &r, V3 inv, int ni, float tmin, float tmax) const {
   const KdNode &n = nodes[ni];
   if (n.left < 0) {
      for (uint32_t i = 0; i < n.count; ++i) if (primIntersects(leafPrims[n.first + i], r)) return true;
      return false;
   }
   float oa = r.o[n.axis], da = r.d[n.axis];
   float tp = (n.sp - oa) * inv[n.axis];
   bool lf = (oa < n.sp) || (oa == n.sp && da <= 0);
   int fc = lf ? n.left : n.right, sc = lf ? n.right : n.left;
   if (tp > tmax || tp <= 0) return travAny(r, inv, fc, tmin, tmax);
   if (tp < tmin) return travAny(r, inv, sc, tmin, tmax);
   return travAny(r, inv, fc, tmin, tp) || travAny(r, inv, sc, tp, tmax);
}

// KdTree.hs:244-246
bool Geometry::kdOccluded(const Ray &r) const {
   float tn, tf;
   if (root < 0 || !intersectAABB(bounds, r, tn, tf)) return false;
   V3 inv = mk(1 / r.d.x, 1 / r.d.y, 1 / r.d.z);
   return travAny(r, inv, root, tn, tf);
}

}  // namespace orc
