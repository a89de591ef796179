// ORACLE (test infrastructure, NOT product code). See oracle_math.h header. PARITY UNPINNED.
// Geometry: triangles, analytic shapes, kd-tree (build + traversal), brute-force nearest hit.
#pragma once
#include "oracle_math.h"
#include "../include/blingcu.h"

namespace orc {

// DifferentialGeometry.hs:25-35 (dndu/dndv omitted: only bump mapping reads them, which is out of scope)
struct DG {
   V3 p, n;
   float u, v;
   V3 dpdu, dpdv;
   bool tri; float b1, b2;
};
// DifferentialGeometry.hs:40-51
static inline DG mkDg(V3 p, float u, float v, V3 dpdu, V3 dpdv) {
   return DG{p, normalize(cross(dpdu, dpdv)), u, v, dpdu, dpdv, false, 0, 0};
}
static inline DG mkDgN(V3 p, V3 n) {  // mkDg'
   Frame f = coordinateSystem(n);
   return DG{p, n, 0, 0, f.s, f.t, false, 0, 0};
}
// DifferentialGeometry.hs:72-81 ; o2w matrix and its inverse
static inline DG transDg(const float *m, const float *mi, const DG &d) {
   DG r = d;
   r.p = transPoint(m, d.p);
   r.n = normalize(transNormalInv(mi, d.n));
   r.dpdu = transVector(m, d.dpdu);
   r.dpdv = transVector(m, d.dpdv);
   return r;
}

struct ShapeHit { float t, eps; DG dg; };

// Shape.hs:81-229
static inline bool shapeIntersect(const blingcu_shape &s, const Ray &ray, ShapeHit &out) {
   const float *P = s.p;
   const V3 ro = ray.o, rd = ray.d;
   const float tmin = ray.tmin, tmax = ray.tmax;
   switch (s.kind) {
   case BLINGCU_SHAPE_BOX: {  // :81-107
      V3 pmin = mk(P[0], P[1], P[2]), pmax = mk(P[3], P[4], P[5]);
      float nearT = -kInf, farT = kInf; int dd = 0;
      for (int dim = 0; dim < 3; ++dim) {
         if (nearT > farT) return false;
         float oc = ro[dim], dInv = 1.0f / rd[dim];
         float t1p = (pmax[dim] - oc) * dInv, t2p = (pmin[dim] - oc) * dInv;
         float t1, t2;
         if (t1p > t2p) { t1 = t2p; t2 = t1p; } else { t1 = t1p; t2 = t2p; }
         if (nearT < t1) dd = dim;
         nearT = hmax(nearT, t1);
         farT = hmin(farT, t2);
      }
      if (nearT > farT) return false;
      float t0 = hmin(nearT, farT), t1 = hmax(nearT, farT);
      if (t0 > tmax || t0 < tmin) return false;
      float t = (t0 < tmin) ? t1 : t0;
      if (t > tmax) return false;
      V3 p = rayAt(ray, t);
      float half = (pmin[dd] + pmax[dd]) / 2;
      float dir = (p[dd] > half) ? 1.0f : -1.0f;
      V3 n = normalize(setc(dd, dir, mk(0, 0, 0)));
      out = ShapeHit{t, 5e-4f * t, mkDgN(p, n)};
      return true;
   }
   case BLINGCU_SHAPE_CYLINDER: {  // :109-139
      float r = P[0], zmin = P[1], zmax = P[2], phimax = P[3];
      float a = rd.x * rd.x + rd.y * rd.y;
      float b = 2 * (rd.x * ro.x + rd.y * ro.y);
      float c = ro.x * ro.x + ro.y * ro.y - r * r;
      float t0, t1;
      if (!solveQuadric(a, b, c, t0, t1)) return false;
      if (t0 > tmax) return false;
      if (t1 < tmin) return false;
      V3 h0 = rayAt(ray, t0), h1 = rayAt(ray, t1);
      float phi0 = atan2p(h0.y, h0.x), phi1 = atan2p(h1.y, h1.x);
      V3 pHit; float t;
      if (t0 > tmin && h0.z > zmin && h0.z < zmax && phi0 <= phimax) { pHit = h0; t = t0; }
      else if (t1 <= tmax && h1.z > zmin && h1.z < zmax && phi1 <= phimax) { pHit = h1; t = t1; }
      else return false;
      V3 dpdu = mk(-phimax * pHit.y, phimax * pHit.x, 0), dpdv = mk(0, 0, zmax - zmin);
      V3 n = normalize(cross(dpdu, dpdv));
      out = ShapeHit{t, 5e-4f * t, mkDgN(pHit, n)};
      return true;
   }
   case BLINGCU_SHAPE_DISK: {  // :141-155
      float h = P[0], rad = P[1], irad = P[2], phimax = P[3];
      if (std::fabs(rd.z) < 1e-7f) return false;
      float t = (h - ro.z) / rd.z;
      if (t < tmin || t > tmax) return false;
      V3 p = rayAt(ray, t);
      float d2 = p.x * p.x + p.y * p.y;
      if (d2 > rad * rad || d2 < irad * irad) return false;
      if (atan2p(p.y, p.x) > phimax) return false;
      out = ShapeHit{t, 5e-4f * t, mkDgN(p, mk(0, 0, -1))};
      return true;
   }
   case BLINGCU_SHAPE_QUAD: {  // :157-172
      float sx = P[0], sy = P[1];
      if (std::fabs(rd.z) < 1e-7f) return false;
      float t = -(ro.z) / rd.z;
      if (t < tmin || t > tmax) return false;
      V3 p = rayAt(ray, t);
      if (std::fabs(p.x) > sx || std::fabs(p.y) > sy) return false;
      float u = (sx + p.x) / (2 * sx), v = (sy + p.y) / (2 * sy);
      out = ShapeHit{t, 5e-4f * t, mkDg(p, u, v, mk(sx, 0, 0), mk(0, sy, 0))};
      return true;
   }
   case BLINGCU_SHAPE_SPHERE: {  // :174-229
      float r = P[0];
      float a = sqLen(rd), b = 2 * dot(ro, rd), c = sqLen(ro) - (r * r);
      float t1, t2;
      if (!solveQuadric(a, b, c, t1, t2)) return false;
      if (t1 > tmax) return false;
      if (t2 < tmin) return false;
      float t = (t1 < tmin) ? t2 : t1;
      if (t > tmax) return false;
      const float thetaMin = kPi, thetaMax = 0, phiMax = kTwoPi;
      V3 p = rayAt(ray, t);
      float phi = atan2p(p.y, p.x);
      float u = phi / phiMax;
      float theta = std::acos(clampf(p.z / r, -1, 1));
      float v = (theta - thetaMin) / (thetaMax - thetaMin);
      float zradius = std::sqrt(p.x * p.x + p.y * p.y);
      float invz = 1 / zradius;
      float cosphi = p.x * invz, sinphi = p.y * invz;
      V3 dpdu = mk(-phiMax * p.y, phiMax * p.x, 0);
      float dth = thetaMax - thetaMin;
      V3 dpdv = mk(p.z * cosphi, p.z * sinphi, -r * std::sin(theta)) * mk(dth, dth, dth);
      out = ShapeHit{t, 5e-4f * t, mkDg(p, u, v, dpdu, dpdv)};
      return true;
   }
   }
   return false;
}

// Shape.hs:231-284
static inline bool shapeIntersects(const blingcu_shape &s, const Ray &ray) {
   const float *P = s.p;
   const V3 ro = ray.o, rd = ray.d;
   const float tmin = ray.tmin, tmax = ray.tmax;
   switch (s.kind) {
   case BLINGCU_SHAPE_BOX: {
      float a, b;
      return intersectAABB(AABB{mk(P[0], P[1], P[2]), mk(P[3], P[4], P[5])}, ray, a, b);
   }
   case BLINGCU_SHAPE_CYLINDER: {
      float r = P[0], zmin = P[1], zmax = P[2], phimax = P[3];
      float a = rd.x * rd.x + rd.y * rd.y;
      float b = 2 * (rd.x * ro.x + rd.y * ro.y);
      float c = ro.x * ro.x + ro.y * ro.y - r * r;
      float t0, t1;
      if (!solveQuadric(a, b, c, t0, t1)) return false;
      if (t0 > tmax) return false;
      if (t1 < tmin) return false;
      V3 h0 = rayAt(ray, t0), h1 = rayAt(ray, t1);
      float phi0 = atan2p(h0.y, h0.x), phi1 = atan2p(h1.y, h1.x);
      if (t0 > tmin && h0.z > zmin && h0.z < zmax && phi0 <= phimax) return true;
      if (t1 < tmax && h1.z > zmin && h1.z < zmax && phi1 <= phimax && t1 <= tmax) return true;
      return false;
   }
   case BLINGCU_SHAPE_DISK: {
      float h = P[0], rad = P[1], irad = P[2], phimax = P[3];
      if (std::fabs(rd.z) < 1e-7f) return false;
      float t = (h - ro.z) / rd.z;
      if (t < tmin || t > tmax) return false;
      V3 p = rayAt(ray, t);
      float d2 = p.x * p.x + p.y * p.y;
      if (d2 > rad * rad || d2 < irad * irad) return false;
      if (atan2p(p.y, p.x) > phimax) return false;
      return true;
   }
   case BLINGCU_SHAPE_QUAD: {
      float sx = P[0], sy = P[1];
      if (std::fabs(rd.z) < 1e-7f) return false;
      float t = -(ro.z) / rd.z;
      if (t < tmin || t > tmax) return false;
      V3 p = rayAt(ray, t);
      if (std::fabs(p.x) > sx || std::fabs(p.y) > sy) return false;
      return true;
   }
   case BLINGCU_SHAPE_SPHERE: {
      float rad = P[0];
      float a = sqLen(rd), b = 2 * dot(ro, rd), c = sqLen(ro) - (rad * rad);
      float t0, t1;
      if (!solveQuadric(a, b, c, t0, t1)) return false;
      if (t0 > tmax || t1 < tmin) return false;
      if (t0 < tmin) return t1 < tmax;
      return true;
   }
   }
   return false;
}

static inline AABB shapeObjectBounds(const blingcu_shape &s) {  // Shape.hs:297-311
   const float *P = s.p;
   switch (s.kind) {
   case BLINGCU_SHAPE_BOX: return AABB{mk(P[0], P[1], P[2]), mk(P[3], P[4], P[5])};
   case BLINGCU_SHAPE_CYLINDER: return AABB{mk(-P[0], -P[0], P[1]), mk(P[0], P[0], P[2])};
   case BLINGCU_SHAPE_DISK: return AABB{mk(-P[1], -P[1], P[0]), mk(P[1], P[1], P[0])};
   case BLINGCU_SHAPE_QUAD: return AABB{mk(-P[0], -P[1], 0), mk(P[0], P[1], 0)};
   default: return AABB{mk(-P[0], -P[0], -P[0]), mk(P[0], P[0], P[0])};
   }
}
static inline float shapeArea(const blingcu_shape &s) {  // Shape.hs:314-328
   const float *P = s.p;
   switch (s.kind) {
   case BLINGCU_SHAPE_BOX: { float h = P[3] - P[0], w = P[4] - P[1], l = P[5] - P[2]; return 2 * (h * w + h * l + w * l); }
   case BLINGCU_SHAPE_CYLINDER: return 2 * kPi * P[0] * (P[2] - P[1]);
   case BLINGCU_SHAPE_DISK: return kPi * (P[1] * P[1] - P[2] * P[2]);
   case BLINGCU_SHAPE_QUAD: return 4 * P[0] * P[1];
   default: return P[0] * P[0] * 4 * kPi;
   }
}

// ---------------------------------------------------------------------------------------------
// primitives
// ---------------------------------------------------------------------------------------------
struct Hit {  // Primitive.hs:49-55 Intersection (without the lazy Bsdf)
   float t, eps;
   DG dg;       // geometric DG in world space
   int prim;    // prim id
   bool valid;
};

struct Prim {
   bool is_tri;
   uint32_t idx;  // triangle index or shape index
   AABB wb;
};

struct Tri { V3 p1, p2, p3; float uv[6]; };

// TriangleMesh.hs:160-207
static inline bool triangleIntersect(const Tri &T, const Ray &r, float &tOut, DG &dg) {
   V3 e1 = T.p2 - T.p1, e2 = T.p3 - T.p1;
   V3 s1 = cross(r.d, e2);
   float divisor = dot(s1, e1);
   if (divisor == 0) return false;
   float invDiv = 1 / divisor;
   V3 d = r.o - T.p1;
   float b1 = dot(d, s1) * invDiv;
   if (b1 < 0 || b1 > 1) return false;
   V3 s2 = cross(d, e1);
   float b2 = dot(r.d, s2) * invDiv;
   if (b2 < 0 || b1 + b2 > 1) return false;
   float t = dot(e2, s2) * invDiv;
   if (t < r.tmin || t > r.tmax) return false;
   V3 n = normalize(cross(e1, e2));
   const float *uv = T.uv;
   float du1 = uv[0] - uv[4], du2 = uv[2] - uv[4], dv1 = uv[1] - uv[5], dv2 = uv[3] - uv[5];
   V3 dp1 = T.p1 - T.p3, dp2 = T.p2 - T.p3;
   float det = du1 * dv2 - dv1 * du2;
   V3 dpdu, dpdv;
   if (det == 0) { Frame f = coordinateSystem(n); dpdu = f.s; dpdv = f.t; }
   else {
      float invDet = 1 / det;
      dpdu = scl(invDet, scl(dv2, dp1) - scl(dv1, dp2));
      dpdv = scl(invDet, scl(-du2, dp1) + scl(du1, dp2));
   }
   float b0 = 1 - b1 - b2;
   float tu = b0 * uv[0] + b1 * uv[2] + b2 * uv[4];
   float tv = b0 * uv[1] + b1 * uv[3] + b2 * uv[5];
   dg = DG{rayAt(r, t), normalize(cross(dpdu, dpdv)), tu, tv, dpdu, dpdv, true, b1, b2};
   tOut = t;
   return true;
}
// TriangleMesh.hs:140-158
static inline bool triangleIntersects(const Tri &T, const Ray &r) {
   V3 e1 = T.p2 - T.p1, e2 = T.p3 - T.p1;
   V3 s1 = cross(r.d, e2);
   float divisor = dot(s1, e1);
   if (divisor == 0) return false;
   float invDiv = 1 / divisor;
   V3 d = r.o - T.p1;
   float b1 = dot(d, s1) * invDiv;
   if (b1 < 0 || b1 > 1) return false;
   V3 s2 = cross(d, e1);
   float b2 = dot(r.d, s2) * invDiv;
   if (b2 < 0 || b1 + b2 > 1) return false;
   float t = dot(e2, s2) * invDiv;
   if (t < r.tmin || t > r.tmax) return false;
   return true;
}

struct KdNode {  // KdTree.hs:31-33
   int left, right;   // interior: child node indices; leaf: left = -1
   float sp; int axis;
   uint32_t first, count;  // leaf: range in leafPrims
};

struct Geometry {
   std::vector<Tri> tris;
   std::vector<blingcu_shape> shapes;
   std::vector<Prim> prims;          // indexed by prim id
   // kd-tree
   AABB bounds;
   std::vector<KdNode> nodes;
   std::vector<uint32_t> leafPrims;
   int root = -1;
   bool kd_built = false;

   // Primitive.hs:29-43 near / Geometry.hs:33-36 / TriangleMesh.hs:160
   inline bool primIntersect(uint32_t pid, const Ray &r, Hit &h) const {
      const Prim &p = prims[pid];
      if (p.is_tri) {
         float t; DG dg;
         if (!triangleIntersect(tris[p.idx], r, t, dg)) return false;
         h = Hit{t, 1e-3f * t, dg, (int)pid, true};
         return true;
      }
      const blingcu_shape &s = shapes[p.idx];
      ShapeHit sh;
      if (!shapeIntersect(s, transRay(s.w2o, r), sh)) return false;
      h = Hit{sh.t, sh.eps, transDg(s.o2w, s.w2o, sh.dg), (int)pid, true};
      return true;
   }
   inline bool primIntersects(uint32_t pid, const Ray &r) const {
      const Prim &p = prims[pid];
      if (p.is_tri) return triangleIntersects(tris[p.idx], r);
      const blingcu_shape &s = shapes[p.idx];
      return shapeIntersects(s, transRay(s.w2o, r));
   }

   // ground truth: fold `near` over every primitive in prim-id order (Primitive.hs:29-43)
   Hit bruteNearest(Ray r) const {
      Hit best; best.valid = false; best.prim = -1; best.t = 0; best.eps = 0;
      for (uint32_t i = 0; i < prims.size(); ++i) {
         Hit h;
         if (primIntersect(i, r, h)) { r.tmax = h.t; best = h; }
      }
      return best;
   }
   bool bruteOccluded(const Ray &r) const {
      for (uint32_t i = 0; i < prims.size(); ++i) if (primIntersects(i, r)) return true;
      return false;
   }

   void buildKd();
   // KdTree.hs:223-246, with the counters of dbgTraverse (:260-281)
   Hit kdNearest(const Ray &r, uint64_t *nodesTraversed = nullptr, uint64_t *intersections = nullptr) const;
   bool kdOccluded(const Ray &r) const;

private:
   int buildTree(const AABB &b, std::vector<uint32_t> &ps, int depth);
   void trav(Ray &r, Hit &best, V3 inv, int node, float tmin, float tmax, uint64_t *nt, uint64_t *ni) const;
   bool travAny(const Ray &r, V3 inv, int node, float tmin, float tmax) const;
};

}  // namespace orc
