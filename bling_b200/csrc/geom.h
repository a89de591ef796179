// geom.h -- device scene layout, ray/primitive intersection and differential geometry.
// Replaces: TriangleMesh.hs:122-207, Shape.hs:81-284, Primitive/Geometry.hs:14-36,
// DifferentialGeometry.hs:40-81. Op order follows the reference so hit decisions are bit-identical.
#pragma once
#include "hd.h"
#include "../../include/blingcu.h"

namespace bl {

struct alignas(16) F4 { float x, y, z, w; };
struct alignas(8) F2 { float x, y; };

#if defined(__CUDA_ARCH__)
__device__ __forceinline__ F4 ld4(const F4 *p) { float4 v = __ldg(reinterpret_cast<const float4 *>(p)); F4 r; r.x = v.x; r.y = v.y; r.z = v.z; r.w = v.w; return r; }
__device__ __forceinline__ int f2i(float f) { return __float_as_int(f); }
__device__ __forceinline__ float i2f(int i) { return __int_as_float(i); }
#else
inline F4 ld4(const F4 *p) { return *p; }
inline int f2i(float f) { int i; __builtin_memcpy(&i, &f, 4); return i; }
inline float i2f(int i) { float f; __builtin_memcpy(&f, &i, 4); return f; }
#endif

// DifferentialGeometry.hs:25-35 (dndu/dndv are only read by bump mapping: out of scope)
struct DG { V3 p, n; float u, v; V3 dpdu, dpdv; };
HD DG mkDg(V3 p, float u, float v, V3 dpdu, V3 dpdv) { DG d; d.p = p; d.n = normalize3(cross3(dpdu, dpdv)); d.u = u; d.v = v; d.dpdu = dpdu; d.dpdv = dpdv; return d; }
HD DG mkDgN(V3 p, V3 n) { Frame f = coordinateSystem(n); DG d; d.p = p; d.n = n; d.u = 0; d.v = 0; d.dpdu = f.s; d.dpdv = f.t; return d; }
HD DG transDg(const float *m, const float *mi, const DG &d) {   // DifferentialGeometry.hs:72-81
   DG r = d;
   r.p = transPoint(m, d.p); r.n = normalize3(transNormalInv(mi, d.n));
   r.dpdu = transVector(m, d.dpdu); r.dpdv = transVector(m, d.dpdv);
   return r;
}

// ---------------------------------------------------------------------------------------- triangle
// Moeller-Trumbore exactly as TriangleMesh.hs:160-181. e1 = p2-p1, e2 = p3-p1 are precomputed on the host in
// f32 (same subtraction, same bits). Accepts t == tmax so the later-tested primitive wins ties.
HD bool triHit(V3 p1, V3 e1, V3 e2, const Ray &r, float &t, float &b1, float &b2) {
   V3 s1 = cross3(r.d, e2);
   float divisor = dot3(s1, e1);
   if (divisor == 0.0f) return false;
   float invDiv = 1.0f / divisor;
   V3 d = r.o - p1;
   b1 = dot3(d, s1) * invDiv;
   if (b1 < 0.0f || b1 > 1.0f) return false;
   V3 s2 = cross3(d, e1);
   b2 = dot3(r.d, s2) * invDiv;
   if (b2 < 0.0f || b1 + b2 > 1.0f) return false;
   t = dot3(e2, s2) * invDiv;
   if (t < r.tmin || t > r.tmax) return false;
   return true;
}
// geometric DG of a triangle hit (TriangleMesh.hs:183-205, mkDgTri)
HD DG triDG(V3 p1, V3 p2, V3 p3, const float *uv, V3 pHit, float b1, float b2) {
   V3 e1 = p2 - p1, e2 = p3 - p1;
   V3 n = normalize3(cross3(e1, e2));
   float du1 = uv[0] - uv[4], du2 = uv[2] - uv[4], dv1 = uv[1] - uv[5], dv2 = uv[3] - uv[5];
   V3 dp1 = p1 - p3, dp2 = p2 - p3;
   float det = du1 * dv2 - dv1 * du2;
   V3 dpdu, dpdv;
   if (det == 0.0f) { Frame f = coordinateSystem(n); dpdu = f.s; dpdv = f.t; }
   else {
      float invDet = 1.0f / det;
      dpdu = scl(invDet, scl(dv2, dp1) - scl(dv1, dp2));
      dpdv = scl(invDet, scl(-du2, dp1) + scl(du1, dp2));
   }
   float b0 = 1.0f - b1 - b2;
   float tu = b0 * uv[0] + b1 * uv[2] + b2 * uv[4];
   float tv = b0 * uv[1] + b1 * uv[3] + b2 * uv[5];
   return mkDg(pHit, tu, tv, dpdu, dpdv);
}

// ---------------------------------------------------------------------------------------- analytic shapes
// Shape.hs:81-229 `intersect` in object space. WANT_DG=false is the traversal variant (t only).
template <bool WANT_DG>
HD bool shapeIntersect(const blingcu_shape &s, const Ray &ray, float &tOut, DG &dg) {
   const float *P = s.p;
   const V3 ro = ray.o, rd = ray.d;
   const float tmin = ray.tmin, tmax = ray.tmax;
   switch (s.kind) {
   case BLINGCU_SHAPE_BOX: {   // :81-107 (a ray starting inside never hits: t0 < tmin is rejected)
      V3 pmin = mk3(P[0], P[1], P[2]), pmax = mk3(P[3], P[4], P[5]);
      float nearT = -BL_INF, farT = BL_INF; int dd = 0;
      for (int dim = 0; dim < 3; ++dim) {
         if (nearT > farT) return false;
         float oc = comp(ro, dim), dInv = 1.0f / comp(rd, dim);
         float t1p = (comp(pmax, dim) - oc) * dInv, t2p = (comp(pmin, dim) - oc) * dInv;
         float t1, t2;
         if (t1p > t2p) { t1 = t2p; t2 = t1p; } else { t1 = t1p; t2 = t2p; }
         if (nearT < t1) dd = dim;
         nearT = hmaxf(nearT, t1); farT = hminf(farT, t2);
      }
      if (nearT > farT) return false;
      float t0 = hminf(nearT, farT), t1 = hmaxf(nearT, farT);
      if (t0 > tmax || t0 < tmin) return false;
      float t = (t0 < tmin) ? t1 : t0;
      if (t > tmax) return false;
      tOut = t;
      if (WANT_DG) {
         V3 p = rayAt(ray, t);
         float half = (comp(pmin, dd) + comp(pmax, dd)) / 2;
         float dir = (comp(p, dd) > half) ? 1.0f : -1.0f;
         dg = mkDgN(p, normalize3(setc(dd, dir, mk3(0, 0, 0))));
      }
      return true;
   }
   case BLINGCU_SHAPE_CYLINDER: {   // :109-139
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
      tOut = t;
      if (WANT_DG) {
         V3 dpdu = mk3(-phimax * pHit.y, phimax * pHit.x, 0), dpdv = mk3(0, 0, zmax - zmin);
         dg = mkDgN(pHit, normalize3(cross3(dpdu, dpdv)));
      }
      return true;
   }
   case BLINGCU_SHAPE_DISK: {   // :141-155
      float h = P[0], rad = P[1], irad = P[2], phimax = P[3];
      if (fabsf(rd.z) < 1e-7f) return false;
      float t = (h - ro.z) / rd.z;
      if (t < tmin || t > tmax) return false;
      V3 p = rayAt(ray, t);
      float d2 = p.x * p.x + p.y * p.y;
      if (d2 > rad * rad || d2 < irad * irad) return false;
      if (atan2p(p.y, p.x) > phimax) return false;
      tOut = t;
      if (WANT_DG) dg = mkDgN(p, mk3(0, 0, -1));
      return true;
   }
   case BLINGCU_SHAPE_QUAD: {   // :157-172
      float sx = P[0], sy = P[1];
      if (fabsf(rd.z) < 1e-7f) return false;
      float t = -(ro.z) / rd.z;
      if (t < tmin || t > tmax) return false;
      V3 p = rayAt(ray, t);
      if (fabsf(p.x) > sx || fabsf(p.y) > sy) return false;
      tOut = t;
      if (WANT_DG) dg = mkDg(p, (sx + p.x) / (2 * sx), (sy + p.y) / (2 * sy), mk3(sx, 0, 0), mk3(0, sy, 0));
      return true;
   }
   default: {   // sphere :174-229
      float r = P[0];
      float a = sqLen(rd), b = 2 * dot3(ro, rd), c = sqLen(ro) - (r * r);
      float t1, t2;
      if (!solveQuadric(a, b, c, t1, t2)) return false;
      if (t1 > tmax) return false;
      if (t2 < tmin) return false;
      float t = (t1 < tmin) ? t2 : t1;
      if (t > tmax) return false;
      tOut = t;
      if (WANT_DG) {
         const float thetaMin = BL_PI, thetaMax = 0.0f, phiMax = BL_TWOPI;
         V3 p = rayAt(ray, t);
         float phi = atan2p(p.y, p.x);
         float u = phi / phiMax;
         float theta = acosf(clampf(p.z / r, -1.0f, 1.0f));
         float v = (theta - thetaMin) / (thetaMax - thetaMin);
         float zradius = sqrtf(p.x * p.x + p.y * p.y);
         float invz = 1.0f / zradius;
         float cosphi = p.x * invz, sinphi = p.y * invz;
         V3 dpdu = mk3(-phiMax * p.y, phiMax * p.x, 0);
         float dth = thetaMax - thetaMin;
         V3 dpdv = mk3(p.z * cosphi, p.z * sinphi, -r * sinf(theta)) * mk3(dth, dth, dth);
         dg = mkDg(p, u, v, dpdu, dpdv);
      }
      return true;
   }
   }
}

// Shape.hs:231-284 `intersects` (shadow rays). NOT the same predicate as `intersect` for boxes/spheres/cylinders.
HD bool shapeIntersects(const blingcu_shape &s, const Ray &ray) {
   const float *P = s.p;
   const V3 ro = ray.o, rd = ray.d;
   const float tmin = ray.tmin, tmax = ray.tmax;
   switch (s.kind) {
   case BLINGCU_SHAPE_BOX: {   // intersectAABB (AABB.hs:79-94)
      V3 lo = mk3(P[0], P[1], P[2]), hi = mk3(P[3], P[4], P[5]);
      float nearT = tmin, farT = tmax;
      for (int dim = 0; dim < 3; ++dim) {
         if (nearT > farT) return false;
         float oc = comp(ro, dim), dInv = 1.0f / comp(rd, dim);
         float tFar = (comp(hi, dim) - oc) * dInv, tNear = (comp(lo, dim) - oc) * dInv;
         float n2, f2;
         if (tNear > tFar) { n2 = tFar; f2 = tNear; } else { n2 = tNear; f2 = tFar; }
         nearT = hmaxf(nearT, n2); farT = hminf(farT, f2);
      }
      return !(nearT > farT);
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
      if (fabsf(rd.z) < 1e-7f) return false;
      float t = (h - ro.z) / rd.z;
      if (t < tmin || t > tmax) return false;
      V3 p = rayAt(ray, t);
      float d2 = p.x * p.x + p.y * p.y;
      if (d2 > rad * rad || d2 < irad * irad) return false;
      return !(atan2p(p.y, p.x) > phimax);
   }
   case BLINGCU_SHAPE_QUAD: {
      float sx = P[0], sy = P[1];
      if (fabsf(rd.z) < 1e-7f) return false;
      float t = -(ro.z) / rd.z;
      if (t < tmin || t > tmax) return false;
      V3 p = rayAt(ray, t);
      return !(fabsf(p.x) > sx || fabsf(p.y) > sy);
   }
   default: {
      float rad = P[0];
      float a = sqLen(rd), b = 2 * dot3(ro, rd), c = sqLen(ro) - (rad * rad);
      float t0, t1;
      if (!solveQuadric(a, b, c, t0, t1)) return false;
      if (t0 > tmax || t1 < tmin) return false;
      if (t0 < tmin) return t1 < tmax;
      return true;
   }
   }
}

HD float shapeArea(const blingcu_shape &s) {   // Shape.hs:314-328
   const float *P = s.p;
   switch (s.kind) {
   case BLINGCU_SHAPE_BOX: { float h = P[3] - P[0], w = P[4] - P[1], l = P[5] - P[2]; return 2 * (h * w + h * l + w * l); }
   case BLINGCU_SHAPE_CYLINDER: return 2 * BL_PI * P[0] * (P[2] - P[1]);
   case BLINGCU_SHAPE_DISK: return BL_PI * (P[1] * P[1] - P[2] * P[2]);
   case BLINGCU_SHAPE_QUAD: return 4 * P[0] * P[1];
   default: return P[0] * P[0] * 4 * BL_PI;
   }
}

}  // namespace bl
