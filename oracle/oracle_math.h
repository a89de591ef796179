// ORACLE (test infrastructure, NOT product code): CPU restatement of waldheinz/bling.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may use it.
// PARITY UNPINNED: the reference ships no golden vectors for this path and GHC is
// absent, so this restatement is validated only by analytic self-checks (tests/test_oracle_*.py).
//
// Math, spectra, RNG. Every function cites the reference file:line it follows
// (paths relative to /root/reference/src/lib/Graphics/Bling).
// Build: g++ -O2 -ffp-contract=off (no FMA contraction; GHC does not fuse either).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <algorithm>
#include <vector>

namespace orc {

static const float kInf = std::numeric_limits<float>::infinity();
static const float kPi = 3.14159265358979323846f;   // Haskell `pi :: Float`
static const float kTwoPi = 2.0f * kPi;              // Math.hs:62-64
static const float kInvPi = 1.0f / kPi;              // Math.hs:54-56
static const float kInvTwoPi = 1.0f / (2.0f * kPi);  // Math.hs:58-60

// Haskell Ord Float: max x y = if x <= y then y else x ; min x y = if x <= y then x else y (Q12)
static inline float hmax(float x, float y) { return (x <= y) ? y : x; }
static inline float hmin(float x, float y) { return (x <= y) ? x : y; }
// Math.hs:80-89
static inline float clampf(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }
// Math.hs:116-118
static inline float lerpf(float t, float a, float b) { return (1.0f - t) * a + t * b; }

struct V3 {
   float x, y, z;
   float operator[](int d) const { return d == 0 ? x : (d == 1 ? y : z); }  // Math.hs:319-324
};
static inline V3 mk(float x, float y, float z) { return V3{x, y, z}; }
static inline V3 operator+(V3 a, V3 b) { return V3{a.x + b.x, a.y + b.y, a.z + b.z}; }
static inline V3 operator-(V3 a, V3 b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
static inline V3 operator*(V3 a, V3 b) { return V3{a.x * b.x, a.y * b.y, a.z * b.z}; }
static inline V3 operator-(V3 a) { return V3{-a.x, -a.y, -a.z}; }
// Math.hs:226-228  f *# v = vpromote f * v
static inline V3 scl(float f, V3 v) { return V3{f * v.x, f * v.y, f * v.z}; }
static inline V3 setc(int d, float t, V3 v) {  // Math.hs:330-335
   if (d == 0) return V3{t, v.y, v.z};
   if (d == 1) return V3{v.x, t, v.z};
   return V3{v.x, v.y, t};
}
static inline float sqLen(V3 v) { return v.x * v.x + v.y * v.y + v.z * v.z; }          // Math.hs:337-339
static inline float len(V3 v) { return std::sqrt(sqLen(v)); }                           // Math.hs:341-343
static inline float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }       // Math.hs:350-352
static inline float absDot(V3 a, V3 b) { return std::fabs(dot(a, b)); }
static inline V3 cross(V3 u, V3 w) {  // Math.hs:345-348
   return V3{u.y * w.z - u.z * w.y, -(u.x * w.z - u.z * w.x), u.x * w.y - u.y * w.x};
}
static inline V3 normalize(V3 v) {  // Math.hs:358-362
   if (sqLen(v) != 0.0f) { float il = 1.0f / len(v); return V3{v.x * il, v.y * il, v.z * il}; }
   return V3{0, 1, 0};
}

struct Ray { V3 o, d; float tmin, tmax; };
static inline V3 rayAt(const Ray &r, float t) { return r.o + scl(t, r.d); }  // Math.hs:400-402 (d * vpromote t)

struct Frame { V3 s, t, n; };  // LocalCoordinates sn tn nn
// Math.hs:424-437
static inline Frame coordinateSystem(V3 v) {
   if (std::fabs(v.x) > std::fabs(v.y)) {
      float il = 1.0f / std::sqrt(v.x * v.x + v.z * v.z);
      V3 v2 = V3{-v.z * il, 0, v.x * il};
      return Frame{v2, cross(v, v2), v};
   }
   float il = 1.0f / std::sqrt(v.y * v.y + v.z * v.z);
   V3 v2 = V3{0, v.z * il, -v.y * il};
   return Frame{v2, cross(v, v2), v};
}
static inline V3 worldToLocal(const Frame &f, V3 v) { return V3{dot(v, f.s), dot(v, f.t), dot(v, f.n)}; }  // Math.hs:452-454
static inline V3 localToWorld(const Frame &f, V3 v) {  // Math.hs:456-462
   return V3{f.s.x * v.x + f.t.x * v.y + f.n.x * v.z, f.s.y * v.x + f.t.y * v.y + f.n.y * v.z,
             f.s.z * v.x + f.t.z * v.y + f.n.z * v.z};
}
// Math.hs:66-75
static inline float atan2p(float y, float x) { float a = std::atan2(y, x); return a < 0 ? a + kTwoPi : a; }
// Math.hs:126-139
static inline bool solveQuadric(float a, float b, float c, float &t0, float &t1) {
   float discrim = b * b - 4 * a * c;
   if (discrim < 0) return false;
   float root = std::sqrt(discrim);
   float q = (b < 0) ? -0.5f * (b - root) : -0.5f * (b + root);
   float x0 = q / a, x1 = c / q;
   t0 = hmin(x0, x1); t1 = hmax(x0, x1);
   return true;
}
static inline V3 sphericalDirection(float sint, float cost, float phi) {  // Math.hs:141-143
   return V3{sint * std::cos(phi), sint * std::sin(phi), cost};
}
static inline float sphericalTheta(V3 v) { return std::acos(hmax(-1.0f, hmin(1.0f, v.z))); }  // Math.hs:160-162
static inline float sphericalPhi(V3 v) { float p = std::atan2(v.y, v.x); return p < 0 ? p + 2 * kPi : p; }  // Math.hs:164-170

// Transform.hs:246-272 on a row-major 4x4 (mi m r c = m[r*4+c])
static inline V3 transPoint(const float *m, V3 p) {
   float xp = m[0] * p.x + m[1] * p.y + m[2] * p.z + m[3];
   float yp = m[4] * p.x + m[5] * p.y + m[6] * p.z + m[7];
   float zp = m[8] * p.x + m[9] * p.y + m[10] * p.z + m[11];
   float wp = m[12] * p.x + m[13] * p.y + m[14] * p.z + m[15];
   if (wp == 1.0f) return V3{xp, yp, zp};
   return V3{xp / wp, yp / wp, zp / wp};
}
static inline V3 transVector(const float *m, V3 v) {
   return V3{m[0] * v.x + m[1] * v.y + m[2] * v.z, m[4] * v.x + m[5] * v.y + m[6] * v.z,
             m[8] * v.x + m[9] * v.y + m[10] * v.z};
}
// transNormal uses the transpose of the INVERSE matrix: pass the inverse here (Transform.hs:267-272)
static inline V3 transNormalInv(const float *mi, V3 n) {
   return V3{mi[0] * n.x + mi[4] * n.y + mi[8] * n.z, mi[1] * n.x + mi[5] * n.y + mi[9] * n.z,
             mi[2] * n.x + mi[6] * n.y + mi[10] * n.z};
}
static inline Ray transRay(const float *m, const Ray &r) {  // Transform.hs:275-278
   return Ray{transPoint(m, r.o), transVector(m, r.d), r.tmin, r.tmax};
}

struct AABB { V3 lo, hi; };
static inline AABB emptyBox() { return AABB{V3{kInf, kInf, kInf}, V3{-kInf, -kInf, -kInf}}; }  // AABB.hs:26-30
static inline AABB extendP(AABB b, V3 p) {  // AABB.hs:46-49
   return AABB{V3{hmin(b.lo.x, p.x), hmin(b.lo.y, p.y), hmin(b.lo.z, p.z)},
               V3{hmax(b.hi.x, p.x), hmax(b.hi.y, p.y), hmax(b.hi.z, p.z)}};
}
static inline AABB extendB(AABB a, AABB b) {  // AABB.hs:36-44
   return AABB{V3{hmin(a.lo.x, b.lo.x), hmin(a.lo.y, b.lo.y), hmin(a.lo.z, b.lo.z)},
               V3{hmax(a.hi.x, b.hi.x), hmax(a.hi.y, b.hi.y), hmax(a.hi.z, b.hi.z)}};
}
static inline float surfaceArea(const AABB &b) {  // AABB.hs:72-75
   V3 d = b.hi - b.lo;
   return 2 * (d.x * d.y + d.x * d.z + d.y * d.z);
}
static inline int dominant(V3 v) {  // Math.hs:296-305
   float ax = std::fabs(v.x), ay = std::fabs(v.y), az = std::fabs(v.z);
   if (ax > ay && ax > az) return 0;
   if (ay > az) return 1;
   return 2;
}
// AABB.hs:79-94
static inline bool intersectAABB(const AABB &b, const Ray &r, float &tn, float &tf) {
   float nearT = r.tmin, farT = r.tmax;
   for (int dim = 0; dim < 3; ++dim) {
      if (nearT > farT) return false;
      float oc = r.o[dim];
      float dInv = 1.0f / r.d[dim];
      float tFar = (b.hi[dim] - oc) * dInv;
      float tNear = (b.lo[dim] - oc) * dInv;
      float n2, f2;
      if (tNear > tFar) { n2 = tFar; f2 = tNear; } else { n2 = tNear; f2 = tFar; }
      nearT = hmax(nearT, n2);
      farT = hmin(farT, f2);
   }
   if (nearT > farT) return false;
   tn = nearT; tf = farT;
   return true;
}
static inline AABB transBox(const float *m, const AABB &b) {  // Transform.hs:281-292
   AABB r = emptyBox();
   float xs[2] = {b.lo.x, b.hi.x}, ys[2] = {b.lo.y, b.hi.y}, zs[2] = {b.lo.z, b.hi.z};
   for (int i = 0; i < 2; ++i) for (int j = 0; j < 2; ++j) for (int k = 0; k < 2; ++k)
      r = extendP(r, transPoint(m, V3{xs[i], ys[j], zs[k]}));
   return r;
}

// ---------------------------------------------------------------------------------------------
// Spectrum.hs:421-469: 16-band element-wise arithmetic
// ---------------------------------------------------------------------------------------------
static const int NB = 16;
struct Spec {
   float v[NB];
};
static inline Spec sConst(float c) { Spec s; for (int i = 0; i < NB; ++i) s.v[i] = c; return s; }
static inline Spec operator+(const Spec &a, const Spec &b) { Spec s; for (int i = 0; i < NB; ++i) s.v[i] = a.v[i] + b.v[i]; return s; }
static inline Spec operator-(const Spec &a, const Spec &b) { Spec s; for (int i = 0; i < NB; ++i) s.v[i] = a.v[i] - b.v[i]; return s; }
static inline Spec operator*(const Spec &a, const Spec &b) { Spec s; for (int i = 0; i < NB; ++i) s.v[i] = a.v[i] * b.v[i]; return s; }
static inline Spec operator/(const Spec &a, const Spec &b) { Spec s; for (int i = 0; i < NB; ++i) s.v[i] = a.v[i] / b.v[i]; return s; }
static inline Spec sScale(const Spec &a, float f) { Spec s; for (int i = 0; i < NB; ++i) s.v[i] = a.v[i] * f; return s; }  // Spectrum.hs:447-449
static inline bool isBlack(const Spec &a) { for (int i = 0; i < NB; ++i) if (!(a.v[i] == 0.0f)) return false; return true; }  // :443-445
static inline bool sNaN(const Spec &a) { for (int i = 0; i < NB; ++i) if (std::isnan(a.v[i])) return true; return false; }
static inline bool sInfinite(const Spec &a) { for (int i = 0; i < NB; ++i) if (std::isinf(a.v[i])) return true; return false; }
static inline Spec sClamp(float lo, float hi, const Spec &a) {  // Spectrum.hs:452-455  max smin $ min smax x
   Spec s; for (int i = 0; i < NB; ++i) s.v[i] = hmax(lo, hmin(hi, a.v[i])); return s;
}

// ---------------------------------------------------------------------------------------------
// RNG. The reference draws from mwc-random seeded from system entropy per tile per pass
// (Rendering.hs:128, Random.hs:61-62), so sequences cannot be matched. The oracle and
// the GPU share this counter-based SPEC instead (DESIGN.md "Sampler"), implemented independently
// on each side, which makes per-sample radiance comparable between the two.
// ---------------------------------------------------------------------------------------------
static inline uint64_t mix64(uint64_t x) {  // splitmix64 finaliser
   x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ULL; x ^= x >> 27; x *= 0x94d049bb133111ebULL; x ^= x >> 31;
   return x;
}
static inline uint32_t hash32(uint32_t x) {  // "lowbias32"
   x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
   return x;
}
static inline uint64_t pixelKey(uint64_t seed, uint32_t pass, uint32_t pix) {
   return mix64(mix64(seed ^ ((uint64_t)(pass + 1u) * 0x9E3779B97F4A7C15ULL)) + (uint64_t)pix * 0xD1B54A32D192ED03ULL);
}
static inline uint32_t dimKey(uint64_t kp, uint32_t dim) {
   uint32_t k = hash32((uint32_t)kp ^ (dim * 0x9E3779B9U)) + (uint32_t)(kp >> 32);
   return hash32(k);
}
static inline float u01(uint32_t h) { return (float)(h >> 8) * (1.0f / 16777216.0f); }
static inline uint32_t sampleHash(uint32_t kd, uint32_t s) { return hash32(kd + s * 0x9E3779B9U + 0x7F4A7C15U); }
// Kensler, "Correlated Multi-Jittered Sampling" (2013): stateless permutation of [0,l)
static inline uint32_t permute(uint32_t i, uint32_t l, uint32_t p) {
   if (l <= 1) return 0;
   uint32_t w = l - 1;
   w |= w >> 1; w |= w >> 2; w |= w >> 4; w |= w >> 8; w |= w >> 16;
   do {
      i ^= p; i *= 0xe170893dU; i ^= p >> 16; i ^= (i & w) >> 4; i ^= p >> 8; i *= 0x0929eb3fU; i ^= p >> 23;
      i ^= (i & w) >> 1; i *= 1 | p >> 27; i *= 0x6935fa69U; i ^= (i & w) >> 11; i *= 0x74dcb303U;
      i ^= (i & w) >> 2; i *= 0x9e501cc3U; i ^= (i & w) >> 2; i *= 0xc860a3dfU; i &= w; i ^= i >> 5;
   } while (i >= l);
   return (i + p) % l;
}
static const float kAlmostOne = 0.9999999403953552f;  // Sampling.hs:154-155

// dimension ids of the SPEC
enum { DIM_IMAGE = 0, DIM_LENS = 1, DIM_1D_BASE = 16, DIM_2D_BASE = 4096 };

struct SampleCtx {  // one camera sample of one pixel
   uint64_t kp;
   uint32_t s;      // sample index within the pixel, 0..nu*nv-1
   int nu, nv;
   int n1d, n2d;    // precomputed (stratified) dimension counts: 4*sd, 3*sd (Path.hs:18-36)
   bool stratified; // false: `Random` sampler, everything is a plain uniform (Sampling.hs:101-110)
};
// Sampling.hs:157-160 stratified1D + the shuffle of `fill` (:134-152)
static inline float rnd1D(const SampleCtx &c, int n) {  // Sampling.hs:203-211 rnd'
   uint32_t kd = dimKey(c.kp, DIM_1D_BASE + (uint32_t)n);
   uint32_t h = sampleHash(kd, c.s);
   if (!c.stratified || n >= c.n1d) return u01(h);
   uint32_t N = (uint32_t)(c.nu * c.nv);
   uint32_t i = permute(c.s, N, kd);
   float du = 1.0f / (float)N;
   return hmin(kAlmostOne, ((float)i + u01(h)) * du);
}
// Sampling.hs:163-171 stratified2D (Q9: quotRem i nu for both axes)
static inline void strat2D(uint32_t i, int nu, int nv, float ju, float jv, float &u, float &v) {
   float du = 1.0f / (float)nu, dv = 1.0f / (float)nv;
   uint32_t q = i / (uint32_t)nu, r = i % (uint32_t)nu;
   u = hmin(kAlmostOne, ((float)q + ju) * du);
   v = hmin(kAlmostOne, ((float)r + jv) * dv);
}
static inline void rnd2D(const SampleCtx &c, int n, float &u, float &v) {  // Sampling.hs:213-221 rnd2D'
   uint32_t kd = dimKey(c.kp, DIM_2D_BASE + (uint32_t)n);
   uint32_t h = sampleHash(kd, c.s);
   uint32_t h2 = hash32(h ^ 0x85ebca6bU);
   if (!c.stratified || n >= c.n2d) { u = u01(h); v = u01(h2); return; }
   uint32_t i = permute(c.s, (uint32_t)(c.nu * c.nv), kd);
   strat2D(i, c.nu, c.nv, u01(h), u01(h2), u, v);
}
// camera sample: image offsets are NOT shuffled (Sampling.hs:117,128), lens samples are (:118-121)
static inline void cameraSample(const SampleCtx &c, float &ox, float &oy, float &lu, float &lv) {
   uint32_t ki = dimKey(c.kp, DIM_IMAGE), kl = dimKey(c.kp, DIM_LENS);
   uint32_t h = sampleHash(ki, c.s), h2 = hash32(h ^ 0x85ebca6bU);
   uint32_t g = sampleHash(kl, c.s), g2 = hash32(g ^ 0x85ebca6bU);
   if (!c.stratified) { ox = u01(h); oy = u01(h2); lu = u01(g); lv = u01(g2); return; }
   strat2D(c.s, c.nu, c.nv, u01(h), u01(h2), ox, oy);
   strat2D(permute(c.s, (uint32_t)(c.nu * c.nv), kl), c.nu, c.nv, u01(g), u01(g2), lu, lv);
}

}  // namespace orc
