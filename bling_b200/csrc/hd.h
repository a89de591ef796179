// hd.h -- scalar/vector/spectrum math shared by every kernel body.
//
// All bodies are `HD` (host+device inline) so that the CUDA backend (the product, cuda_backend.cu)
// and the CPU kernel-body emulator used by the no-GPU CI tests (tests/emu/, never linked into
// libblingcu.so) drive the SAME per-item logic. The product library has no CPU execution path.
//
// Arithmetic contract: f32 everywhere (Types.hs:3), NO fused multiply-add in geometry and shading
// (compiled with -fmad=false / -ffp-contract=off) so that primitive hits are bit-identical to the
// reference's unfused arithmetic. Traversal box tests are the only place that may round differently
// (they are conservative, see bvh.h).
#pragma once
#include <stdint.h>
#include <math.h>
#include <float.h>

#ifdef __CUDACC__
#define HD __host__ __device__ __forceinline__
#define HDNI inline __host__ __device__ __noinline__
#else
#define HD inline
#define HDNI inline
#endif

namespace bl {

#define BL_INF (__builtin_huge_valf())
#define BL_PI 3.14159265358979323846f
#define BL_TWOPI (2.0f * BL_PI)
#define BL_INVPI (1.0f / BL_PI)
#define BL_INVTWOPI (1.0f / (2.0f * BL_PI))
#define NB 16

// Haskell Ord Float semantics (Q12): max x y = if x <= y then y else x
HD float hmaxf(float x, float y) { return (x <= y) ? y : x; }
HD float hminf(float x, float y) { return (x <= y) ? x : y; }
HD float clampf(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }   // Math.hs:80-89
HD float lerpf(float t, float a, float b) { return (1.0f - t) * a + t * b; }                // Math.hs:116-118
HD int imin(int a, int b) { return a < b ? a : b; }
HD int imax(int a, int b) { return a > b ? a : b; }

struct V3 { float x, y, z; };
HD V3 mk3(float x, float y, float z) { V3 v; v.x = x; v.y = y; v.z = z; return v; }
HD float comp(V3 v, int d) { return d == 0 ? v.x : (d == 1 ? v.y : v.z); }
HD V3 setc(int d, float t, V3 v) { if (d == 0) v.x = t; else if (d == 1) v.y = t; else v.z = t; return v; }
HD V3 operator+(V3 a, V3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
HD V3 operator-(V3 a, V3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
HD V3 operator*(V3 a, V3 b) { return mk3(a.x * b.x, a.y * b.y, a.z * b.z); }
HD V3 operator-(V3 a) { return mk3(-a.x, -a.y, -a.z); }
HD V3 scl(float f, V3 v) { return mk3(f * v.x, f * v.y, f * v.z); }                         // (*#), Math.hs:226-228
HD float sqLen(V3 v) { return v.x * v.x + v.y * v.y + v.z * v.z; }
HD float len3(V3 v) { return sqrtf(sqLen(v)); }
HD float dot3(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
HD float absDot(V3 a, V3 b) { return fabsf(dot3(a, b)); }
HD V3 cross3(V3 u, V3 w) { return mk3(u.y * w.z - u.z * w.y, -(u.x * w.z - u.z * w.x), u.x * w.y - u.y * w.x); }  // Math.hs:345-348
HD V3 normalize3(V3 v) {                                                                    // Math.hs:358-362
   if (sqLen(v) != 0.0f) { float il = 1.0f / len3(v); return mk3(v.x * il, v.y * il, v.z * il); }
   return mk3(0, 1, 0);
}

struct Ray { V3 o; float tmin; V3 d; float tmax; };
HD V3 rayAt(const Ray &r, float t) { return r.o + scl(t, r.d); }

struct Frame { V3 s, t, n; };
HD Frame coordinateSystem(V3 v) {                                                           // Math.hs:424-437
   Frame f;
   if (fabsf(v.x) > fabsf(v.y)) {
      float il = 1.0f / sqrtf(v.x * v.x + v.z * v.z);
      f.s = mk3(-v.z * il, 0, v.x * il);
   } else {
      float il = 1.0f / sqrtf(v.y * v.y + v.z * v.z);
      f.s = mk3(0, v.z * il, -v.y * il);
   }
   f.t = cross3(v, f.s); f.n = v;
   return f;
}
HD V3 worldToLocal(const Frame &f, V3 v) { return mk3(dot3(v, f.s), dot3(v, f.t), dot3(v, f.n)); }
HD V3 localToWorld(const Frame &f, V3 v) {
   return mk3(f.s.x * v.x + f.t.x * v.y + f.n.x * v.z, f.s.y * v.x + f.t.y * v.y + f.n.y * v.z,
              f.s.z * v.x + f.t.z * v.y + f.n.z * v.z);
}
HD float atan2p(float y, float x) { float a = atan2f(y, x); return a < 0 ? a + BL_TWOPI : a; }  // Math.hs:66-75
HD bool solveQuadric(float a, float b, float c, float &t0, float &t1) {                     // Math.hs:126-139
   float discrim = b * b - 4 * a * c;
   if (discrim < 0) return false;
   float root = sqrtf(discrim);
   float q = (b < 0) ? -0.5f * (b - root) : -0.5f * (b + root);
   float x0 = q / a, x1 = c / q;
   t0 = hminf(x0, x1); t1 = hmaxf(x0, x1);
   return true;
}
HD V3 sphericalDirection(float sint, float cost, float phi) { return mk3(sint * cosf(phi), sint * sinf(phi), cost); }
HD float sphericalTheta(V3 v) { return acosf(hmaxf(-1.0f, hminf(1.0f, v.z))); }
HD float sphericalPhi(V3 v) { float p = atan2f(v.y, v.x); return p < 0 ? p + 2 * BL_PI : p; }

// Transform.hs:246-278 on row-major 4x4
HD V3 transPoint(const float *m, V3 p) {
   float xp = m[0] * p.x + m[1] * p.y + m[2] * p.z + m[3];
   float yp = m[4] * p.x + m[5] * p.y + m[6] * p.z + m[7];
   float zp = m[8] * p.x + m[9] * p.y + m[10] * p.z + m[11];
   float wp = m[12] * p.x + m[13] * p.y + m[14] * p.z + m[15];
   if (wp == 1.0f) return mk3(xp, yp, zp);
   return mk3(xp / wp, yp / wp, zp / wp);
}
HD V3 transVector(const float *m, V3 v) {
   return mk3(m[0] * v.x + m[1] * v.y + m[2] * v.z, m[4] * v.x + m[5] * v.y + m[6] * v.z, m[8] * v.x + m[9] * v.y + m[10] * v.z);
}
HD V3 transNormalInv(const float *mi, V3 n) {   // transpose of the inverse (Transform.hs:267-272)
   return mk3(mi[0] * n.x + mi[4] * n.y + mi[8] * n.z, mi[1] * n.x + mi[5] * n.y + mi[9] * n.z, mi[2] * n.x + mi[6] * n.y + mi[10] * n.z);
}
HD Ray transRay(const float *m, const Ray &r) { Ray o; o.o = transPoint(m, r.o); o.d = transVector(m, r.d); o.tmin = r.tmin; o.tmax = r.tmax; return o; }

// ------------------------------------------------------------------------------- 16-band spectrum
struct Spec { float v[NB]; };
#ifdef __CUDACC__
#define BL_UNROLL _Pragma("unroll")
#else
#define BL_UNROLL
#endif
HD Spec sConst(float c) { Spec s; BL_UNROLL for (int i = 0; i < NB; ++i) s.v[i] = c; return s; }
HD Spec operator+(const Spec &a, const Spec &b) { Spec s; BL_UNROLL for (int i = 0; i < NB; ++i) s.v[i] = a.v[i] + b.v[i]; return s; }
HD Spec operator-(const Spec &a, const Spec &b) { Spec s; BL_UNROLL for (int i = 0; i < NB; ++i) s.v[i] = a.v[i] - b.v[i]; return s; }
HD Spec operator*(const Spec &a, const Spec &b) { Spec s; BL_UNROLL for (int i = 0; i < NB; ++i) s.v[i] = a.v[i] * b.v[i]; return s; }
HD Spec operator/(const Spec &a, const Spec &b) { Spec s; BL_UNROLL for (int i = 0; i < NB; ++i) s.v[i] = a.v[i] / b.v[i]; return s; }
HD Spec sScale(const Spec &a, float f) { Spec s; BL_UNROLL for (int i = 0; i < NB; ++i) s.v[i] = a.v[i] * f; return s; }
HD bool isBlack(const Spec &a) { bool b = true; BL_UNROLL for (int i = 0; i < NB; ++i) b = b && (a.v[i] == 0.0f); return b; }
HD bool sBad(const Spec &a) { bool b = false; BL_UNROLL for (int i = 0; i < NB; ++i) b = b || isnan(a.v[i]) || isinf(a.v[i]); return b; }
HD Spec sClamp01(const Spec &a) { Spec s; BL_UNROLL for (int i = 0; i < NB; ++i) s.v[i] = hmaxf(0.0f, hminf(1.0f, a.v[i])); return s; }
HD Spec loadSpec(const float *p) { Spec s; BL_UNROLL for (int i = 0; i < NB; ++i) s.v[i] = p[i]; return s; }

// ------------------------------------------------------------------------------- sampler SPEC (DESIGN.md "Sampler")
// Counter-based replacement for mwc-random (Random.hs:92-96): a pixel key from (seed, pass, pixel), a
// 32-bit key per dimension, Kensler's permutation for the per-dimension shuffle of the stratified sets
// (Sampling.hs:112-171). Written independently of oracle/oracle_math.h from the same SPEC.
HD uint64_t mix64(uint64_t x) { x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ULL; x ^= x >> 27; x *= 0x94d049bb133111ebULL; x ^= x >> 31; return x; }
HD uint32_t hash32(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }
HD uint64_t pixelKey(uint64_t seed, uint32_t pass, uint32_t pix) {
   return mix64(mix64(seed ^ ((uint64_t)(pass + 1u) * 0x9E3779B97F4A7C15ULL)) + (uint64_t)pix * 0xD1B54A32D192ED03ULL);
}
HD uint32_t dimKey(uint64_t kp, uint32_t dim) { return hash32(hash32((uint32_t)kp ^ (dim * 0x9E3779B9U)) + (uint32_t)(kp >> 32)); }
HD float u01(uint32_t h) { return (float)(h >> 8) * (1.0f / 16777216.0f); }
HD uint32_t sampleHash(uint32_t kd, uint32_t s) { return hash32(kd + s * 0x9E3779B9U + 0x7F4A7C15U); }
// Kensler's permutation of [0, l). `w` = (next power of two >= l) - 1 is a per-scene constant, and so is `pow2`
// (l is a power of two: every config's strata count is); both only shorten the arithmetic, the value is the SPEC's.
HD uint32_t permuteW(uint32_t i, uint32_t l, uint32_t p, uint32_t w, bool pow2) {
   if (l <= 1) return 0;
   do {
      i ^= p; i *= 0xe170893dU; i ^= p >> 16; i ^= (i & w) >> 4; i ^= p >> 8; i *= 0x0929eb3fU; i ^= p >> 23;
      i ^= (i & w) >> 1; i *= 1 | p >> 27; i *= 0x6935fa69U; i ^= (i & w) >> 11; i *= 0x74dcb303U;
      i ^= (i & w) >> 2; i *= 0x9e501cc3U; i ^= (i & w) >> 2; i *= 0xc860a3dfU; i &= w; i ^= i >> 5;
   } while (i >= l);
   return pow2 ? ((i + p) & w) : ((i + p) % l);
}
HD uint32_t smear(uint32_t l) { uint32_t w = l - 1; w |= w >> 1; w |= w >> 2; w |= w >> 4; w |= w >> 8; w |= w >> 16; return w; }
HD uint32_t permute(uint32_t i, uint32_t l, uint32_t p) { return l <= 1 ? 0u : permuteW(i, l, p, smear(l), (l & (l - 1)) == 0); }
#define BL_ALMOST_ONE 0.9999999403953552f
enum { DIM_IMAGE = 0, DIM_LENS = 1, DIM_1D_BASE = 16, DIM_2D_BASE = 4096 };

// per-scene constants of the stratified sampler, computed once on the host (pipeline.h) with the same f32 operations
// the per-sample code used to repeat: 1/nu, 1/nv, 1/(nu nv) are IEEE divisions, so the bits are identical
struct SamplerConst {
   int nu, nv, n1d, n2d, stratified;
   uint32_t N, wmask;      // strata per pixel, smear(N)
   int pow2N;              // N is a power of two
   int nuShift;            // log2(nu) when nu is a power of two, else -1
   float du, dv, invN;
};
// n1d / n2d = the integrator's sampleCount1D / sampleCount2D (Path.hs:18-36: 4 sd, 3 sd; DirectLighting.hs:18-19: 2 md, 2 md)
HD SamplerConst mkSamplerConst(int nu, int nv, int n1d, int n2d, int stratified) {
   SamplerConst k; k.nu = nu; k.nv = nv; k.n1d = n1d; k.n2d = n2d; k.stratified = stratified;
   k.N = (uint32_t)(nu * nv); k.wmask = k.N > 1 ? smear(k.N) : 0u; k.pow2N = (k.N & (k.N - 1)) == 0;
   k.nuShift = -1; for (int b = 0; b < 31; ++b) if ((1u << b) == (uint32_t)nu) k.nuShift = b;
   k.du = 1.0f / (float)nu; k.dv = 1.0f / (float)nv; k.invN = 1.0f / (float)k.N;
   return k;
}
struct Sampler {   // per-sample view of the pixel's stratified sets
   uint64_t kp; uint32_t s; const SamplerConst *k;
};
HD void strat2D(uint32_t i, const SamplerConst &k, float ju, float jv, float &u, float &v) {   // Sampling.hs:163-171 (Q9)
   uint32_t q, r;
   if (k.nuShift >= 0) { q = i >> k.nuShift; r = i & ((uint32_t)k.nu - 1u); } else { q = i / (uint32_t)k.nu; r = i % (uint32_t)k.nu; }
   u = hminf(BL_ALMOST_ONE, ((float)q + ju) * k.du);
   v = hminf(BL_ALMOST_ONE, ((float)r + jv) * k.dv);
}
HD float rnd1D(const Sampler &c, int n) {                                               // rnd' Sampling.hs:203-211
   const SamplerConst &k = *c.k;
   uint32_t kd = dimKey(c.kp, DIM_1D_BASE + (uint32_t)n);
   uint32_t h = sampleHash(kd, c.s);
   if (!k.stratified || n >= k.n1d) return u01(h);
   uint32_t i = permuteW(c.s, k.N, kd, k.wmask, k.pow2N != 0);
   return hminf(BL_ALMOST_ONE, ((float)i + u01(h)) * k.invN);
}
HD void rnd2D(const Sampler &c, int n, float &u, float &v) {                            // rnd2D' Sampling.hs:213-221
   const SamplerConst &k = *c.k;
   uint32_t kd = dimKey(c.kp, DIM_2D_BASE + (uint32_t)n);
   uint32_t h = sampleHash(kd, c.s), h2 = hash32(h ^ 0x85ebca6bU);
   if (!k.stratified || n >= k.n2d) { u = u01(h); v = u01(h2); return; }
   strat2D(permuteW(c.s, k.N, kd, k.wmask, k.pow2N != 0), k, u01(h), u01(h2), u, v);
}
HD void cameraSample(const Sampler &c, float &ox, float &oy, float &lu, float &lv) {    // Sampling.hs:112-132
   const SamplerConst &k = *c.k;
   uint32_t ki = dimKey(c.kp, DIM_IMAGE), kl = dimKey(c.kp, DIM_LENS);
   uint32_t h = sampleHash(ki, c.s), h2 = hash32(h ^ 0x85ebca6bU);
   uint32_t g = sampleHash(kl, c.s), g2 = hash32(g ^ 0x85ebca6bU);
   if (!k.stratified) { ox = u01(h); oy = u01(h2); lu = u01(g); lv = u01(g2); return; }
   strat2D(c.s, k, u01(h), u01(h2), ox, oy);
   strat2D(permuteW(c.s, k.N, kl, k.wmask, k.pow2N != 0), k, u01(g), u01(g2), lu, lv);
}

}  // namespace bl
