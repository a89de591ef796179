// bodies.h -- wavefront path state and the per-item kernel bodies (functors).
// Replaces: a2 runSample (Sampling.hs:112-132), a3 fireRay (Camera.hs:49-76),
// a10 nextVertex (Integrator/Path.hs:41-87), a11 sampleOneLight/estimateDirect (Scene.hs:61-118),
// a17 addSample/addTile (Image.hs:178-299).
//
// One path = one camera sample. State is SoA over `cap` slots; slot = sl * NPIX + pix where pix is the linear
// pixel index in the sample extent and sl the sample's position in the batch [s0, s0+k). Queues hold slots.
#pragma once
#include "shading.h"

namespace bl {

enum { C_ACTIVE = 0, C_NEXT = 1, C_SHADOW = 2, C_MIS = 3, C_MISANY = 4, C_DROPPED = 5, C_MISCULL = 6, C_EXTCULL = 7, C_MAT0 = 8, C_SPAWN = 8 + 32, C_OVERFLOW, N_COUNTERS };   // C_MAT0 + kind: 0 = miss, 1.. = 1 + shade kind
// Shade queues ("slots"; the slot of a primitive's material, minus one, is the 5-bit field of its hit reference, so
// classification never touches geometry or materials): 0 = miss, 1 + kind for the nine material kinds, and from N_PLAIN_SLOTS on
// ONE SLOT PER MATERIAL whose textures compute (textures.h), each launched with its kind's textured kernel: the texture code of
// different materials in one launch starves instruction fetch (profiles/r01_general_shade.md).
enum { N_PLAIN_SLOTS = 1 + BLINGCU_MAT_KINDS, MAX_TEX_SLOTS = 32 - N_PLAIN_SLOTS, N_SHADE_KINDS = 32 };
enum { S_SAMPLES = 0, S_CAM, S_EXT, S_MIS, S_SHADOW, S_DROPPED, S_MISCULL, S_MISANY, S_EXTCULL, N_STATS = 12 };

#define BL_MIS_NONFINITE 0x40000000   // miInfo.y: the BSDF-MIS weight is not finite (directAtVertex, ResolveMisBodyT)
struct PathState {
   uint32_t cap;
   F4 *rayO, *rayD;        // extension ray: (o, tmin) (d, tmax)
   F4 *hit;                // (t, b1, b2, prim bits)
   F4 *T, *L;              // throughput / radiance, quarter q of slot i at [q*cap + i]
   F4 *shO, *shD, *PS;     // NEE shadow ray + pending contribution (already x T x nLights)
   uint8_t *occl, *occlM;  // shadow-ray / any-hit MIS-ray occlusion flags
   F4 *miO, *miD, *mihit, *PM;   // BSDF-MIS ray, its hit, pending T x f x nLights
   F2 *miInfo;             // (bsdf pdf, light index bits | BL_MIS_NONFINITE)
   uint32_t *meta;         // depth | spec << 8
   uint64_t *kp;           // pixel key of the sampler
   uint32_t *sidx;         // sample index within the pixel
   F2 *spos;               // image position of the sample
   F4 *xyz;                // finalised sample: X, Y, Z, valid
   uint32_t *qA, *qB;      // active queues (ping-pong)
   uint32_t *qShadow, *qMis, *qMisAny;   // qMis: nearest-hit MIS rays (area lights); qMisAny: any-hit MIS rays (infinite lights)
   uint32_t *qMat;         // (slots in use) * cap
   uint32_t *root;         // direct-lighting integrator: slot of the camera sample a spawned branch belongs to
   uint32_t *counters;     // N_COUNTERS
   unsigned long long *stats;   // N_STATS
};

// where quarter q (4 bands) of slot i's spectrum lives. Product layout since round 2: ONE 64-BYTE RECORD per slot and spectrum
// (two full 32-byte sectors whatever the queue looks like). Round 1 kept four float4 planes (`[q * cap + i]`, coalesced while
// the queue is dense): after the first bounces the queue entries are increasing but sparse, a plane access then uses half of each
// sector, and ncu had the shade kernels on the DRAM sector rate of exactly these scattered accesses. Measured on the B200
// (tools/gpu_r02_f.sh, profiles/r02_shade_layout.md): shade -14 .. -23 %, the named scenes +6 .. +14 %, films bit-identical.
// BL_SPEC_PLANES rebuilds the old layout for A/B (tools/ab_libs.py).
#if defined(__CUDA_ARCH__)
__device__ __forceinline__ void pfL2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
#endif
HD size_t spec4At(uint32_t cap, uint32_t i, int q) {
#ifndef BL_SPEC_PLANES
   (void)cap; return (size_t)i * 4 + q;
#else
   return (size_t)q * cap + i;
#endif
}
// The 64-byte record moves as TWO 32-byte accesses (LDG.E.256 / STG.E.256) instead of four 16-byte ones: the shade and resolve
// kernels sit on the request rate of their scattered accesses, and a vertex reads the throughput record three times and writes three
// more records (BL_SPEC_V8 0 rebuilds the four-access version for A/B). Records are 64-byte aligned (each array is its own cudaMalloc).
#ifndef BL_SPEC_V8
#define BL_SPEC_V8 1
#endif
HD Spec loadSpec4(const F4 *base, uint32_t cap, uint32_t i) {
   Spec s;
#if defined(__CUDA_ARCH__) && BL_SPEC_V8 && !defined(BL_SPEC_PLANES)
   (void)cap;
   BL_UNROLL for (int h = 0; h < 2; ++h)
      asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                   : "=f"(s.v[8 * h]), "=f"(s.v[8 * h + 1]), "=f"(s.v[8 * h + 2]), "=f"(s.v[8 * h + 3]), "=f"(s.v[8 * h + 4]), "=f"(s.v[8 * h + 5]), "=f"(s.v[8 * h + 6]), "=f"(s.v[8 * h + 7])
                   : "l"(base + (size_t)i * 4 + 2 * h));
#else
   BL_UNROLL for (int q = 0; q < 4; ++q) { F4 v = base[spec4At(cap, i, q)]; s.v[4 * q] = v.x; s.v[4 * q + 1] = v.y; s.v[4 * q + 2] = v.z; s.v[4 * q + 3] = v.w; }
#endif
   return s;
}
HD void storeSpec4(F4 *base, uint32_t cap, uint32_t i, const Spec &s) {
#if defined(__CUDA_ARCH__) && BL_SPEC_V8 && !defined(BL_SPEC_PLANES)
   (void)cap;
   BL_UNROLL for (int h = 0; h < 2; ++h)
      asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(base + (size_t)i * 4 + 2 * h),
                   "f"(s.v[8 * h]), "f"(s.v[8 * h + 1]), "f"(s.v[8 * h + 2]), "f"(s.v[8 * h + 3]), "f"(s.v[8 * h + 4]), "f"(s.v[8 * h + 5]), "f"(s.v[8 * h + 6]), "f"(s.v[8 * h + 7]) : "memory");
#else
   BL_UNROLL for (int q = 0; q < 4; ++q) { F4 v; v.x = s.v[4 * q]; v.y = s.v[4 * q + 1]; v.z = s.v[4 * q + 2]; v.w = s.v[4 * q + 3]; base[spec4At(cap, i, q)] = v; }
#endif
}
// A ray is ONE 32-byte record {(o, tmin), (d, tmax)} -- the C ABI's blingcu_ray as it is -- so that a scattered access touches one
// full 32-byte sector (round 1: separate `O` and `D` arrays = two half-used sectors per ray, in the traversal kernels and four
// times over in every shade kernel, which sit on the sector rate of exactly such accesses). Every (o, d) pointer pair in the
// kernels points INTO one interleaved array: d == o + 1, record i at o[2 i] / d[2 i].
HD size_t rayAt2(uint32_t i) { return 2 * (size_t)i; }
HD void storeRay(F4 *o, F4 *d, uint32_t i, const Ray &r) {
   F4 a, b; a.x = r.o.x; a.y = r.o.y; a.z = r.o.z; a.w = r.tmin; b.x = r.d.x; b.y = r.d.y; b.z = r.d.z; b.w = r.tmax;
#if defined(__CUDA_ARCH__)
   (void)d;   // one 32-byte store (STG.E.256 on sm_100)
   asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(o + rayAt2(i)), "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w), "f"(b.x), "f"(b.y), "f"(b.z), "f"(b.w) : "memory");
#else
   o[rayAt2(i)] = a; d[rayAt2(i)] = b;
#endif
}
HD Ray loadRay(const F4 *o, const F4 *d, uint32_t i) {
   F4 a, b;
#if defined(__CUDA_ARCH__)
   (void)d;   // one 32-byte load (LDG.E.256)
   asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(o + rayAt2(i)));
#else
   a = o[rayAt2(i)]; b = d[rayAt2(i)];
#endif
   Ray r; r.o = mk3(a.x, a.y, a.z); r.tmin = a.w; r.d = mk3(b.x, b.y, b.z); r.tmax = b.w; return r;
}

// L[slot] += s with atomics (direct-lighting integrator: several branch slots feed one camera sample)
#if defined(__CUDA_ARCH__)
__device__ __forceinline__ void addSpec4Atomic(F4 *base, uint32_t cap, uint32_t i, const Spec &s) {
   BL_UNROLL for (int q = 0; q < 4; ++q) { float *p = (float *)(base + spec4At(cap, i, q)); BL_UNROLL for (int k = 0; k < 4; ++k) atomicAdd(p + k, s.v[4 * q + k]); }
}
#else
inline void addSpec4Atomic(F4 *base, uint32_t cap, uint32_t i, const Spec &s) { storeSpec4(base, cap, i, loadSpec4(base, cap, i) + s); }
#endif

// queue append: warp-aggregated atomic on the device, plain increment in the single-threaded emulator
#if defined(__CUDA_ARCH__)
__device__ __forceinline__ void qPush(uint32_t *q, uint32_t *counter, uint32_t v) {
   // lanes of one warp may push to DIFFERENT queues at the same call site (classify: one queue per material
   // kind), so the aggregation groups lanes by counter address
   unsigned m = __match_any_sync(__activemask(), (unsigned long long)(uintptr_t)counter);
   unsigned lane = threadIdx.x & 31u;
   int leader = __ffs(m) - 1;
   uint32_t base = 0;
   if ((int)lane == leader) base = atomicAdd(counter, (uint32_t)__popc(m));
   base = __shfl_sync(m, base, leader);
   q[base + __popc(m & ((1u << lane) - 1u))] = v;
}
__device__ __forceinline__ void statAdd(unsigned long long *p, unsigned long long v) { atomicAdd(p, v); }
__device__ __forceinline__ void cntAdd(uint32_t *p, uint32_t v) { atomicAdd(p, v); }
#else
inline void qPush(uint32_t *q, uint32_t *counter, uint32_t v) { q[(*counter)++] = v; }
inline void statAdd(unsigned long long *p, unsigned long long v) { *p += v; }
inline void cntAdd(uint32_t *p, uint32_t v) { *p += v; }
#endif

HD Sampler mkSampler(const DScene &sc, uint64_t kp, uint32_t s) { Sampler c; c.kp = kp; c.s = s; c.k = &sc.smp; return c; }

// ------------------------------------------------------------------------------------------ K1 raygen
struct RaygenBody {
   const DScene *sc; PathState ps;
   uint64_t seed; uint32_t pass, s0, npix;
   const int32_t *px, *py; const uint32_t *smp;   // explicit sample list (render_samples) or null
   HD void operator()(uint32_t i) const {
      const DScene &S = *sc;
      int ix, iy; uint32_t s, pix;
      if (px) { ix = px[i]; iy = py[i]; s = smp[i]; pix = (uint32_t)(iy - S.ey0) * (uint32_t)S.EW + (uint32_t)(ix - S.ex0); }
      else { pix = i % npix; s = s0 + i / npix; ix = S.ex0 + (int)(pix % (uint32_t)S.EW); iy = S.ey0 + (int)(pix / (uint32_t)S.EW); }
      uint64_t kp = pixelKey(seed, pass, pix);
      Sampler c = mkSampler(S, kp, s);
      float ox, oy, lu, lv; cameraSample(c, ox, oy, lu, lv);
      float sx = (float)ix + ox, sy = (float)iy + oy;
      Ray r = fireRay(S.cam, sx, sy, lu, lv);
      storeRay(ps.rayO, ps.rayD, i, r);
      storeSpec4(ps.T, ps.cap, i, sConst(1)); storeSpec4(ps.L, ps.cap, i, sConst(0));   // Path.hs:38-39: t = white, l = black
      ps.meta[i] = 0u | (1u << 8);   // depth 0, spec = True
      ps.kp[i] = kp; ps.sidx[i] = s;
      F2 p; p.x = sx; p.y = sy; ps.spos[i] = p;
      ps.qA[i] = i;
   }
};

// ------------------------------------------------------------------------------------------ trace (v1 bodies; the CUDA backend has its own kernels)
struct TraceNearestBody {
   const DScene *sc; const F4 *o, *d; F4 *hit;
   HD void operator()(uint32_t i) const {
      HitRec h = traceNearest<false>(sc->bvh, loadRay(o, d, i), 0, 0);
      F4 v; v.x = h.t; v.y = h.b1; v.z = h.b2; v.w = i2f(h.prim); hit[i] = v;
   }
};
struct TraceAnyBody {
   const DScene *sc; const F4 *o, *d; uint8_t *occl;
   HD void operator()(uint32_t i) const { occl[i] = traceAny(sc->bvh, loadRay(o, d, i)) ? 1 : 0; }
};

// ------------------------------------------------------------------------------------------ K4 classify: material-sorted queues
struct ClassifyBody {
   const DScene *sc; PathState ps;
   HD void operator()(uint32_t i) const {
      const DScene &S = *sc;
      const int href = f2i(ps.hit[i].w);
      int kind;
      if (href == BL_REF_MISS) kind = 0;
      else {
         if ((int)(ps.meta[i] & 0xffu) == S.max_depth) return;   // Path.hs:51: depth == md -> return l
         kind = 1 + refKind(href);                                // the shade slot travels in the hit record
      }
      qPush(ps.qMat + (size_t)kind * ps.cap, ps.counters + C_MAT0 + kind, i);
   }
};

// ------------------------------------------------------------------------------------------ K5 shade (miss)
struct ShadeMissBody {   // Path.hs:43-47
   const DScene *sc; PathState ps;
   HD void operator()(uint32_t i) const {
      const DScene &S = *sc;
      if (!((ps.meta[i] >> 8) & 1u)) return;   // non-specular bounce: nothing
      F4 d = ps.rayD[rayAt2(i)];
      Spec sum = sConst(0);
      for (int l = 0; l < S.n_lights; ++l) sum = sum + lightLe(S, S.lights[l], mk3(d.x, d.y, d.z));
      Spec L = loadSpec4(ps.L, ps.cap, i), T = loadSpec4(ps.T, ps.cap, i);
      storeSpec4(ps.L, ps.cap, i, L + T * sum);
   }
};

// ------------------------------------------------------------------------------------------ K5 shade (hit)
// sampleOneLight (Scene.hs:110-118) at one vertex, shared by the path and the direct-lighting shade bodies: queues the
// NEE shadow ray and the BSDF-MIS ray of slot i with their pending contributions (already x T x nLights).
// T (16 registers) is re-read from L1/L2 at each use instead of being kept live: the shade kernels are register-bound.
template <class M>
HD void directAtVertex(const DScene &S, const PathState &ps, uint32_t i, const Bsdf &bsdf, V3 wo, V3 p, V3 n, float eps,
                       float lNumU, float lD1, float lD2, float bCompU, float bD1, float bD2) {
   int lc = S.n_lights;
   if (lc > 0) {   // sampleOneLight (Scene.hs:110-118)
      int ln = (lc == 1) ? 0 : imin((int)floorf(lNumU * (float)lc), lc - 1);
      const blingcu_light &lt = S.lights[ln];
      float lcf = (lc == 1) ? 1.0f : (float)lc;
      {   // sampleLightMis (Scene.hs:61-69)
         LightSample ls; lightSampleOf<M>(S, lt, p, eps, n, lD1, lD2, ls);
         if (ls.pdf != 0 && !isBlack(ls.de)) {
            Spec f = evalBsdfOf<M>(bsdf, wo, ls.wi);
            if (!isBlack(f)) {
               float w = ls.delta ? 1 / ls.pdf : powerHeuristic(ls.pdf, bsdfPdfOf<M>(bsdf, wo, ls.wi)) / ls.pdf;
               Spec c = sScale(f * ls.de, w);
               if (lc > 1) c = sScale(c, lcf);
               storeSpec4(ps.PS, ps.cap, i, loadSpec4(ps.T, ps.cap, i) * c);
               storeRay(ps.shO, ps.shD, i, ls.testRay);
               qPush(ps.qShadow, ps.counters + C_SHADOW, i);
            }
         }
      }
      {   // sampleBsdfMis (Scene.hs:71-82): the ray is traced now, the light lookup happens in the resolve bodies.
         // The reference traces a nearest-hit ray and keeps the sample only if the hit primitive IS the chosen
         // light, or, on a miss, adds `le l ray`. Same result with less traversal:
         //   infinite light  -> only hit/miss matters: any-hit query (qMisAny) -- unless the scene holds a Box shape:
         //                      the reference's Box answers `intersects` for a ray that starts inside it but not
         //                      `intersect` (Shape.hs:86-93 vs :235), so there the nearest-hit query is kept;
         //   area light      -> a ray that does not even reach the light's own shape contributes nothing: culled;
         //   delta lights    -> never hit, `le` is black: culled.
         // All of that holds for a FINITE weight f and pdf. A non-finite one (the glossy sample / eval asymmetry Q7 with extreme
         // parameters) times a black `le` is NaN in the reference (`sc (le l ray)` on a miss, `sc (intLe ..)` on the back of
         // the light), and addSample then drops the whole sample (Image.hs:253-256): such a ray is always traced as the
         // reference's nearest-hit query and carries BL_MIS_NONFINITE to the resolve body, which keeps the black cases.
         BsdfSample bs; sampleBsdfOf<M>(bsdf, wo, bCompU, bD1, bD2, bs);
         if (bs.pdf != 0 && !isBlack(bs.f)) {
            Ray mr; mr.o = p; mr.d = bs.wi; mr.tmin = eps; mr.tmax = BL_INF;
            const bool inf = lt.kind == BLINGCU_LIGHT_INFINITE;
            const float bp2 = bs.pdf * bs.pdf;   // powerHeuristic squares the pdf: inf / inf beyond 1.8e19, 0 / 0 below 1e-23 (delta light)
            const bool nf = sBad(bs.f) || !(bp2 > 0.0f && bp2 <= 3.402823466e38f);
            bool any = inf && !S.has_box && !nf, keep = inf || nf;
            if (lt.kind == BLINGCU_LIGHT_AREA && !nf) {
               const blingcu_shape &ls = S.shapes[lt.shape];
               float tl; DG dgl;
               keep = shapeIntersect<false>(ls, transRay(ls.w2o, mr), tl, dgl);
            }
            if (keep) {
               Spec c = bs.f;
               if (lc > 1) c = sScale(c, lcf);
               storeSpec4(ps.PM, ps.cap, i, loadSpec4(ps.T, ps.cap, i) * c);
               storeRay(ps.miO, ps.miD, i, mr);
               F2 info; info.x = bs.pdf; info.y = i2f(nf ? (ln | BL_MIS_NONFINITE) : ln); ps.miInfo[i] = info;
               if (any) qPush(ps.qMisAny, ps.counters + C_MISANY, i);
               else qPush(ps.qMis, ps.counters + C_MIS, i);
            } else cntAdd(ps.counters + C_MISCULL, 1u);
         }
      }
   }
}

template <int MATKIND>
struct ShadeHitBody {   // Path.hs:49-87 + Scene.hs:61-118; one instantiation per material kind (material-sorted queues)
   typedef MatOf<MATKIND> M;
   const DScene *sc; PathState ps; uint32_t *qNext;
#if defined(__CUDA_ARCH__)
   // Cross-item software pipeline of kRunQueueHeavy (cuda_backend.cu): while a thread shades slot k it already has the records of
   // its NEXT slot on their way into L2. The shade kernels run at 20-25 % occupancy (128-168 registers: 16-band spectra) and ncu
   // shows them waiting on three levels of dependent, scattered loads (queue entry -> path state -> geometry); more warps are not
   // to be had, so the independence comes from the next loop iteration instead.
   __device__ __forceinline__ void prefetchSlot(uint32_t s) const {          // level 1: what operator() reads by slot
      pfL2(ps.rayO + rayAt2(s)); pfL2(ps.meta + s); pfL2(ps.kp + s); pfL2(ps.sidx + s);
      pfL2(ps.T + spec4At(ps.cap, s, 0)); pfL2(ps.T + spec4At(ps.cap, s, 2));
   }
   __device__ __forceinline__ int peekHit(uint32_t s) const { return f2i(ps.hit[s].w); }   // a real load: its value addresses level 2
   __device__ __forceinline__ void prefetchSurface(int href) const {         // level 2: the geometry surfaceAt() reads
      if (href == BL_REF_MISS || refIsShape(href)) return;
      const DScene &S = *sc;
      const uint32_t ref = refIndex(href);
      const F4 *tp = S.tri_p + BL_TRI_F4 * (size_t)ref; pfL2(tp); pfL2(tp + 2);
      if (S.tri_n) { const F4 *N = S.tri_n + 3 * (size_t)ref; pfL2(N); pfL2(N + 2); }
   }
#endif
   HD void operator()(uint32_t i) const {
      const DScene &S = *sc;
      Ray ray = loadRay(ps.rayO, ps.rayD, i);
      F4 hv = ps.hit[i];
      uint32_t meta = ps.meta[i];
      int depth = (int)(meta & 0xffu); bool spec = ((meta >> 8) & 1u) != 0;
      Sampler smp = mkSampler(S, ps.kp[i], ps.sidx[i]);
      SurfaceHit sh; DG dgs;
      surfaceAt(S, ray, hv.x, hv.y, hv.z, f2i(hv.w), sh, dgs);
      Spec texScratch[M::TX ? 4 : 1];   // computed texture values of the textured instantiation (unused otherwise)
      Bsdf bsdf; makeBsdfOf<M>(S, sh, dgs, bsdf, texScratch);
      // T (16 registers) is re-read from L1/L2 at each use instead of being kept live across the whole body: the
      // kernel is register-bound (occupancy), not bandwidth-bound
#define BL_T() loadSpec4(ps.T, ps.cap, i)
      V3 rd = ray.d, wo = -rd;
      // emitted light, only after specular bounces / from the camera; Q1: tested against the RAY direction
      if (spec && sh.light >= 0) {
         const blingcu_light &el = S.lights[sh.light];
         if (el.kind == BLINGCU_LIGHT_AREA && areaEmits(sh.dgg.n, rd)) {
            Spec L = loadSpec4(ps.L, ps.cap, i);
            storeSpec4(ps.L, ps.cap, i, L + BL_T() * loadSpec(el.s.v));
         }
      }
      // A throughput that is not finite (meta bit 9, set where T is updated below) makes `l + t * lHere` (Path.hs:65) NaN or
      // infinite at this vertex whatever lHere is -- also where lHere is black and nothing gets added here -- and addSample
      // drops the sample (Image.hs:253-256): t * 0 reproduces that in exactly the bands the reference loses.
      if (meta & (1u << 9)) storeSpec4(ps.L, ps.cap, i, loadSpec4(ps.L, ps.cap, i) + BL_T() * sConst(0));
      V3 n = bsdf.cs.n, p = bsdf.p; float eps = sh.eps;
      float lNumU = rnd1D(smp, 1 + 4 * depth);
      float lD1, lD2; rnd2D(smp, 1 + 3 * depth, lD1, lD2);
      float bCompU = rnd1D(smp, 2 + 4 * depth);
      float bD1, bD2; rnd2D(smp, 2 + 3 * depth, bD1, bD2);
      directAtVertex<M>(S, ps, i, bsdf, wo, p, n, eps, lNumU, lD1, lD2, bCompU, bD1, bD2);
      // Russian roulette (Path.hs:68-72)
      float pc = (depth <= 7) ? 1.0f : hminf(0.75f, sY(S, BL_T()));
      float x = rnd1D(smp, 3 + 4 * depth);
      if (x > pc) return;
      float uc = rnd1D(smp, 0 + 4 * depth);
      float ud1, ud2; rnd2D(smp, 0 + 3 * depth, ud1, ud2);
      BsdfSample s; sampleBsdfOf<M>(bsdf, wo, uc, ud1, ud2, s);
      if (s.pdf == 0 || isBlack(s.f)) return;
      // The vertex this ray would find has depth == maxDepth: nextVertex returns l there (Path.hs:51), and a miss only
      // adds light after a SPECULAR bounce (Path.hs:43-47). After a non-specular sample the ray decides nothing, so
      // it is not traced (counted in rays_ext_culled; the reference evaluates the intersection and discards it).
      if (depth + 1 == S.max_depth && !(s.type & BX_SPECULAR)) { cntAdd(ps.counters + C_EXTCULL, 1u); return; }
      Ray nr; nr.o = p; nr.d = s.wi; nr.tmin = eps; nr.tmax = BL_INF;
      storeRay(ps.rayO, ps.rayD, i, nr);
      Spec tNext = sScale(s.f * BL_T(), 1 / pc);                   // Path.hs:82: no pdf / cosine factor, the weight carries them
      storeSpec4(ps.T, ps.cap, i, tNext);
      ps.meta[i] = (uint32_t)(depth + 1) | (((s.type & BX_SPECULAR) ? 1u : 0u) << 8) | ((sBad(tNext) ? 1u : 0u) << 9);
      qPush(qNext, ps.counters + C_NEXT, i);
#undef BL_T
   }
};

// ------------------------------------------------------------------------------------------ K6 resolve
// On the GPU the any-hit traversal kernel does this when it retires an unoccluded ray (trace_kernels.cuh, fused NEE resolve);
// this body is what the CPU emulator runs behind its any-hit loop.
struct ResolveShadowBody {   // Scene.hs:64: `occluded scene ray` -> black
   PathState ps;
   HD void operator()(uint32_t i) const {
      if (ps.occl[i]) return;
      storeSpec4(ps.L, ps.cap, i, loadSpec4(ps.L, ps.cap, i) + loadSpec4(ps.PS, ps.cap, i));
   }
};
// where a slot's radiance goes: itself (path integrator; camera samples of the direct-lighting integrator), or the camera
// sample a spawned branch belongs to
HD uint32_t rootOf(const PathState &ps, uint32_t nRoot, uint32_t i) { return i < nRoot ? i : ps.root[i]; }
struct DlResolveShadowBody {
   PathState ps; uint32_t nRoot;
   HD void operator()(uint32_t i) const {
      if (ps.occl[i]) return;
      addSpec4Atomic(ps.L, ps.cap, rootOf(ps, nRoot, i), loadSpec4(ps.PS, ps.cap, i));
   }
};
template <bool DL>
struct ResolveMisBodyT {   // Scene.hs:75-82
   const DScene *sc; PathState ps; uint32_t nRoot;
   HD void operator()(uint32_t i) const {
      const DScene &S = *sc;
      F4 hv = ps.mihit[i];
      F2 info = ps.miInfo[i];
      const bool nf = (f2i(info.y) & BL_MIS_NONFINITE) != 0;          // non-finite weight: black `le` must still reach L (NaN)
      int ln = f2i(info.y) & ~BL_MIS_NONFINITE;
      const blingcu_light &l = S.lights[ln];
      Ray ray = loadRay(ps.miO, ps.miD, i);
      const int href = f2i(hv.w);
      Spec li;
      if (href != BL_REF_MISS) {
         if (!refIsShape(href)) return;                             // triangles carry no light (TriangleMesh.hs:105)
         const blingcu_shape &s = S.shapes[refIndex(href)];
         if (s.light != ln || l.kind != BLINGCU_LIGHT_AREA) return;   // Eq Light: same area-light id (Light.hs:48-50)
         SurfaceHit sh; DG dgs;
         surfaceAt(S, ray, hv.x, hv.y, hv.z, href, sh, dgs);
         if (areaEmits(sh.dgg.n, -ray.d)) li = loadSpec(l.s.v);       // intLe int (-wi)
         else if (nf) li = sConst(0);
         else return;
      } else {
         li = lightLe(S, l, ray.d);
         if (isBlack(li) && !nf) return;
      }
      float w = powerHeuristic(info.x, lightPdf(S, l, ray.o, ray.d));   // Q3: also for specular samples
      Spec c = sScale(loadSpec4(ps.PM, ps.cap, i) * li, w);
      if (DL) addSpec4Atomic(ps.L, ps.cap, rootOf(ps, nRoot, i), c);
      else storeSpec4(ps.L, ps.cap, i, loadSpec4(ps.L, ps.cap, i) + c);
   }
};
typedef ResolveMisBodyT<false> ResolveMisBody;
typedef ResolveMisBodyT<true> DlResolveMisBody;

template <bool DL>
struct ResolveMisAnyBodyT {   // Scene.hs:75-82, miss branch: `le l ray` of an infinite light
   const DScene *sc; PathState ps; uint32_t nRoot;
   HD void operator()(uint32_t i) const {
      if (ps.occlM[i]) return;
      const DScene &S = *sc;
      F2 info = ps.miInfo[i];
      const blingcu_light &l = S.lights[f2i(info.y)];
      Ray ray = loadRay(ps.miO, ps.miD, i);
      Spec li = lightLe(S, l, ray.d);
      if (isBlack(li)) return;
      float w = powerHeuristic(info.x, lightPdf(S, l, ray.o, ray.d));   // Q3: also for specular samples
      Spec c = sScale(loadSpec4(ps.PM, ps.cap, i) * li, w);
      if (DL) addSpec4Atomic(ps.L, ps.cap, rootOf(ps, nRoot, i), c);
      else storeSpec4(ps.L, ps.cap, i, loadSpec4(ps.L, ps.cap, i) + c);
   }
};
typedef ResolveMisAnyBodyT<false> ResolveMisAnyBody;
typedef ResolveMisAnyBodyT<true> DlResolveMisAnyBody;

// ------------------------------------------------------------------------------------------ direct-lighting integrator
// Integrator/DirectLighting.hs:23-58 on the same wavefront: one launch per depth over the active queue.
//   directLighting d: miss -> black; hit -> sampleOneLight + intLe int wo + f_r * L(reflected) + f_t * L(transmitted)
// where the two continuations follow only SPECULAR components (`cont`, :47-58: sampleBsdf' t bsdf wo 0.5 (0.5, 0.5)) and stop
// at d + 1 == maxDepth. The recursion is a tree: the first continuation stays in the slot, the second is spawned into a
// fresh slot (C_SPAWN); every slot adds its radiance to the camera sample it descends from (rootOf) with atomics.
// One instantiation for all material kinds (MatOf<SK_GENERAL>: every BxDF, computing textures, out-of-line general copies).
struct DlShadeBody {
   typedef MatOf<SK_GENERAL> M;
   const DScene *sc; PathState ps; uint32_t *qNext; uint32_t nRoot;
   HD void operator()(uint32_t i) const {
      const DScene &S = *sc;
      F4 hv = ps.hit[i];
      if (f2i(hv.w) == BL_REF_MISS) return;
      Ray ray = loadRay(ps.rayO, ps.rayD, i);
      const int depth = (int)(ps.meta[i] & 0xffu);
      Sampler smp = mkSampler(S, ps.kp[i], ps.sidx[i]);
      SurfaceHit sh; DG dgs;
      surfaceAt(S, ray, hv.x, hv.y, hv.z, f2i(hv.w), sh, dgs);
      Spec texScratch[4];
      Bsdf bsdf; makeBsdfOf<M>(S, sh, dgs, bsdf, texScratch);
      const uint32_t root = rootOf(ps, nRoot, i);
      V3 wo = -ray.d, n = bsdf.cs.n, p = bsdf.p; float eps = sh.eps;
      if (sh.light >= 0) {   // intLe int wo (:45): unlike the path integrator (Q1) the test uses wo
         const blingcu_light &el = S.lights[sh.light];
         if (el.kind == BLINGCU_LIGHT_AREA && areaEmits(sh.dgg.n, wo)) addSpec4Atomic(ps.L, ps.cap, root, loadSpec4(ps.T, ps.cap, i) * loadSpec(el.s.v));
      }
      float uln = rnd1D(smp, 0 + 2 * depth);
      float uld1, uld2; rnd2D(smp, 0 + 2 * depth, uld1, uld2);
      float ubc = rnd1D(smp, 1 + 2 * depth);
      float ubd1, ubd2; rnd2D(smp, 1 + 2 * depth, ubd1, ubd2);
      directAtVertex<M>(S, ps, i, bsdf, wo, p, n, eps, uln, uld1, uld2, ubc, ubd1, ubd2);
      if (depth + 1 == S.max_depth) return;   // cont: d == md -> black
      const Spec T = loadSpec4(ps.T, ps.cap, i);
      const uint64_t kp = ps.kp[i]; const uint32_t sidx = ps.sidx[i];
      bool first = true;
      for (int c = 0; c < 2; ++c) {
         BsdfSample s; sampleBsdfSpecularGeneral(bsdf, BX_SPECULAR | (c == 0 ? BX_REFLECTION : BX_TRANSMISSION), wo, s);
         if (s.pdf == 0) continue;
         uint32_t j = i;
         if (!first) {
#if defined(__CUDA_ARCH__)
            j = atomicAdd(ps.counters + C_SPAWN, 1u);
#else
            j = ps.counters[C_SPAWN]++;
#endif
            if (j >= ps.cap) { ps.counters[C_OVERFLOW] = 1u; continue; }   // the host re-runs the batch with more head-room
            ps.root[j] = root; ps.kp[j] = kp; ps.sidx[j] = sidx;
         }
         first = false;
         Ray nr; nr.o = p; nr.d = s.wi; nr.tmin = eps; nr.tmax = BL_INF;
         storeRay(ps.rayO, ps.rayD, j, nr);
         const Spec tNext = s.f * T;
         storeSpec4(ps.T, ps.cap, j, tNext);
         // `f * l` (:58) of a weight that is not finite is NaN or infinite whatever the continuation finds (black on a miss
         // included): poison the camera sample now, in the bands the reference loses (cf. ShadeHitBody, meta bit 9)
         if (sBad(tNext)) addSpec4Atomic(ps.L, ps.cap, root, tNext * sConst(0));
         ps.meta[j] = (uint32_t)(depth + 1) | (1u << 8);
         qPush(qNext, ps.counters + C_NEXT, j);
      }
   }
};

// `debug normals` (mkNormalMap, Integrator/Debug.hs:23-33): the shading normal of the first hit as a reflectance spectrum
struct NormalMapBody {
   typedef MatOf<SK_GENERAL> M;   // bump mapping moves the shading normal
   const DScene *sc; PathState ps;
   HD void operator()(uint32_t i) const {
      const DScene &S = *sc;
      F4 hv = ps.hit[i];
      if (f2i(hv.w) == BL_REF_MISS) return;   // Nothing -> black (raygen left L = 0)
      Ray ray = loadRay(ps.rayO, ps.rayD, i);
      SurfaceHit sh; DG dgs;
      surfaceAt(S, ray, hv.x, hv.y, hv.z, f2i(hv.w), sh, dgs);
      Spec texScratch[4];
      Bsdf bsdf; makeBsdfOf<M>(S, sh, dgs, bsdf, texScratch);
      V3 n = bsdf.cs.n;
      storeSpec4(ps.L, ps.cap, i, rgbToSpectrumBasis(S.refl, (1.0f + n.x) / 2, (1.0f + n.y) / 2, (1.0f + n.z) / 2));
   }
};

// one thread: fold the queue counters into the statistics and rotate the queues (end of a bounce)
struct AdvanceBody {
   PathState ps;
   HD void operator()(uint32_t) const {
      uint32_t *c = ps.counters;
      statAdd(ps.stats + S_EXT, c[C_NEXT]); statAdd(ps.stats + S_SHADOW, c[C_SHADOW]); statAdd(ps.stats + S_MIS, c[C_MIS] + c[C_MISANY]);
      statAdd(ps.stats + S_MISCULL, c[C_MISCULL]); statAdd(ps.stats + S_MISANY, c[C_MISANY]); statAdd(ps.stats + S_EXTCULL, c[C_EXTCULL]);
      c[C_ACTIVE] = c[C_NEXT]; c[C_NEXT] = 0; c[C_SHADOW] = 0; c[C_MIS] = 0; c[C_MISANY] = 0; c[C_MISCULL] = 0; c[C_EXTCULL] = 0;
      for (int k = 0; k < N_SHADE_KINDS; ++k) c[C_MAT0 + k] = 0;
   }
};
struct BeginBatchBody {
   PathState ps; uint32_t n;
   HD void operator()(uint32_t) const {
      uint32_t *c = ps.counters;
      for (int k = 0; k < N_COUNTERS; ++k) if (k != C_DROPPED) c[k] = 0;
      c[C_ACTIVE] = n; c[C_SPAWN] = n;
      statAdd(ps.stats + S_CAM, n); statAdd(ps.stats + S_SAMPLES, n);
   }
};

// ------------------------------------------------------------------------------------------ finalize: spectrum -> XYZ (Image.hs:253-258)
struct FinalizeBody {
   const DScene *sc; PathState ps;
   HD void operator()(uint32_t i) const {
      Spec L = loadSpec4(ps.L, ps.cap, i);
      F4 o; o.x = o.y = o.z = o.w = 0;
      if (sBad(L)) { statAdd(ps.stats + S_DROPPED, 1); }   // NaN / infinite samples are skipped
      else { spectrumToXYZ(*sc, L, o.x, o.y, o.z); o.w = 1.0f; }
      ps.xyz[i] = o;
   }
};

// ------------------------------------------------------------------------------------------ K7 film: atomic-free gather, one thread per film pixel
// Reproduces addSample on per-tile images + addTile (Image.hs:108-120,178-199,250-299, Q10): a sample only
// reaches pixels of ITS 16x16 sample window's tile image, which has no left/top apron and extends
// floor(0.5+f) pixels right/bottom.
// shared pieces of the two film kernels (FilmBody: one thread per pixel straight from global memory; kFilmTile in
// cuda_backend.cu: one CTA per 16x16 pixel tile with the candidate samples staged in shared memory)
struct FilmGeom {   // per-scene constants of the gather
   float fw, fh, ifw, ifh; int extx, exty;
};
HD FilmGeom filmGeom(const DScene &S) {
   FilmGeom g; g.fw = S.fw; g.fh = S.fh; g.ifw = 1 / S.fw; g.ifh = 1 / S.fw;   // Q10: ifh = 1 / fw
   g.extx = (int)floorf(0.5f + S.fw); g.exty = (int)floorf(0.5f + S.fh);
   return g;
}
// sample pixels whose samples can reach film pixel column/row v: a sample of pixel i sits at d = i + o - 0.5 with o in
// [0,1] (the sum may round up to i + 1), and covers v iff ceil(d - f) <= v <= floor(d + f)
//   => i in [ceil(v - f - 0.5), floor(v + f + 0.5)]   (5x5 pixels for a radius-2 filter, not 9x9)
HD int filmCellLo(int v, float f) { return (int)ceilf((float)v - f - 0.5f); }
HD int filmCellHi(int v, float f) { return (int)floorf((float)v + f + 0.5f); }
// the tile image of the 16x16 sample window that holds sample pixel i covers [to, tmax] along one axis (Image.hs:108-120:
// no left/top apron, `ext` pixels to the right/bottom)
HD void filmTileSpan(int i, int e0, int e1, int ext, int &to, int &tmax) {
   int t0 = e0 + ((i - e0) >> 4) * 16, t1 = imin(t0 + 15, e1);
   to = imax(0, t0); tmax = t1 + ext - 1;
}
// addSample of one sample into pixel (x, y) of its tile image (Image.hs:250-299)
HD void filmAddSample(const DScene &S, const FilmGeom &g, int x, int y, int tox, int txmax, int toy, int tymax, const F4 &c, const F2 &sp,
                      float &aw, float &ax, float &ay, float &az) {
   if (c.w == 0) return;
   float dx = sp.x - 0.5f, dy = sp.y - 0.5f;
   int x0 = imax(tox, (int)ceilf(dx - g.fw)), x1 = imin(txmax, (int)floorf(dx + g.fw));
   int y0 = imax(toy, (int)ceilf(dy - g.fh)), y1 = imin(tymax, (int)floorf(dy + g.fh));
   if (x < x0 || x > x1 || y < y0 || y > y1) return;
   int tix = imin((int)floorf(fabsf(((float)x - dx) * g.ifw * 16.0f)), 15);
   int tiy = imin((int)floorf(fabsf(((float)y - dy) * g.ifh * 16.0f)), 15);
   float w = S.ftbl[tiy * 16 + tix];
   aw = aw + w; ax = ax + c.x * w; ay = ay + c.y * w; az = az + c.z * w;
}

struct FilmBody {
   const DScene *sc; PathState ps; F4 *film; uint32_t k, npix;
   HD void operator()(uint32_t fp) const {
      const DScene &S = *sc;
      int x = (int)(fp % (uint32_t)S.W), y = (int)(fp / (uint32_t)S.W);
      const FilmGeom g = filmGeom(S);
      int ixlo = filmCellLo(x, g.fw), ixhi = filmCellHi(x, g.fw), iylo = filmCellLo(y, g.fh), iyhi = filmCellHi(y, g.fh);
      float aw = 0, ax = 0, ay = 0, az = 0;
      for (int iy = imax(S.ey0, iylo); iy <= imin(S.ey1, iyhi); ++iy) {
         int toy, tymax; filmTileSpan(iy, S.ey0, S.ey1, g.exty, toy, tymax);
         if (y < toy || y > tymax) continue;
         for (int ix = imax(S.ex0, ixlo); ix <= imin(S.ex1, ixhi); ++ix) {
            int tox, txmax; filmTileSpan(ix, S.ex0, S.ex1, g.extx, tox, txmax);
            if (x < tox || x > txmax) continue;
            uint32_t pix = (uint32_t)(iy - S.ey0) * (uint32_t)S.EW + (uint32_t)(ix - S.ex0);
            for (uint32_t sl = 0; sl < k; ++sl) {
               uint32_t slot = sl * npix + pix;
               filmAddSample(S, g, x, y, tox, txmax, toy, tymax, ps.xyz[slot], ps.spos[slot], aw, ax, ay, az);
            }
         }
      }
      F4 f = film[fp];
      f.x = f.x + aw; f.y = f.y + ax; f.z = f.z + ay; f.w = f.w + az;
      film[fp] = f;
   }
};

// explicit ray batches: ABI layout <-> kernel layout
struct HitToAbiBody {   // kernel hit (t, b1, b2, ref) -> ABI hit {t, prim id, b1, b2}
   const DScene *sc; F4 *h;
   HD void operator()(uint32_t i) const {
      F4 v = h[i]; const int href = f2i(v.w);
      int prim = -1;
      if (href != BL_REF_MISS) prim = refIsShape(href) ? sc->shapes[refIndex(href)].prim_id : sc->tri_prim[refIndex(href)];
      F4 r; r.x = v.x; r.y = i2f(prim); r.z = v.y; r.w = v.z; h[i] = r;
   }
};

// blingcu_eval_texture: one texture-table entry at explicit points
struct EvalTextureBody {
   const DScene *sc; int tex; const float *p, *uv; float *out;
   HD void operator()(uint32_t i) const {
      const DScene &S = *sc;
      DG dg; dg.p = mk3(p[3 * i], p[3 * i + 1], p[3 * i + 2]); dg.u = uv[2 * i]; dg.v = uv[2 * i + 1];
      dg.n = mk3(0, 0, 1); dg.dpdu = mk3(1, 0, 0); dg.dpdv = mk3(0, 1, 0);
      Spec s = sConst(0);
      if (S.textures[tex].kind >= BLINGCU_STEX_CONSTANT) s.v[0] = evalScalarTexture(S, tex, dg);
      else s = SpectrumValue<BL_BLEND_DEPTH>::eval(S, tex, dg);
      for (int k = 0; k < NB; ++k) out[(size_t)NB * i + k] = s.v[k];
   }
};

struct AddFilmBody { F4 *dst; const F4 *src; HD void operator()(uint32_t i) const { F4 a = dst[i], b = src[i]; a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; dst[i] = a; } };

}  // namespace bl
#include "lighttrace.h"
#include "bidir.h"   // the light tracer on the same kernels
#include "kdtree.h"   // the host's kd-tree as an alternative accelerator input (needs loadRay above)
