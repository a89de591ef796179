// ORACLE (test infrastructure, NOT product code). See oracle_math.h header. PARITY UNPINNED.
// Materials, lights, the path integrator, camera, film and the C API of the oracle.
// Restates Material.hs, Texture.hs, Light.hs, Scene.hs, Integrator/Path.hs, Camera.hs, Image.hs,
// Rendering.hs (tile decomposition) of /root/reference/src/lib/Graphics/Bling.
#include "oracle_shade.h"
#include "oracle_texture.h"
#include "oracle.h"
#include <atomic>
#include <chrono>
#include <string>
#include <thread>

namespace orc {

struct Scene {
   Geometry geo;
   std::vector<blingcu_material> materials;
   std::vector<blingcu_texture> textures;
   std::vector<blingcu_image> images;             // data pointers into imageData
   std::vector<std::vector<float>> imageData;
   Spec refl[7];                                  // rgbReflectance basis (Spectrum.hs:128-133)
   TexEnv texEnv() const { return TexEnv{textures, images}; }
   std::vector<blingcu_light> lights;
   std::vector<blingcu_envmap> envs;
   std::vector<std::vector<float>> envData;  // owns copies of env arrays
   std::vector<int> triMaterial;
   std::vector<float> triNormals;  // 9 per tri or empty
   blingcu_camera cam;
   SpectralTables T;
   int W, H;
   float fw, fh;
   float ftbl[256];
   int samplerKind, nu, nv, maxDepth, sampleDepth;
   int integrator = BLINGCU_INTEGRATOR_PATH;
   bool useKd = true;
   // film
   std::vector<float> film;  // H*W*4
   // stats
   std::atomic<uint64_t> nSamples{0}, rCam{0}, rExt{0}, rMis{0}, rShadow{0}, dropped{0};
   // light tracer (Renderer/LightTracer.hs): the splat buffer _imgS = [H][W]{X, Y, Z} and its counters
   std::vector<float> splat;
   std::atomic<uint64_t> nPhotons{0}, rLight{0}, rConnect{0}, nSplats{0};
};

struct RayCounters { uint64_t cam = 0, ext = 0, mis = 0, shadow = 0; };

// ----------------------------------------------------------------------------- textures (Texture.hs:159-207)
static Spec evalSpectrumTexture(const Scene &sc, int id, const DG &dg) {
   const blingcu_texture &t = sc.textures[id];
   switch (t.kind) {
   case BLINGCU_TEX_CONSTANT: return fromC(t.s);
   case BLINGCU_TEX_CHECKER: {   // checkerBoard (Texture.hs:209-221): (floor x*sx + floor y*sy + floor z*sz) `mod` 2 == 0
      long q = (long)std::floor(dg.p.x * t.f[0]) + (long)std::floor(dg.p.y * t.f[1]) + (long)std::floor(dg.p.z * t.f[2]);
      return evalSpectrumTexture(sc, (((q % 2) + 2) % 2 == 0) ? t.child[0] : t.child[1], dg);
   }
   case BLINGCU_TEX_GRAPHPAPER: {   // graphPaper lw m p l (Texture.hs:191-207); m = uvMapping (:166-170) or planarMapping (:172-181)
      float lw = t.f[0];
      float x, z;
      if (t.aux == 0) { x = t.f[1] * dg.u + t.f[3]; z = t.f[2] * dg.v + t.f[4]; } else mapping2d(t.s.v, dg, x, z);
      float xi = std::trunc(x), zi = std::trunc(z);  // properFraction: integer part truncates toward zero
      float xpp = x - xi, zpp = z - zi;
      float xp = std::fabs(xpp), zp = std::fabs(zpp);
      float lo = lw / 2, hi = 1.0f - lo;
      return evalSpectrumTexture(sc, (xp < lo || zp < lo || xp > hi || zp > hi) ? t.child[1] : t.child[0], dg);
   }
   case BLINGCU_TEX_BLEND: {   // spectrumBlend (Texture.hs:129-141)
      Spec v1 = evalSpectrumTexture(sc, t.child[0], dg), v2 = evalSpectrumTexture(sc, t.child[1], dg);
      float x = evalScalarTexture(sc.texEnv(), t.aux, dg);
      if (x <= 0) return v1;
      if (x >= 1) return v2;
      return sScale(v1, 1 - x) + sScale(v2, x);
   }
   case BLINGCU_TEX_GRADIENT: {   // gradient (Texture.hs:239-253); the IR holds gradCols sorted (mkGradient :232-237)
      float f = evalScalarTexture(sc.texEnv(), t.aux, dg);
      const blingcu_texture *cols = &sc.textures[t.child[0]]; int n = t.child[1];
      float gmin = cols[0].f[0], gmax = cols[n - 1].f[0];
      if (f <= gmin) return fromC(cols[0].s);
      if (f >= gmax) return fromC(cols[n - 1].s);
      int idx = 0; while (!(cols[idx].f[0] > f)) ++idx;   // fromJust $ V.findIndex ((> f) . fst)
      const blingcu_texture &e0 = cols[idx - 1], &e1 = cols[idx];
      float weight = (f - e0.f[0]) / (e1.f[0] - e0.f[0]);
      return sScale(fromC(e0.s), 1 - weight) + sScale(fromC(e1.s), weight);
   }
   case BLINGCU_TEX_IMAGE: {   // imageTexture tm mapping dg = texMapEval tm (mapping dg); getPixel + pixelSpectrum (Texture.hs:87-101,125-126)
      float x, y; mapping2d(t.s.v, dg, x, y);
      const float *px = imagePixelAt(sc.images[t.aux], x, y);   // already (fromIntegral c / 255) ** 2.2 (unGamma, host side)
      return rgbToSpectrum(sc.refl, px[0], px[1], px[2]);
   }
   default: return sConst(0);
   }
}

// ----------------------------------------------------------------------------- materials (Material.hs:32-96)
static BxDF mkLambertian(const Spec &r) { BxDF b{}; b.kind = K_LAMBERT; b.type = BX_REFLECTION | BX_DIFFUSE; b.r = r; return b; }
static BxDF mkOrenNayar(const Spec &r, float sig) {  // Diffuse.hs:29-36
   BxDF b{}; b.kind = K_ORENNAYAR; b.type = BX_REFLECTION | BX_DIFFUSE; b.r = r;
   float s = clampf(sig, 0, 1); float sig2 = s * s;
   b.a = 1 - (sig2 / (2 * (sig2 + 0.33f)));
   b.b = 0.45f * sig2 / (sig2 + 0.09f);
   return b;
}
static float fixExponent(float e) { return (e > 10000 || std::isnan(e)) ? 10000 : e; }  // Microfacet.hs:127-129

// Primitive.hs:57-65 mkIntersection: bsdf = mat dg (shadingGeometry p dg mempty); Reflection.hs:209-225 mkBsdf'
static Bsdf makeBsdf(const Scene &sc, const Hit &hit) {
   const Prim &pr = sc.geo.prims[hit.prim];
   DG dgs = hit.dg;
   int matId;
   if (pr.is_tri) {
      matId = sc.triMaterial[pr.idx];
      if (!sc.triNormals.empty()) {  // triangleShadingGeometry (TriangleMesh.hs:122-134) with o2w = mempty
         const float *N = &sc.triNormals[9 * (size_t)pr.idx];
         V3 n0 = mk(N[0], N[1], N[2]), n1 = mk(N[3], N[4], N[5]), n2 = mk(N[6], N[7], N[8]);
         bool flat = true;   // blingcu.h: nine zeros = a mesh without normals (Nothing in TriangleMesh.hs:39-60) beside smooth ones
         for (int k = 0; k < 9; ++k) flat = flat && N[k] == 0;
         float b1 = hit.dg.b1, b2 = hit.dg.b2, b0 = 1 - b1 - b2;
         if (!flat) {
            V3 nsp = (scl(b0, n0) + scl(b1, n1)) + scl(b2, n2);
            V3 ns = normalize(nsp);  // transNormal identity
            V3 ssp = normalize(hit.dg.dpdu);
            V3 tsp = cross(ssp, ns);
            if (sqLen(tsp) > 0) { dgs.dpdu = cross(normalize(tsp), ns); dgs.dpdv = normalize(tsp); }
            else { Frame f = coordinateSystem(ns); dgs.dpdu = f.s; dgs.dpdv = f.t; }
            dgs.n = ns;
         }
      }
   } else matId = sc.geo.shapes[pr.idx].material;
   const blingcu_material &m0 = sc.materials[matId];
   if (m0.bump) dgs = bump(sc.texEnv(), m0.bump - 1, hit.dg, dgs);   // bumpMapped d mat dgg dgs = mat dgg $ bump d dgg dgs (Reflection.hs:344-345)
   blingcu_material m = m0;   // ScalarTexture parameters (sigma, ior, rough, ...) are evaluated at the shading geometry
   for (int i = 0; i < 3; ++i) if (m0.ftex[i]) m.f[i] = evalScalarTexture(sc.texEnv(), m0.ftex[i] - 1, dgs);
   Bsdf b; b.n = 0;
   switch (m.kind) {
   case BLINGCU_MAT_MATTE: {
      Spec r = evalSpectrumTexture(sc, m.tex[0], dgs);
      float s = m.f[0];
      b.bx[0] = (s == 0) ? mkLambertian(r) : mkOrenNayar(r, s);
      b.n = 1; break;
   }
   case BLINGCU_MAT_GLASS: {
      Spec r = sClamp(0, 1, evalSpectrumTexture(sc, m.tex[0], dgs));
      Spec t = sClamp(0, 1, evalSpectrumTexture(sc, m.tex[1], dgs));
      float ior = m.f[0];
      BxDF refl{}; refl.kind = K_SPECREFL; refl.type = BX_REFLECTION | BX_SPECULAR; refl.r = r; refl.fr = FR_DIELECTRIC; refl.etai = 1; refl.etat = ior;
      BxDF tr{}; tr.kind = K_SPECTRANS; tr.type = BX_TRANSMISSION | BX_SPECULAR; tr.r = t; tr.etai = 1; tr.etat = ior;
      b.bx[0] = refl; b.bx[1] = tr; b.n = 2; break;
   }
   case BLINGCU_MAT_MIRROR: {
      BxDF refl{}; refl.kind = K_SPECREFL; refl.type = BX_REFLECTION | BX_SPECULAR; refl.fr = FR_NOOP;
      refl.r = sClamp(0, 1, evalSpectrumTexture(sc, m.tex[0], dgs));
      b.bx[0] = refl; b.n = 1; break;
   }
   case BLINGCU_MAT_PLASTIC: {
      Spec rd = evalSpectrumTexture(sc, m.tex[0], dgs), rs = evalSpectrumTexture(sc, m.tex[1], dgs);
      BxDF spec{}; spec.kind = K_MICROFACET; spec.type = BX_REFLECTION | BX_GLOSSY; spec.r = rs;
      spec.fr = FR_DIELECTRIC; spec.etai = 1.0f; spec.etat = 1.5f; spec.e = fixExponent(1 / m.f[0]);
      b.bx[0] = mkLambertian(rd); b.bx[1] = spec; b.n = 2; break;
   }
   case BLINGCU_MAT_METAL: {
      BxDF spec{}; spec.kind = K_MICROFACET; spec.type = BX_REFLECTION | BX_GLOSSY; spec.r = sConst(1);
      spec.fr = FR_CONDUCTOR; spec.eta = evalSpectrumTexture(sc, m.tex[0], dgs); spec.k = evalSpectrumTexture(sc, m.tex[1], dgs);
      spec.e = fixExponent(1 / m.f[0]);
      b.bx[0] = spec; b.n = 1; break;
   }
   case BLINGCU_MAT_SHINYMETAL: {   // mkShinyMetal (Material.hs:98-108). The IR's four textures are ks / kr passed through
      // frApproxEta / frApproxK (Fresnel.hs:72-78) leaf by leaf on the host (bling_b200/host/spectra.py)
      BxDF diff{}; diff.kind = K_MICROFACET; diff.type = BX_REFLECTION | BX_GLOSSY; diff.r = sConst(1); diff.fr = FR_CONDUCTOR;
      diff.eta = evalSpectrumTexture(sc, m.tex[0], dgs); diff.k = evalSpectrumTexture(sc, m.tex[1], dgs); diff.e = fixExponent(1 / m.f[0]);
      BxDF spec{}; spec.kind = K_SPECREFL; spec.type = BX_REFLECTION | BX_SPECULAR; spec.r = sConst(1); spec.fr = FR_CONDUCTOR;
      spec.eta = evalSpectrumTexture(sc, m.tex[2], dgs); spec.k = evalSpectrumTexture(sc, m.tex3, dgs);
      b.bx[0] = diff; b.bx[1] = spec; b.n = 2; break;
   }
   case BLINGCU_MAT_TRANSMATTE: {   // translucentMatte (Material.hs:43-53)
      Spec r = sClamp(0, 1, evalSpectrumTexture(sc, m.tex[0], dgs));
      Spec t = sClamp(0, 1, evalSpectrumTexture(sc, m.tex[1], dgs)) * (sConst(1) - r);
      float s = m.f[0];
      BxDF refl = (s == 0) ? mkLambertian(r) : mkOrenNayar(r, s);
      BxDF trans = (s == 0) ? mkLambertian(t) : mkOrenNayar(t, s);
      trans.flip = true; trans.type = BX_TRANSMISSION | BX_DIFFUSE;   // bxdfTypeFlip [Reflection, Transmission]
      b.bx[0] = refl; b.bx[1] = trans; b.n = 2; break;
   }
   case BLINGCU_MAT_SUBSTRATE: {   // mkSubstrate (Material.hs:110-127)
      BxDF fb{}; fb.kind = K_FRESNELBLEND; fb.type = BX_REFLECTION | BX_GLOSSY;
      fb.r = sClamp(0, 1, evalSpectrumTexture(sc, m.tex[0], dgs));
      fb.rs = sClamp(0, 1, evalSpectrumTexture(sc, m.tex[1], dgs));
      fb.ra = sClamp(0, 1, evalSpectrumTexture(sc, m.tex[2], dgs));
      float u = hmax(0, m.f[0]), v = hmax(0, m.f[1]);
      fb.ex = fixExponent(1 / u); fb.ey = fixExponent(1 / v); fb.depth = m.f[2];
      b.bx[0] = fb; b.n = 1; break;
   }
   default: break;  // blackbody: no BxDFs
   }
   // mkBsdf (Reflection.hs:209-218)
   V3 nn = dgs.n, sn = normalize(dgs.dpdu);
   b.cs = Frame{sn, cross(nn, sn), nn};
   b.p = dgs.p;
   b.ng = hit.dg.n;
   return b;
}

// ----------------------------------------------------------------------------- lights (Light.hs)
static int hitLight(const Scene &sc, const Hit &h) {  // intLight
   const Prim &pr = sc.geo.prims[h.prim];
   return pr.is_tri ? -1 : sc.geo.shapes[pr.idx].light;
}
// intLe (Primitive.hs:68-76) + lEmit (Light.hs:85-96)
static Spec intLe(const Scene &sc, const Hit &h, V3 wo) {
   int li = hitLight(sc, h);
   if (li < 0) return sConst(0);
   const blingcu_light &l = sc.lights[li];
   if (l.kind != BLINGCU_LIGHT_AREA) return sConst(0);
   return (dot(h.dg.n, wo) > 0) ? fromC(l.s) : sConst(0);
}
// le (Light.hs:98-106)
static Spec lightLe(const Scene &sc, const blingcu_light &l, const Ray &ray) {
   if (l.kind != BLINGCU_LIGHT_INFINITE) return sConst(0);
   const blingcu_envmap &e = sc.envs[l.env];
   V3 wh = normalize(transVector(e.w2l, ray.d));
   float phi = sphericalPhi(wh), theta = sphericalTheta(wh);
   return envEval(sc.T, e, phi / (2 * kPi), theta / kPi);   // sphToCart, Types.hs:36-39
}
struct LightSample { Spec de; V3 wi; Ray testRay; float pdf; bool delta; };
// Light.hs:122-160
static LightSample lightSample(const Scene &sc, const blingcu_light &l, V3 p, float eps, V3 n, float u1, float u2) {
   LightSample black{sConst(0), mk(0, 1, 0), Ray{mk(0, 0, 0), mk(0, 1, 0), 0, 1}, 0, false};
   switch (l.kind) {
   case BLINGCU_LIGHT_INFINITE: {
      const blingcu_envmap &e = sc.envs[l.env];
      float u, v, mapPdf;
      sampleContinuous2D(e, u1, u2, u, v, mapPdf);
      if (mapPdf == 0) return black;
      float phi = u * 2 * kPi, theta = v * kPi;  // cartToSph, Types.hs:31-33
      float sint = std::sin(theta);
      if (sint == 0) return black;
      Spec ls = envEval(sc.T, e, u, v);
      V3 wi = transVector(e.l2w, sphericalDirection(std::sin(theta), std::cos(theta), phi));
      return LightSample{ls, wi, Ray{p, wi, eps, kInf}, mapPdf / (2 * kPi * kPi * sint), false};
   }
   case BLINGCU_LIGHT_DIRECTIONAL: {
      V3 d = mk(l.v[0], l.v[1], l.v[2]);
      return LightSample{sScale(fromC(l.s), absDot(n, d)), d, Ray{p, d, eps, kInf}, 1, true};
   }
   case BLINGCU_LIGHT_POINT: {
      V3 pos = mk(l.v[0], l.v[1], l.v[2]);
      return LightSample{sScale(fromC(l.s), 1 / sqLen(pos - p)), normalize(pos - p), Ray{p, pos - p, eps, kInf}, 1, true};
   }
   default: {  // AreaLight: sample in light-local space (Q11)
      const blingcu_shape &s = sc.geo.shapes[l.shape];
      V3 pl = transPoint(s.w2o, p);
      V3 ps, ns; sampleShape(s, pl, u1, u2, ps, ns);
      V3 wi = normalize(ps - pl);
      float pd = shapePdf(s, pl, wi);
      Ray ray{pl, wi, eps, len(ps - pl) - eps};
      Spec r = (dot(ns, wi) < 0) ? fromC(l.s) : sConst(0);
      return LightSample{r, transVector(s.o2w, wi), transRay(s.o2w, ray), pd, false};
   }
   }
}
// Light.hs:215-229
static float lightPdf(const Scene &sc, const blingcu_light &l, V3 p, V3 wiW) {
   switch (l.kind) {
   case BLINGCU_LIGHT_INFINITE: {
      const blingcu_envmap &e = sc.envs[l.env];
      V3 w = transVector(e.w2l, wiW);
      float phi = sphericalPhi(w), theta = sphericalTheta(w);
      float sint = std::sin(theta);
      if (sint == 0) return 0;
      return pdfDist2D(e, phi / (2 * kPi), theta / kPi) / (2 * kPi * kPi * sint);
   }
   case BLINGCU_LIGHT_AREA: {
      const blingcu_shape &s = sc.geo.shapes[l.shape];
      return shapePdf(s, transPoint(s.w2o, p), transVector(s.w2o, wiW));
   }
   default: return 0;
   }
}

static Hit sceneIntersect(const Scene &sc, const Ray &r) { return sc.useKd ? sc.geo.kdNearest(r) : sc.geo.bruteNearest(r); }  // Scene.hs:49-51
static bool sceneOccluded(const Scene &sc, const Ray &r) { return sc.useKd ? sc.geo.kdOccluded(r) : sc.geo.bruteOccluded(r); }  // Scene.hs:45-47

// Scene.hs:61-69
static Spec sampleLightMis(const Scene &sc, const LightSample &ls, const Bsdf &bsdf, V3 wo, RayCounters &rc) {
   if (ls.pdf == 0 || isBlack(ls.de)) return sConst(0);
   Spec f = evalBsdf(bsdf, wo, ls.wi);
   if (isBlack(f)) return sConst(0);
   rc.shadow++;
   if (sceneOccluded(sc, ls.testRay)) return sConst(0);
   if (ls.delta) return sScale(f * ls.de, 1 / ls.pdf);
   float weight = powerHeuristic(1, ls.pdf, 1, bsdfPdf(bsdf, wo, ls.wi));
   return sScale(f * ls.de, weight / ls.pdf);
}
// Scene.hs:71-82
static Spec sampleBsdfMis(const Scene &sc, int lightIdx, const BsdfSample &bs, V3 p, float epsilon, RayCounters &rc) {
   if (bs.pdf == 0 || isBlack(bs.f)) return sConst(0);
   const blingcu_light &l = sc.lights[lightIdx];
   Ray ray{p, bs.wi, epsilon, kInf};
   rc.mis++;
   Hit lint = sceneIntersect(sc, ray);
   Spec li;
   if (lint.valid) {
      int hl = hitLight(sc, lint);
      if (hl < 0) return sConst(0);
      // Eq Light: only two area lights with the same id are equal (Light.hs:48-50)
      if (!(sc.lights[hl].kind == BLINGCU_LIGHT_AREA && l.kind == BLINGCU_LIGHT_AREA && hl == lightIdx)) return sConst(0);
      li = intLe(sc, lint, -bs.wi);
   } else li = lightLe(sc, l, ray);
   float lPdf = lightPdf(sc, l, p, bs.wi);
   return sScale(bs.f * li, powerHeuristic(1, bs.pdf, 1, lPdf));  // Q3: also for specular samples
}

// ----------------------------------------------------------------------------- camera (Camera.hs:49-76)
static Ray fireRay(const Scene &sc, float ix, float iy, float lu, float lv) {
   const blingcu_camera &c = sc.cam;
   if (c.kind == BLINGCU_CAM_ENVIRONMENT) {
      float t = kPi * iy / c.env_sy, p = 2 * kPi * ix / c.env_sx;
      Ray r{mk(0, 0, 0), mk(std::sin(t) * std::cos(p), std::cos(t), std::sin(t) * std::sin(p)), 0, kInf};
      return transRay(c.cam2world, r);
   }
   V3 pCamera = transPoint(c.raster2cam, mk(ix, iy, 0));
   Ray ray{mk(0, 0, 0), normalize(pCamera), 0, kInf};
   if (c.lens_radius > 0) {
      float dx, dy; concentricSampleDisk(lu, lv, dx, dy);
      V3 ro = mk(dx * c.lens_radius, dy * c.lens_radius, 0);
      V3 pFocus = rayAt(ray, c.focal_distance / ray.d.z);
      ray = Ray{ro, normalize(pFocus - ro), 0, kInf};
   }
   return transRay(c.cam2world, ray);
}

// ----------------------------------------------------------------------------- Integrator/Path.hs:41-87
static Spec pathLi(const Scene &sc, const SampleCtx &smp, Ray ray, RayCounters &rc) {
   const int md = sc.maxDepth;
   Spec t = sConst(1), l = sConst(0);
   bool spec = true;
   rc.cam++;
   Hit hit = sceneIntersect(sc, ray);
   for (int depth = 0;; ++depth) {
      if (!hit.valid) {
         if (spec) {  // :44
            Spec s = sConst(0);
            for (const blingcu_light &lt : sc.lights) s = s + lightLe(sc, lt, ray);
            return l + t * s;
         }
         return l;  // :47
      }
      if (depth == md) return l;  // :51
      float lNumU = rnd1D(smp, 1 + 4 * depth);
      float lDir1, lDir2; rnd2D(smp, 1 + 3 * depth, lDir1, lDir2);
      float bCompU = rnd1D(smp, 2 + 4 * depth);
      float bDir1, bDir2; rnd2D(smp, 2 + 3 * depth, bDir1, bDir2);
      V3 rd = ray.d;
      Spec intl = spec ? intLe(sc, hit, rd) : sConst(0);  // Q1: ray direction, not wo
      V3 wo = -rd;
      Bsdf bsdf = makeBsdf(sc, hit);
      V3 n = bsdf.cs.n, p = bsdf.p;
      float eps = hit.eps;
      // sampleOneLight (Scene.hs:110-118) / estimateDirect (:84-99)
      Spec direct = sConst(0);
      int lc = (int)sc.lights.size();
      if (lc > 0) {
         int ln = (lc == 1) ? 0 : std::min((int)std::floor(lNumU * (float)lc), lc - 1);
         const blingcu_light &lt = sc.lights[ln];
         Spec ls = sampleLightMis(sc, lightSample(sc, lt, p, eps, n, lDir1, lDir2), bsdf, wo, rc);
         Spec bs = sampleBsdfMis(sc, ln, sampleBsdf(bsdf, wo, bCompU, bDir1, bDir2), p, eps, rc);
         direct = ls + bs;
         if (lc > 1) direct = sScale(direct, (float)lc);
      }
      Spec lHere = intl + direct;
      l = l + t * lHere;
      float pc = (depth <= 7) ? 1.0f : hmin(0.75f, sY(sc.T, t));
      float x = rnd1D(smp, 3 + 4 * depth);
      if (x > pc) return l;
      float uc = rnd1D(smp, 0 + 4 * depth);
      float ud1, ud2; rnd2D(smp, 0 + 3 * depth, ud1, ud2);
      BsdfSample s = sampleBsdf(bsdf, wo, uc, ud1, ud2);
      if (s.pdf == 0 || isBlack(s.f)) return l;
      ray = Ray{p, s.wi, eps, kInf};
      spec = (s.type & BX_SPECULAR) != 0;
      t = sScale(s.f * t, 1 / pc);
      rc.ext++;
      hit = sceneIntersect(sc, ray);
   }
}

// ----------------------------------------------------------------------------- Integrator/DirectLighting.hs:23-58
static Spec directLighting(const Scene &sc, const SampleCtx &smp, int d, const Ray &r, RayCounters &rc);
static Spec dlCont(const Scene &sc, const SampleCtx &smp, float e, int d, const Bsdf &bsdf, V3 wo, int t, RayCounters &rc) {   // :47-58
   if (d == sc.maxDepth) return sConst(0);
   BsdfSample bs = sampleBsdfFlags(bsdf, t, wo, 0.5f, 0.5f, 0.5f);
   if (bs.pdf == 0) return sConst(0);
   Ray ray{bsdf.p, bs.wi, e, kInf};
   rc.ext++;
   Spec l = directLighting(sc, smp, d, ray, rc);
   return bs.f * l;
}
static Spec directLighting(const Scene &sc, const SampleCtx &smp, int d, const Ray &r, RayCounters &rc) {   // :25-45
   Hit hit = sceneIntersect(sc, r);
   if (!hit.valid) return sConst(0);   // maybe (return black)
   float uln = rnd1D(smp, 0 + 2 * d);
   float uld1, uld2; rnd2D(smp, 0 + 2 * d, uld1, uld2);
   float ubc = rnd1D(smp, 1 + 2 * d);
   float ubd1, ubd2; rnd2D(smp, 1 + 2 * d, ubd1, ubd2);
   Bsdf bsdf = makeBsdf(sc, hit);
   V3 p = bsdf.p, n = bsdf.cs.n, wo = -r.d;
   float e = hit.eps;
   Spec l = sConst(0);   // sampleOneLight (Scene.hs:110-118)
   int lc = (int)sc.lights.size();
   if (lc > 0) {
      int ln = (lc == 1) ? 0 : std::min((int)std::floor(uln * (float)lc), lc - 1);
      const blingcu_light &lt = sc.lights[ln];
      Spec ls = sampleLightMis(sc, lightSample(sc, lt, p, e, n, uld1, uld2), bsdf, wo, rc);
      Spec bs = sampleBsdfMis(sc, ln, sampleBsdf(bsdf, wo, ubc, ubd1, ubd2), p, e, rc);
      l = ls + bs;
      if (lc > 1) l = sScale(l, (float)lc);
   }
   Spec re = dlCont(sc, smp, e, d + 1, bsdf, wo, BX_SPECULAR | BX_REFLECTION, rc);
   Spec tr = dlCont(sc, smp, e, d + 1, bsdf, wo, BX_SPECULAR | BX_TRANSMISSION, rc);
   return ((l + re) + tr) + intLe(sc, hit, wo);
}

// ----------------------------------------------------------------------------- film (Image.hs)
struct Window { int x0, x1, y0, y1; };  // inclusive (Sampling.hs:36-41)
static Window sampleExtent(const Scene &sc) {  // Image.hs:162-168
   return Window{(int)std::floor(0.5f - sc.fw), (int)std::floor(0.5f + (float)sc.W + sc.fw),
                 (int)std::floor(0.5f - sc.fh), (int)std::floor(0.5f + (float)sc.H + sc.fh)};
}
struct TileImage { int w, h, ox, oy; std::vector<float> px; };
static TileImage mkImageTile(const Scene &sc, const Window &wnd) {  // Image.hs:108-120
   TileImage t;
   t.ox = std::max(0, wnd.x0); t.oy = std::max(0, wnd.y0);
   t.w = wnd.x1 - t.ox + (int)std::floor(0.5f + sc.fw);
   t.h = wnd.y1 - t.oy + (int)std::floor(0.5f + sc.fh);
   t.px.assign((size_t)std::max(0, t.w) * std::max(0, t.h) * 4, 0.0f);
   return t;
}
// Image.hs:250-299
static bool addSample(const Scene &sc, TileImage &img, float sx, float sy, const Spec &ss) {
   if (sNaN(ss) || sInfinite(ss)) return false;
   float smx, smy, smz; spectrumToXYZ(sc.T, ss, smx, smy, smz);
   float fw = sc.fw, fh = sc.fh;
   float ifw = 1 / fw, ifh = 1 / fw;  // Q10: ifh = 1 / fw
   float dx = sx - 0.5f, dy = sy - 0.5f;
   int x0 = std::max(img.ox, (int)std::ceil(dx - fw)), x1 = std::min(img.ox + img.w - 1, (int)std::floor(dx + fw));
   int y0 = std::max(img.oy, (int)std::ceil(dy - fh)), y1 = std::min(img.oy + img.h - 1, (int)std::floor(dy + fh));
   if ((x1 - x0) < 0 || (y1 - y0) < 0) return true;
   for (int y = y0; y <= y1; ++y) {
      float fy = std::fabs(((float)y - dy) * ifh * 16.0f);
      int iy = std::min((int)std::floor(fy), 15);
      for (int x = x0; x <= x1; ++x) {
         float fx = std::fabs(((float)x - dx) * ifw * 16.0f);
         int ix = std::min((int)std::floor(fx), 15);
         float w = sc.ftbl[iy * 16 + ix];
         float *p = &img.px[4 * ((size_t)(x - img.ox) + (size_t)(y - img.oy) * img.w)];
         p[0] = p[0] + w;
         p[1] = p[1] + smx * w;
         p[2] = p[2] + smy * w;
         p[3] = p[3] + smz * w;
      }
   }
   return true;
}
static void addTile(Scene &sc, const TileImage &t) {  // Image.hs:178-199
   for (int y = 0; y < t.h; ++y) for (int x = 0; x < t.w; ++x) {
      int fx = x + t.ox, fy = y + t.oy;
      if (fy >= sc.H || fx >= sc.W) continue;
      float *d = &sc.film[4 * ((size_t)fy * sc.W + fx)];
      const float *s = &t.px[4 * ((size_t)y * t.w + x)];
      for (int o = 0; o < 4; ++o) d[o] = d[o] + s[o];
   }
}

static SampleCtx mkSampleCtx(const Scene &sc, const Window &ext, uint64_t seed, uint32_t pass, int ix, int iy, uint32_t s) {
   uint32_t ew = (uint32_t)(ext.x1 - ext.x0 + 1);
   uint32_t pix = (uint32_t)(iy - ext.y0) * ew + (uint32_t)(ix - ext.x0);
   SampleCtx c;
   c.kp = pixelKey(seed, pass, pix); c.s = s; c.nu = sc.nu; c.nv = sc.nv;
   if (sc.integrator == BLINGCU_INTEGRATOR_DIRECT) { c.n1d = 2 * sc.maxDepth; c.n2d = 2 * sc.maxDepth; }   // DirectLighting.hs:18-19
   else if (sc.integrator == BLINGCU_INTEGRATOR_BIDIR) { c.n1d = 4 * sc.sampleDepth * 3 + 1; c.n2d = 3 * sc.sampleDepth * 3 + 2; }   // BidirPath.hs:44-48
   else if (sc.integrator == BLINGCU_INTEGRATOR_NORMALS) { c.n1d = 0; c.n2d = 0; }                            // Debug.hs:24
   else { c.n1d = 4 * sc.sampleDepth; c.n2d = 3 * sc.sampleDepth; }
   c.stratified = sc.samplerKind == BLINGCU_SAMPLER_STRATIFIED;
   return c;
}
static Spec bidirLi(const Scene &sc, const SampleCtx &smp, const Ray &r, RayCounters &rc);
// one iteration of the `tile` body (Rendering.hs:142-150): fireRay >>= surfaceLi, then the camera sample position
static Spec renderSample(const Scene &sc, const Window &ext, uint64_t seed, uint32_t pass, int ix, int iy, uint32_t s,
                         float &sx, float &sy, RayCounters &rc) {
   SampleCtx c = mkSampleCtx(sc, ext, seed, pass, ix, iy, s);
   float ox, oy, lu, lv; cameraSample(c, ox, oy, lu, lv);
   sx = (float)ix + ox; sy = (float)iy + oy;
   Ray r = fireRay(sc, sx, sy, lu, lv);
   if (sc.integrator == BLINGCU_INTEGRATOR_DIRECT) { rc.cam++; return directLighting(sc, c, 0, r, rc); }
   if (sc.integrator == BLINGCU_INTEGRATOR_BIDIR) return bidirLi(sc, c, r, rc);
   if (sc.integrator == BLINGCU_INTEGRATOR_NORMALS) {   // mkNormalMap (Integrator/Debug.hs:23-33)
      rc.cam++;
      Hit hit = sceneIntersect(sc, r);
      if (!hit.valid) return sConst(0);
      V3 n = makeBsdf(sc, hit).cs.n;                    // bsdfShadingNormal
      V3 v = mk((1.0f + n.x) / 2, (1.0f + n.y) / 2, (1.0f + n.z) / 2);   // (vpromote 1 + n) / 2
      return rgbToSpectrum(sc.refl, v.x, v.y, v.z);     // rgbToSpectrumRefl
   }
   return pathLi(sc, c, r, rc);
}


// ----------------------------------------------------------------------------- Renderer/LightTracer.hs
struct LightRay { Spec li; Ray ray; V3 nl; float pdf; };
// boundingSphere (AABB.hs:62-66): centroid and the distance from it to pmax
static void boundingSphere(const AABB &b, V3 &c, float &r) { c = scl(0.5f, b.lo + b.hi); r = len(b.hi - c); }
// cosineSampleHemisphere' (Montecarlo.hs:152-161)
static V3 cosineSampleHemisphereAround(V3 n, float u1, float u2) { return localToWorld(coordinateSystem(n), cosineSampleHemisphere(u1, u2)); }
// Light.sample' (Light.hs:166-213): an outgoing ray of one light
static LightRay lightSampleRay(const Scene &sc, const blingcu_light &l, float uo1, float uo2, float ud1, float ud2) {
   LightRay empty{sConst(0), Ray{mk(0, 0, 0), mk(0, 1, 0), 0, 0}, mk(0, 1, 0), 0};
   const AABB &bounds = sc.geo.bounds;
   switch (l.kind) {
   case BLINGCU_LIGHT_AREA: {   // :174-180
      const blingcu_shape &sh = sc.geo.shapes[l.shape];
      V3 orgL, nsL; sampleShapeAny(sh, uo1, uo2, orgL, nsL);
      V3 org = transPoint(sh.o2w, orgL), ns = normalize(transNormalInv(sh.w2o, nsL));
      V3 wi = cosineSampleHemisphereAround(ns, ud1, ud2);
      float pd = kInvPi * (1 / shapeArea(sh)) * absDot(ns, wi);   // invPi * shapePdf' s org * ns `absDot` wi
      return LightRay{fromC(l.s), Ray{org, wi, 1e-3f, kInf}, ns, pd};
   }
   case BLINGCU_LIGHT_DIRECTIONAL: {   // :182-189
      V3 n = mk(l.v[0], l.v[1], l.v[2]);
      V3 wc; float wr; boundingSphere(bounds, wc, wr);
      Frame f = coordinateSystem(n);   // coordinateSystem'' n = (du, dv)
      float d1, d2; concentricSampleDisk(uo1, uo2, d1, d2);
      V3 pdisk = wc + scl(wr, scl(d1, f.s) + scl(d2, f.t));
      V3 ns = -n;
      return LightRay{fromC(l.s), Ray{pdisk + scl(wr, n), ns, 0, kInf}, ns, 1 / (kPi * wr * wr)};
   }
   case BLINGCU_LIGHT_INFINITE: {   // :191-208
      const blingcu_envmap &e = sc.envs[l.env];
      float u, v, pdMap; sampleContinuous2D(e, ud1, ud2, u, v, pdMap);
      if (pdMap == 0) return empty;
      Spec ls = envEval(sc.T, e, u, v);
      float phi = u * 2 * kPi, theta = v * kPi;
      V3 d = transVector(e.l2w, sphericalDirection(std::sin(theta), std::cos(theta), phi));
      V3 wc; float wr; boundingSphere(bounds, wc, wr);
      Frame f = coordinateSystem(-d);
      float d1, d2; concentricSampleDisk(uo1, uo2, d1, d2);
      V3 pDisk = wc + scl(wr, scl(d1, f.s) + scl(d2, f.t));
      float sint = std::sin(theta);
      float pdDir = pdMap / (2 * kPi * kPi * sint), pdArea = 1 / (kPi * wr * wr);
      float pd = (sint == 0) ? 0 : pdDir * pdArea;
      return LightRay{ls, Ray{pDisk + scl(wr, d), -d, 0, kInf}, d, pd};
   }
   default: {   // PointLight :210-213 (Q5: uniformSpherePdf = 1 / (2 pi))
      V3 d = uniformSampleSphere(ud1, ud2);
      return LightRay{fromC(l.s), Ray{mk(l.v[0], l.v[1], l.v[2]), d, 0, kInf}, d, 1 / (2 * kPi)};
   }
   }
}
// sampleLightRay (Scene.hs:121-136)
static LightRay sampleLightRay(const Scene &sc, float uL, float uo1, float uo2, float ud1, float ud2) {
   int lc = (int)sc.lights.size();
   if (lc == 0) return LightRay{sConst(0), Ray{mk(0, 0, 0), mk(0, 1, 0), 0, 0}, mk(0, 1, 0), 0};
   if (lc == 1) return lightSampleRay(sc, sc.lights[0], uo1, uo2, ud1, ud2);
   int ln = std::min((int)std::floor(uL * (float)lc), lc - 1);
   LightRay r = lightSampleRay(sc, sc.lights[ln], uo1, uo2, ud1, ud2);
   r.pdf = r.pdf / (float)lc;
   return r;
}
// ----------------------------------------------------------------------------- Integrator/BidirPath.hs
// sampleOneLight (Scene.hs:110-118) as estimateDirect of BidirPath.hs:151-164 calls it (the path integrator has it inline)
static Spec sampleOneLight(const Scene &sc, V3 p, float eps, V3 n, V3 wo, const Bsdf &bsdf, float lNumU, float lDir1, float lDir2,
                           float bCompU, float bDir1, float bDir2, RayCounters &rc) {
   int lc = (int)sc.lights.size();
   if (lc == 0) return sConst(0);
   int ln = (lc == 1) ? 0 : std::min((int)std::floor(lNumU * (float)lc), lc - 1);
   const blingcu_light &lt = sc.lights[ln];
   Spec ls = sampleLightMis(sc, lightSample(sc, lt, p, eps, n, lDir1, lDir2), bsdf, wo, rc);
   Spec bs = sampleBsdfMis(sc, ln, sampleBsdf(bsdf, wo, bCompU, bDir1, bDir2), p, eps, rc);
   Spec direct = ls + bs;
   if (lc > 1) direct = sScale(direct, (float)lc);
   return direct;
}
struct BdVertex { V3 wi, wo; Hit hit; int type; Spec alpha; };   // data Vertex (:20-26)
// nextVertex (:184-214). f1d d = base1 + smps1D * d * 3, f2d d = base2 + smps2D * d * 3 (smps1D = 4, smps2D = 3, :30-34). rrProb = 1
// (:203), so the roulette never ends a path; the hit behind the LAST vertex (depth + 1 == md) decides nothing and is not traced.
static std::vector<BdVertex> bdNextVertex(const Scene &sc, bool adj, V3 wi, Hit hit, Spec alpha, int base1, int base2,
                                          const SampleCtx &smp, RayCounters &rc) {
   std::vector<BdVertex> path;
   for (int depth = 0;; ++depth) {
      if (!hit.valid) break;                 // :186
      if (depth == sc.maxDepth) break;       // :188
      float ubc = rnd1D(smp, base1 + 12 * depth);
      float ub1, ub2; rnd2D(smp, base2 + 9 * depth, ub1, ub2);
      float rr = rnd1D(smp, 1 + base1 + 12 * depth);
      Bsdf bsdf = makeBsdf(sc, hit);
      BsdfSample bs = adj ? sampleAdjBsdf(bsdf, wi, ubc, ub1, ub2) : sampleBsdf(bsdf, wi, ubc, ub1, ub2);
      path.push_back(BdVertex{wi, bs.wi, hit, bs.type, alpha});
      if (isBlack(bs.f) || bs.pdf == 0) break;   // :209-211
      const float rrProb = 1;
      if (rr > rrProb) break;
      alpha = sScale(bs.f * alpha, 1 / rrProb);
      if (depth + 1 == sc.maxDepth) break;
      rc.ext++;
      hit = sceneIntersect(sc, Ray{bsdf.p, bs.wi, hit.eps, kInf});
      wi = -bs.wi;
   }
   return path;
}
// contrib False md (:50-104). The reference's `connect` (:124-149) is called with the LIGHT vertex first and binds the second
// field of each vertex (the SAMPLED direction _vwo) where its names say wi: restated as written, not as meant.
static Spec bidirLi(const Scene &sc, const SampleCtx &smp, const Ray &r, RayCounters &rc) {
   float ul = rnd1D(smp, 0), ulo1, ulo2, uld1, uld2; rnd2D(smp, 0, ulo1, ulo2); rnd2D(smp, 1, uld1, uld2);
   // lightPath (:175-182)
   LightRay lr = sampleLightRay(sc, ul, ulo1, ulo2, uld1, uld2);
   V3 lwo = -lr.ray.d;
   Spec li = sScale(lr.li, absDot(lr.nl, lwo) / lr.pdf);
   rc.ext++;
   std::vector<BdVertex> lp = bdNextVertex(sc, true, lwo, sceneIntersect(sc, lr.ray), li, 3 + 1, 3 + 2, smp, rc);
   // eyePath (:168-172)
   rc.cam++;
   std::vector<BdVertex> ep = bdNextVertex(sc, false, -r.d, sceneIntersect(sc, r), sConst(1), 2 + 1, 2 + 2, smp, rc);
   // countSpec (:111-122)
   std::vector<float> nspec(ep.size() + lp.size() + 2, 0.0f);
   for (size_t i = 0; i < ep.size(); ++i)
      for (size_t j = 0; j < lp.size(); ++j)
         if ((ep[i].type & BX_SPECULAR) || (lp[j].type & BX_SPECULAR)) nspec[i + j + 2] += 1;
   // S1 subpaths (:69-72): estimateDirect (:151-164) at every eye vertex
   Spec ld = sConst(0);
   for (size_t i = 0; i < ep.size(); ++i) {
      const BdVertex &v = ep[i]; const int depth = (int)i;
      float lNumU = rnd1D(smp, 0 + 1 + 12 * depth);
      float lDir1, lDir2; rnd2D(smp, 0 + 2 + 9 * depth, lDir1, lDir2);
      float bCompU = rnd1D(smp, 1 + 1 + 12 * depth);
      float bDir1, bDir2; rnd2D(smp, 1 + 2 + 9 * depth, bDir1, bDir2);
      Bsdf bsdf = makeBsdf(sc, v.hit);
      Spec lHere = sampleOneLight(sc, bsdf.p, v.hit.eps, bsdf.cs.n, v.wi, bsdf, lNumU, lDir1, lDir2, bCompU, bDir1, bDir2, rc);
      Spec d = lHere * v.alpha;
      ld = ld + sScale(d, 1 / (1 + (float)i - nspec[i + 1]));
   }
   // S0 subpaths (:79-83): emitters seen directly or through specular bounces; intLe is asked for the SAMPLED direction _vwo
   Spec le = sConst(0);
   for (size_t i = 0; i < ep.size(); ++i) {
      bool prevSpec = (i == 0) ? true : (ep[i - 1].type & BX_SPECULAR) != 0;
      if (prevSpec) le = le + ep[i].alpha * intLe(sc, ep[i].hit, ep[i].wo);
   }
   if (ep.empty() || lp.empty()) return ld + le;
   // connections (:91-93, :124-149)
   Spec l = sConst(0);
   for (size_t s_ = 0; s_ < lp.size(); ++s_) {
      Spec row = sConst(0);
      for (size_t t_ = 0; t_ < ep.size(); ++t_) {
         const BdVertex &a = lp[s_], &b = ep[t_];   // a plays the pattern's "eye vertex" (i = s), b its "light vertex" (j = t)
         Spec c = sConst(0);
         if (!(a.type & BX_SPECULAR) && !(b.type & BX_SPECULAR)) {
            Bsdf bsdfe = makeBsdf(sc, a.hit), bsdfl = makeBsdf(sc, b.hit);
            V3 pe = bsdfe.p, pl = bsdfl.p;
            V3 d = pl - pe;
            float wl = 0; V3 w = d;                           // normLen (Math.hs:356-363)
            if (sqLen(d) != 0) { wl = len(d); w = scl(1 / wl, d); }
            float g = 1 / sqLen(d);
            Spec fe = evalBsdf(bsdfe, a.wo, w);
            Spec fl = evalAdjBsdf(bsdfl, b.wo, -w);
            if (!(isBlack(fe) || isBlack(fl))) {
               rc.shadow++;
               if (!sceneOccluded(sc, Ray{pe, w, a.hit.eps, wl - b.hit.eps})) {
                  float pathWt = 1 / ((float)(s_ + t_ + 2) - nspec[s_ + t_ + 2]);
                  c = sScale(a.alpha * fe * b.alpha * fl, g * pathWt);
               }
            }
         }
         row = row + c;
      }
      l = l + row;
   }
   return ld + le + l;
}

// sampleCam (Camera.hs:78-103), projective cameras only (the reference `error`s otherwise)
struct CamSample { V3 pLens; float px, py, pdf; };
static CamSample sampleCam(const Scene &sc, V3 p) {
   const blingcu_camera &c = sc.cam;
   V3 pRas = transPoint(c.world2raster, p);
   V3 pLens = transPoint(c.cam2world, mk(0, 0, 0));
   float cost = std::fabs(normalize(transPoint(c.raster2cam, pRas)).z);
   return CamSample{pLens, pRas.x, pRas.y, c.pixel_area * (cost * cost * cost)};
}
struct SplatRec { uint32_t photon, depth; float px, py, X, Y, Z; };
// splatSample (Image.hs:201-221): floor to the pixel, drop NaN / infinite spectra and positions outside the image
static bool splatPixel(const Scene &sc, float sx, float sy, const Spec &ss, int &px, int &py) {
   px = (int)std::floor(sx); py = (int)std::floor(sy);
   if (px >= sc.W || py >= sc.H || px < 0 || py < 0) return false;
   for (int i = 0; i < NB; ++i) if (std::isnan(ss.v[i]) || std::isinf(ss.v[i])) return false;
   return true;
}
// oneRay + nextVertex + connectCam (LightTracer.hs:53-108). Random numbers: the counter-based stream of photon `index` of the pass
// (oracle_math.h SPEC, plain uniforms): 1-D dims 0 = light choice, 1 + 2 d = BSDF component, 2 + 2 d = Russian roulette;
// 2-D dims 0 = position on the light, 1 = direction, 2 + d = BSDF direction.
static void lightPath(const Scene &sc, uint64_t seed, uint32_t pass, uint32_t index, std::vector<SplatRec> &out, uint64_t &nLight, uint64_t &nConnect) {
   SampleCtx c; c.kp = pixelKey(seed, pass, index); c.s = 0; c.nu = 1; c.nv = 1; c.n1d = 0; c.n2d = 0; c.stratified = false;
   float ul = rnd1D(c, 0), uo1, uo2, ud1, ud2; rnd2D(c, 0, uo1, uo2); rnd2D(c, 1, ud1, ud2);
   LightRay lr = sampleLightRay(sc, ul, uo1, uo2, ud1, ud2);
   if (!(lr.pdf > 0)) return;
   V3 wo = normalize(lr.ray.d);
   Spec li = sScale(lr.li, absDot(lr.nl, wo) / lr.pdf);
   if (isBlack(li)) return;
   V3 wi = -wo;
   nLight++;
   Hit hit = sceneIntersect(sc, lr.ray);
   for (int depth = 0;; ++depth) {
      if (!hit.valid) return;
      if (isBlack(li)) return;
      float ubc = rnd1D(c, 1 + 2 * depth), ub1, ub2; rnd2D(c, 2 + depth, ub1, ub2);
      float pcont = (depth > 3) ? 0.8f : 1.0f;
      Bsdf bsdf = makeBsdf(sc, hit);
      V3 p = bsdf.p;
      BsdfSample bs = sampleAdjBsdf(bsdf, wi, ubc, ub1, ub2);
      Spec liNext = sScale(li * bs.f, 1 / pcont);
      // connectCam (:62-76)
      {
         CamSample cs = sampleCam(sc, p);
         V3 dCam = cs.pLens - p;
         V3 we = normalize(dCam);
         Spec f = evalAdjBsdf(bsdf, wi, we);
         float dCam2 = sqLen(dCam);
         if (!(isBlack(f) || cs.pdf == 0)) {
            nConnect++;
            if (!sceneOccluded(sc, Ray{p, we, hit.eps, std::sqrt(dCam2)})) {
               Spec lrS = sScale(li * f, 1 / (cs.pdf * dCam2));
               int px, py;
               if (splatPixel(sc, cs.px, cs.py, lrS, px, py)) {
                  float X, Y, Z; spectrumToXYZ(sc.T, lrS, X, Y, Z);
                  out.push_back(SplatRec{index, (uint32_t)depth, cs.px, cs.py, X, Y, Z});
               }
            }
         }
      }
      if (isBlack(bs.f) || bs.pdf == 0) return;
      float x = rnd1D(c, 2 + 2 * depth);
      if (x > pcont) return;
      nLight++;
      hit = sceneIntersect(sc, Ray{p, bs.wi, hit.eps, kInf});
      wi = -bs.wi;
      li = liNext;
   }
}
}  // namespace orc

// =============================================================================================
// C API
// =============================================================================================
using namespace orc;
struct oracle_ctx { Scene sc; std::string err; };

extern "C" {

int oracle_create(const blingcu_scene *ir, int build_kdtree, oracle_ctx **out) {
   oracle_ctx *c = new oracle_ctx();
   Scene &sc = c->sc;
   size_t nt = (size_t)ir->n_triangles;
   sc.geo.tris.resize(nt);
   sc.triMaterial.resize(nt);
   size_t nprims = nt + ir->n_shapes;
   sc.geo.prims.resize(nprims);
   std::vector<char> seen(nprims, 0);
   for (size_t i = 0; i < nt; ++i) {
      const float *v = ir->tri_verts + 9 * i;
      Tri &T = sc.geo.tris[i];
      T.p1 = mk(v[0], v[1], v[2]); T.p2 = mk(v[3], v[4], v[5]); T.p3 = mk(v[6], v[7], v[8]);
      for (int k = 0; k < 6; ++k) T.uv[k] = ir->tri_uvs[6 * i + k];
      sc.triMaterial[i] = ir->tri_material[i];
      size_t pid = ir->tri_prim_id ? (size_t)ir->tri_prim_id[i] : (size_t)ir->tri_prim_id_base + i;
      if (pid >= nprims || seen[pid]) { delete c; return BLINGCU_EINVAL; }
      seen[pid] = 1;
      AABB wb = extendP(extendP(extendP(emptyBox(), T.p1), T.p2), T.p3);  // TriangleMesh.hs:136-138
      sc.geo.prims[pid] = Prim{true, (uint32_t)i, wb};
   }
   if (ir->tri_normals) sc.triNormals.assign(ir->tri_normals, ir->tri_normals + 9 * nt);
   sc.geo.shapes.assign(ir->shapes, ir->shapes + ir->n_shapes);
   for (uint32_t i = 0; i < ir->n_shapes; ++i) {
      size_t pid = (size_t)ir->shapes[i].prim_id;
      if (pid >= nprims || seen[pid]) { delete c; return BLINGCU_EINVAL; }
      seen[pid] = 1;
      sc.geo.prims[pid] = Prim{false, i, transBox(ir->shapes[i].o2w, shapeObjectBounds(ir->shapes[i]))};  // Shape.hs:287-294
   }
   sc.materials.assign(ir->materials, ir->materials + ir->n_materials);
   sc.textures.assign(ir->textures, ir->textures + ir->n_textures);
   for (uint32_t i = 0; i < ir->n_images; ++i) {   // own the pixel data
      blingcu_image im = ir->images[i];
      sc.imageData.emplace_back(im.data, im.data + (size_t)im.width * im.height * im.channels);
      sc.images.push_back(im);
   }
   for (size_t i = 0; i < sc.images.size(); ++i) sc.images[i].data = sc.imageData[i].data();
   for (int b = 0; b < 7; ++b) sc.refl[b] = fromC(ir->refl_basis[b]);
   sc.lights.assign(ir->lights, ir->lights + ir->n_lights);
   sc.envs.assign(ir->envs, ir->envs + ir->n_envs);
   for (blingcu_envmap &e : sc.envs) {  // own the arrays
      auto own = [&](const float *&p, size_t n) {
         if (!p) return;
         sc.envData.emplace_back(p, p + n);
         p = sc.envData.back().data();
      };
      size_t nu = (size_t)e.nu, nv = (size_t)e.nv;
      if (e.kind == BLINGCU_ENV_RGBTABLE) own(e.rgb, nu * nv * 3);
      own(e.cond_func, nu * nv); own(e.cond_cdf, (nu + 1) * nv); own(e.cond_int, nv);
      own(e.marg_func, nv); own(e.marg_cdf, nv + 1);
   }
   sc.cam = ir->camera;
   sc.T.cieX = fromC(ir->cie_x); sc.T.cieY = fromC(ir->cie_y); sc.T.cieZ = fromC(ir->cie_z); sc.T.ySum = ir->cie_y_sum;
   for (int i = 0; i < 7; ++i) sc.T.illum[i] = fromC(ir->illum_basis[i]);
   sc.W = ir->width; sc.H = ir->height; sc.fw = ir->filter_w; sc.fh = ir->filter_h;
   std::memcpy(sc.ftbl, ir->filter_table, sizeof(sc.ftbl));
   sc.samplerKind = ir->sampler_kind; sc.nu = ir->nu; sc.nv = ir->nv;
   sc.maxDepth = ir->max_depth; sc.sampleDepth = ir->sample_depth; sc.integrator = ir->integrator_kind;
   sc.film.assign((size_t)sc.W * sc.H * 4, 0.0f);
   sc.splat.assign((size_t)sc.W * sc.H * 3, 0.0f);
   sc.useKd = build_kdtree != 0;
   if (!sc.useKd) { sc.geo.bounds = emptyBox(); for (const Prim &p : sc.geo.prims) sc.geo.bounds = extendB(sc.geo.bounds, p.wb); }   // worldBounds of the scene
   if (sc.useKd) sc.geo.buildKd();
   *out = c;
   return 0;
}
void oracle_destroy(oracle_ctx *c) { delete c; }

static Ray toRay(const blingcu_ray &r) { return Ray{mk(r.o[0], r.o[1], r.o[2]), mk(r.d[0], r.d[1], r.d[2]), r.tmin, r.tmax}; }

// mode 0 = brute force (ground truth), 1 = kd-tree
int oracle_trace_nearest(oracle_ctx *c, const blingcu_ray *rays, size_t n, blingcu_hit *out, int mode,
                         uint64_t *nodes_traversed, uint64_t *intersections) {
   if (mode == 1 && !c->sc.geo.kd_built) return BLINGCU_ESTATE;
   uint64_t nt = 0, ni = 0;
   for (size_t i = 0; i < n; ++i) {
      Ray r = toRay(rays[i]);
      Hit h = (mode == 1) ? c->sc.geo.kdNearest(r, &nt, &ni) : c->sc.geo.bruteNearest(r);
      out[i].t = h.valid ? h.t : 0; out[i].prim = h.valid ? h.prim : -1;
      out[i].b1 = (h.valid && h.dg.tri) ? h.dg.b1 : 0; out[i].b2 = (h.valid && h.dg.tri) ? h.dg.b2 : 0;
   }
   if (nodes_traversed) *nodes_traversed = nt;
   if (intersections) *intersections = ni;
   return 0;
}
int oracle_debug_eval(int what, const float *in, float *out) {
   auto spec = [](const float *p) { Spec r; for (int i = 0; i < NB; ++i) r.v[i] = p[i]; return r; };
   auto put = [](float *o, const Spec &sp) { for (int i = 0; i < NB; ++i) o[i] = sp.v[i]; };
   auto v3 = [](const float *p) { return mk(p[0], p[1], p[2]); };
   auto mfBx = [&](float e, const Spec &r) { BxDF b{}; b.kind = K_MICROFACET; b.type = BX_REFLECTION | BX_GLOSSY; b.r = r; b.fr = FR_DIELECTRIC; b.etai = 1; b.etat = 1.5f; b.e = e; b.flip = false; return b; };
   switch (what) {
   case 1: put(out, frDielectric(in[0], in[1], in[2])); return 0;
   case 2: put(out, frConductor(spec(in), spec(in + 16), in[32])); return 0;
   case 3: { V3 wh = v3(in + 1); out[0] = blinnD(in[0], wh); out[1] = blinnPdf(in[0], wh); out[2] = mfG(v3(in + 4), v3(in + 7), wh); return 0; }
   case 4: {
      BxDF b = mfBx(in[0], spec(in + 1));
      V3 wo = v3(in + 17), wi = v3(in + 20);
      put(out, bxdfEval(b, wo, wi)); out[16] = bxdfPdf(b, wo, wi);
      Spec f = sConst(0); V3 ws = mk(0, 0, 0); float pdf = 0;
      bxdfSample(b, wo, in[23], in[24], f, ws, pdf);
      put(out + 17, f); out[33] = ws.x; out[34] = ws.y; out[35] = ws.z; out[36] = pdf;
      return 0;
   }
   case 5: {
      Bsdf bs{}; bs.n = 2;
      bs.bx[0] = BxDF{}; bs.bx[0].kind = K_LAMBERT; bs.bx[0].type = BX_REFLECTION | BX_DIFFUSE; bs.bx[0].r = spec(in); bs.bx[0].flip = false;
      bs.bx[1] = mfBx(in[32], spec(in + 16));
      bs.cs.s = v3(in + 33); bs.cs.t = v3(in + 36); bs.cs.n = v3(in + 39); bs.ng = v3(in + 42); bs.p = mk(0, 0, 0);
      V3 woW = v3(in + 45), wiW = v3(in + 48);
      put(out, evalBsdf(bs, woW, wiW)); out[16] = bsdfPdf(bs, woW, wiW);
      BsdfSample smp = sampleBsdf(bs, woW, in[51], in[52], in[53]);
      out[17] = (float)smp.type; out[18] = smp.pdf; put(out + 19, smp.f); out[35] = smp.wi.x; out[36] = smp.wi.y; out[37] = smp.wi.z;
      return 0;
   }
   case 6: {
      blingcu_shape sh{}; sh.kind = (int)in[0]; for (int i = 0; i < 6; ++i) sh.p[i] = in[1 + i];
      for (int i = 0; i < 16; ++i) { sh.o2w[i] = (i % 5 == 0) ? 1.0f : 0.0f; sh.w2o[i] = sh.o2w[i]; }
      V3 pt = v3(in + 7), ps, ns;
      sampleShape(sh, pt, in[10], in[11], ps, ns);
      out[0] = ps.x; out[1] = ps.y; out[2] = ps.z; out[3] = ns.x; out[4] = ns.y; out[5] = ns.z;
      out[6] = shapePdf(sh, pt, v3(in + 12));
      return 0;
   }
   default: return BLINGCU_EINVAL;
   }
}
int oracle_export_kdtree(oracle_ctx *c, blingcu_kdnode *nodes, uint32_t *n_nodes, uint32_t *leaf_prims, size_t *n_leaf_prims, int32_t *root, float bounds[6]) {
   const Geometry &g = c->sc.geo;
   if (!g.kd_built) return BLINGCU_ESTATE;
   *n_nodes = (uint32_t)g.nodes.size(); *n_leaf_prims = g.leafPrims.size(); *root = g.root;
   bounds[0] = g.bounds.lo.x; bounds[1] = g.bounds.lo.y; bounds[2] = g.bounds.lo.z; bounds[3] = g.bounds.hi.x; bounds[4] = g.bounds.hi.y; bounds[5] = g.bounds.hi.z;
   if (nodes) for (size_t i = 0; i < g.nodes.size(); ++i) {
      const KdNode &n = g.nodes[i];
      nodes[i].left = n.left; nodes[i].right = n.right; nodes[i].split = n.sp; nodes[i].axis = n.axis; nodes[i].first = n.first; nodes[i].count = n.count;
   }
   if (leaf_prims) for (size_t i = 0; i < g.leafPrims.size(); ++i) leaf_prims[i] = g.leafPrims[i];
   return 0;
}
int oracle_trace_kd_stats(oracle_ctx *c, const blingcu_ray *rays, size_t n, blingcu_hit *out, uint32_t *nodes_traversed, uint32_t *intersections) {
   if (!c->sc.geo.kd_built) return BLINGCU_ESTATE;
   for (size_t i = 0; i < n; ++i) {
      uint64_t nt = 0, ni = 0;
      Hit h = c->sc.geo.kdNearest(toRay(rays[i]), &nt, &ni);
      out[i].t = h.valid ? h.t : 0; out[i].prim = h.valid ? h.prim : -1;
      out[i].b1 = (h.valid && h.dg.tri) ? h.dg.b1 : 0; out[i].b2 = (h.valid && h.dg.tri) ? h.dg.b2 : 0;
      nodes_traversed[i] = (uint32_t)nt; intersections[i] = (uint32_t)ni;
   }
   return 0;
}
int oracle_trace_occluded(oracle_ctx *c, const blingcu_ray *rays, size_t n, uint8_t *out, int mode) {
   if (mode == 1 && !c->sc.geo.kd_built) return BLINGCU_ESTATE;
   for (size_t i = 0; i < n; ++i) {
      Ray r = toRay(rays[i]);
      out[i] = (mode == 1) ? c->sc.geo.kdOccluded(r) : c->sc.geo.bruteOccluded(r);
   }
   return 0;
}

int oracle_sample_extent(oracle_ctx *c, int32_t *x0, int32_t *x1, int32_t *y0, int32_t *y1) {
   Window w = sampleExtent(c->sc);
   *x0 = w.x0; *x1 = w.x1; *y0 = w.y0; *y1 = w.y1;
   return 0;
}

int oracle_render_samples(oracle_ctx *c, uint32_t pass, uint64_t seed, const int32_t *px, const int32_t *py,
                          const uint32_t *sample, size_t n, float *out_L, float *out_xy) {
   Window ext = sampleExtent(c->sc);
   RayCounters rc;
   for (size_t i = 0; i < n; ++i) {
      float sx, sy;
      Spec L = renderSample(c->sc, ext, seed, pass, px[i], py[i], sample[i], sx, sy, rc);
      std::memcpy(out_L + 16 * i, L.v, sizeof(L.v));
      out_xy[2 * i] = sx; out_xy[2 * i + 1] = sy;
   }
   return 0;
}

// one pass slice with the reference's decomposition: 16x16 sample windows (Sampling.hs:55-58), one worker per
// tile (parBuffer numCapabilities, Rendering.hs:118), sequential addTile (:130-134).
int oracle_render_slice(oracle_ctx *c, uint32_t pass, uint64_t seed, uint32_t s_begin, uint32_t s_end, int nthreads) {
   Scene &sc = c->sc;
   Window ext = sampleExtent(sc);
   std::vector<Window> wnds;
   for (int y = ext.y0; y <= ext.y1; y += 16) for (int x = ext.x0; x <= ext.x1; x += 16)
      wnds.push_back(Window{x, std::min(x + 15, ext.x1), y, std::min(y + 15, ext.y1)});
   std::vector<TileImage> tiles(wnds.size());
   std::atomic<size_t> next{0};
   auto worker = [&]() {
      RayCounters rc; uint64_t ns = 0, nd = 0;
      for (;;) {
         size_t ti = next.fetch_add(1);
         if (ti >= wnds.size()) break;
         const Window &w = wnds[ti];
         TileImage img = mkImageTile(sc, w);
         for (int iy = w.y0; iy <= w.y1; ++iy) for (int ix = w.x0; ix <= w.x1; ++ix)
            for (uint32_t s = s_begin; s < s_end; ++s) {
               float sx, sy;
               Spec L = renderSample(sc, ext, seed, pass, ix, iy, s, sx, sy, rc);
               if (!addSample(sc, img, sx, sy, L)) nd++;
               ns++;
            }
         tiles[ti] = std::move(img);
      }
      sc.nSamples += ns; sc.dropped += nd;
      sc.rCam += rc.cam; sc.rExt += rc.ext; sc.rMis += rc.mis; sc.rShadow += rc.shadow;
   };
   if (nthreads <= 1) worker();
   else {
      std::vector<std::thread> th;
      for (int i = 0; i < nthreads; ++i) th.emplace_back(worker);
      for (auto &t : th) t.join();
   }
   for (const TileImage &t : tiles) addTile(sc, t);
   return 0;
}

int oracle_read_film(oracle_ctx *c, float *wxyz) { std::memcpy(wxyz, c->sc.film.data(), c->sc.film.size() * sizeof(float)); return 0; }
int oracle_clear_film(oracle_ctx *c) { std::fill(c->sc.film.begin(), c->sc.film.end(), 0.0f); std::fill(c->sc.splat.begin(), c->sc.splat.end(), 0.0f); return 0; }
int oracle_get_stats(oracle_ctx *c, blingcu_stats *s) {
   std::memset(s, 0, sizeof(*s));
   s->samples = c->sc.nSamples; s->rays_camera = c->sc.rCam; s->rays_extension = c->sc.rExt;
   s->rays_mis = c->sc.rMis; s->rays_shadow = c->sc.rShadow; s->dropped_samples = c->sc.dropped;
   s->bvh_nodes = c->sc.geo.nodes.size(); s->bvh_leaf_items = c->sc.geo.leafPrims.size();
   s->photons = c->sc.nPhotons; s->rays_light = c->sc.rLight; s->rays_connect = c->sc.rConnect; s->splats = c->sc.nSplats;
   return 0;
}
int oracle_reset_stats(oracle_ctx *c) {
   c->sc.nSamples = 0; c->sc.rCam = 0; c->sc.rExt = 0; c->sc.rMis = 0; c->sc.rShadow = 0; c->sc.dropped = 0;
   c->sc.nPhotons = 0; c->sc.rLight = 0; c->sc.rConnect = 0; c->sc.nSplats = 0;
   return 0;
}

// light tracer: photons [first, first + n) of pass `pass`; splats are added in photon order (deterministic); optional records
int oracle_light_trace(oracle_ctx *c, uint32_t pass, uint64_t seed, uint64_t first, uint32_t n, float *records, size_t maxRecords, size_t *nRecords) {
   Scene &sc = c->sc;
   if (sc.cam.kind != BLINGCU_CAM_PERSPECTIVE || sc.cam.pixel_area == 0) return BLINGCU_EINVAL;
   std::vector<SplatRec> recs;
   uint64_t nl = 0, nc = 0;
   for (uint32_t i = 0; i < n; ++i) lightPath(sc, seed, pass, (uint32_t)(first + i), recs, nl, nc);
   size_t k = 0;
   for (const SplatRec &r : recs) {
      int px = (int)std::floor(r.px), py = (int)std::floor(r.py);
      float *d = &sc.splat[3 * ((size_t)py * sc.W + px)];
      d[0] = d[0] + r.X; d[1] = d[1] + r.Y; d[2] = d[2] + r.Z;
      if (records && k < maxRecords) { float *o = records + 7 * k; o[0] = (float)r.photon; o[1] = (float)r.depth; o[2] = r.px; o[3] = r.py; o[4] = r.X; o[5] = r.Y; o[6] = r.Z; k++; }
   }
   if (nRecords) *nRecords = recs.size();
   sc.nPhotons += n; sc.rLight += nl; sc.rConnect += nc; sc.nSplats += recs.size();
   return 0;
}
int oracle_read_splat(oracle_ctx *c, float *xyz) { std::memcpy(xyz, c->sc.splat.data(), c->sc.splat.size() * sizeof(float)); return 0; }

int oracle_eval_texture(oracle_ctx *c, int32_t tex, const float *p, const float *uv, size_t n, float *out) {
   if (tex < 0 || (size_t)tex >= c->sc.textures.size()) return 1;
   for (size_t i = 0; i < n; ++i) {
      DG dg = mkDg(mk(p[3 * i], p[3 * i + 1], p[3 * i + 2]), uv[2 * i], uv[2 * i + 1], mk(1, 0, 0), mk(0, 1, 0));
      Spec s = sConst(0);
      if (c->sc.textures[tex].kind >= BLINGCU_STEX_CONSTANT) s.v[0] = evalScalarTexture(c->sc.texEnv(), tex, dg);
      else s = evalSpectrumTexture(c->sc, tex, dg);
      std::memcpy(out + 16 * i, s.v, sizeof(s.v));
   }
   return 0;
}

// unit-test hooks: BSDF sampling/evaluation of a material at a canonical frame, filter/film on a single tile
int oracle_add_sample_tile(oracle_ctx *c, int wx0, int wx1, int wy0, int wy1, float sx, float sy, const float *L16,
                           float *out_tile, int *ox, int *oy, int *w, int *h) {
   Window wnd{wx0, wx1, wy0, wy1};
   TileImage img = mkImageTile(c->sc, wnd);
   Spec s; std::memcpy(s.v, L16, sizeof(s.v));
   addSample(c->sc, img, sx, sy, s);
   *ox = img.ox; *oy = img.oy; *w = img.w; *h = img.h;
   if (out_tile) std::memcpy(out_tile, img.px.data(), img.px.size() * sizeof(float));
   return 0;
}

}  // extern "C"
