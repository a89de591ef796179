// shading.h -- BSDFs, materials, textures, lights and environment maps (kernel-body building blocks).
// Replaces: Material.hs:32-96, Texture.hs:159-207, Reflection.hs:209-332,
// Reflection/{Diffuse,Specular,Microfacet}.hs, Fresnel.hs:31-70, Light.hs:85-229, Shape.hs:314-409,
// Montecarlo.hs:34-205, SunSky.hs:12-94, Spectrum.hs:146-168,349-373. Quirks Q1-Q7, Q11 are kept as written.
//
// Design note: every leaf spectrum texture in scope is a constant, so a texture lookup resolves to a POINTER
// into the texture table instead of a 64-byte value; a BSDF is then a few scalars + pointers and the 16-band
// arithmetic streams through registers one spectrum at a time.
#pragma once
#include "bvh.h"

namespace bl {

struct DScene {
   Bvh bvh;
   // shading geometry of a triangle = ONE 64-byte record (two 32-byte loads; until the end of round 2: 3 x 16 bytes of vertices and
   // 3 x 8 bytes of uvs in two arrays = six requests per shaded vertex, in kernels that sit on their request rate):
   //   (p1.xyz, material) (p2.xyz, uv1.x) (p3.xyz, uv1.y) (uv2.x, uv2.y, uv3.x, uv3.y)
   const F4 *tri_p;            // BL_TRI_F4 per triangle
   const F4 *tri_n;            // vertex normals, 3 per triangle (n1.xyz, -)(n2.xyz, -)(n3.xyz, -): three 16-byte loads instead of nine scalar ones; or null
   const int32_t *tri_prim;    // per triangle: primitive id in mkScene's list (only the C-ABI hit conversion reads it)
   const blingcu_shape *shapes;
   const blingcu_material *materials;
   const blingcu_texture *textures;
   const blingcu_light *lights;
   const blingcu_envmap *envs; // pointers inside are device pointers
   const float *ftbl;          // 16x16 filter table
   int n_lights;
   int has_box;                // some shape is a Box: BSDF-MIS rays towards infinite lights keep the nearest-hit query (bodies.h)
   blingcu_camera cam;
   int W, H; float fw, fh;
   int ex0, ex1, ey0, ey1, EW, EH;   // sample extent (Image.hs:162-168)
   int sampler_kind, nu, nv, max_depth, sample_depth;
   int integrator;             // BLINGCU_INTEGRATOR_*
   SamplerConst smp;           // per-scene sampler constants (hd.h)
   SamplerConst smpUniform;    // the same SPEC with no stratified dimensions: light paths draw plain uniforms (lighttrace.h)
   float bounds_lo[3], bounds_hi[3];   // worldBounds of the scene's primitives (Light.sample' aims at their bounding sphere)
   float cieX[NB], cieY[NB], cieZ[NB], ySum;
   float illum[7][NB];         // r g b c m y w
   const blingcu_image *images;   // image textures (data pointers are device pointers)
   float refl[7][NB];          // reflectance basis r g b c m y w (rgbToSpectrumRefl), image textures only
   uint8_t perm[256];          // noisePerms (Texture.hs:400-414), filled by upload; read by textures.h only
};

}  // namespace bl
#include "textures.h"
namespace bl {

// ------------------------------------------------------------------------------------------ Montecarlo.hs
HD void concentricSampleDisk(float u1, float u2, float &dx, float &dy) {   // :164-181
   float sx = u1 * 2 - 1, sy = u2 * 2 - 1;
   if (sx == 0 && sy == 0) { dx = 0; dy = 0; return; }
   float r, thp;
   if (sx >= -sy) {
      if (sx > sy) { r = sx; thp = (sy > 0) ? sy / sx : 8 + sy / sx; }
      else { r = sy; thp = 2 - sx / sy; }
   } else if (sx <= sy) { r = -sx; thp = 4 - sy / (-sx); }
   else { r = -sy; thp = 6 + sx / (-sy); }
   float theta = thp * BL_PI / 4;
   dx = r * cosf(theta); dy = r * sinf(theta);
}
HD V3 cosineSampleHemisphere(float u1, float u2) { float x, y; concentricSampleDisk(u1, u2, x, y); return mk3(x, y, sqrtf(hmaxf(0, 1 - x * x - y * y))); }
HD float uniformConePdf(float c) { return (c >= 1) ? 0.0f : 1 / (BL_TWOPI * (1 - c)); }
HD V3 uniformSampleCone(const Frame &f, float cosThetaMax, float u1, float u2) {
   float cosTheta = lerpf(u1, cosThetaMax, 1.0f);
   float sinTheta = sqrtf(1 - cosTheta * cosTheta);
   float phi = u2 * BL_TWOPI;
   float a = cosf(phi) * sinTheta, b = sinf(phi) * sinTheta;
   return (f.s * mk3(a, a, a) + f.t * mk3(b, b, b)) + f.n * mk3(cosTheta, cosTheta, cosTheta);
}
HD V3 uniformSampleSphere(float u1, float u2) { float u = u1 * 2 - 1; float s = sqrtf(1 - (u * u)); float om = u2 * 2 * BL_PI; return mk3(s * cosf(om), s * sinf(om), u); }
HD float powerHeuristic(float fPdf, float gPdf) { return (fPdf * fPdf) / (fPdf * fPdf + gPdf * gPdf); }   // :113-116 with nf = ng = 1
HD void remapRand(int segs, float u, int &seg, float &up) { float sf = (float)segs; seg = imin(segs - 1, (int)floorf(u * sf)); up = (u - (float)seg / sf) * sf; }

// ------------------------------------------------------------------------------------------ Shape.hs:330-409
HD bool insideSphere(float r, V3 pt) { return sqLen(pt) - r * r < 1e-4f; }
HD void sampleShapeAny(const blingcu_shape &s, float u1, float u2, V3 &p, V3 &n) {
   const float *P = s.p;
   switch (s.kind) {
   case BLINGCU_SHAPE_BOX: {
      V3 pmin = mk3(P[0], P[1], P[2]), pmax = mk3(P[3], P[4], P[5]);
      int axis, nf; float u1p, u2p;
      remapRand(3, u1, axis, u1p); remapRand(2, u2, nf, u2p);
      n = setc(axis, (float)nf * 2 - 1, mk3(0, 0, 0));
      int oa0 = (axis + 1) % 3, oa1 = (axis + 2) % 3;
      V3 base = (nf == 0) ? pmin : pmax;
      p = setc(oa0, lerpf(u1p, comp(pmin, oa0), comp(pmax, oa0)), setc(oa1, lerpf(u2p, comp(pmin, oa1), comp(pmax, oa1)), base));
      return;
   }
   case BLINGCU_SHAPE_CYLINDER: {
      float z = lerpf(u1, P[1], P[2]), phi = lerpf(u2, 0, BL_TWOPI);
      p = mk3(P[0] * cosf(phi), P[0] * sinf(phi), z); n = normalize3(mk3(p.x, p.y, 0));
      return;
   }
   case BLINGCU_SHAPE_DISK: {
      float r = lerpf(u1, P[2], P[1]), phi = lerpf(u2, 0, P[3]);
      p = mk3(r * cosf(phi), r * sinf(phi), P[0]); n = mk3(0, 0, -1);
      return;
   }
   case BLINGCU_SHAPE_QUAD: p = mk3(lerpf(u1, -P[0], P[0]), lerpf(u2, -P[1], P[1]), 0); n = mk3(0, 0, -1); return;   // Q2
   default: { V3 q = uniformSampleSphere(u1, u2); p = q * mk3(P[0], P[0], P[0]); n = q; return; }
   }
}
HD void sampleShape(const blingcu_shape &s, V3 p, float u1, float u2, V3 &ps, V3 &ns) {
   if (s.kind == BLINGCU_SHAPE_SPHERE && !insideSphere(s.p[0], p)) {
      float r = s.p[0];
      V3 dn = normalize3(-p);
      Frame cs = coordinateSystem(dn);
      float cosThetaMax = sqrtf(hmaxf(0, 1 - (r * r) / sqLen(p)));
      Ray ray; ray.o = p; ray.d = uniformSampleCone(cs, cosThetaMax, u1, u2); ray.tmin = 0; ray.tmax = BL_INF;
      float t; DG dg;
      if (shapeIntersect<false>(s, ray, t, dg)) ps = rayAt(ray, t); else ps = dn * mk3(r, r, r);
      ns = normalize3(ps);
      return;
   }
   sampleShapeAny(s, u1, u2, ps, ns);
}
HD float generalPdf(const blingcu_shape &s, V3 p, V3 wi) {
   Ray r; r.o = p; r.d = wi; r.tmin = 1e-3f; r.tmax = BL_INF;
   float t; DG dg;
   if (!shapeIntersect<true>(s, r, t, dg)) return 0;
   float pd = sqLen(p - rayAt(r, t)) / (absDot(dg.n, -wi) * shapeArea(s));
   return isinf(pd) ? 0.0f : pd;
}
HD float shapePdf(const blingcu_shape &s, V3 p, V3 wi) {
   if (s.kind == BLINGCU_SHAPE_SPHERE && !insideSphere(s.p[0], p)) {
      float r = s.p[0];
      return uniformConePdf(sqrtf(hmaxf(0, 1 - r * r / sqLen(p))));
   }
   return generalPdf(s, p, wi);
}

// ------------------------------------------------------------------------------------------ Fresnel.hs
HD float frDielectric(float etai, float etat, float cosi) {   // :31-55 -- uniform over the bands, so a scalar
   float c0 = hmaxf(0, 1 - cosi * cosi);
   float costp = (cosi > 0) ? c0 / (etat * etat) : c0 * (etat * etat);
   float cost = sqrtf(1 - clampf(costp, 0, 1));
   float acosi = fabsf(cosi);
   float eta = etat / etai;
   float rParlP = eta * acosi;
   float rParl = (cost - rParlP) / (cost + rParlP);
   float rPerpP = eta * cost;
   float rPerp = (acosi - rPerpP) / (acosi + rPerpP);
   return (rParl * rParl + rPerp * rPerp) * 0.5f;
}
HD float frConductorBand(float eta, float k, float acosi) {   // :58-70, one band
   float ec2 = eta * (2 * acosi);
   float tmpF = eta * eta + k * k;
   float tmp = tmpF * (acosi * acosi);
   float rPer2 = (tmpF - ec2 + acosi * acosi) / (tmpF + ec2 + acosi * acosi);
   float rPar2 = (tmp - ec2 + 1.0f) / (tmp + ec2 + 1.0f);
   return (rPer2 + rPar2) / 2.0f;
}

// ------------------------------------------------------------------------------------------ BxDFs
HD float cosTheta(V3 v) { return v.z; }
HD float absCosTheta(V3 v) { return fabsf(v.z); }
HD float sinTheta2(V3 v) { return hmaxf(0, 1 - v.z * v.z); }
HD float sinTheta(V3 v) { return sqrtf(sinTheta2(v)); }
HD float cosPhi(V3 v) { float s = sinTheta(v); return s == 0 ? 1.0f : clampf(v.x / s, -1, 1); }
HD float sinPhi(V3 v) { float s = sinTheta(v); return s == 0 ? 0.0f : clampf(v.y / s, -1, 1); }
HD bool sameHemisphere(V3 a, V3 b) { return a.z * b.z > 0; }
HD V3 toSameHemisphere(V3 wo, V3 wi) { return wo.z < 0 ? mk3(wi.x, wi.y, -wi.z) : wi; }

enum { BX_REFLECTION = 1, BX_TRANSMISSION = 2, BX_DIFFUSE = 4, BX_GLOSSY = 8, BX_SPECULAR = 16 };
enum { K_LAMBERT = 0, K_ORENNAYAR, K_SPECREFL, K_SPECTRANS, K_MICROFACET, K_FRESNELBLEND };
enum { FR_NOOP = 0, FR_DIELECTRIC, FR_CONDUCTOR };

// Compile-time description of what a material can contain. The shade kernel is launched once per material KIND
// (material-sorted queues), so each launch is instantiated for its kind and the BxDF code of every other kind drops
// out (fewer registers, more resident warps). AnyMat keeps everything (resolve kernels, generic callers).
// TX: the material's textures may COMPUTE (scalar textures, blends, gradients, images, bump mapping; textures.h): upload gives
//     such materials the shade kind SK_TEX0 + kind, shaded by MatOf<SK_TEX0 + kind> = the kind's own traits with TX set.
// GEN: the ANY-kind instantiation (direct-lighting and normal-map bodies): BSDF / light code goes through the out-of-line
//     general copies below instead of being inlined (instruction fetch, profiles/r01_general_shade.md).
enum { SK_GENERAL = BLINGCU_MAT_KINDS, SK_TEX0 = 16 };
struct AnyMat { static const unsigned KM = 0x3fu, FM = 0x7u; static const int NC = 2, MK = -1; static const bool TX = false, GEN = false; };
// GenMat: AnyMat whose per-component BxDF calls go to ONE out-of-line copy each (bxdf*General below) -- used by the general
// shade kernels, whose code size is what limits them (see "out-of-line general versions")
struct GenMat : AnyMat {};
template <class M> struct IsGen { static const bool v = false; };
template <> struct IsGen<GenMat> { static const bool v = true; };
template <int MATKIND> struct MatOf : AnyMat {};
template <> struct MatOf<SK_GENERAL> : AnyMat { static const bool TX = true, GEN = true; };
#define BL_K(k) (1u << (k))
template <> struct MatOf<BLINGCU_MAT_MATTE> { static const unsigned KM = BL_K(0) | BL_K(1), FM = 0u; static const int NC = 1, MK = BLINGCU_MAT_MATTE; static const bool TX = false, GEN = false; };
template <> struct MatOf<BLINGCU_MAT_GLASS> { static const unsigned KM = BL_K(2) | BL_K(3), FM = BL_K(1); static const int NC = 2, MK = BLINGCU_MAT_GLASS; static const bool TX = false, GEN = false; };
template <> struct MatOf<BLINGCU_MAT_MIRROR> { static const unsigned KM = BL_K(2), FM = BL_K(0); static const int NC = 1, MK = BLINGCU_MAT_MIRROR; static const bool TX = false, GEN = false; };
template <> struct MatOf<BLINGCU_MAT_PLASTIC> { static const unsigned KM = BL_K(0) | BL_K(4), FM = BL_K(1); static const int NC = 2, MK = BLINGCU_MAT_PLASTIC; static const bool TX = false, GEN = false; };
template <> struct MatOf<BLINGCU_MAT_METAL> { static const unsigned KM = BL_K(4), FM = BL_K(2); static const int NC = 1, MK = BLINGCU_MAT_METAL; static const bool TX = false, GEN = false; };
template <> struct MatOf<BLINGCU_MAT_SHINYMETAL> { static const unsigned KM = BL_K(2) | BL_K(4), FM = BL_K(2); static const int NC = 2, MK = BLINGCU_MAT_SHINYMETAL; static const bool TX = false, GEN = false; };
template <> struct MatOf<BLINGCU_MAT_TRANSMATTE> { static const unsigned KM = BL_K(0) | BL_K(1), FM = 0u; static const int NC = 2, MK = BLINGCU_MAT_TRANSMATTE; static const bool TX = false, GEN = false; };
template <> struct MatOf<BLINGCU_MAT_SUBSTRATE> { static const unsigned KM = BL_K(5), FM = 0u; static const int NC = 1, MK = BLINGCU_MAT_SUBSTRATE; static const bool TX = false, GEN = false; };
template <> struct MatOf<BLINGCU_MAT_BLACKBODY> { static const unsigned KM = 0u, FM = 0u; static const int NC = 1, MK = BLINGCU_MAT_BLACKBODY; static const bool TX = false, GEN = false; };

#define BL_TEXMAT(K) template <> struct MatOf<SK_TEX0 + K> : MatOf<K> { static const bool TX = true; };
BL_TEXMAT(BLINGCU_MAT_MATTE) BL_TEXMAT(BLINGCU_MAT_GLASS) BL_TEXMAT(BLINGCU_MAT_MIRROR) BL_TEXMAT(BLINGCU_MAT_PLASTIC) BL_TEXMAT(BLINGCU_MAT_METAL)
BL_TEXMAT(BLINGCU_MAT_BLACKBODY) BL_TEXMAT(BLINGCU_MAT_SHINYMETAL) BL_TEXMAT(BLINGCU_MAT_TRANSMATTE) BL_TEXMAT(BLINGCU_MAT_SUBSTRATE)
#undef BL_TEXMAT

struct BxDF {
   int kind, type, fr, clamp01;   // clamp01: sClamp' applied to r on read (glass, mirror; Material.hs:63-64,71)
   int flip;                      // brdfToBtdf (Reflection.hs:188-195): the BRDF seen through the other hemisphere
   const float *r;                // reflectance / transmittance spectrum; null = white
   const float *r2;               // translucentMatte: r is scaled by (1 - clamp01 r2) (Material.hs:51)
   const float *eta, *k;          // conductor; FresnelBlend: eta = specular rs, k = absorption ra (both clamped on read)
   float a, b, e, etai, etat;     // OrenNayar A,B ; Blinn exponent (FresnelBlend: Anisotropic ex) ; dielectric indices
   float ey, depth;               // FresnelBlend: Anisotropic ey, coating depth
};
HD float bxR(const BxDF &b, int i) {
   float v = b.r ? b.r[i] : 1.0f;
   if (b.clamp01) v = hmaxf(0.0f, hminf(1.0f, v));
   if (b.r2) v = v * (1.0f - hmaxf(0.0f, hminf(1.0f, b.r2[i])));
   return v;
}
HD V3 flipZ(V3 w) { return mk3(w.x, w.y, -w.z); }
HD Spec bxScaledR(const BxDF &b, float f) { Spec s; BL_UNROLL for (int i = 0; i < NB; ++i) s.v[i] = bxR(b, i) * f; return s; }
// r * fr(cosi) (spectral product, then the caller scales)
template <class M = AnyMat>
HD Spec bxRFresnel(const BxDF &b, float cosi) {
   Spec s;
   if ((M::FM & BL_K(FR_CONDUCTOR)) && b.fr == FR_CONDUCTOR) { float ac = fabsf(cosi); BL_UNROLL for (int i = 0; i < NB; ++i) s.v[i] = bxR(b, i) * frConductorBand(b.eta[i], b.k[i], ac); }
   else { float f = ((M::FM & BL_K(FR_DIELECTRIC)) && b.fr == FR_DIELECTRIC) ? frDielectric(b.etai, b.etat, cosi) : 1.0f; BL_UNROLL for (int i = 0; i < NB; ++i) s.v[i] = bxR(b, i) * f; }
   return s;
}
HD float mfG(V3 wo, V3 wi, V3 wh) {   // Microfacet.hs:113-120
   float nh = absCosTheta(wh), no = absCosTheta(wo), ni = absCosTheta(wi), oh = absDot(wo, wh);
   return hminf(1, hminf(2 * nh * no / oh, 2 * nh * ni / oh));
}
HD float orenNayarF(const BxDF &b, V3 wo, V3 wi) {   // Diffuse.hs:52-65 (scalar factor on r)
   float sinti = sinTheta(wi), sinto = sinTheta(wo);
   float sina, tanb;
   if (absCosTheta(wi) > absCosTheta(wo)) { sina = sinto; tanb = sinti / absCosTheta(wi); }
   else { sina = sinti; tanb = sinto / absCosTheta(wo); }
   float maxcos = 0;
   if (sinti > 1e-4f && sinto > 1e-4f) maxcos = hmaxf(0, cosPhi(wi) * cosPhi(wo) + sinPhi(wi) * sinPhi(wo));
   return b.a + b.b * maxcos * sina * tanb;
}
// Anisotropic distribution (Microfacet.hs:140-173,184-192)
HD float anisoE(float ex, float ey, V3 wh, float d) { return (ex * wh.x * wh.x + ey * wh.y * wh.y) / d; }
HD float anisoPdf(float ex, float ey, V3 wh) {
   float costh = absCosTheta(wh);
   float e = anisoE(ex, ey, wh, hmaxf(0, 1 - costh * costh));
   return sqrtf((ex + 1) * (ey + 1)) * BL_INVTWOPI * powf(costh, e);
}
HD float anisoD(float ex, float ey, V3 wh) {
   float costh = absCosTheta(wh);
   float d = 1 - costh * costh;
   if (d == 0) return 0;
   return sqrtf((ex + 2) * (ey + 2)) * BL_INVTWOPI * powf(costh, anisoE(ex, ey, wh, d));
}
HD void anisoQuad(float ex, float ey, float u, float u2, float &p, float &c) {   // smpFirstQuadrand
   p = (ex == ey) ? BL_PI * u * 0.5f : atanf(sqrtf((ex + 1) / (ey + 1)) * tanf(BL_PI * u * 0.5f));
   float cp = cosf(p), sp = sinf(p);
   c = powf(u2, 1 / (ex * cp * cp + ey * sp * sp + 1));
}
HD void anisoSample(float ex, float ey, float u1, float u2, V3 &wh, float &pdf) {
   float phi, cost, p;
   if (u1 < 0.25f) { anisoQuad(ex, ey, 4 * u1, u2, p, cost); phi = p; }
   else if (u1 < 0.50f) { anisoQuad(ex, ey, 4 * (0.5f - u1), u2, p, cost); phi = BL_PI - p; }
   else if (u1 < 0.75f) { anisoQuad(ex, ey, 4 * (u1 - 0.5f), u2, p, cost); phi = p + BL_PI; }
   else { anisoQuad(ex, ey, 4 * (1 - u1), u2, p, cost); phi = 2 * BL_PI - p; }
   float sint = sqrtf(hmaxf(0, 1 - cost * cost));
   wh = sphericalDirection(sint, cost, phi);
   float e = anisoE(ex, ey, wh, 1 - cost * cost);
   pdf = sqrtf((ex + 1) * (ey + 1)) * (BL_INVTWOPI * powf(cost, e));
}
HD V3 halfUp(V3 wi, V3 wo) { V3 w = normalize3(wi + wo); return w.z < 0 ? -w : w; }
// FresnelBlend `e wo wi` (Microfacet.hs:65-84): r = rd, eta = rs, k = ra, all clamped to [0,1] (mkSubstrate)
HD Spec fresnelBlendEval(const BxDF &b, V3 wo, V3 wi) {
   float costi = absCosTheta(wi), costo = absCosTheta(wo);
   float asc = -(b.depth * (costi + costo) / (costi * costo));
   float ds = (costo * 28 / 23 * BL_PI) * (1 - powf(1 - 0.5f * costi, 5.0f)) * (1 - powf(1 - 0.5f * costo, 5.0f));
   V3 wh = halfUp(wi, wo);
   float costih = absDot(wi, wh);
   float sch = powf(1 - costih, 5.0f);
   float ss = anisoD(b.e, b.ey, wh) * costo / (4 * costih * hmaxf(costi, costo));
   Spec f;
   BL_UNROLL for (int i = 0; i < NB; ++i) {
      float rd = hmaxf(0.0f, hminf(1.0f, b.r[i])), rs = hmaxf(0.0f, hminf(1.0f, b.eta[i])), ra = hmaxf(0.0f, hminf(1.0f, b.k[i]));
      float a = (b.depth > 0) ? expf(ra * asc) : 1.0f;
      f.v[i] = ((a * rd) * (1.0f - rs)) * ds + (rs + (1.0f - rs) * sch) * ss;
   }
   return f;
}
// bxdfEval b wo wi -- callers pass flipped arguments for the non-adjoint case (Reflection.hs:310,330)
template <class M = AnyMat>
HD Spec bxdfEval(const BxDF &b, V3 wo, V3 wi) {
   if (M::MK == BLINGCU_MAT_TRANSMATTE || M::MK < 0) { if (b.flip) wi = flipZ(wi); }
   if ((M::KM & BL_K(K_LAMBERT)) && b.kind == K_LAMBERT) return bxScaledR(b, BL_INVPI * absCosTheta(wo));
   if ((M::KM & BL_K(K_ORENNAYAR)) && b.kind == K_ORENNAYAR) return sScale(bxScaledR(b, orenNayarF(b, wo, wi)), BL_INVPI * absCosTheta(wo));
   if ((M::KM & BL_K(K_MICROFACET)) && b.kind == K_MICROFACET) {   // Microfacet.hs:20-33
      float costo = absCosTheta(wo), costi = absCosTheta(wi);
      if (costi == 0 || costo == 0) return sConst(0);
      V3 whp = wi + wo;
      if (whp.x == 0 && whp.y == 0 && whp.z == 0) return sConst(0);
      V3 wh = normalize3(whp);
      if (cosTheta(wh) < 0) return sConst(0);
      float costh = dot3(wi, wh);
      float x = (b.e + 2) * BL_INVTWOPI * powf(absCosTheta(wh), b.e) * mfG(wo, wi, wh) / (4 * costi);
      return sScale(bxRFresnel<M>(b, costh), x);
   }
   if ((M::KM & BL_K(K_FRESNELBLEND)) && b.kind == K_FRESNELBLEND) return fresnelBlendEval(b, wo, wi);
   return sConst(0);
}
HD float cosPdf(V3 wo, V3 wi) { return sameHemisphere(wo, wi) ? BL_INVPI * absCosTheta(wi) : 0.0f; }
template <class M = AnyMat>
HD float bxdfPdf(const BxDF &b, V3 wo, V3 wi) {
   if (M::MK == BLINGCU_MAT_TRANSMATTE || M::MK < 0) { if (b.flip) wi = flipZ(wi); }
   if ((M::KM & (BL_K(K_LAMBERT) | BL_K(K_ORENNAYAR))) && (b.kind == K_LAMBERT || b.kind == K_ORENNAYAR)) return cosPdf(wo, wi);
   if ((M::KM & BL_K(K_MICROFACET)) && b.kind == K_MICROFACET) {   // Microfacet.hs:35-41
      V3 whp = wo + wi;
      if (sqLen(whp) == 0) return 0;
      V3 wh = normalize3(whp);
      if (cosTheta(wh) < 0) return 0;
      return (b.e + 1) * powf(absCosTheta(wh), b.e) * BL_INVTWOPI / (4 * absDot(wo, wh));
   }
   if ((M::KM & BL_K(K_FRESNELBLEND)) && b.kind == K_FRESNELBLEND) {   // Microfacet.hs:103-108
      if (!sameHemisphere(wo, wi)) return 0;
      V3 wh = halfUp(wi, wo);
      return 0.5f * (absCosTheta(wi) * BL_INVPI + anisoPdf(b.e, b.ey, wh) / (4 * absDot(wo, wh)));
   }
   return 0;
}
template <class M = AnyMat>
HD void bxdfSample(const BxDF &b, V3 wo, float u1, float u2, Spec &f, V3 &wi, float &pdf) {   // adj = False
   if ((M::KM & (BL_K(K_LAMBERT) | BL_K(K_ORENNAYAR))) && (b.kind == K_LAMBERT || b.kind == K_ORENNAYAR)) {   // Diffuse.hs:14-22,38-42
      wi = toSameHemisphere(wo, cosineSampleHemisphere(u1, u2));
      if (sameHemisphere(wo, wi)) { f = bxScaledR(b, b.kind == K_LAMBERT ? 1.0f : orenNayarF(b, wo, wi)); pdf = cosPdf(wo, wi); }
      else { f = sConst(0); pdf = 0; }
      if (M::MK == BLINGCU_MAT_TRANSMATTE || M::MK < 0) { if (b.flip) wi = flipZ(wi); }
      return;
   }
   if ((M::KM & BL_K(K_SPECREFL)) && b.kind == K_SPECREFL) { f = bxRFresnel<M>(b, cosTheta(wo)); wi = mk3(-wo.x, -wo.y, wo.z); pdf = 1; return; }   // Specular.hs:11-20
   if ((M::KM & BL_K(K_SPECTRANS)) && b.kind == K_SPECTRANS) {   // Specular.hs:34-57
      bool entering = cosTheta(wo) > 0;
      float ei = entering ? b.etai : b.etat, et = entering ? b.etat : b.etai;
      float eta = ei / et, eta2 = eta * eta, sint2 = eta2 * sinTheta2(wo);
      if (sint2 >= 1) { f = sConst(0); wi = wo; pdf = 0; return; }
      float c = sqrtf(hmaxf(0, 1 - sint2));
      float cost = entering ? -c : c;
      wi = mk3(eta * (-wo.x), eta * (-wo.y), cost);
      float fr = frDielectric(ei, et, cost);   // Q4
      BL_UNROLL for (int i = 0; i < NB; ++i) f.v[i] = ((1.0f - fr) * bxR(b, i)) * eta2;
      pdf = 1;
      return;
   }
   if ((M::KM & BL_K(K_FRESNELBLEND)) && b.kind == K_FRESNELBLEND) {   // Microfacet.hs:86-101, adj = False
      float pdfp; V3 wh;
      if (u1 < 0.5f) {
         wi = toSameHemisphere(wo, cosineSampleHemisphere(u1 * 2, u2));
         wh = halfUp(wi, wo);
         pdfp = anisoPdf(b.e, b.ey, wh);
      } else {
         anisoSample(b.e, b.ey, 2 * (u1 - 0.5f), u2, wh, pdfp);
         wi = scl(2, scl(dot3(wo, wh), wh)) - wo;
      }
      if (pdfp == 0) { f = sConst(0); pdf = 0; return; }
      pdf = 0.5f * (absCosTheta(wi) * BL_INVPI + pdfp / (4 * absDot(wo, wh)));
      f = sScale(fresnelBlendEval(b, wo, wi), 1 / pdf);
      return;
   }
   if ((M::KM & BL_K(K_MICROFACET)) && (M::MK >= 0 || b.kind == K_MICROFACET)) {   // microfacet, Blinn (Microfacet.hs:43-54,175-182)
      float cost = powf(u1, 1 / (b.e + 1));
      float sint = sqrtf(hmaxf(0, 1 - cost * cost));
      V3 whp = sphericalDirection(sint, cost, u2 * 2 * BL_PI);
      float ff = powf(cost, b.e) * BL_INVTWOPI;
      float d = (b.e + 2) * ff, dpdf = (b.e + 1) * ff;
      V3 wh = (cosTheta(whp) < 0) ? -whp : whp;
      float costH = dot3(wo, wh);
      wi = scl(2 * costH, wh) - wo;
      if (!sameHemisphere(wo, wi)) { f = sConst(0); wi = wo; pdf = 0; return; }
      float fact = d * fabsf(costH) / dpdf * mfG(wo, wi, wh);
      f = sScale(bxRFresnel<M>(b, costH), fact / absCosTheta(wi));   // Q7
      pdf = dpdf / (4 * fabsf(costH));
      return;
   }
   f = sConst(0); wi = wo; pdf = 0;
}

HDNI void bxdfEvalGeneral(const BxDF &b, V3 wo, V3 wi, Spec &f) { f = bxdfEval<AnyMat>(b, wo, wi); }
HDNI float bxdfPdfGeneral(const BxDF &b, V3 wo, V3 wi) { return bxdfPdf<AnyMat>(b, wo, wi); }
HDNI void bxdfSampleGeneral(const BxDF &b, V3 wo, float u1, float u2, Spec &f, V3 &wi, float &pdf) { bxdfSample<AnyMat>(b, wo, u1, u2, f, wi, pdf); }
template <class M> HD Spec bxdfEvalOf(const BxDF &b, V3 wo, V3 wi) { if (IsGen<M>::v) { Spec f; bxdfEvalGeneral(b, wo, wi, f); return f; } return bxdfEval<M>(b, wo, wi); }
template <class M> HD float bxdfPdfOf(const BxDF &b, V3 wo, V3 wi) { if (IsGen<M>::v) return bxdfPdfGeneral(b, wo, wi); return bxdfPdf<M>(b, wo, wi); }
template <class M> HD void bxdfSampleOf(const BxDF &b, V3 wo, float u1, float u2, Spec &f, V3 &wi, float &pdf) { if (IsGen<M>::v) bxdfSampleGeneral(b, wo, u1, u2, f, wi, pdf); else bxdfSample<M>(b, wo, u1, u2, f, wi, pdf); }

struct Bsdf { int n; BxDF bx[2]; Frame cs; V3 p, ng; };
struct BsdfSample { int type; float pdf; Spec f; V3 wi; };

HD bool bxMatch(const BxDF &b, bool wantTrans) { return (b.type & (wantTrans ? BX_TRANSMISSION : BX_REFLECTION)) != 0; }

// Reflection.hs:278-316, adj = False, flags = bxdfAll
template <class M = AnyMat>
HD void sampleBsdf(const Bsdf &bsdf, V3 woW, float uComp, float u1, float u2, BsdfSample &out) {
   out.type = BX_REFLECTION | BX_DIFFUSE; out.pdf = 0; out.f = sConst(0); out.wi = mk3(0, 1, 0);
   const int cntm = (M::NC == 1) ? imin(bsdf.n, 1) : bsdf.n;
   if (M::KM == 0u || cntm == 0) return;
   V3 wo = worldToLocal(bsdf.cs, woW);
   float cntf = (float)cntm, invCnt = 1 / cntf;
   const int sNum = (M::NC == 1) ? 0 : imax(0, imin(cntm - 1, (int)floorf(uComp * cntf)));
   const BxDF &bx = bsdf.bx[sNum];
   Spec fS; V3 wi = mk3(0, 1, 0); float pdfp = 0;
   bxdfSampleOf<M>(bx, wo, u1, u2, fS, wi, pdfp);
   V3 wiW = localToWorld(bsdf.cs, wi);
   float sideTest = dot3(wiW, bsdf.ng) / dot3(woW, bsdf.ng);
   if (pdfp == 0 || sideTest == 0) return;
   bool wantTrans = sideTest < 0;
   if (!bxMatch(bx, wantTrans)) return;
   out.type = bx.type; out.wi = wiW;
   if (bx.type & BX_SPECULAR) { out.pdf = pdfp * invCnt; out.f = sScale(fS, cntf); return; }
   if (M::NC == 1 || cntm == 1) { out.pdf = pdfp; out.f = fS; return; }
   const BxDF &o = bsdf.bx[1 - sNum];
   float pdf = (pdfp + bxdfPdfOf<M>(o, wo, wi)) * invCnt;
   Spec fOthers = sConst(0);
   if (bxMatch(o, wantTrans)) fOthers = fOthers + bxdfEvalOf<M>(o, wi, wo);
   out.pdf = pdf; out.f = sScale(sScale(fS, pdfp) + fOthers, 1 / pdf);
}
// sampleBsdf' flags (Reflection.hs:271-272,278-316), adj = False, for the SPECULAR component filters of the direct-lighting
// integrator (DirectLighting.hs:50-52: [Specular, Reflection] and [Specular, Transmission]). bxdfMatches b flags =
// (type b .&. flags) == type b (:119-121,185-186); every material in scope has at most ONE component per such filter,
// so the sample is that component's: pdf' / 1, f * 1.
template <class M = AnyMat>
HD void sampleBsdfSpecular(const Bsdf &bsdf, int flags, V3 woW, BsdfSample &out) {
   out.type = BX_REFLECTION | BX_DIFFUSE; out.pdf = 0; out.f = sConst(0); out.wi = mk3(0, 1, 0);
   int pick = -1;
   BL_UNROLL for (int i = 0; i < 2; ++i) if (i < bsdf.n && pick < 0 && (bsdf.bx[i].type & flags) == bsdf.bx[i].type) pick = i;
   if (pick < 0) return;
   const BxDF &bx = bsdf.bx[pick];
   V3 wo = worldToLocal(bsdf.cs, woW);
   Spec fS; V3 wi = mk3(0, 1, 0); float pdfp = 0;
   bxdfSampleOf<M>(bx, wo, 0.5f, 0.5f, fS, wi, pdfp);
   V3 wiW = localToWorld(bsdf.cs, wi);
   float sideTest = dot3(wiW, bsdf.ng) / dot3(woW, bsdf.ng);
   if (pdfp == 0 || sideTest == 0) return;
   if (!bxMatch(bx, sideTest < 0)) return;
   out.type = bx.type; out.wi = wiW; out.pdf = pdfp; out.f = fS;
}
// Reflection.hs:318-332, adj = False
template <class M = AnyMat>
HD Spec evalBsdf(const Bsdf &bsdf, V3 woW, V3 wiW) {
   if (M::KM == 0u) return sConst(0);
   float cosWo = dot3(woW, bsdf.ng);
   float sideTest = dot3(wiW, bsdf.ng) / cosWo;
   if (sideTest == 0) return sConst(0);
   if (fabsf(cosWo) < 1e-5f) return sConst(0);
   bool wantTrans = sideTest < 0;
   V3 wo = worldToLocal(bsdf.cs, woW), wi = worldToLocal(bsdf.cs, wiW);
   Spec f = sConst(0);
   BL_UNROLL for (int i = 0; i < M::NC; ++i) if (i < bsdf.n && bxMatch(bsdf.bx[i], wantTrans)) f = f + bxdfEvalOf<M>(bsdf.bx[i], wi, wo);
   return f;
}
template <class M = AnyMat>
HD float bsdfPdf(const Bsdf &bsdf, V3 woW, V3 wiW) {   // Reflection.hs:251-257 (Q6)
   if (M::KM == 0u || bsdf.n == 0) return 0;
   V3 wo = worldToLocal(bsdf.cs, woW), wi = worldToLocal(bsdf.cs, wiW);
   float s = 0;
   BL_UNROLL for (int i = 0; i < M::NC; ++i) if (i < bsdf.n) s = s + bxdfPdfOf<M>(bsdf.bx[i], wo, wi);
   return s / (float)bsdf.n;
}

// ------------------------------------------------------------------------------------------ textures + materials
HD const float *evalSpectrumTexture(const DScene &sc, int id, const DG &dg) {   // Texture.hs:159-207
   for (int guard = 0; guard < 16; ++guard) {
      const blingcu_texture &t = sc.textures[id];
      if (t.kind == BLINGCU_TEX_CONSTANT) return t.s.v;
      if (t.kind == BLINGCU_TEX_CHECKER) {   // Texture.hs:209-221; Haskell `mod` 2 of a sum of floors: parity of the sum
         int q = (int)floorf(dg.p.x * t.f[0]) + (int)floorf(dg.p.y * t.f[1]) + (int)floorf(dg.p.z * t.f[2]);
         id = ((q & 1) == 0) ? t.child[0] : t.child[1];
         continue;
      }
      float x, z;
      if (t.aux == 0) { x = t.f[1] * dg.u + t.f[3]; z = t.f[2] * dg.v + t.f[4]; }   // uvMapping :166-170
      else map2d(t.s.v, dg, x, z);
      float xp = fabsf(x - truncf(x)), zp = fabsf(z - truncf(z));      // properFraction
      float lo = t.f[0] / 2, hi = 1.0f - lo;
      id = (xp < lo || zp < lo || xp > hi || zp > hi) ? t.child[1] : t.child[0];
   }
   return sc.textures[id].s.v;
}
HD float fixExponent(float e) { return (e > 10000 || isnan(e)) ? 10000.0f : e; }

struct SurfaceHit {   // what mkIntersection carries (Primitive.hs:49-65)
   DG dgg;            // geometric DG, world space
   float eps;
   int material, light;
};

// geometric DG of a hit from (ray, t, b1, b2, prim): re-derives what the reference stores in Intersection
HD void surfaceAt(const DScene &sc, const Ray &ray, float t, float b1, float b2, int href, SurfaceHit &sh, DG &dgs) {
   const uint32_t ref = refIndex(href);
   if (!refIsShape(href)) {
      const F4 *tp = sc.tri_p + BL_TRI_F4 * (size_t)ref;
      F4 a, b, c, u;
#if defined(__CUDA_ARCH__)
      ld8(tp, a, b); ld8(tp + 2, c, u);
#else
      a = ld4(tp); b = ld4(tp + 1); c = ld4(tp + 2); u = ld4(tp + 3);
#endif
      float uv[6] = {b.w, c.w, u.x, u.y, u.z, u.w};
      sh.dgg = triDG(mk3(a.x, a.y, a.z), mk3(b.x, b.y, b.z), mk3(c.x, c.y, c.z), uv, rayAt(ray, t), b1, b2);
      sh.eps = 1e-3f * t;   // TriangleMesh.hs:206
      sh.material = f2i(a.w); sh.light = -1;
      dgs = sh.dgg;
      if (sc.tri_n) {   // triangleShadingGeometry (TriangleMesh.hs:122-134), o2w = mempty
         const F4 *np = sc.tri_n + 3 * (size_t)ref;
         const F4 n1 = ld4(np), n2 = ld4(np + 1), n3 = ld4(np + 2);
         const float N[9] = {n1.x, n1.y, n1.z, n2.x, n2.y, n2.z, n3.x, n3.y, n3.z};
         // nine zeros: a triangle of a mesh WITHOUT normals in a scene that also holds smooth meshes keeps its geometric frame
         if (N[0] == 0 && N[1] == 0 && N[2] == 0 && N[3] == 0 && N[4] == 0 && N[5] == 0 && N[6] == 0 && N[7] == 0 && N[8] == 0) return;
         float b0 = 1 - b1 - b2;
         V3 ns = normalize3((scl(b0, mk3(N[0], N[1], N[2])) + scl(b1, mk3(N[3], N[4], N[5]))) + scl(b2, mk3(N[6], N[7], N[8])));
         V3 tsp = cross3(normalize3(sh.dgg.dpdu), ns);
         if (sqLen(tsp) > 0) { dgs.dpdu = cross3(normalize3(tsp), ns); dgs.dpdv = normalize3(tsp); }
         else { Frame f = coordinateSystem(ns); dgs.dpdu = f.s; dgs.dpdv = f.t; }
         dgs.n = ns;
      }
      return;
   }
   const blingcu_shape &s = sc.shapes[ref];
   Ray ro = transRay(s.w2o, ray); ro.tmax = t;   // same arithmetic as the traversal => same t
   float t2; DG dgo;
   shapeIntersect<true>(s, ro, t2, dgo);
   sh.dgg = transDg(s.o2w, s.w2o, dgo);
   sh.eps = 5e-4f * t;    // Shape.hs:95,137,150,164,185
   sh.material = s.material; sh.light = s.light;
   dgs = sh.dgg;
}

// Material.hs:32-96 + mkBsdf' (Reflection.hs:209-225)
// Texture access of makeBsdf. Fast path: a pointer into the texture table and the constant f[i]. Textured path (M::TX):
// the spectrum is computed into the caller's scratch (one Spec per texture slot) and scalars come from ftex[i].
template <class M>
HD const float *matSpectrum(const DScene &sc, int tex, const DG &dg, Spec *scratch, int slot) {
   if (M::TX) { scratch[slot] = SpectrumValue<BL_BLEND_DEPTH>::eval(sc, tex, dg); return scratch[slot].v; }
   return evalSpectrumTexture(sc, tex, dg);
}
template <class M>
HD float matScalar(const DScene &sc, const blingcu_material &m, int i, const DG &dg) {
   if (M::TX) { if (m.ftex[i]) return evalScalarTexture(sc, m.ftex[i] - 1, dg); }
   return m.f[i];
}
template <class M = AnyMat>
HD void makeBsdf(const DScene &sc, const SurfaceHit &sh, const DG &dgsIn, Bsdf &b, Spec *scratch = 0) {
   const blingcu_material &m = sc.materials[sh.material];
   const int mk = (M::MK >= 0) ? M::MK : m.kind;   // compile-time constant in the per-kind shade kernels
   DG dgs = dgsIn;
   if (M::TX) { if (m.bump) dgs = bumpDG(sc, m.bump - 1, sh.dgg, dgsIn); }   // bumpMapped d mat dgg dgs = mat dgg (bump d dgg dgs)
#define BL_TEX(t, slot) matSpectrum<M>(sc, (t), dgs, scratch, (slot))
#define BL_F(i) matScalar<M>(sc, m, (i), dgs)
   b.n = 0;
   BL_UNROLL for (int i = 0; i < 2; ++i) { BxDF &x = b.bx[i]; x.kind = 0; x.type = 0; x.fr = FR_NOOP; x.clamp01 = 0; x.flip = 0; x.r = 0; x.r2 = 0; x.eta = 0; x.k = 0; x.a = x.b = x.e = 0; x.etai = x.etat = 1; x.ey = 0; x.depth = 0; }
   switch (mk) {
   case BLINGCU_MAT_MATTE: {
      BxDF &x = b.bx[0]; x.r = BL_TEX(m.tex[0], 0); x.type = BX_REFLECTION | BX_DIFFUSE;
      float s = BL_F(0);
      if (s == 0) x.kind = K_LAMBERT;
      else {   // Diffuse.hs:29-36
         x.kind = K_ORENNAYAR; float sg = clampf(s, 0, 1), sig2 = sg * sg;
         x.a = 1 - (sig2 / (2 * (sig2 + 0.33f))); x.b = 0.45f * sig2 / (sig2 + 0.09f);
      }
      b.n = 1; break;
   }
   case BLINGCU_MAT_GLASS: {
      BxDF &r = b.bx[0], &t = b.bx[1];
      const float ior = BL_F(0);
      r.kind = K_SPECREFL; r.type = BX_REFLECTION | BX_SPECULAR; r.r = BL_TEX(m.tex[0], 0); r.clamp01 = 1; r.fr = FR_DIELECTRIC; r.etai = 1; r.etat = ior;
      t.kind = K_SPECTRANS; t.type = BX_TRANSMISSION | BX_SPECULAR; t.r = BL_TEX(m.tex[1], 1); t.clamp01 = 1; t.etai = 1; t.etat = ior;
      b.n = 2; break;
   }
   case BLINGCU_MAT_MIRROR: {
      BxDF &r = b.bx[0]; r.kind = K_SPECREFL; r.type = BX_REFLECTION | BX_SPECULAR; r.r = BL_TEX(m.tex[0], 0); r.clamp01 = 1; r.fr = FR_NOOP;
      b.n = 1; break;
   }
   case BLINGCU_MAT_PLASTIC: {
      BxDF &d = b.bx[0], &s = b.bx[1];
      d.kind = K_LAMBERT; d.type = BX_REFLECTION | BX_DIFFUSE; d.r = BL_TEX(m.tex[0], 0);
      s.kind = K_MICROFACET; s.type = BX_REFLECTION | BX_GLOSSY; s.r = BL_TEX(m.tex[1], 1);
      s.fr = FR_DIELECTRIC; s.etai = 1.0f; s.etat = 1.5f; s.e = fixExponent(1 / BL_F(0));
      b.n = 2; break;
   }
   case BLINGCU_MAT_METAL: {
      BxDF &s = b.bx[0]; s.kind = K_MICROFACET; s.type = BX_REFLECTION | BX_GLOSSY; s.r = 0; s.fr = FR_CONDUCTOR;
      s.eta = BL_TEX(m.tex[0], 0); s.k = BL_TEX(m.tex[1], 1); s.e = fixExponent(1 / BL_F(0));
      b.n = 1; break;
   }
   case BLINGCU_MAT_SHINYMETAL: {   // Material.hs:98-108; eta / k textures already carry frApproxEta / frApproxK (host)
      BxDF &d = b.bx[0], &sp = b.bx[1];
      d.kind = K_MICROFACET; d.type = BX_REFLECTION | BX_GLOSSY; d.r = 0; d.fr = FR_CONDUCTOR;
      d.eta = BL_TEX(m.tex[0], 0); d.k = BL_TEX(m.tex[1], 1); d.e = fixExponent(1 / BL_F(0));
      sp.kind = K_SPECREFL; sp.type = BX_REFLECTION | BX_SPECULAR; sp.r = 0; sp.fr = FR_CONDUCTOR;
      sp.eta = BL_TEX(m.tex[2], 2); sp.k = BL_TEX(m.tex3, 3);
      b.n = 2; break;
   }
   case BLINGCU_MAT_SUBSTRATE: {   // mkSubstrate (Material.hs:110-127)
      BxDF &fb = b.bx[0]; fb.kind = K_FRESNELBLEND; fb.type = BX_REFLECTION | BX_GLOSSY;
      fb.r = BL_TEX(m.tex[0], 0); fb.eta = BL_TEX(m.tex[1], 1); fb.k = BL_TEX(m.tex[2], 2);
      fb.e = fixExponent(1 / hmaxf(0.0f, BL_F(0))); fb.ey = fixExponent(1 / hmaxf(0.0f, BL_F(1))); fb.depth = BL_F(2);
      b.n = 1; break;
   }
   case BLINGCU_MAT_TRANSMATTE: {   // Material.hs:43-53
      BxDF &rf = b.bx[0], &tr = b.bx[1];
      const float *kr = BL_TEX(m.tex[0], 0), *kt = BL_TEX(m.tex[1], 1);
      float sg = BL_F(0), a = 0, bb = 0; int kind = K_LAMBERT;
      if (sg != 0) { kind = K_ORENNAYAR; float c = clampf(sg, 0, 1), sig2 = c * c; a = 1 - (sig2 / (2 * (sig2 + 0.33f))); bb = 0.45f * sig2 / (sig2 + 0.09f); }
      rf.kind = kind; rf.type = BX_REFLECTION | BX_DIFFUSE; rf.r = kr; rf.clamp01 = 1; rf.a = a; rf.b = bb;
      tr.kind = kind; tr.type = BX_TRANSMISSION | BX_DIFFUSE; tr.r = kt; tr.clamp01 = 1; tr.r2 = kr; tr.flip = 1; tr.a = a; tr.b = bb;
      b.n = 2; break;
   }
   default: break;
   }
   V3 nn = dgs.n, sn = normalize3(dgs.dpdu);
   b.cs.s = sn; b.cs.t = cross3(nn, sn); b.cs.n = nn;
   b.p = dgs.p; b.ng = sh.dgg.n;
#undef BL_TEX
#undef BL_F
}

// ------------------------------------------------------------------------------------------ spectra conversions
HD Spec rgbToSpectrumBasis(const float (*basis)[NB], float r, float g, float b) {   // rgbToSpectrum (Spectrum.hs:146-159)
   const float *rb = basis[0], *gb = basis[1], *bb = basis[2], *cb = basis[3], *mb = basis[4], *yb = basis[5], *wb = basis[6];
   const float *A, *B; float w, fa, fb;
   if (r <= g && r <= b) { w = r; if (g <= b) { A = cb; fa = g - r; B = bb; fb = b - g; } else { A = cb; fa = b - r; B = gb; fb = g - b; } }
   else if (g <= r && g <= b) { w = g; if (r <= b) { A = mb; fa = r - g; B = bb; fb = b - r; } else { A = mb; fa = b - g; B = rb; fb = r - b; } }
   else { w = b; if (r <= b) { A = yb; fa = r - b; B = gb; fb = g - r; } else { A = yb; fa = g - b; B = rb; fb = r - g; } }
   Spec s; BL_UNROLL for (int i = 0; i < NB; ++i) s.v[i] = wb[i] * w + (A[i] * fa + B[i] * fb);
   return s;
}
HD Spec rgbToSpectrumIllum(const DScene &sc, float r, float g, float b) { return rgbToSpectrumBasis(sc.illum, r, g, b); }
HD void spectrumToXYZ(const DScene &sc, const Spec &s, float &X, float &Y, float &Z) {   // Spectrum.hs:349-355
   float a = 0, b = 0, c = 0;
   BL_UNROLL for (int i = 0; i < NB; ++i) { a = a + sc.cieX[i] * s.v[i]; b = b + sc.cieY[i] * s.v[i]; c = c + sc.cieZ[i] * s.v[i]; }
   X = a / sc.ySum; Y = b / sc.ySum; Z = c / sc.ySum;
}
HD float sY(const DScene &sc, const Spec &s) { float a = 0; BL_UNROLL for (int i = 0; i < NB; ++i) a = a + s.v[i] * sc.cieY[i]; return a / sc.ySum; }

// ------------------------------------------------------------------------------------------ environment maps
HD float perez(const float *p, float sunT, float t, float g, float lvz) {   // SunSky.hs:81-86
   float csg = cosf(g), cst = cosf(sunT);
   float num = (1 + p[0] * expf(p[1] / cosf(t))) * (1 + p[2] * expf(p[3] * g)) + p[4] * csg * csg;
   float den = (1 + p[0] * expf(p[1])) * (1 + p[2] * expf(p[3] * sunT)) + p[4] * cst * cst;
   return lvz * num / den;
}
HD Spec sunSkyEval(const DScene &sc, const blingcu_sunsky &k, V3 dir) {   // SunSky.hs:12-24,67-94
   Spec sky = sConst(0);
   float dz = -dir.z;
   if (!(dz < 1e-4f)) {
      V3 sunDir = mk3(k.sun_dir[0], k.sun_dir[1], k.sun_dir[2]);
      float theta = acosf(dz);
      float gamma = acosf(clampf(dot3(dir, sunDir), -1, 1));
      float x = perez(k.perez_x, k.sun_theta, theta, gamma, k.zenith_x);
      float y = perez(k.perez_y, k.sun_theta, theta, gamma, k.zenith_y);
      float yp = perez(k.perez_Y, k.sun_theta, theta, gamma, k.zenith_Y) * 1e-4f;
      float m1 = (-1.3515f - 1.7703f * x + 5.9114f * y) / (0.0241f + 0.2562f * x - 0.7341f * y);
      float m2 = (0.03f - 31.4424f * x + 30.0717f * y) / (0.0241f + 0.2562f * x - 0.7341f * y);
      float cx = k.s0xyz[0] + m1 * k.s1xyz[0] + m2 * k.s2xyz[0];
      float cy = k.s0xyz[1] + m1 * k.s1xyz[1] + m2 * k.s2xyz[1];
      float cz = k.s0xyz[2] + m1 * k.s1xyz[2] + m2 * k.s2xyz[2];
      float xp = cx * yp / cy, zp = cz * yp / cy;
      float r = 3.240479f * xp - 1.537150f * yp - 0.498535f * zp;
      float g = (-0.969256f) * xp + 1.875991f * yp + 0.041556f * zp;
      float b = 0.055648f * xp - 0.204043f * yp + 1.057311f * zp;
      sky = rgbToSpectrumIllum(sc, r, g, b);
   }
   float d = dot3(mk3(k.sun_disc_dir[0], k.sun_disc_dir[1], k.sun_disc_dir[2]) * mk3(1, 1, -1), dir);
   float stm = sqrtf(hmaxf(0, 1 - 6.955e5f / 1.496e8f));
   if (d > stm) sky = sky + loadSpec(k.sun_radiance.v);
   return sky;
}
HD Spec envEval(const DScene &sc, const blingcu_envmap &e, float u, float v) {   // texMapEval
   if (e.kind == BLINGCU_ENV_CONSTANT) return loadSpec(e.s.v);
   if (e.kind == BLINGCU_ENV_RGBTABLE) {   // IO/Bitmap.hs:22-29
      int x = imax(0, imin(e.nu - 1, (int)floorf((1 - u) * (float)e.nu)));
      int y = imax(0, imin(e.nv - 1, (int)floorf((1 - v) * (float)e.nv)));
      const float *px = e.rgb + 3 * ((size_t)y * e.nu + x);
      return rgbToSpectrumIllum(sc, px[0], px[1], px[2]);
   }
   float phi = u * 2 * BL_PI, theta = v * BL_PI;
   return sunSkyEval(sc, e.sky, sphericalDirection(sinf(theta), cosf(theta), phi));
}
// Montecarlo.hs:53-54: first index with cdf >= u, minus one, clamped. The reference scans linearly; the cdf is
// non-decreasing so a binary search returns the same index.
HD int upperBound(const float *cdf, int len, float u) {
   int lo = 0, hi = len;   // first i in [0,len) with cdf[i] >= u (len if none)
   while (lo < hi) { int mid = (lo + hi) >> 1; if (cdf[mid] >= u) hi = mid; else lo = mid + 1; }
   int idx = (lo == len) ? len - 1 : lo - 1;
   return imin(len - 2, imax(0, idx));
}
HD void sampleContinuous1D(const float *func, const float *cdf, float fi, int n, float u, float &x, float &pdf, int &off) {   // :66-71
   off = upperBound(cdf, n + 1, u);
   pdf = (fi == 0) ? 0.0f : func[off] / fi;
   float du = (u - cdf[off]) / (cdf[off + 1] - cdf[off]);
   x = ((float)off + du) / (float)n;
}
HD void sampleContinuous2D(const blingcu_envmap &e, float u0, float u1, float &u, float &v, float &pdf) {   // :89-92
   float pdf1, pdf0; int imarg, dummy;
   sampleContinuous1D(e.marg_func, e.marg_cdf, e.marg_int, e.nv, u1, v, pdf1, imarg);
   sampleContinuous1D(e.cond_func + (size_t)imarg * e.nu, e.cond_cdf + (size_t)imarg * (e.nu + 1), e.cond_int[imarg], e.nu, u0, u, pdf0, dummy);
   pdf = pdf0 * pdf1;
}
HD float pdfDist2D(const blingcu_envmap &e, float u, float v) {   // :94-104
   int iu = imax(0, imin(e.nu - 1, (int)floorf(u * (float)e.nu)));
   int iv = imax(0, imin(e.nv - 1, (int)floorf(v * (float)e.nv)));
   if (e.marg_int * e.cond_int[iv] == 0) return 0;
   return (e.cond_func[(size_t)iv * e.nu + iu] * e.marg_func[iv]) / (e.cond_int[iv] * e.marg_int);
}

// ------------------------------------------------------------------------------------------ lights (Light.hs)
HD Spec lightLe(const DScene &sc, const blingcu_light &l, V3 rayDir) {   // le :98-106
   if (l.kind != BLINGCU_LIGHT_INFINITE) return sConst(0);
   const blingcu_envmap &e = sc.envs[l.env];
   V3 wh = normalize3(transVector(e.w2l, rayDir));
   return envEval(sc, e, sphericalPhi(wh) / (2 * BL_PI), sphericalTheta(wh) / BL_PI);
}
// emission of an area light seen along direction `wo` from a surface with geometric normal n (lEmit :85-96)
HD bool areaEmits(V3 n, V3 wo) { return dot3(n, wo) > 0; }

struct LightSample { Spec de; V3 wi; Ray testRay; float pdf; bool delta; };
HD void lightSample(const DScene &sc, const blingcu_light &l, V3 p, float eps, V3 n, float u1, float u2, LightSample &o) {   // sample :122-160
   o.de = sConst(0); o.wi = mk3(0, 1, 0); o.pdf = 0; o.delta = false;
   o.testRay.o = mk3(0, 0, 0); o.testRay.d = mk3(0, 1, 0); o.testRay.tmin = 0; o.testRay.tmax = 1;
   switch (l.kind) {
   case BLINGCU_LIGHT_INFINITE: {
      const blingcu_envmap &e = sc.envs[l.env];
      float u, v, mapPdf;
      sampleContinuous2D(e, u1, u2, u, v, mapPdf);
      if (mapPdf == 0) return;
      float phi = u * 2 * BL_PI, theta = v * BL_PI;
      float sint = sinf(theta);
      if (sint == 0) return;
      o.de = envEval(sc, e, u, v);
      o.wi = transVector(e.l2w, sphericalDirection(sinf(theta), cosf(theta), phi));
      o.testRay.o = p; o.testRay.d = o.wi; o.testRay.tmin = eps; o.testRay.tmax = BL_INF;
      o.pdf = mapPdf / (2 * BL_PI * BL_PI * sint);
      return;
   }
   case BLINGCU_LIGHT_DIRECTIONAL: {
      V3 d = mk3(l.v[0], l.v[1], l.v[2]);
      o.de = sScale(loadSpec(l.s.v), absDot(n, d)); o.wi = d;
      o.testRay.o = p; o.testRay.d = d; o.testRay.tmin = eps; o.testRay.tmax = BL_INF;
      o.pdf = 1; o.delta = true;
      return;
   }
   case BLINGCU_LIGHT_POINT: {
      V3 pos = mk3(l.v[0], l.v[1], l.v[2]);
      o.de = sScale(loadSpec(l.s.v), 1 / sqLen(pos - p)); o.wi = normalize3(pos - p);
      o.testRay.o = p; o.testRay.d = pos - p; o.testRay.tmin = eps; o.testRay.tmax = BL_INF;
      o.pdf = 1; o.delta = true;
      return;
   }
   default: {   // area light, sampled in light space (Q11)
      const blingcu_shape &s = sc.shapes[l.shape];
      V3 pl = transPoint(s.w2o, p);
      V3 ps, ns; sampleShape(s, pl, u1, u2, ps, ns);
      V3 wi = normalize3(ps - pl);
      o.pdf = shapePdf(s, pl, wi);
      Ray ray; ray.o = pl; ray.d = wi; ray.tmin = eps; ray.tmax = len3(ps - pl) - eps;
      if (dot3(ns, wi) < 0) o.de = loadSpec(l.s.v);
      o.wi = transVector(s.o2w, wi);
      o.testRay = transRay(s.o2w, ray);
      return;
   }
   }
}
HD float lightPdf(const DScene &sc, const blingcu_light &l, V3 p, V3 wiW) {   // pdf :215-229
   if (l.kind == BLINGCU_LIGHT_INFINITE) {
      const blingcu_envmap &e = sc.envs[l.env];
      V3 w = transVector(e.w2l, wiW);
      float phi = sphericalPhi(w), theta = sphericalTheta(w);
      float sint = sinf(theta);
      if (sint == 0) return 0;
      return pdfDist2D(e, phi / (2 * BL_PI), theta / BL_PI) / (2 * BL_PI * BL_PI * sint);
   }
   if (l.kind == BLINGCU_LIGHT_AREA) {
      const blingcu_shape &s = sc.shapes[l.shape];
      return shapePdf(s, transPoint(s.w2o, p), transVector(s.w2o, wiW));
   }
   return 0;
}

// ------------------------------------------------------------------------------------------ out-of-line general versions
// The general (any material kind) shade kernels -- textured queue, direct-lighting integrator -- inline every BxDF with its
// 16-band arithmetic at each call site: ~700 KB of code, and ncu shows them starved by instruction fetch
// (stall no_instruction 47 warps per issue, profiles/r01_texshade.md). They call these shared copies instead; the
// per-kind kernels (M::GEN == false, textured or not) keep the inlined, specialised code.
HDNI void sampleBsdfGeneral(const Bsdf &b, V3 wo, float uc, float u1, float u2, BsdfSample &o) { sampleBsdf<GenMat>(b, wo, uc, u1, u2, o); }
HDNI void sampleBsdfSpecularGeneral(const Bsdf &b, int flags, V3 wo, BsdfSample &o) { sampleBsdfSpecular<GenMat>(b, flags, wo, o); }
HDNI void evalBsdfGeneral(const Bsdf &b, V3 wo, V3 wi, Spec &f) { f = evalBsdf<GenMat>(b, wo, wi); }
HDNI float bsdfPdfGeneral(const Bsdf &b, V3 wo, V3 wi) { return bsdfPdf<GenMat>(b, wo, wi); }
HDNI void lightSampleGeneral(const DScene &sc, const blingcu_light &l, V3 p, float eps, V3 n, float u1, float u2, LightSample &o) { lightSample(sc, l, p, eps, n, u1, u2, o); }
HDNI void makeBsdfGeneral(const DScene &sc, const SurfaceHit &sh, const DG &dgs, Bsdf &b, Spec *scratch) { makeBsdf<MatOf<SK_GENERAL> >(sc, sh, dgs, b, scratch); }
// dispatchers: M::GEN is a compile-time constant
template <class M> HD void sampleBsdfOf(const Bsdf &b, V3 wo, float uc, float u1, float u2, BsdfSample &o) { if (M::GEN) sampleBsdfGeneral(b, wo, uc, u1, u2, o); else sampleBsdf<M>(b, wo, uc, u1, u2, o); }
template <class M> HD Spec evalBsdfOf(const Bsdf &b, V3 wo, V3 wi) { if (M::GEN) { Spec f; evalBsdfGeneral(b, wo, wi, f); return f; } return evalBsdf<M>(b, wo, wi); }
template <class M> HD float bsdfPdfOf(const Bsdf &b, V3 wo, V3 wi) { if (M::GEN) return bsdfPdfGeneral(b, wo, wi); return bsdfPdf<M>(b, wo, wi); }
template <class M> HD void lightSampleOf(const DScene &sc, const blingcu_light &l, V3 p, float eps, V3 n, float u1, float u2, LightSample &o) { if (M::GEN) lightSampleGeneral(sc, l, p, eps, n, u1, u2, o); else lightSample(sc, l, p, eps, n, u1, u2, o); }
template <class M> HD void makeBsdfOf(const DScene &sc, const SurfaceHit &sh, const DG &dgs, Bsdf &b, Spec *scratch) { if (M::GEN) makeBsdfGeneral(sc, sh, dgs, b, scratch); else makeBsdf<M>(sc, sh, dgs, b, scratch); }

// ------------------------------------------------------------------------------------------ camera (Camera.hs:49-76)
HD Ray fireRay(const blingcu_camera &c, float ix, float iy, float lu, float lv) {
   Ray ray; ray.tmin = 0; ray.tmax = BL_INF;
   if (c.kind == BLINGCU_CAM_ENVIRONMENT) {
      float t = BL_PI * iy / c.env_sy, p = 2 * BL_PI * ix / c.env_sx;
      ray.o = mk3(0, 0, 0); ray.d = mk3(sinf(t) * cosf(p), cosf(t), sinf(t) * sinf(p));
      return transRay(c.cam2world, ray);
   }
   ray.o = mk3(0, 0, 0); ray.d = normalize3(transPoint(c.raster2cam, mk3(ix, iy, 0)));
   if (c.lens_radius > 0) {
      float dx, dy; concentricSampleDisk(lu, lv, dx, dy);
      V3 ro = mk3(dx * c.lens_radius, dy * c.lens_radius, 0);
      V3 pFocus = rayAt(ray, c.focal_distance / ray.d.z);
      ray.o = ro; ray.d = normalize3(pFocus - ro);
   }
   return transRay(c.cam2world, ray);
}

}  // namespace bl
