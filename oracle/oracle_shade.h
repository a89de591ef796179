// ORACLE (test infrastructure, NOT product code). See oracle_math.h header. PARITY UNPINNED.
// BxDFs, BSDF, materials, textures, lights, environment maps -- restating Reflection*.hs,
// Fresnel.hs, Material.hs, Texture.hs, Light.hs, Montecarlo.hs, SunSky.hs, Spectrum.hs.
#pragma once
#include "oracle_scene.h"

namespace orc {

// ----------------------------------------------------------------------------- Montecarlo.hs
static inline void concentricSampleDisk(float u1, float u2, float &dx, float &dy) {  // :164-181
   float sx = u1 * 2 - 1, sy = u2 * 2 - 1;
   if (sx == 0 && sy == 0) { dx = 0; dy = 0; return; }
   float r, thp;
   if (sx >= -sy) {
      if (sx > sy) { r = sx; thp = (sy > 0) ? sy / sx : 8 + sy / sx; }
      else { r = sy; thp = 2 - sx / sy; }
   } else if (sx <= sy) { r = -sx; thp = 4 - sy / (-sx); }
   else { r = -sy; thp = 6 + sx / (-sy); }
   float theta = thp * kPi / 4;
   dx = r * std::cos(theta); dy = r * std::sin(theta);
}
static inline V3 cosineSampleHemisphere(float u1, float u2) {  // :147-150
   float x, y; concentricSampleDisk(u1, u2, x, y);
   return mk(x, y, std::sqrt(hmax(0, 1 - x * x - y * y)));
}
static inline float uniformConePdf(float cosThetaMax) {  // :127-131
   if (cosThetaMax >= 1) return 0;
   return 1 / (kTwoPi * (1 - cosThetaMax));
}
static inline V3 uniformSampleCone(const Frame &f, float cosThetaMax, float u1, float u2) {  // :133-145
   float cosTheta = lerpf(u1, cosThetaMax, 1.0f);
   float sinTheta = std::sqrt(1 - cosTheta * cosTheta);
   float phi = u2 * kTwoPi;
   float a = std::cos(phi) * sinTheta, b = std::sin(phi) * sinTheta;
   return (f.s * mk(a, a, a) + f.t * mk(b, b, b)) + f.n * mk(cosTheta, cosTheta, cosTheta);
}
static inline V3 uniformSampleSphere(float u1, float u2) {  // :183-188
   float u = u1 * 2 - 1;
   float s = std::sqrt(1 - (u * u));
   float omega = u2 * 2 * kPi;
   return mk(s * std::cos(omega), s * std::sin(omega), u);
}
static inline float powerHeuristic(int nf, float fPdf, int ng, float gPdf) {  // :113-116
   float f = (float)nf * fPdf, g = (float)ng * gPdf;
   return (f * f) / (f * f + g * g);
}
static inline void remapRand(int segs, float u, int &seg, float &up) {  // Math.hs:120-124
   float segsf = (float)segs;
   seg = std::min(segs - 1, (int)std::floor(u * segsf));
   up = (u - (float)seg / segsf) * segsf;
}

// ----------------------------------------------------------------------------- Shape.hs sampling
static inline bool insideSphere(float r, V3 pt) { return sqLen(pt) - r * r < 1e-4f; }  // :330-331

static inline void sampleShapeAny(const blingcu_shape &s, float u1, float u2, V3 &p, V3 &n) {  // sampleShape' :379-409
   const float *P = s.p;
   switch (s.kind) {
   case BLINGCU_SHAPE_BOX: {
      V3 pmin = mk(P[0], P[1], P[2]), pmax = mk(P[3], P[4], P[5]);
      int axis, nf; float u1p, u2p;
      remapRand(3, u1, axis, u1p);
      remapRand(2, u2, nf, u2p);
      n = setc(axis, (float)nf * 2 - 1, mk(0, 0, 0));
      int oa0 = (axis + 1) % 3, oa1 = (axis + 2) % 3;
      V3 base = (nf == 0) ? pmin : pmax;
      p = setc(oa0, lerpf(u1p, pmin[oa0], pmax[oa0]), setc(oa1, lerpf(u2p, pmin[oa1], pmax[oa1]), base));
      return;
   }
   case BLINGCU_SHAPE_CYLINDER: {
      float r = P[0], z0 = P[1], z1 = P[2];
      float z = lerpf(u1, z0, z1), phi = lerpf(u2, 0, kTwoPi);
      p = mk(r * std::cos(phi), r * std::sin(phi), z);
      n = normalize(mk(p.x, p.y, 0));
      return;
   }
   case BLINGCU_SHAPE_DISK: {
      float h = P[0], rmax = P[1], rmin = P[2], phiMax = P[3];
      float r = lerpf(u1, rmin, rmax), phi = lerpf(u2, 0, phiMax);
      p = mk(r * std::cos(phi), r * std::sin(phi), h);
      n = mk(0, 0, -1);
      return;
   }
   case BLINGCU_SHAPE_QUAD: {
      p = mk(lerpf(u1, -P[0], P[0]), lerpf(u2, -P[1], P[1]), 0);
      n = mk(0, 0, -1);  // Q2: sampled normal is -z while the hit normal is +z
      return;
   }
   default: {
      V3 q = uniformSampleSphere(u1, u2);
      p = q * mk(P[0], P[0], P[0]); n = q;
      return;
   }
   }
}
static inline void sampleShape(const blingcu_shape &s, V3 p, float u1, float u2, V3 &ps, V3 &ns) {  // :333-377
   if (s.kind == BLINGCU_SHAPE_SPHERE) {
      float r = s.p[0];
      if (insideSphere(r, p)) { sampleShapeAny(s, u1, u2, ps, ns); return; }
      V3 dn = normalize(-p);
      Frame cs = coordinateSystem(dn);
      float cosThetaMax = std::sqrt(hmax(0, 1 - (r * r) / sqLen(p)));
      V3 d = uniformSampleCone(cs, cosThetaMax, u1, u2);
      Ray ray{p, d, 0, kInf};
      ShapeHit sh;
      if (shapeIntersect(s, ray, sh)) ps = rayAt(ray, sh.t);
      else ps = dn * mk(r, r, r);
      ns = normalize(ps);
      return;
   }
   sampleShapeAny(s, u1, u2, ps, ns);
}
static inline float generalPdf(const blingcu_shape &s, V3 p, V3 wi) {  // :346-350
   Ray r{p, wi, 1e-3f, kInf};
   ShapeHit sh;
   if (!shapeIntersect(s, r, sh)) return 0;
   float pd = sqLen(p - rayAt(r, sh.t)) / (absDot(sh.dg.n, -wi) * shapeArea(s));
   return std::isinf(pd) ? 0 : pd;
}
static inline float shapePdf(const blingcu_shape &s, V3 p, V3 wi) {  // :333-344
   if (s.kind == BLINGCU_SHAPE_SPHERE) {
      float r = s.p[0];
      if (insideSphere(r, p)) return generalPdf(s, p, wi);
      float sinThetaMax2 = r * r / sqLen(p);
      float cosThetaMax = std::sqrt(hmax(0, 1 - sinThetaMax2));
      return uniformConePdf(cosThetaMax);
   }
   return generalPdf(s, p, wi);
}

// ----------------------------------------------------------------------------- Fresnel.hs
static inline Spec frDielectric(float etai, float etat, float cosi) {  // :31-55
   float c0 = hmax(0, 1 - cosi * cosi);
   float costp = (cosi > 0) ? c0 / (etat * etat) : c0 * (etat * etat);
   float cost = std::sqrt(1 - clampf(costp, 0, 1));
   float acosi = std::fabs(cosi);
   Spec eta = sConst(etat) / sConst(etai);
   Spec costS = sConst(cost);
   Spec rParlP = sScale(eta, acosi);
   Spec rParl = (costS - rParlP) / (costS + rParlP);
   Spec rPerpP = eta * costS;
   Spec rPerp = (sConst(acosi) - rPerpP) / (sConst(acosi) + rPerpP);
   return sScale(rParl * rParl + rPerp * rPerp, 0.5f);
}
static inline Spec frConductor(const Spec &eta, const Spec &k, float cosi) {  // :58-70
   float acosi = std::fabs(cosi);
   Spec ec2 = sScale(eta, 2 * acosi);
   Spec tmpF = eta * eta + k * k;
   Spec tmp = sScale(tmpF, acosi * acosi);
   Spec rPer2 = (tmpF - ec2 + sConst(acosi * acosi)) / (tmpF + ec2 + sConst(acosi * acosi));
   Spec rPar2 = (tmp - ec2 + sConst(1)) / (tmp + ec2 + sConst(1));
   return (rPer2 + rPar2) / sConst(2);
}

// ----------------------------------------------------------------------------- Reflection.hs helpers
static inline float cosTheta(V3 v) { return v.z; }
static inline float absCosTheta(V3 v) { return std::fabs(v.z); }
static inline float sinTheta2(V3 v) { return hmax(0, 1 - v.z * v.z); }
static inline float sinTheta(V3 v) { return std::sqrt(sinTheta2(v)); }
static inline float cosPhi(V3 v) { float s = sinTheta(v); return s == 0 ? 1 : clampf(v.x / s, -1, 1); }
static inline float sinPhi(V3 v) { float s = sinTheta(v); return s == 0 ? 0 : clampf(v.y / s, -1, 1); }
static inline bool sameHemisphere(V3 a, V3 b) { return a.z * b.z > 0; }
static inline V3 toSameHemisphere(V3 wo, V3 wi) { return wo.z < 0 ? mk(wi.x, wi.y, -wi.z) : wi; }

enum { BX_REFLECTION = 1, BX_TRANSMISSION = 2, BX_DIFFUSE = 4, BX_GLOSSY = 8, BX_SPECULAR = 16 };  // :94-107
enum { K_LAMBERT, K_ORENNAYAR, K_SPECREFL, K_SPECTRANS, K_MICROFACET, K_FRESNELBLEND };
enum { FR_NOOP, FR_DIELECTRIC, FR_CONDUCTOR };

struct BxDF {
   int kind; int type;
   Spec r;            // reflectance / transmittance
   float a, b;        // OrenNayar A,B
   int fr; float etai, etat; Spec eta, k;  // Fresnel
   float e;           // Blinn exponent
   Spec rs, ra; float ex, ey, depth;   // FresnelBlend (Microfacet.hs:56-108): r = rd, specular, absorption, Anisotropic ex ey, coat depth
   bool flip;         // brdfToBtdf (Reflection.hs:188-195): evaluated / sampled through the other hemisphere
};
static inline V3 otherHemisphere(V3 w) { return mk(w.x, w.y, -w.z); }

static inline Spec fresnel(const BxDF &b, float cosi) {
   if (b.fr == FR_DIELECTRIC) return frDielectric(b.etai, b.etat, cosi);
   if (b.fr == FR_CONDUCTOR) return frConductor(b.eta, b.k, cosi);
   return sConst(1);
}

// Microfacet.hs:113-120
static inline float mfG(V3 wo, V3 wi, V3 wh) {
   float nDotWh = absCosTheta(wh), nDotWo = absCosTheta(wo), nDotWi = absCosTheta(wi), woDotWh = absDot(wo, wh);
   return hmin(1, hmin(2 * nDotWh * nDotWo / woDotWh, 2 * nDotWh * nDotWi / woDotWh));
}
static inline float blinnD(float e, V3 wh) { return (e + 2) * kInvTwoPi * std::pow(absCosTheta(wh), e); }      // :194-195
static inline float blinnPdf(float e, V3 wh) { return (e + 1) * std::pow(absCosTheta(wh), e) * kInvTwoPi; }    // :146-147

// Anisotropic distribution (Microfacet.hs:140-173,184-192)
static inline float anisoE(float ex, float ey, V3 wh, float d) { return (ex * wh.x * wh.x + ey * wh.y * wh.y) / d; }
static inline float anisoPdf(float ex, float ey, V3 wh) {
   float costh = absCosTheta(wh);
   float e = anisoE(ex, ey, wh, hmax(0, 1 - costh * costh));
   return std::sqrt((ex + 1) * (ey + 1)) * kInvTwoPi * std::pow(costh, e);
}
static inline float anisoD(float ex, float ey, V3 wh) {
   float costh = absCosTheta(wh);
   float d = 1 - costh * costh;
   if (d == 0) return 0;
   return std::sqrt((ex + 2) * (ey + 2)) * kInvTwoPi * std::pow(costh, anisoE(ex, ey, wh, d));
}
static inline void anisoSample(float ex, float ey, float u1, float u2, V3 &wh, float &pdf) {
   auto quad = [&](float u, float &p, float &c) {   // smpFirstQuadrand
      p = (ex == ey) ? kPi * u * 0.5f : std::atan(std::sqrt((ex + 1) / (ey + 1)) * std::tan(kPi * u * 0.5f));
      float cp = std::cos(p), sp = std::sin(p);
      c = std::pow(u2, 1 / (ex * cp * cp + ey * sp * sp + 1));
   };
   float phi, cost, p;
   if (u1 < 0.25f) { quad(4 * u1, p, cost); phi = p; }
   else if (u1 < 0.50f) { quad(4 * (0.5f - u1), p, cost); phi = kPi - p; }
   else if (u1 < 0.75f) { quad(4 * (u1 - 0.5f), p, cost); phi = p + kPi; }
   else { quad(4 * (1 - u1), p, cost); phi = 2 * kPi - p; }
   float sint = std::sqrt(hmax(0, 1 - cost * cost));
   wh = sphericalDirection(sint, cost, phi);
   float ds = 1 - cost * cost;
   float e = anisoE(ex, ey, wh, ds);
   pdf = std::sqrt((ex + 1) * (ey + 1)) * (kInvTwoPi * std::pow(cost, e));
}
static inline V3 halfUp(V3 wi, V3 wo) { V3 w = normalize(wi + wo); return w.z < 0 ? -w : w; }
// FresnelBlend `e wo wi` (Microfacet.hs:65-84)
static inline Spec fresnelBlendEval(const BxDF &b, V3 wo, V3 wi) {
   float costi = absCosTheta(wi), costo = absCosTheta(wo);
   Spec a = sConst(1);
   if (b.depth > 0) { float sc = -(b.depth * (costi + costo) / (costi * costo)); for (int i = 0; i < NB; ++i) a.v[i] = std::exp(b.ra.v[i] * sc); }
   float ds = (costo * 28 / 23 * kPi) * (1 - std::pow(1 - 0.5f * costi, 5.0f)) * (1 - std::pow(1 - 0.5f * costo, 5.0f));
   Spec diff = sScale(a * b.r * (sConst(1) - b.rs), ds);
   V3 wh = halfUp(wi, wo);
   float costih = absDot(wi, wh);
   Spec schlick = b.rs + sScale(sConst(1) - b.rs, std::pow(1 - costih, 5.0f));
   Spec spec = sScale(schlick, anisoD(b.ex, b.ey, wh) * costo / (4 * costih * hmax(costi, costo)));
   return diff + spec;
}

static inline Spec orenNayar(const BxDF &b, V3 wo, V3 wi) {  // Diffuse.hs:52-65
   float sinti = sinTheta(wi), sinto = sinTheta(wo);
   float sina, tanb;
   if (absCosTheta(wi) > absCosTheta(wo)) { sina = sinto; tanb = sinti / absCosTheta(wi); }
   else { sina = sinti; tanb = sinto / absCosTheta(wo); }
   float maxcos = 0;
   if (sinti > 1e-4f && sinto > 1e-4f) {
      float sinpi = sinPhi(wi), cospi = cosPhi(wi), sinpo = sinPhi(wo), cospo = cosPhi(wo);
      maxcos = hmax(0, cospi * cospo + sinpi * sinpo);
   }
   return sScale(b.r, b.a + b.b * maxcos * sina * tanb);
}

// bxdfEval b wo wi (the CALLER flips for non-adjoint, Reflection.hs:310,330)
static inline Spec bxdfEval(const BxDF &b, V3 wo, V3 wi) {
   if (b.flip) wi = otherHemisphere(wi);   // e wo wi = bxdfEval brdf wo (otherHemisphere wi)
   switch (b.kind) {
   case K_LAMBERT: return sScale(b.r, kInvPi * absCosTheta(wo));                          // Diffuse.hs:24-26
   case K_ORENNAYAR: return sScale(orenNayar(b, wo, wi), kInvPi * absCosTheta(wo));      // Diffuse.hs:50
   case K_MICROFACET: {                                                                  // Microfacet.hs:20-33
      float costo = absCosTheta(wo), costi = absCosTheta(wi);
      if (costi == 0 || costo == 0) return sConst(0);
      V3 whp = wi + wo;
      if (whp.x == 0 && whp.y == 0 && whp.z == 0) return sConst(0);
      V3 wh = normalize(whp);
      if (cosTheta(wh) < 0) return sConst(0);
      float costh = dot(wi, wh);
      float x = blinnD(b.e, wh) * mfG(wo, wi, wh) / (4 * costi);
      return sScale(b.r * fresnel(b, costh), x);
   }
   case K_FRESNELBLEND: return fresnelBlendEval(b, wo, wi);
   default: return sConst(0);  // specular: Specular.hs:17,31
   }
}
static inline float cosPdf(V3 wo, V3 wi) { return sameHemisphere(wo, wi) ? kInvPi * absCosTheta(wi) : 0; }  // Diffuse.hs:9-12
static inline float bxdfPdf(const BxDF &b, V3 wo, V3 wi) {
   if (b.flip) wi = otherHemisphere(wi);
   switch (b.kind) {
   case K_LAMBERT: case K_ORENNAYAR: return cosPdf(wo, wi);
   case K_MICROFACET: {  // Microfacet.hs:35-41
      V3 whp = wo + wi;
      if (sqLen(whp) == 0) return 0;
      V3 wh = normalize(whp);
      if (cosTheta(wh) < 0) return 0;
      return blinnPdf(b.e, wh) / (4 * absDot(wo, wh));
   }
   case K_FRESNELBLEND: {   // Microfacet.hs:103-108
      if (!sameHemisphere(wo, wi)) return 0;
      V3 wh = halfUp(wi, wo);
      return 0.5f * (absCosTheta(wi) * kInvPi + anisoPdf(b.ex, b.ey, wh) / (4 * absDot(wo, wh)));
   }
   default: return 0;
   }
}
// bxdfSample b adj=False wo u -> (f, wi, pdf)
static inline void bxdfSample(const BxDF &b, V3 wo, float u1, float u2, Spec &f, V3 &wi, float &pdf) {
   switch (b.kind) {
   case K_LAMBERT: {  // Diffuse.hs:14-22
      wi = toSameHemisphere(wo, cosineSampleHemisphere(u1, u2));
      if (sameHemisphere(wo, wi)) { f = b.r; pdf = cosPdf(wo, wi); }
      else { f = sConst(0); wi = wo; pdf = 0; }
      if (b.flip) wi = otherHemisphere(wi);   // s adj wo u = (f, otherHemisphere wi, pdf)
      return;
   }
   case K_ORENNAYAR: {  // Diffuse.hs:38-42
      wi = toSameHemisphere(wo, cosineSampleHemisphere(u1, u2));
      if (sameHemisphere(wo, wi)) { f = orenNayar(b, wo, wi); pdf = cosPdf(wo, wi); }
      else { f = sConst(0); pdf = 0; }
      if (b.flip) wi = otherHemisphere(wi);
      return;
   }
   case K_SPECREFL: {  // Specular.hs:11-20
      f = b.r * fresnel(b, cosTheta(wo));
      wi = mk(-wo.x, -wo.y, wo.z);
      pdf = 1;
      return;
   }
   case K_SPECTRANS: {  // Specular.hs:34-57 (adj = False)
      bool entering = cosTheta(wo) > 0;
      float ei = entering ? b.etai : b.etat, et = entering ? b.etat : b.etai;
      float sini2 = sinTheta2(wo);
      float eta = ei / et, eta2 = eta * eta, sint2 = eta2 * sini2;
      if (sint2 >= 1) { f = sConst(0); wi = wo; pdf = 0; return; }
      float c = std::sqrt(hmax(0, 1 - sint2));
      float cost = entering ? -c : c;
      wi = mk(eta * (-wo.x), eta * (-wo.y), cost);
      Spec fr = frDielectric(ei, et, cost);  // Q4: transmitted cosine in the non-adjoint case
      Spec fp = (sConst(1) - fr) * b.r;
      f = sScale(fp, eta2);
      pdf = 1;
      return;
   }
   case K_FRESNELBLEND: {   // Microfacet.hs:86-101, adj = False
      float pdfp; V3 wh;
      if (u1 < 0.5f) {
         wi = toSameHemisphere(wo, cosineSampleHemisphere(u1 * 2, u2));
         wh = halfUp(wi, wo);
         pdfp = anisoPdf(b.ex, b.ey, wh);
      } else {
         anisoSample(b.ex, b.ey, 2 * (u1 - 0.5f), u2, wh, pdfp);
         wi = scl(2, scl(dot(wo, wh), wh)) - wo;   // 2 * wo `dot` wh *# wh - wo  (infixl 9 for both `dot` and *#)
      }
      if (pdfp == 0) { f = sConst(0); pdf = 0; return; }
      pdf = 0.5f * (absCosTheta(wi) * kInvPi + pdfp / (4 * absDot(wo, wh)));
      f = sScale(fresnelBlendEval(b, wo, wi), 1 / pdf);
      return;
   }
   case K_MICROFACET: {  // Microfacet.hs:43-54, Blinn sample :175-182
      float cost = std::pow(u1, 1 / (b.e + 1));
      float sint = std::sqrt(hmax(0, 1 - cost * cost));
      float phi = u2 * 2 * kPi;
      V3 whp = sphericalDirection(sint, cost, phi);
      float ff = std::pow(cost, b.e) * kInvTwoPi;
      float d = (b.e + 2) * ff, dpdf = (b.e + 1) * ff;
      V3 wh = (cosTheta(whp) < 0) ? -whp : whp;
      float costH = dot(wo, wh);
      wi = scl(2 * costH, wh) - wo;   // (2 * wo `dot` wh) *# wh - wo
      if (!sameHemisphere(wo, wi)) { f = sConst(0); wi = wo; pdf = 0; return; }
      float fact = d * std::fabs(costH) / dpdf * mfG(wo, wi, wh);
      Spec fp = b.r * fresnel(b, costH);
      f = sScale(fp, fact / absCosTheta(wi));   // Q7
      pdf = dpdf / (4 * std::fabs(costH));
      return;
   }
   }
}

// Reflection.hs:191-225
struct Bsdf {
   int n;
   BxDF bx[2];
   Frame cs;
   V3 p, ng;
};
struct BsdfSample { int type; float pdf; Spec f; V3 wi; };

static inline bool isRefl(const BxDF &b) { return (b.type & BX_REFLECTION) != 0; }
static inline bool isTrans(const BxDF &b) { return (b.type & BX_TRANSMISSION) != 0; }
static inline bool isSpec(const BxDF &b) { return (b.type & BX_SPECULAR) != 0; }

// Reflection.hs:278-316 with adj = False, flags = bxdfAll
static inline BsdfSample sampleBsdf(const Bsdf &bsdf, V3 woW, float uComp, float uDir1, float uDir2) {
   BsdfSample empty{BX_REFLECTION | BX_DIFFUSE, 0, sConst(0), mk(0, 1, 0)};
   int cntm = bsdf.n;
   if (cntm == 0) return empty;
   V3 wo = worldToLocal(bsdf.cs, woW);
   float cntf = (float)cntm, invCnt = 1 / cntf;
   int sNum = std::max(0, std::min(cntm - 1, (int)std::floor(uComp * cntf)));
   const BxDF &bx = bsdf.bx[sNum];
   Spec fSample = sConst(0); V3 wi = mk(0, 1, 0); float pdfp = 0;
   bxdfSample(bx, wo, uDir1, uDir2, fSample, wi, pdfp);
   V3 wiW = localToWorld(bsdf.cs, wi);
   float sideTest = dot(wiW, bsdf.ng) / dot(woW, bsdf.ng);
   if (pdfp == 0 || sideTest == 0) return empty;
   bool wantTrans = sideTest < 0;
   if (!(wantTrans ? isTrans(bx) : isRefl(bx))) return empty;
   if (isSpec(bx)) return BsdfSample{bx.type, pdfp * invCnt, sScale(fSample, cntf), wiW};
   if (cntm == 1) return BsdfSample{bx.type, pdfp, fSample, wiW};
   float pdfSum = 0; Spec fOthers = sConst(0);
   for (int i = 0; i < cntm; ++i) {
      if (i == sNum) continue;
      pdfSum = pdfSum + bxdfPdf(bsdf.bx[i], wo, wi);
      if (wantTrans ? isTrans(bsdf.bx[i]) : isRefl(bsdf.bx[i])) fOthers = fOthers + bxdfEval(bsdf.bx[i], wi, wo);
   }
   float pdf = (pdfp + pdfSum) * invCnt;
   Spec fSum = sScale(sScale(fSample, pdfp) + fOthers, 1 / pdf);
   return BsdfSample{bx.type, pdf, fSum, wiW};
}
// Reflection.hs:278-316 with adj = False and a component filter: bsm = filter (`bxdfMatches` flags) bs, where
// bxdfMatches b flags = (type b .&. flags) == type b (:119-121,185-186). DirectLighting.hs:50-52 calls it with
// [Specular, Reflection] and [Specular, Transmission].
static inline BsdfSample sampleBsdfFlags(const Bsdf &bsdf, int flags, V3 woW, float uComp, float uDir1, float uDir2) {
   BsdfSample empty{BX_REFLECTION | BX_DIFFUSE, 0, sConst(0), mk(0, 1, 0)};
   const BxDF *bsm[2]; int cntm = 0;
   for (int i = 0; i < bsdf.n; ++i) if ((bsdf.bx[i].type & flags) == bsdf.bx[i].type) bsm[cntm++] = &bsdf.bx[i];
   if (cntm == 0) return empty;
   V3 wo = worldToLocal(bsdf.cs, woW);
   float cntf = (float)cntm, invCnt = 1 / cntf;
   int sNum = std::max(0, std::min(cntm - 1, (int)std::floor(uComp * cntf)));
   const BxDF &bx = *bsm[sNum];
   Spec fSample = sConst(0); V3 wi = mk(0, 1, 0); float pdfp = 0;
   bxdfSample(bx, wo, uDir1, uDir2, fSample, wi, pdfp);
   V3 wiW = localToWorld(bsdf.cs, wi);
   float sideTest = dot(wiW, bsdf.ng) / dot(woW, bsdf.ng);
   if (pdfp == 0 || sideTest == 0) return empty;
   bool wantTrans = sideTest < 0;
   if (!(wantTrans ? isTrans(bx) : isRefl(bx))) return empty;
   if (isSpec(bx)) return BsdfSample{bx.type, pdfp * invCnt, sScale(fSample, cntf), wiW};
   if (cntm == 1) return BsdfSample{bx.type, pdfp, fSample, wiW};
   float pdfSum = 0; Spec fOthers = sConst(0);
   for (int i = 0; i < cntm; ++i) {
      if (i == sNum) continue;
      pdfSum = pdfSum + bxdfPdf(*bsm[i], wo, wi);
      if (wantTrans ? isTrans(*bsm[i]) : isRefl(*bsm[i])) fOthers = fOthers + bxdfEval(*bsm[i], wi, wo);
   }
   float pdf = (pdfp + pdfSum) * invCnt;
   return BsdfSample{bx.type, pdf, sScale(sScale(fSample, pdfp) + fOthers, 1 / pdf), wiW};
}
// Reflection.hs:318-332 with adj = False
static inline Spec evalBsdf(const Bsdf &bsdf, V3 woW, V3 wiW) {
   float cosWo = dot(woW, bsdf.ng);
   float sideTest = dot(wiW, bsdf.ng) / cosWo;
   if (sideTest == 0) return sConst(0);
   if (std::fabs(cosWo) < 1e-5f) return sConst(0);
   bool wantTrans = sideTest < 0;
   V3 wo = worldToLocal(bsdf.cs, woW), wi = worldToLocal(bsdf.cs, wiW);
   Spec f = sConst(0);
   for (int i = 0; i < bsdf.n; ++i)
      if (wantTrans ? isTrans(bsdf.bx[i]) : isRefl(bsdf.bx[i])) f = f + bxdfEval(bsdf.bx[i], wi, wo);
   return f;
}
// Reflection.hs:251-257 (Q6)
static inline float bsdfPdf(const Bsdf &bsdf, V3 woW, V3 wiW) {
   if (bsdf.n == 0) return 0;
   V3 wo = worldToLocal(bsdf.cs, woW), wi = worldToLocal(bsdf.cs, wiW);
   float s = 0;
   for (int i = 0; i < bsdf.n; ++i) s = s + bxdfPdf(bsdf.bx[i], wo, wi);
   return s / (float)bsdf.n;
}

// ----------------------------------------------------------------------------- adjoint (light-transport) variants: SURVEY 8(f)4
// bxdfSample b adj=True wo u: what the light tracer draws (Renderer/LightTracer.hs:84 sampleAdjBsdf). Differences from adj = False,
// BxDF by BxDF: Lambertian / OrenNayar scale by |cos wo / cos wi| (Diffuse.hs:20-22,44-49); specTrans takes the Fresnel term at
// the INCIDENT cosine and scales by |cos wo / cost| instead of eta^2 (Specular.hs:52-57); the microfacet weight divides by
// |cos wo| instead of |cos wi| (Microfacet.hs:52-54); FresnelBlend evaluates `e wi wo` (:92); specRefl is the same (Specular.hs:19).
static inline void bxdfSampleAdj(const BxDF &b, V3 wo, float u1, float u2, Spec &f, V3 &wi, float &pdf) {
   switch (b.kind) {
   case K_LAMBERT: case K_ORENNAYAR: {
      wi = toSameHemisphere(wo, cosineSampleHemisphere(u1, u2));
      if (sameHemisphere(wo, wi)) {
         Spec r = (b.kind == K_LAMBERT) ? b.r : orenNayar(b, wo, wi);
         f = sScale(r, std::fabs(cosTheta(wo) / cosTheta(wi))); pdf = cosPdf(wo, wi);
      } else { f = sConst(0); if (b.kind == K_LAMBERT) wi = wo; pdf = 0; }
      if (b.flip) wi = otherHemisphere(wi);
      return;
   }
   case K_SPECTRANS: {
      bool entering = cosTheta(wo) > 0;
      float ei = entering ? b.etai : b.etat, et = entering ? b.etat : b.etai;
      float sini2 = sinTheta2(wo);
      float eta = ei / et, eta2 = eta * eta, sint2 = eta2 * sini2;
      if (sint2 >= 1) { f = sConst(0); wi = wo; pdf = 0; return; }
      float c = std::sqrt(hmax(0, 1 - sint2));
      float cost = entering ? -c : c;
      wi = mk(eta * (-wo.x), eta * (-wo.y), cost);
      Spec fr = frDielectric(ei, et, cosTheta(wo));
      Spec fp = (sConst(1) - fr) * b.r;
      f = sScale(fp, std::fabs(cosTheta(wo) / cost));
      pdf = 1;
      return;
   }
   case K_FRESNELBLEND: {
      float pdfp; V3 wh;
      if (u1 < 0.5f) {
         wi = toSameHemisphere(wo, cosineSampleHemisphere(u1 * 2, u2));
         wh = halfUp(wi, wo);
         pdfp = anisoPdf(b.ex, b.ey, wh);
      } else {
         anisoSample(b.ex, b.ey, 2 * (u1 - 0.5f), u2, wh, pdfp);
         wi = scl(2, scl(dot(wo, wh), wh)) - wo;
      }
      if (pdfp == 0) { f = sConst(0); pdf = 0; return; }
      pdf = 0.5f * (absCosTheta(wi) * kInvPi + pdfp / (4 * absDot(wo, wh)));
      f = sScale(fresnelBlendEval(b, wi, wo), 1 / pdf);
      return;
   }
   case K_MICROFACET: {
      float cost = std::pow(u1, 1 / (b.e + 1));
      float sint = std::sqrt(hmax(0, 1 - cost * cost));
      float phi = u2 * 2 * kPi;
      V3 whp = sphericalDirection(sint, cost, phi);
      float ff = std::pow(cost, b.e) * kInvTwoPi;
      float d = (b.e + 2) * ff, dpdf = (b.e + 1) * ff;
      V3 wh = (cosTheta(whp) < 0) ? -whp : whp;
      float costH = dot(wo, wh);
      wi = scl(2 * costH, wh) - wo;
      if (!sameHemisphere(wo, wi)) { f = sConst(0); wi = wo; pdf = 0; return; }
      float fact = d * std::fabs(costH) / dpdf * mfG(wo, wi, wh);
      Spec fp = b.r * fresnel(b, costH);
      f = sScale(fp, fact / absCosTheta(wo));
      pdf = dpdf / (4 * std::fabs(costH));
      return;
   }
   default: bxdfSample(b, wo, u1, u2, f, wi, pdf); return;   // specRefl
   }
}
// sampleBsdf'' True bxdfAll (Reflection.hs:278-316): the other components are evaluated UNflipped (`eval b = bxdfEval b`, :310)
// and every weight is scaled by |sideTest| (`fAdj`, :315-316)
static inline BsdfSample sampleAdjBsdf(const Bsdf &bsdf, V3 woW, float uComp, float uDir1, float uDir2) {
   BsdfSample empty{BX_REFLECTION | BX_DIFFUSE, 0, sConst(0), mk(0, 1, 0)};
   int cntm = bsdf.n;
   if (cntm == 0) return empty;
   V3 wo = worldToLocal(bsdf.cs, woW);
   float cntf = (float)cntm, invCnt = 1 / cntf;
   int sNum = std::max(0, std::min(cntm - 1, (int)std::floor(uComp * cntf)));
   const BxDF &bx = bsdf.bx[sNum];
   Spec fSample = sConst(0); V3 wi = mk(0, 1, 0); float pdfp = 0;
   bxdfSampleAdj(bx, wo, uDir1, uDir2, fSample, wi, pdfp);
   V3 wiW = localToWorld(bsdf.cs, wi);
   float sideTest = dot(wiW, bsdf.ng) / dot(woW, bsdf.ng);
   if (pdfp == 0 || sideTest == 0) return empty;
   bool wantTrans = sideTest < 0;
   if (!(wantTrans ? isTrans(bx) : isRefl(bx))) return empty;
   float as = std::fabs(sideTest);
   if (isSpec(bx)) return BsdfSample{bx.type, pdfp * invCnt, sScale(sScale(fSample, as), cntf), wiW};
   if (cntm == 1) return BsdfSample{bx.type, pdfp, sScale(fSample, as), wiW};
   float pdfSum = 0; Spec fOthers = sConst(0);
   for (int i = 0; i < cntm; ++i) {
      if (i == sNum) continue;
      pdfSum = pdfSum + bxdfPdf(bsdf.bx[i], wo, wi);
      if (wantTrans ? isTrans(bsdf.bx[i]) : isRefl(bsdf.bx[i])) fOthers = fOthers + bxdfEval(bsdf.bx[i], wo, wi);
   }
   float pdf = (pdfp + pdfSum) * invCnt;
   Spec fSum = sScale(sScale(fSample, pdfp) + fOthers, 1 / pdf);
   return BsdfSample{bx.type, pdf, sScale(fSum, as), wiW};
}
// evalBsdf True (Reflection.hs:318-332)
static inline Spec evalAdjBsdf(const Bsdf &bsdf, V3 woW, V3 wiW) {
   float cosWo = dot(woW, bsdf.ng);
   float sideTest = dot(wiW, bsdf.ng) / cosWo;
   if (sideTest == 0) return sConst(0);
   if (std::fabs(cosWo) < 1e-5f) return sConst(0);
   bool wantTrans = sideTest < 0;
   V3 wo = worldToLocal(bsdf.cs, woW), wi = worldToLocal(bsdf.cs, wiW);
   Spec f = sConst(0);
   for (int i = 0; i < bsdf.n; ++i)
      if (wantTrans ? isTrans(bsdf.bx[i]) : isRefl(bsdf.bx[i])) f = f + bxdfEval(bsdf.bx[i], wo, wi);
   return sScale(f, std::fabs(sideTest));
}

// ----------------------------------------------------------------------------- Spectrum.hs conversions
struct SpectralTables { Spec cieX, cieY, cieZ; float ySum; Spec illum[7]; };  // illum: r g b c m y w

static inline Spec rgbToSpectrum(const Spec *B, float r, float g, float b) {  // Spectrum.hs:146-159
   const Spec &rb = B[0], &gb = B[1], &bb = B[2], &cb = B[3], &mb = B[4], &yb = B[5], &wb = B[6];
   if (r <= g && r <= b) {
      if (g <= b) return sScale(wb, r) + (sScale(cb, g - r) + sScale(bb, b - g));
      return sScale(wb, r) + (sScale(cb, b - r) + sScale(gb, g - b));
   }
   if (g <= r && g <= b) {
      if (r <= b) return sScale(wb, g) + (sScale(mb, r - g) + sScale(bb, b - r));
      return sScale(wb, g) + (sScale(mb, b - g) + sScale(rb, r - b));
   }
   if (r <= b) return sScale(wb, b) + (sScale(yb, r - b) + sScale(gb, g - r));
   return sScale(wb, b) + (sScale(yb, g - b) + sScale(rb, r - g));
}
static inline void xyzToRgb(float x, float y, float z, float &r, float &g, float &b) {  // :162-168
   r = 3.240479f * x - 1.537150f * y - 0.498535f * z;
   g = (-0.969256f) * x + 1.875991f * y + 0.041556f * z;
   b = 0.055648f * x - 0.204043f * y + 1.057311f * z;
}
static inline void spectrumToXYZ(const SpectralTables &T, const Spec &s, float &X, float &Y, float &Z) {  // :349-355
   float a = 0, b = 0, c = 0;
   for (int i = 0; i < NB; ++i) { a = a + T.cieX.v[i] * s.v[i]; b = b + T.cieY.v[i] * s.v[i]; c = c + T.cieZ.v[i] * s.v[i]; }
   X = a / T.ySum; Y = b / T.ySum; Z = c / T.ySum;
}
static inline float sY(const SpectralTables &T, const Spec &s) {  // :371-373
   float a = 0;
   for (int i = 0; i < NB; ++i) a = a + s.v[i] * T.cieY.v[i];
   return a / T.ySum;
}

// ----------------------------------------------------------------------------- SunSky.hs
static inline float perez(const float *p, float sunT, float t, float g, float lvz) {  // :81-86
   float csg = std::cos(g), cst = std::cos(sunT);
   float num = (1 + p[0] * std::exp(p[1] / std::cos(t))) * (1 + p[2] * std::exp(p[3] * g)) + p[4] * csg * csg;
   float den = (1 + p[0] * std::exp(p[1])) * (1 + p[2] * std::exp(p[3] * sunT)) + p[4] * cst * cst;
   return lvz * num / den;
}
static inline Spec sunSkyEval(const SpectralTables &T, const blingcu_sunsky &k, V3 dir) {  // :12-24, 67-94
   Spec sky = sConst(0);
   float dz = -dir.z;
   V3 sunDir = mk(k.sun_dir[0], k.sun_dir[1], k.sun_dir[2]);
   if (!(dz < 1e-4f)) {
      float theta = std::acos(dz);
      float gamma = std::acos(clampf(dot(dir, sunDir), -1, 1));
      float x = perez(k.perez_x, k.sun_theta, theta, gamma, k.zenith_x);
      float y = perez(k.perez_y, k.sun_theta, theta, gamma, k.zenith_y);
      float yp = perez(k.perez_Y, k.sun_theta, theta, gamma, k.zenith_Y) * 1e-4f;
      // chromaticityToXYZ (Spectrum.hs:226-244)
      float m1 = (-1.3515f - 1.7703f * x + 5.9114f * y) / (0.0241f + 0.2562f * x - 0.7341f * y);
      float m2 = (0.03f - 31.4424f * x + 30.0717f * y) / (0.0241f + 0.2562f * x - 0.7341f * y);
      float cx = k.s0xyz[0] + m1 * k.s1xyz[0] + m2 * k.s2xyz[0];
      float cy = k.s0xyz[1] + m1 * k.s1xyz[1] + m2 * k.s2xyz[1];
      float cz = k.s0xyz[2] + m1 * k.s1xyz[2] + m2 * k.s2xyz[2];
      float xp = cx * yp / cy, zp = cz * yp / cy;
      float r, g, b; xyzToRgb(xp, yp, zp, r, g, b);
      sky = rgbToSpectrum(T.illum, r, g, b);   // xyzToSpectrum (Spectrum.hs:357-358)
   }
   Spec sun = sConst(0);
   float d = dot(mk(k.sun_disc_dir[0], k.sun_disc_dir[1], k.sun_disc_dir[2]) * mk(1, 1, -1), dir);
   float sint2 = 6.955e5f / 1.496e8f;
   float stm = std::sqrt(hmax(0, 1 - sint2));
   if (d > stm) { for (int i = 0; i < NB; ++i) sun.v[i] = k.sun_radiance.v[i]; }
   return sky + sun;
}

// ----------------------------------------------------------------------------- env maps + Dist2D
static inline Spec fromC(const blingcu_spectrum &s) { Spec r; for (int i = 0; i < NB; ++i) r.v[i] = s.v[i]; return r; }

static inline Spec envEval(const SpectralTables &T, const blingcu_envmap &e, float u, float v) {  // texMapEval
   switch (e.kind) {
   case BLINGCU_ENV_CONSTANT: return fromC(e.s);
   case BLINGCU_ENV_RGBTABLE: {  // IO/Bitmap.hs:22-29
      int w = e.nu, h = e.nv;
      int x = std::max(0, std::min(w - 1, (int)std::floor((1 - u) * (float)w)));
      int y = std::max(0, std::min(h - 1, (int)std::floor((1 - v) * (float)h)));
      const float *px = e.rgb + 3 * ((size_t)y * w + x);
      return rgbToSpectrum(T.illum, px[0], px[1], px[2]);
   }
   default: {  // SunSky.hs:18-20: sphToDir (cartToSph cc)
      float phi = u * 2 * kPi, theta = v * kPi;
      V3 dir = sphericalDirection(std::sin(theta), std::cos(theta), phi);
      return sunSkyEval(T, e.sky, dir);
   }
   }
}
// Montecarlo.hs:53-54 upperBound: linear findIndex (>= u)
static inline int upperBound(const float *cdf, int len, float u) {
   int idx = len - 1;
   for (int i = 0; i < len; ++i) if (cdf[i] >= u) { idx = i - 1; break; }
   return std::min(len - 2, std::max(0, idx));
}
// Montecarlo.hs:66-71
static inline void sampleContinuous1D(const float *func, const float *cdf, float fi, int n, float u, float &x, float &pdf, int &off) {
   off = upperBound(cdf, n + 1, u);
   pdf = (fi == 0) ? 0 : func[off] / fi;
   float du = (u - cdf[off]) / (cdf[off + 1] - cdf[off]);
   x = ((float)off + du) / (float)n;
}
static inline void sampleContinuous2D(const blingcu_envmap &e, float u0, float u1, float &u, float &v, float &pdf) {  // :89-92
   float pdf1, pdf0; int imarg, dummy;
   sampleContinuous1D(e.marg_func, e.marg_cdf, e.marg_int, e.nv, u1, v, pdf1, imarg);
   sampleContinuous1D(e.cond_func + (size_t)imarg * e.nu, e.cond_cdf + (size_t)imarg * (e.nu + 1), e.cond_int[imarg], e.nu, u0, u, pdf0, dummy);
   pdf = pdf0 * pdf1;
}
static inline float pdfDist2D(const blingcu_envmap &e, float u, float v) {  // :94-104
   int iu = std::max(0, std::min(e.nu - 1, (int)std::floor(u * (float)e.nu)));
   int iv = std::max(0, std::min(e.nv - 1, (int)std::floor(v * (float)e.nv)));
   if (e.marg_int * e.cond_int[iv] == 0) return 0;
   return (e.cond_func[(size_t)iv * e.nu + iu] * e.marg_func[iv]) / (e.cond_int[iv] * e.marg_int);
}

}  // namespace orc
