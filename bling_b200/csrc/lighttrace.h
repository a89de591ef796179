// lighttrace.h -- SURVEY.md §8(f)4: the light tracer (Renderer/LightTracer.hs:1-110) as a sibling of the path integrator on the
// same traversal, material and spectrum code. Replaces oneRay / nextVertex / connectCam (LightTracer.hs:53-108), sampleLightRay
// (Scene.hs:121-136), Light.sample' (Light.hs:166-213), sampleAdjBsdf / evalBsdf True (Reflection.hs:263-332 with adj = True and
// the adjoint branches of Diffuse.hs:14-22,44-49, Specular.hs:52-57, Microfacet.hs:52-54,92), sampleCam (Camera.hs:78-103) and
// splatSample (Image.hs:201-221).
//
// Wavefront: one light path ("photon") per slot. gen -> { trace_nearest -> vertex (adjoint BSDF sample + camera connection ray)
// -> trace_any on the connection rays -> collect the unoccluded ones as splat records -> splat } per bounce; there is no depth
// limit in the reference, only Russian roulette (0.8 beyond depth 3), so the host loops until the queue is empty.
// Splatting is ATOMIC-FREE on the floats and deterministic: the records of one bounce are grouped by pixel with an integer
// counting sort (count, exclusive scan, scatter) and every pixel adds its own records in slot order.
#pragma once

namespace bl {

enum { S_PHOTONS = 9, S_RAYS_LIGHT = 10, S_RAYS_CONNECT = 11 };   // stats slots (bodies.h N_STATS = 12); splats are counted on the host side of the flush
enum { C_LT_RECORDS = C_MIS };   // counter reused: number of splat records of the current bounce

// ---------------------------------------------------------------------------------------------------- adjoint BSDF
// bxdfSample b True wo u, any material kind. Differences from adj = False are listed BxDF by BxDF in the file header.
HDNI void bxdfSampleAdjGeneral(const BxDF &b, V3 wo, float u1, float u2, Spec &f, V3 &wi, float &pdf) {
   if (b.kind == K_LAMBERT || b.kind == K_ORENNAYAR) {
      wi = toSameHemisphere(wo, cosineSampleHemisphere(u1, u2));
      if (sameHemisphere(wo, wi)) { f = bxScaledR(b, (b.kind == K_LAMBERT ? 1.0f : orenNayarF(b, wo, wi)) * fabsf(cosTheta(wo) / cosTheta(wi))); pdf = cosPdf(wo, wi); }
      else { f = sConst(0); pdf = 0; }
      if (b.flip) wi = flipZ(wi);
      return;
   }
   if (b.kind == K_SPECTRANS) {
      bool entering = cosTheta(wo) > 0;
      float ei = entering ? b.etai : b.etat, et = entering ? b.etat : b.etai;
      float eta = ei / et, eta2 = eta * eta, sint2 = eta2 * sinTheta2(wo);
      if (sint2 >= 1) { f = sConst(0); wi = wo; pdf = 0; return; }
      float c = sqrtf(hmaxf(0, 1 - sint2));
      float cost = entering ? -c : c;
      wi = mk3(eta * (-wo.x), eta * (-wo.y), cost);
      float fr = frDielectric(ei, et, cosTheta(wo));           // the INCIDENT cosine when adjoint (Specular.hs:52)
      float sc_ = fabsf(cosTheta(wo) / cost);                  // instead of eta^2 (:55-57)
      BL_UNROLL for (int i = 0; i < NB; ++i) f.v[i] = ((1.0f - fr) * bxR(b, i)) * sc_;
      pdf = 1;
      return;
   }
   if (b.kind == K_FRESNELBLEND) {
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
      f = sScale(fresnelBlendEval(b, wi, wo), 1 / pdf);         // `e wi wo` when adjoint (Microfacet.hs:92)
      return;
   }
   if (b.kind == K_MICROFACET) {
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
      f = sScale(bxRFresnel<AnyMat>(b, costH), fact / absCosTheta(wo));   // |cos wo| when adjoint (Microfacet.hs:52-54)
      pdf = dpdf / (4 * fabsf(costH));
      return;
   }
   bxdfSampleGeneral(b, wo, u1, u2, f, wi, pdf);   // specRefl: no adjoint branch (Specular.hs:19)
}
// sampleBsdf'' True bxdfAll (Reflection.hs:278-316)
HDNI void sampleAdjBsdfGeneral(const Bsdf &bsdf, V3 woW, float uComp, float u1, float u2, BsdfSample &out) {
   out.type = BX_REFLECTION | BX_DIFFUSE; out.pdf = 0; out.f = sConst(0); out.wi = mk3(0, 1, 0);
   const int cntm = bsdf.n;
   if (cntm == 0) return;
   V3 wo = worldToLocal(bsdf.cs, woW);
   float cntf = (float)cntm, invCnt = 1 / cntf;
   const int sNum = imax(0, imin(cntm - 1, (int)floorf(uComp * cntf)));
   const BxDF &bx = bsdf.bx[sNum];
   Spec fS; V3 wi = mk3(0, 1, 0); float pdfp = 0;
   bxdfSampleAdjGeneral(bx, wo, u1, u2, fS, wi, pdfp);
   V3 wiW = localToWorld(bsdf.cs, wi);
   float sideTest = dot3(wiW, bsdf.ng) / dot3(woW, bsdf.ng);
   if (pdfp == 0 || sideTest == 0) return;
   bool wantTrans = sideTest < 0;
   if (!bxMatch(bx, wantTrans)) return;
   const float as = fabsf(sideTest);                          // fAdj (:315-316)
   out.type = bx.type; out.wi = wiW;
   if (bx.type & BX_SPECULAR) { out.pdf = pdfp * invCnt; out.f = sScale(sScale(fS, as), cntf); return; }
   if (cntm == 1) { out.pdf = pdfp; out.f = sScale(fS, as); return; }
   const BxDF &o = bsdf.bx[1 - sNum];
   float pdf = (pdfp + bxdfPdfGeneral(o, wo, wi)) * invCnt;
   Spec fOthers = sConst(0);
   if (bxMatch(o, wantTrans)) { Spec e; bxdfEvalGeneral(o, wo, wi, e); fOthers = fOthers + e; }   // NOT flipped when adjoint (:310)
   out.pdf = pdf; out.f = sScale(sScale(sScale(fS, pdfp) + fOthers, 1 / pdf), as);
}
// evalBsdf True (Reflection.hs:318-332)
HDNI void evalAdjBsdfGeneral(const Bsdf &bsdf, V3 woW, V3 wiW, Spec &f) {
   f = sConst(0);
   float cosWo = dot3(woW, bsdf.ng);
   float sideTest = dot3(wiW, bsdf.ng) / cosWo;
   if (sideTest == 0) return;
   if (fabsf(cosWo) < 1e-5f) return;
   bool wantTrans = sideTest < 0;
   V3 wo = worldToLocal(bsdf.cs, woW), wi = worldToLocal(bsdf.cs, wiW);
   for (int i = 0; i < 2; ++i) if (i < bsdf.n && bxMatch(bsdf.bx[i], wantTrans)) { Spec e; bxdfEvalGeneral(bsdf.bx[i], wo, wi, e); f = f + e; }
   f = sScale(f, fabsf(sideTest));
}

// ---------------------------------------------------------------------------------------------------- lights, camera
struct LightRay { Spec li; Ray ray; V3 nl; float pdf; };
HD void boundingSphere(const DScene &S, V3 &c, float &r) {   // AABB.hs:62-66 of the scene's world bounds
   V3 lo = mk3(S.bounds_lo[0], S.bounds_lo[1], S.bounds_lo[2]), hi = mk3(S.bounds_hi[0], S.bounds_hi[1], S.bounds_hi[2]);
   c = scl(0.5f, lo + hi); r = len3(hi - c);
}
// Light.sample' (Light.hs:166-213)
HDNI void lightSampleRay(const DScene &S, const blingcu_light &l, float uo1, float uo2, float ud1, float ud2, LightRay &o) {
   o.li = sConst(0); o.ray.o = mk3(0, 0, 0); o.ray.d = mk3(0, 1, 0); o.ray.tmin = 0; o.ray.tmax = 0; o.nl = mk3(0, 1, 0); o.pdf = 0;
   if (l.kind == BLINGCU_LIGHT_AREA) {
      const blingcu_shape &sh = S.shapes[l.shape];
      V3 orgL, nsL; sampleShapeAny(sh, uo1, uo2, orgL, nsL);
      V3 org = transPoint(sh.o2w, orgL), ns = normalize3(transNormalInv(sh.w2o, nsL));
      V3 wi = localToWorld(coordinateSystem(ns), cosineSampleHemisphere(ud1, ud2));
      o.li = loadSpec(l.s.v); o.ray.o = org; o.ray.d = wi; o.ray.tmin = 1e-3f; o.ray.tmax = BL_INF; o.nl = ns;
      o.pdf = BL_INVPI * (1 / shapeArea(sh)) * absDot(ns, wi);
      return;
   }
   if (l.kind == BLINGCU_LIGHT_DIRECTIONAL) {
      V3 n = mk3(l.v[0], l.v[1], l.v[2]);
      V3 wc; float wr; boundingSphere(S, wc, wr);
      Frame f = coordinateSystem(n);
      float d1, d2; concentricSampleDisk(uo1, uo2, d1, d2);
      V3 pdisk = wc + scl(wr, scl(d1, f.s) + scl(d2, f.t));
      o.li = loadSpec(l.s.v); o.ray.o = pdisk + scl(wr, n); o.ray.d = -n; o.ray.tmin = 0; o.ray.tmax = BL_INF; o.nl = -n;
      o.pdf = 1 / (BL_PI * wr * wr);
      return;
   }
   if (l.kind == BLINGCU_LIGHT_INFINITE) {
      const blingcu_envmap &e = S.envs[l.env];
      float u, v, pdMap; sampleContinuous2D(e, ud1, ud2, u, v, pdMap);
      if (pdMap == 0) return;
      float phi = u * 2 * BL_PI, theta = v * BL_PI;
      V3 d = transVector(e.l2w, sphericalDirection(sinf(theta), cosf(theta), phi));
      V3 wc; float wr; boundingSphere(S, wc, wr);
      Frame f = coordinateSystem(-d);
      float d1, d2; concentricSampleDisk(uo1, uo2, d1, d2);
      V3 pDisk = wc + scl(wr, scl(d1, f.s) + scl(d2, f.t));
      float sint = sinf(theta);
      float pdDir = pdMap / (2 * BL_PI * BL_PI * sint), pdArea = 1 / (BL_PI * wr * wr);
      o.li = envEval(S, e, u, v); o.ray.o = pDisk + scl(wr, d); o.ray.d = -d; o.ray.tmin = 0; o.ray.tmax = BL_INF; o.nl = d;
      o.pdf = (sint == 0) ? 0.0f : pdDir * pdArea;
      return;
   }
   V3 d = uniformSampleSphere(ud1, ud2);   // PointLight; uniformSpherePdf = 1 / (2 pi) (Q5)
   o.li = loadSpec(l.s.v); o.ray.o = mk3(l.v[0], l.v[1], l.v[2]); o.ray.d = d; o.ray.tmin = 0; o.ray.tmax = BL_INF; o.nl = d;
   o.pdf = 1 / (2 * BL_PI);
}
// sampleCam (Camera.hs:78-103)
HD void sampleCam(const blingcu_camera &c, V3 p, V3 &pLens, float &px, float &py, float &pdf) {
   V3 pRas = transPoint(c.world2raster, p);
   pLens = transPoint(c.cam2world, mk3(0, 0, 0));
   float cost = fabsf(normalize3(transPoint(c.raster2cam, pRas)).z);
   px = pRas.x; py = pRas.y; pdf = c.pixel_area * (cost * cost * cost);
}

// the sampler of a light path: plain uniforms of the counter-based stream (hd.h), keyed by the photon's index in the pass
HD Sampler photonSampler(const DScene &S, uint64_t kp) { Sampler c; c.kp = kp; c.s = 0; c.k = &S.smpUniform; return c; }

// ---------------------------------------------------------------------------------------------------- kernel bodies
// oneRay up to the first intersection query (LightTracer.hs:53-63)
struct LtGenBody {
   const DScene *sc; PathState ps; uint64_t seed; uint32_t pass; uint64_t first;
   HD void operator()(uint32_t i) const {
      const DScene &S = *sc;
      const uint64_t kp = pixelKey(seed, pass, (uint32_t)(first + i));
      Sampler c = photonSampler(S, kp);
      float ul = rnd1D(c, 0), uo1, uo2, ud1, ud2; rnd2D(c, 0, uo1, uo2); rnd2D(c, 1, ud1, ud2);
      if (S.n_lights == 0) return;
      LightRay lr;
      if (S.n_lights == 1) lightSampleRay(S, S.lights[0], uo1, uo2, ud1, ud2, lr);
      else {   // sampleLightRay (Scene.hs:121-136)
         int ln = imin((int)floorf(ul * (float)S.n_lights), S.n_lights - 1);
         lightSampleRay(S, S.lights[ln], uo1, uo2, ud1, ud2, lr);
         lr.pdf = lr.pdf / (float)S.n_lights;
      }
      if (!(lr.pdf > 0)) return;
      V3 wo = normalize3(lr.ray.d);
      Spec li = sScale(lr.li, absDot(lr.nl, wo) / lr.pdf);
      if (isBlack(li)) return;
      storeRay(ps.rayO, ps.rayD, i, lr.ray);
      storeSpec4(ps.T, ps.cap, i, li);
      F4 w; w.x = -wo.x; w.y = -wo.y; w.z = -wo.z; w.w = 0; ps.miD[rayAt2(i)] = w;   // wi of the first vertex
      ps.meta[i] = 0u; ps.kp[i] = kp; ps.sidx[i] = i;
      qPush(ps.qA, ps.counters + C_ACTIVE, i);
   }
};

// nextVertex + connectCam (LightTracer.hs:62-108)
struct LtVertexBody {
   const DScene *sc; PathState ps; uint32_t *qNext;
   HD void operator()(uint32_t i) const {
      const DScene &S = *sc;
      F4 hv = ps.hit[i];
      if (f2i(hv.w) == BL_REF_MISS) return;
      const Spec li = loadSpec4(ps.T, ps.cap, i);
      if (isBlack(li)) return;
      Ray ray = loadRay(ps.rayO, ps.rayD, i);
      const int depth = (int)ps.meta[i];
      Sampler c = photonSampler(S, ps.kp[i]);
      float ubc = rnd1D(c, 1 + 2 * depth), ub1, ub2; rnd2D(c, 2 + depth, ub1, ub2);
      const float pcont = (depth > 3) ? 0.8f : 1.0f;
      SurfaceHit sh; DG dgs;
      surfaceAt(S, ray, hv.x, hv.y, hv.z, f2i(hv.w), sh, dgs);
      Spec texScratch[4];
      Bsdf bsdf; makeBsdfGeneral(S, sh, dgs, bsdf, texScratch);
      const F4 wv = ps.miD[rayAt2(i)]; const V3 wi = mk3(wv.x, wv.y, wv.z);
      const V3 p = bsdf.p; const float eps = sh.eps;
      BsdfSample bs; sampleAdjBsdfGeneral(bsdf, wi, ubc, ub1, ub2, bs);
      {   // connectCam
         V3 pLens; float px, py, cPdf; sampleCam(S.cam, p, pLens, px, py, cPdf);
         V3 dCam = pLens - p, we = normalize3(dCam);
         Spec f; evalAdjBsdfGeneral(bsdf, wi, we, f);
         if (!(isBlack(f) || cPdf == 0)) {
            const float dCam2 = sqLen(dCam);
            Ray cr; cr.o = p; cr.d = we; cr.tmin = eps; cr.tmax = sqrtf(dCam2);
            storeRay(ps.shO, ps.shD, i, cr);
            storeSpec4(ps.PS, ps.cap, i, sScale(li * f, 1 / (cPdf * dCam2)));
            F4 q; q.x = px; q.y = py; q.z = i2f(depth); q.w = 0; ps.mihit[i] = q;   // raster position and depth of this connection
            qPush(ps.qShadow, ps.counters + C_SHADOW, i);
         }
      }
      if (isBlack(bs.f) || bs.pdf == 0) return;
      if (rnd1D(c, 2 + 2 * depth) > pcont) return;
      Ray nr; nr.o = p; nr.d = bs.wi; nr.tmin = eps; nr.tmax = BL_INF;
      storeRay(ps.rayO, ps.rayD, i, nr);
      storeSpec4(ps.T, ps.cap, i, sScale(li * bs.f, 1 / pcont));
      F4 w; w.x = -bs.wi.x; w.y = -bs.wi.y; w.z = -bs.wi.z; w.w = 0; ps.miD[rayAt2(i)] = w;
      ps.meta[i] = (uint32_t)(depth + 1);
      qPush(qNext, ps.counters + C_NEXT, i);
   }
};

// splat records of one bounce (structure of arrays, grow-only scratch owned by the pipeline)
struct LtRecords { uint32_t *pixel, *key; F4 *xyz; F2 *pos; uint32_t *depth; uint32_t cap; };

// the unoccluded connections become records: splatSample's tests (Image.hs:201-221) and the conversion to XYZ happen here
struct LtCollectBody {
   const DScene *sc; PathState ps; LtRecords rec;
   HD void operator()(uint32_t i) const {
      const DScene &S = *sc;
      if (ps.occl[i]) return;
      const F4 q = ps.mihit[i];
      const int px = (int)floorf(q.x), py = (int)floorf(q.y);
      if (px >= S.W || py >= S.H || px < 0 || py < 0) return;
      const Spec ss = loadSpec4(ps.PS, ps.cap, i);
      if (sBad(ss)) return;
      float X, Y, Z; spectrumToXYZ(S, ss, X, Y, Z);
#if defined(__CUDA_ARCH__)
      const uint32_t k = atomicAdd(ps.counters + C_LT_RECORDS, 1u);
#else
      const uint32_t k = ps.counters[C_LT_RECORDS]++;
#endif
      if (k >= rec.cap) return;   // cannot happen: at most one record per slot per bounce, cap = slots
      rec.pixel[k] = (uint32_t)py * (uint32_t)S.W + (uint32_t)px; rec.key[k] = i;
      F4 v; v.x = X; v.y = Y; v.z = Z; v.w = 0; rec.xyz[k] = v; F2 xy; xy.x = q.x; xy.y = q.y; rec.pos[k] = xy; rec.depth[k] = (uint32_t)f2i(q.z);
   }
};

// ---- deterministic, atomic-free splat of one bounce's records: counting sort by pixel, then every pixel adds its own records in
// key (= slot = photon) order
struct LtCountBody { LtRecords rec; const uint32_t *n; uint32_t *count; HD void operator()(uint32_t k) const { if (k < *n) cntAdd(count + rec.pixel[k], 1u); } };
struct LtScatterBody {
   LtRecords rec; const uint32_t *n; const uint32_t *offset; uint32_t *cursor; uint32_t *sorted;
   HD void operator()(uint32_t k) const {
      if (k >= *n) return;
      const uint32_t p = rec.pixel[k];
#if defined(__CUDA_ARCH__)
      const uint32_t at = offset[p] + atomicAdd(cursor + p, 1u);
#else
      const uint32_t at = offset[p] + cursor[p]++;
#endif
      sorted[at] = k;
   }
};
struct LtGatherBody {   // one item per pixel
   LtRecords rec; const uint32_t *count; const uint32_t *offset; const uint32_t *sorted; float *splat;
   HD void operator()(uint32_t p) const {
      const uint32_t n = count[p];
      if (n == 0) return;
      const uint32_t o = offset[p];
      float X = splat[3 * (size_t)p], Y = splat[3 * (size_t)p + 1], Z = splat[3 * (size_t)p + 2];
      uint32_t last = 0; bool any = false;
      for (uint32_t r = 0; r < n; ++r) {   // selection by increasing key: n is the number of splats ONE bounce puts on ONE pixel
         uint32_t best = 0xffffffffu, bk = 0;
         for (uint32_t j = 0; j < n; ++j) { const uint32_t k = sorted[o + j], key = rec.key[k]; if ((!any || key > last) && key < best) { best = key; bk = k; } }
         const F4 v = rec.xyz[bk];
         X = X + v.x; Y = Y + v.y; Z = Z + v.z;
         last = best; any = true;
      }
      splat[3 * (size_t)p] = X; splat[3 * (size_t)p + 1] = Y; splat[3 * (size_t)p + 2] = Z;
   }
};
// exclusive scan of count[0..n) into offset[0..n): three small kernels over blocks of LT_SCAN_BLOCK items
#define LT_SCAN_BLOCK 4096u
struct LtScanSumBody {   // one item per block: its total
   const uint32_t *count; uint32_t n; uint32_t *blockSum;
   HD void operator()(uint32_t b) const { uint32_t s = 0; const uint32_t e = (b + 1) * LT_SCAN_BLOCK < n ? (b + 1) * LT_SCAN_BLOCK : n; for (uint32_t i = b * LT_SCAN_BLOCK; i < e; ++i) s += count[i]; blockSum[b] = s; }
};
struct LtScanBlocksBody {   // one item: exclusive scan of the block totals (at most a few thousand)
   uint32_t *blockSum; uint32_t nb;
   HD void operator()(uint32_t) const { uint32_t s = 0; for (uint32_t b = 0; b < nb; ++b) { const uint32_t v = blockSum[b]; blockSum[b] = s; s += v; } }
};
struct LtScanApplyBody {   // one item per block
   const uint32_t *count; uint32_t n; const uint32_t *blockSum; uint32_t *offset;
   HD void operator()(uint32_t b) const { uint32_t s = blockSum[b]; const uint32_t e = (b + 1) * LT_SCAN_BLOCK < n ? (b + 1) * LT_SCAN_BLOCK : n; for (uint32_t i = b * LT_SCAN_BLOCK; i < e; ++i) { offset[i] = s; s += count[i]; } }
};
struct LtAdvanceBody {   // end of a bounce: stats, swap of the queues' counters
   PathState ps;
   HD void operator()(uint32_t) const {
      uint32_t *c = ps.counters;
      statAdd(ps.stats + S_RAYS_LIGHT, c[C_NEXT]); statAdd(ps.stats + S_RAYS_CONNECT, c[C_SHADOW]);
      c[C_ACTIVE] = c[C_NEXT]; c[C_NEXT] = 0; c[C_SHADOW] = 0; c[C_LT_RECORDS] = 0;
   }
};
struct LtBeginBody {
   PathState ps; uint32_t n;
   HD void operator()(uint32_t) const { uint32_t *c = ps.counters; for (int k = 0; k < N_COUNTERS; ++k) if (k != C_DROPPED) c[k] = 0; statAdd(ps.stats + S_PHOTONS, n); }
};
struct LtAfterGenBody { PathState ps; HD void operator()(uint32_t) const { statAdd(ps.stats + S_RAYS_LIGHT, ps.counters[C_ACTIVE]); } };

}  // namespace bl
