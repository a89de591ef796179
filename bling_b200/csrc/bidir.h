// bidir.h -- the bidirectional path integrator (SURVEY §8(f)4 "sibling integrators on the same kernels") as wavefront bodies:
// Integrator/BidirPath.hs:44-214 (mkBidirPathIntegrator md sd = contrib False md) on the traversal, BSDF, light-sampling and
// resolve code of the path integrator (bodies.h) and the adjoint BSDF / light-ray sampling of the light tracer (lighttrace.h).
//
// One camera sample = one eye path and one light path of at most md vertices each (nextVertex :184-214), estimateDirect at every
// eye vertex (S1, :69-72, :151-164), emitters seen directly or through specular bounces (S0, :79-83) and every eye vertex
// connected with every light vertex (:91-93, connect :124-149), all weighted by 1 / (path length - specular vertices) (:111-122).
// Schedule (pipeline.h::bouncesBidir): raygen -> md x [trace, BdVertexBody<eye>, shadow / MIS traces, resolves into the depth's
// own plane] -> light rays -> md x [trace, BdVertexBody<light>] -> BdFinishBody (weights of S1) -> md x md x [BdConnectBody,
// any-hit trace, resolve]. The vertices stay in HBM between the stages ([side][depth][slot] records).
//
// Restated AS WRITTEN: `connect` is called with the light vertex first and its patterns bind the second field of a vertex --
// the SAMPLED direction _vwo -- where their names say wi; `le` asks intLe for that sampled direction too; rrProb = 1, so the
// roulette never ends a path; `mkNoDirectBidirIntegrator` (`bidirnod`) is `undefined` in the reference and not offered.
#pragma once

namespace bl {

#define BL_BD_MAXDEPTH 16   // maxDepth of the bidirectional integrator (vertex records per path, local weight tables)

struct BdState {              // record (v, i) with v = side * md + depth at index v * cap + i; side 0 = eye path, 1 = light path
   int md;
   F4 *vRay;                  // x2: the ray that found the vertex (_vwi = -d, the hit point and its epsilon come from ray + hit)
   F4 *vHit;                  // (t, b1, b2, prim bits): _vint
   F4 *vWo;                   // _vwo, the direction sampled at the vertex; .w = the BxDF type bits of the sample (_vtype)
   F4 *vAlpha;                // x4: _valpha
   uint32_t *nVert;           // [side][slot]: vertices of the path
   F4 *D;                     // [depth][slot] x4: estimateDirect at eye vertex `depth` (times alpha, not yet weighted)
};
HD size_t bdAt(const BdState &bd, uint32_t cap, int side, int depth, uint32_t i) { return (size_t)(side * bd.md + depth) * cap + i; }

struct BdBeginBody {
   PathState ps; BdState bd;
   HD void operator()(uint32_t i) const { bd.nVert[i] = 0u; bd.nVert[ps.cap + i] = 0u; }
};

// sampleLightRay + the head of lightPath (Scene.hs:121-136, BidirPath.hs:175-182): the light path of slot i starts in the slot's
// own ray / throughput fields (the eye path is finished and recorded by then); rnd' 0, rnd2D' 0, rnd2D' 1 of the camera sample
struct BdLightGenBody {
   const DScene *sc; PathState ps; uint32_t *q;
   HD void operator()(uint32_t i) const {
      const DScene &S = *sc;
      Sampler c = mkSampler(S, ps.kp[i], ps.sidx[i]);
      float ul = rnd1D(c, 0), uo1, uo2, ud1, ud2; rnd2D(c, 0, uo1, uo2); rnd2D(c, 1, ud1, ud2);
      LightRay lr;
      lr.li = sConst(0); lr.ray.o = mk3(0, 0, 0); lr.ray.d = mk3(0, 1, 0); lr.ray.tmin = 0; lr.ray.tmax = 0; lr.nl = mk3(0, 1, 0); lr.pdf = 0;
      if (S.n_lights == 1) lightSampleRay(S, S.lights[0], uo1, uo2, ud1, ud2, lr);
      else if (S.n_lights > 1) {
         int ln = imin((int)floorf(ul * (float)S.n_lights), S.n_lights - 1);
         lightSampleRay(S, S.lights[ln], uo1, uo2, ud1, ud2, lr);
         lr.pdf = lr.pdf / (float)S.n_lights;
      }
      const V3 wo = -lr.ray.d;
      storeRay(ps.rayO, ps.rayD, i, lr.ray);
      storeSpec4(ps.T, ps.cap, i, sScale(lr.li, absDot(lr.nl, wo) / lr.pdf));   // li' (:179); a zero pdf is the reference's NaN / inf as well
      ps.meta[i] = 0u;
      q[i] = i;
   }
};
struct BdLightStartBody {   // the n light rays are extension rays in the statistics (the oracle counts them the same way)
   PathState ps; uint32_t n;
   HD void operator()(uint32_t) const { ps.counters[C_ACTIVE] = n; ps.counters[C_NEXT] = 0; statAdd(ps.stats + S_EXT, n); }
};

// nextVertex (:184-214) for the vertex the traced ray found; the eye side adds S0 and queues S1 here as well
struct BdVertexBody {
   typedef MatOf<SK_GENERAL> M;
   const DScene *sc; PathState ps; BdState bd; uint32_t *qNext; int side, depth;
   HD void operator()(uint32_t i) const {
      const DScene &S = *sc;
      const F4 hv = ps.hit[i];
      if (f2i(hv.w) == BL_REF_MISS) return;                 // :186 nothing hit: the path ends
      const Ray ray = loadRay(ps.rayO, ps.rayD, i);
      const uint32_t meta = ps.meta[i];
      Sampler smp = mkSampler(S, ps.kp[i], ps.sidx[i]);
      SurfaceHit sh; DG dgs;
      surfaceAt(S, ray, hv.x, hv.y, hv.z, f2i(hv.w), sh, dgs);
      Spec texScratch[4];
      Bsdf bsdf; makeBsdfGeneral(S, sh, dgs, bsdf, texScratch);
      const V3 wi = -ray.d, p = bsdf.p; const float eps = sh.eps;
      const int base1 = side ? 3 + 1 : 2 + 1, base2 = side ? 3 + 2 : 2 + 2;   // f1d / f2d of eyePath (:170) and lightPath (:182)
      const float ubc = rnd1D(smp, base1 + 12 * depth);
      float ub1, ub2; rnd2D(smp, base2 + 9 * depth, ub1, ub2);
      const float rr = rnd1D(smp, 1 + base1 + 12 * depth);
      BsdfSample bs;
      if (side) sampleAdjBsdfGeneral(bsdf, wi, ubc, ub1, ub2, bs); else sampleBsdfGeneral(bsdf, wi, ubc, ub1, ub2, bs);
      {   // vHere = Vert wi wo int t alpha (:199)
         const size_t at = bdAt(bd, ps.cap, side, depth, i);
         bd.vRay[2 * at] = ps.rayO[rayAt2(i)]; bd.vRay[2 * at + 1] = ps.rayD[rayAt2(i)];
         bd.vHit[at] = hv;
         F4 w; w.x = bs.wi.x; w.y = bs.wi.y; w.z = bs.wi.z; w.w = i2f(bs.type); bd.vWo[at] = w;
         for (int q = 0; q < 4; ++q) bd.vAlpha[4 * at + q] = ps.T[spec4At(ps.cap, i, q)];
         bd.nVert[(size_t)side * ps.cap + i] = (uint32_t)(depth + 1);
      }
      if (!side) {
         // S0 (:79-83): `_valpha v * intLe (_vint v) (_vwo v)` where the previous vertex was specular (or there is none)
         if (((meta >> 8) & 1u) && sh.light >= 0) {
            const blingcu_light &el = S.lights[sh.light];
            if (el.kind == BLINGCU_LIGHT_AREA && areaEmits(sh.dgg.n, bs.wi))
               storeSpec4(ps.L, ps.cap, i, loadSpec4(ps.L, ps.cap, i) + loadSpec4(ps.T, ps.cap, i) * loadSpec(el.s.v));
         }
         // S1 (:151-164): sampleOneLight with this vertex's numbers; the resolve bodies add it to the depth's plane bd.D
         const float lNumU = rnd1D(smp, 0 + 1 + 12 * depth);
         float lD1, lD2; rnd2D(smp, 0 + 2 + 9 * depth, lD1, lD2);
         const float bCompU = rnd1D(smp, 1 + 1 + 12 * depth);
         float bD1, bD2; rnd2D(smp, 1 + 2 + 9 * depth, bD1, bD2);
         directAtVertex<M>(S, ps, i, bsdf, wi, p, bsdf.cs.n, eps, lNumU, lD1, lD2, bCompU, bD1, bD2);
      }
      if (isBlack(bs.f) || bs.pdf == 0) return;              // :209-211
      const float rrProb = 1;                                // :203
      if (rr > rrProb) return;
      const Spec aNext = sScale(bs.f * loadSpec4(ps.T, ps.cap, i), 1 / rrProb);
      storeSpec4(ps.T, ps.cap, i, aNext);
      if (depth + 1 == bd.md) return;                        // the next call returns [] whatever its ray finds (:188): not traced
      Ray nr; nr.o = p; nr.d = bs.wi; nr.tmin = eps; nr.tmax = BL_INF;
      storeRay(ps.rayO, ps.rayD, i, nr);
      ps.meta[i] = (uint32_t)(depth + 1) | (((bs.type & BX_SPECULAR) ? 1u : 0u) << 8);
      qPush(qNext, ps.counters + C_NEXT, i);
   }
};

// number of (eye, light) vertex pairs of total length k (= a + b + 2) in which either sample was specular: countSpec (:111-122)
HD float bdSpecCount(const BdState &bd, uint32_t cap, uint32_t i, int nE, int nL, int k) {
   float c = 0;
   for (int a = 0; a < nE; ++a) {
      const int b = k - 2 - a;
      if (b < 0 || b >= nL) continue;
      const int te = f2i(bd.vWo[bdAt(bd, cap, 0, a, i)].w), tl = f2i(bd.vWo[bdAt(bd, cap, 1, b, i)].w);
      if ((te & BX_SPECULAR) || (tl & BX_SPECULAR)) c += 1;
   }
   return c;
}

// ld (:69-72): the S1 estimates of the eye vertices, each over (1 + i - specular pairs of length i + 1)
struct BdFinishBody {
   PathState ps; BdState bd;
   HD void operator()(uint32_t i) const {
      const int nE = (int)bd.nVert[i], nL = (int)bd.nVert[ps.cap + i];
      Spec ld = sConst(0);
      for (int d = 0; d < nE; ++d) {
         const Spec di = loadSpec4(bd.D + (size_t)d * 4 * ps.cap, ps.cap, i);
         ld = ld + sScale(di, 1 / (1 + (float)d - bdSpecCount(bd, ps.cap, i, nE, nL, d + 1)));
      }
      storeSpec4(ps.L, ps.cap, i, ld + loadSpec4(ps.L, ps.cap, i));   // ld + le (:101-104)
   }
};

// connect (:124-149) of light vertex s with eye vertex t: the contribution waits in PS for the any-hit query of the slot
struct BdConnectBody {
   const DScene *sc; PathState ps; BdState bd; int s, t;
   HD void operator()(uint32_t i) const {
      const DScene &S = *sc;
      const int nE = (int)bd.nVert[i], nL = (int)bd.nVert[ps.cap + i];
      if (s >= nL || t >= nE) return;
      const size_t ia = bdAt(bd, ps.cap, 1, s, i), ib = bdAt(bd, ps.cap, 0, t, i);   // a: the pattern's "eye vertex" (i = s), b: its "light vertex" (j = t)
      const F4 wa = bd.vWo[ia], wb = bd.vWo[ib];
      if ((f2i(wa.w) & BX_SPECULAR) || (f2i(wb.w) & BX_SPECULAR)) return;
      Spec texScratch[4];
      V3 pe, pl; float epsE, epsL; Spec fe, fl;
      V3 w; float wl, g;
      {
         Ray ra; { const F4 o = bd.vRay[2 * ia], d = bd.vRay[2 * ia + 1]; ra.o = mk3(o.x, o.y, o.z); ra.tmin = o.w; ra.d = mk3(d.x, d.y, d.z); ra.tmax = d.w; }
         Ray rb; { const F4 o = bd.vRay[2 * ib], d = bd.vRay[2 * ib + 1]; rb.o = mk3(o.x, o.y, o.z); rb.tmin = o.w; rb.d = mk3(d.x, d.y, d.z); rb.tmax = d.w; }
         const F4 ha = bd.vHit[ia], hb = bd.vHit[ib];
         SurfaceHit shA, shB; DG dgA, dgB;
         surfaceAt(S, ra, ha.x, ha.y, ha.z, f2i(ha.w), shA, dgA);
         surfaceAt(S, rb, hb.x, hb.y, hb.z, f2i(hb.w), shB, dgB);
         Bsdf bsdfe; makeBsdfGeneral(S, shA, dgA, bsdfe, texScratch);
         pe = bsdfe.p; epsE = shA.eps;
         // the direction needs both shading points: take the second one's first (makeBsdf fills p from the hit, no texture needed)
         Bsdf bsdfl; Spec texScratchL[4]; makeBsdfGeneral(S, shB, dgB, bsdfl, texScratchL);
         pl = bsdfl.p; epsL = shB.eps;
         const V3 d = pl - pe;
         w = d; wl = 0;                                      // normLen (Math.hs:356-363)
         if (sqLen(d) != 0) { wl = sqrtf(sqLen(d)); w = scl(1 / wl, d); }
         g = 1 / sqLen(d);
         evalBsdfGeneral(bsdfe, mk3(wa.x, wa.y, wa.z), w, fe);
         evalAdjBsdfGeneral(bsdfl, mk3(wb.x, wb.y, wb.z), -w, fl);
      }
      if (isBlack(fe) || isBlack(fl)) return;
      const float pathWt = 1 / ((float)(s + t + 2) - bdSpecCount(bd, ps.cap, i, nE, nL, s + t + 2));
      Spec aE, aL;
      for (int q = 0; q < 4; ++q) {
         const F4 x = bd.vAlpha[4 * ia + q], y = bd.vAlpha[4 * ib + q];
         aE.v[4 * q] = x.x; aE.v[4 * q + 1] = x.y; aE.v[4 * q + 2] = x.z; aE.v[4 * q + 3] = x.w;
         aL.v[4 * q] = y.x; aL.v[4 * q + 1] = y.y; aL.v[4 * q + 2] = y.z; aL.v[4 * q + 3] = y.w;
      }
      storeSpec4(ps.PS, ps.cap, i, sScale(aE * fe * aL * fl, g * pathWt));
      Ray cr; cr.o = pe; cr.d = w; cr.tmin = epsE; cr.tmax = wl - epsL;
      storeRay(ps.shO, ps.shD, i, cr);
      qPush(ps.qShadow, ps.counters + C_SHADOW, i);
   }
};

}  // namespace bl
