// textures.h -- textures that COMPUTE: scalar textures, spectrum blends / gradients, bump mapping.
// Replaces Texture.hs:129-141 (spectrumBlend), :150-181 (mappings), :183-187 (scaleTexture), :223-253 (gradient),
// :255-303 (cellNoise), :305-326 (quasiCrystal), :329-398 (fbm, perlin3d) and Reflection.hs:344-377 (bump).
// Included by shading.h after DScene. Only the "textured" instantiation of the shade kernel (MatOf<BLINGCU_MAT_KINDS>)
// calls into this file: materials whose texture trees only SELECT among constant spectra keep the pointer fast path.
#pragma once

namespace bl {

// ------------------------------------------------------------------------------------------ mappings
HD void map2d(const float *m, const DG &dg, float &x, float &y) {   // uvMapping :166-170, planarMapping :172-181
   if (m[0] == 0.0f) { x = m[1] * dg.u + m[3]; y = m[2] * dg.v + m[4]; return; }
   x = dot3(dg.p, mk3(m[1], m[2], m[3])) + m[7];
   y = dot3(dg.p, mk3(m[4], m[5], m[6])) + m[8];
}
HD V3 map3d(const float *w2t, const DG &dg) { return transPoint(w2t, dg.p); }   // identityMapping3d :150-152

// ------------------------------------------------------------------------------------------ Perlin noise (:343-398)
HD float noiseWeight(float t) { float t3 = t * t * t, t4 = t3 * t; return 6 * t4 * t - 15 * t4 + 10 * t3; }
HD float noiseGrad(const uint8_t *perm, int x, int y, int z, float dx, float dy, float dz) {
   // noisePerms is the 256-entry table twice over, so an index up to 511 wraps with & 255
   int h = perm[(perm[(perm[x & 255] + y) & 255] + z) & 255] & 15;
   float up = (h < 8 || h == 12 || h == 13) ? dx : dy;
   float vp = (h < 4 || h == 12 || h == 13) ? dy : dz;
   float u = (h & 1) ? -up : up, v = (h & 2) ? -vp : vp;
   return u + v;
}
HD float perlin3d(const uint8_t *perm, float x, float y, float z) {
   float fx = floorf(x), fy = floorf(y), fz = floorf(z);
   float dx = x - fx, dy = y - fy, dz = z - fz;
   int ix = (int)fx & 255, iy = (int)fy & 255, iz = (int)fz & 255;
   float w000 = noiseGrad(perm, ix, iy, iz, dx, dy, dz), w100 = noiseGrad(perm, ix + 1, iy, iz, dx - 1, dy, dz);
   float w010 = noiseGrad(perm, ix, iy + 1, iz, dx, dy - 1, dz), w110 = noiseGrad(perm, ix + 1, iy + 1, iz, dx - 1, dy - 1, dz);
   float w001 = noiseGrad(perm, ix, iy, iz + 1, dx, dy, dz - 1), w101 = noiseGrad(perm, ix + 1, iy, iz + 1, dx - 1, dy, dz - 1);
   float w011 = noiseGrad(perm, ix, iy + 1, iz + 1, dx, dy - 1, dz - 1), w111 = noiseGrad(perm, ix + 1, iy + 1, iz + 1, dx - 1, dy - 1, dz - 1);
   float wx = noiseWeight(dx), wy = noiseWeight(dy), wz = noiseWeight(dz);
   float x00 = lerpf(wx, w000, w100), x10 = lerpf(wx, w010, w110), x01 = lerpf(wx, w001, w101), x11 = lerpf(wx, w011, w111);
   return lerpf(wz, lerpf(wy, x00, x10), lerpf(wy, x01, x11));
}
HD float fbm3d(const uint8_t *perm, int octaves, float omega, V3 p) {   // :329-339; `sum` folds from 0 on the left
   float acc = 0, l = 1, o = 1;
   for (int k = 0; k < octaves; ++k) {
      acc = acc + o * perlin3d(perm, p.x * l, p.y * l, p.z * l);
      l = 1.99f * l; o = omega * o;
   }
   return acc;
}

// ------------------------------------------------------------------------------------------ Worley cell noise (:255-303)
// Haskell Int is 64 bits: the products of `hash` wrap there, `abs`, then `rem 2^32`; the LCG state stays below 2^32, so
// its `rem 2^32` is plain unsigned 32-bit wrap-around.
HD uint32_t cellLcg(uint32_t x) { return 1103515245u * x + 12345u; }
HD uint32_t cellHash(int64_t x, int64_t y, int64_t z) {
   int64_t h = (int64_t)((uint64_t)x * 73856093ull) ^ (int64_t)((uint64_t)y * 19349663ull) ^ (int64_t)((uint64_t)z * 83492791ull);
   uint64_t a = h < 0 ? (uint64_t)0 - (uint64_t)h : (uint64_t)h;
   return (uint32_t)(a & 0xffffffffull);
}
HD int cellProb(uint32_t v) {
   return v < 393325350u ? 1 : v < 1022645910u ? 2 : v < 1861739990u ? 3 : v < 2700834071u ? 4 : v < 3372109335u ? 5
        : v < 3819626178u ? 6 : v < 4075350088u ? 7 : v < 4203212043u ? 8 : 9;
}
HD float cellDist(int kind, V3 a, V3 b) {
   V3 d = a - b;
   if (kind == 0) return len3(d);
   if (kind == 1) return sqLen(d);
   if (kind == 2) return fabsf(d.x) + fabsf(d.y) + fabsf(d.z);
   return hmaxf(hmaxf(fabsf(d.x), fabsf(d.y)), fabsf(d.z));
}
HD float cellNoise(int kind, V3 p) {
   int ox = (int)floorf(p.x), oy = (int)floorf(p.y), oz = (int)floorf(p.z);
   float best = BL_INF;
   for (int x = -1; x <= 1; ++x) for (int y = -1; y <= 1; ++y) for (int z = -1; z <= 1; ++z) {
      int cx = x + ox, cy = y + oy, cz = z + oz;
      uint32_t u = cellLcg(cellHash(cx, cy, cz));
      int n = cellProb(u);
      for (int k = 0; k < n; ++k) {
         uint32_t u1 = cellLcg(u), u2 = cellLcg(u1), u3 = cellLcg(u2);
         V3 q = mk3((float)cx + (float)u1 / 4294967296.0f, (float)cy + (float)u2 / 4294967296.0f, (float)cz + (float)u3 / 4294967296.0f);
         best = hminf(best, cellDist(kind, p, q));
         u = u3;
      }
   }
   return best;
}

// ------------------------------------------------------------------------------------------ quasi crystal (:305-326)
HD float quasiCrystal(int octaves, float x, float y) {
   // angles = take n (enumFromThen 0 (pi / n)): numericEnumFromThen n m = n : numericEnumFromThen m (m + m - n)
   float a0 = 0, a1 = BL_PI / (float)octaves, s = 0;
   for (int k = 0; k < octaves; ++k) {
      float cth = cosf(a0), sth = sinf(a0);
      s = s + (cosf(cth * x + sth * y) + 1) / 2;
      float a2 = a1 + a1 - a0; a0 = a1; a1 = a2;
   }
   float ki = truncf(s), v = s - ki;                      // properFraction
   if (v < 0) { ki = ki - 1; v = 1 + v; }
   int k = (int)ki;
   return (k & 1) ? 1 - v : v;
}

// ------------------------------------------------------------------------------------------ image lookups (:82-108)
HD Spec rgbToSpectrumBasis(const float (*basis)[NB], float r, float g, float b);   // shading.h
// mod' a b (:91-94): a - (a `div` b) * b, plus b when negative -- i.e. the mathematical modulus for b > 0
HD int imageMod(float x, int b) { long long a = (long long)x, m = a % (long long)b; return (int)(m < 0 ? m + b : m); }
// getPixel / getPixelScalar: px = mod' (floor (u * w)) w, py = mod' (floor (-v * h)) h, pixelAt i px py
HD const float *imagePixel(const blingcu_image &im, float u, float v) {
   int px = imageMod(floorf(u * (float)im.width), im.width), py = imageMod(floorf((-v) * (float)im.height), im.height);
   return im.data + ((size_t)py * (size_t)im.width + (size_t)px) * (size_t)im.channels;
}

// ------------------------------------------------------------------------------------------ scalar textures
// (out of line: called from many sites of the textured shade kernel)
HDNI float evalScalarTexture(const DScene &sc, int id, const DG &dg) {
   float as[4], ss[4]; int n = 0;
   while (sc.textures[id].kind == BLINGCU_STEX_SCALE && n < 4) { const blingcu_texture &t = sc.textures[id]; as[n] = t.f[0]; ss[n] = t.f[1]; ++n; id = t.child[0]; }
   const blingcu_texture &t = sc.textures[id];
   float v = 0;
   switch (t.kind) {
   case BLINGCU_STEX_CONSTANT: v = t.f[0]; break;
   case BLINGCU_STEX_PERLIN: { V3 q = map3d(t.s.v, dg); v = perlin3d(sc.perm, q.x, q.y, q.z); break; }
   case BLINGCU_STEX_FBM: v = fbm3d(sc.perm, t.aux, t.f[0], map3d(t.s.v, dg)); break;
   case BLINGCU_STEX_CELLNOISE: v = cellNoise(t.aux, map3d(t.s.v, dg)); break;
   case BLINGCU_STEX_CRYSTAL: { float x, y; map2d(t.s.v, dg, x, y); v = quasiCrystal(t.aux, x, y); break; }
   case BLINGCU_STEX_IMAGE: { float x, y; map2d(t.s.v, dg, x, y); v = imagePixel(sc.images[t.aux], x, y)[0]; break; }
   default: break;
   }
   for (int k = n - 1; k >= 0; --k) v = as[k] + ss[k] * v;   // scaleTexture a s t = a + s * t
   return v;
}

// ------------------------------------------------------------------------------------------ spectrum textures, by value
// follows the selecting kinds (graphPaper, checker) down to a node that is constant or computes
HD int selectSpectrumTexture(const DScene &sc, int id, const DG &dg) {
   for (int guard = 0; guard < 16; ++guard) {
      const blingcu_texture &t = sc.textures[id];
      if (t.kind == BLINGCU_TEX_CHECKER) {
         int q = (int)floorf(dg.p.x * t.f[0]) + (int)floorf(dg.p.y * t.f[1]) + (int)floorf(dg.p.z * t.f[2]);
         id = ((q & 1) == 0) ? t.child[0] : t.child[1];
      } else if (t.kind == BLINGCU_TEX_GRAPHPAPER) {
         float x, z;
         if (t.aux == 0) { x = t.f[1] * dg.u + t.f[3]; z = t.f[2] * dg.v + t.f[4]; } else map2d(t.s.v, dg, x, z);
         float xp = fabsf(x - truncf(x)), zp = fabsf(z - truncf(z));
         float lo = t.f[0] / 2, hi = 1.0f - lo;
         id = (xp < lo || zp < lo || xp > hi || zp > hi) ? t.child[1] : t.child[0];
      } else return id;
   }
   return id;
}
HD Spec imageSpectrum(const DScene &sc, const blingcu_texture &t, const DG &dg) {   // pixelSpectrum (:87-89); unGamma on the host
   float x, y; map2d(t.s.v, dg, x, y);
   const float *px = imagePixel(sc.images[t.aux], x, y);
   return rgbToSpectrumBasis(sc.refl, px[0], px[1], px[2]);
}
template <int D> struct SpectrumValue {
   static HDNI Spec eval(const DScene &sc, int id, const DG &dg) {
      id = selectSpectrumTexture(sc, id, dg);
      const blingcu_texture &t = sc.textures[id];
      if (t.kind == BLINGCU_TEX_BLEND) {   // Texture.hs:129-141
         Spec v1 = SpectrumValue<D - 1>::eval(sc, t.child[0], dg), v2 = SpectrumValue<D - 1>::eval(sc, t.child[1], dg);
         float x = evalScalarTexture(sc, t.aux, dg);
         if (x <= 0) return v1;
         if (x >= 1) return v2;
         return sScale(v1, 1 - x) + sScale(v2, x);
      }
      if (t.kind == BLINGCU_TEX_GRADIENT) {   // Texture.hs:239-253
         float f = evalScalarTexture(sc, t.aux, dg);
         const blingcu_texture *st = sc.textures + t.child[0]; int n = t.child[1];
         if (f <= st[0].f[0]) return loadSpec(st[0].s.v);
         if (f >= st[n - 1].f[0]) return loadSpec(st[n - 1].s.v);
         int idx = 1; while (idx < n - 1 && !(st[idx].f[0] > f)) ++idx;   // findIndex ((> f) . fst)
         float w = (f - st[idx - 1].f[0]) / (st[idx].f[0] - st[idx - 1].f[0]);
         return sScale(loadSpec(st[idx - 1].s.v), 1 - w) + sScale(loadSpec(st[idx].s.v), w);
      }
      if (t.kind == BLINGCU_TEX_IMAGE) return imageSpectrum(sc, t, dg);
      return loadSpec(t.s.v);
   }
};
template <> struct SpectrumValue<0> {   // deepest level: upload rejects blends nested further
   static HD Spec eval(const DScene &sc, int id, const DG &dg) {
      const blingcu_texture &t = sc.textures[selectSpectrumTexture(sc, id, dg)];
      if (t.kind == BLINGCU_TEX_IMAGE) return imageSpectrum(sc, t, dg);
      return loadSpec(t.s.v);
   }
};
enum { BL_BLEND_DEPTH = 2 };

// ------------------------------------------------------------------------------------------ bump mapping (Reflection.hs:347-377)
// The displaced copies differ from dgs in p and u (or v) only; their normals involve dndu / dndv but no texture in
// scope reads dgN.
HDNI DG bumpDG(const DScene &sc, int dtex, const DG &dgg, const DG &dgs) {
   const float du = 0.01f, dv = 0.01f;
   float disp = evalScalarTexture(sc, dtex, dgs);
   DG eu = dgs; eu.p = dgs.p + scl(du, dgs.dpdu); eu.u = dgs.u + du;
   DG ev = dgs; ev.p = dgs.p + scl(dv, dgs.dpdv); ev.v = dgs.v + dv;
   float uDisp = evalScalarTexture(sc, dtex, eu), vDisp = evalScalarTexture(sc, dtex, ev);
   float vscale = (vDisp - disp) / dv, uscale = (uDisp - disp) / du;
   DG o = dgs;
   o.dpdv = dgs.dpdv + scl(vscale, dgs.n);
   o.dpdu = dgs.dpdu + scl(uscale, dgs.n);
   V3 nn = normalize3(cross3(o.dpdu, o.dpdv));
   o.n = (dot3(nn, dgg.n) < 0) ? -nn : nn;   // faceForward nn' (dgN dgg)
   return o;
}

}  // namespace bl
