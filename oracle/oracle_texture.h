// ORACLE (test infrastructure, NOT product code). See oracle_math.h header. PARITY UNPINNED.
// Scalar textures, computed spectrum textures and bump mapping -- restating Texture.hs:129-414 and
// Reflection.hs:344-377 of /root/reference/src/lib/Graphics/Bling, in the shape of the Haskell (lists, recursion).
#pragma once
#include "oracle_scene.h"
#include <vector>

namespace orc {

// Texture.hs:400-414 noisePerms = l ++ l
static const int kNoiseL[256] = {
   151,160,137,91,90,15,131,13,201,95,96,53,194,233,7,225,140,36,103,30,69,142,8,99,37,240,21,10,23,190,6,148,247,120,234,75,0,26,
   197,62,94,252,219,203,117,35,11,32,57,177,33,88,237,149,56,87,174,20,125,136,171,168,68,175,74,165,71,134,139,48,27,166,77,146,
   158,231,83,111,229,122,60,211,133,230,220,105,92,41,55,46,245,40,244,102,143,54,65,25,63,161,1,216,80,73,209,76,132,187,208,89,
   18,169,200,196,135,130,116,188,159,86,164,100,109,198,173,186,3,64,52,217,226,250,124,123,5,202,38,147,118,126,255,82,85,212,207,
   206,59,227,47,16,58,17,182,189,28,42,223,183,170,213,119,248,152,2,44,154,163,70,221,153,101,155,167,43,172,9,129,22,39,253,19,98,
   108,110,79,113,224,232,178,185,112,104,218,246,97,228,251,34,242,193,238,210,144,12,191,179,162,241,81,51,145,235,249,14,239,107,
   49,192,214,31,181,199,106,157,184,84,204,176,115,121,50,45,127,4,150,254,138,236,205,93,222,114,67,29,24,72,243,141,128,195,78,66,
   215,61,156,180};
static inline int noisePerms(int i) { return kNoiseL[i % 256]; }   // i in [0, 512)

// Texture.hs:382-385
static inline float noiseWeight(float t) { float t3 = t * t * t; float t4 = t3 * t; return 6 * t4 * t - 15 * t4 + 10 * t3; }
// Texture.hs:387-394
static inline float grad(int x, int y, int z, float dx, float dy, float dz) {
   int hp = noisePerms(noisePerms(noisePerms(x) + y) + z);
   int h = hp & 15;
   float up = (h < 8 || h == 12 || h == 13) ? dx : dy;
   float vp = (h < 4 || h == 12 || h == 13) ? dy : dz;
   float u = ((h & 1) != 0) ? -up : up;
   float v = ((h & 2) != 0) ? -vp : vp;
   return u + v;
}
// Texture.hs:350-380
static inline float perlin3d(float x, float y, float z) {
   long ixp = (long)std::floor(x), iyp = (long)std::floor(y), izp = (long)std::floor(z);
   float dx = x - (float)ixp, dy = y - (float)iyp, dz = z - (float)izp;
   int ix = (int)(ixp & 255), iy = (int)(iyp & 255), iz = (int)(izp & 255);
   float w000 = grad(ix, iy, iz, dx, dy, dz);
   float w100 = grad(ix + 1, iy, iz, dx - 1, dy, dz);
   float w010 = grad(ix, iy + 1, iz, dx, dy - 1, dz);
   float w110 = grad(ix + 1, iy + 1, iz, dx - 1, dy - 1, dz);
   float w001 = grad(ix, iy, iz + 1, dx, dy, dz - 1);
   float w101 = grad(ix + 1, iy, iz + 1, dx - 1, dy, dz - 1);
   float w011 = grad(ix, iy + 1, iz + 1, dx, dy - 1, dz - 1);
   float w111 = grad(ix + 1, iy + 1, iz + 1, dx - 1, dy - 1, dz - 1);
   float wx = noiseWeight(dx), wy = noiseWeight(dy), wz = noiseWeight(dz);
   float x00 = lerpf(wx, w000, w100), x10 = lerpf(wx, w010, w110);
   float x01 = lerpf(wx, w001, w101), x11 = lerpf(wx, w011, w111);
   float y0 = lerpf(wy, x00, x10), y1 = lerpf(wy, x01, x11);
   return lerpf(wz, y0, y1);
}
// Texture.hs:329-339: sum (take octaves [o * perlin3d (p * l) | (l, o) <- zip (iterate (1.99 *) 1) (iterate (omega *) 1)])
static inline float fbm(int octaves, float omega, V3 p) {
   std::vector<float> terms;
   float l = 1, o = 1;
   for (int i = 0; i < octaves; ++i) { terms.push_back(o * perlin3d(p.x * l, p.y * l, p.z * l)); l = 1.99f * l; o = omega * o; }
   float s = 0; for (float t : terms) s = s + t;   // GHC.List.sum = foldl (+) 0
   return s;
}

// Texture.hs:255-303 (Int is 64-bit)
static inline float cellNoise(int distKind, V3 p) {
   auto lcg = [](int64_t x) -> int64_t { return (int64_t)(((__int128)1103515245 * x + 12345) % 4294967296LL); };
   auto hash = [](int64_t x, int64_t y, int64_t z) -> int64_t {
      int64_t v = (int64_t)((uint64_t)x * 73856093ULL) ^ (int64_t)((uint64_t)y * 19349663ULL) ^ (int64_t)((uint64_t)z * 83492791ULL);
      int64_t a = (v == std::numeric_limits<int64_t>::min()) ? v : (v < 0 ? -v : v);   // abs minBound = minBound
      return a % 4294967296LL;                                                         // rem
   };
   auto prob = [](int64_t v) -> int {
      static const int64_t lut[8] = {393325350LL, 1022645910LL, 1861739990LL, 2700834071LL, 3372109335LL, 3819626178LL, 4075350088LL, 4203212043LL};
      for (int i = 0; i < 8; ++i) if (v < lut[i]) return i + 1;
      return 9;
   };
   auto dist = [distKind](V3 a, V3 b) -> float {
      V3 d = a - b;
      switch (distKind) {
      case 0: return len(d);                                             // euclidianDist
      case 1: return sqLen(d);                                           // sqEuclidianDist
      case 2: return std::fabs(d.x) + std::fabs(d.y) + std::fabs(d.z);   // manhattanDist
      default: return hmax(hmax(std::fabs(d.x), std::fabs(d.y)), std::fabs(d.z));   // chebyshevDist
      }
   };
   int64_t ox = (int64_t)std::floor(p.x), oy = (int64_t)std::floor(p.y), oz = (int64_t)std::floor(p.z);
   std::vector<V3> all;
   for (int x = -1; x <= 1; ++x) for (int y = -1; y <= 1; ++y) for (int z = -1; z <= 1; ++z) {
      int64_t cx = x + ox, cy = y + oy, cz = z + oz;
      int64_t us = lcg(hash(cx, cy, cz));
      int n = prob(us);
      int64_t u0 = us;
      for (int k = 0; k < n; ++k) {   // take n $ tail $ iterate go (undefined, us)
         int64_t u1 = lcg(u0), u2 = lcg(u1), u3 = lcg(u2);
         float fx = (float)u1 / 4294967296.0f, fy = (float)u2 / 4294967296.0f, fz = (float)u3 / 4294967296.0f;
         all.push_back(mk((float)cx + fx, (float)cy + fy, (float)cz + fz));
         u0 = u3;
      }
   }
   float best = kInf;
   for (V3 q : all) best = hmin(best, dist(p, q));
   return best;
}

// Texture.hs:305-326. enumFromThen on Float is numericEnumFromThen n m = n : numericEnumFromThen m (m + m - n) (base 4.9)
static inline float quasiCrystal(int octaves, float x, float y) {
   std::vector<float> angles;
   float n = 0, m = kPi / (float)octaves;
   for (int i = 0; i < octaves; ++i) { angles.push_back(n); float nx = m + m - n; n = m; m = nx; }
   float s = 0;
   for (float th : angles) { float cth = std::cos(th), sth = std::sin(th); s = s + (std::cos(cth * x + sth * y) + 1) / 2; }
   float kf = std::trunc(s), v = s - kf;   // properFraction
   long k = (long)kf;
   if (v < 0) { k = k - 1; v = 1 + v; }
   return (k % 2 != 0) ? 1 - v : v;
}

// Texture.hs:150-152,166-181
static inline void mapping2d(const float *m, const DG &dg, float &x, float &y) {
   if (m[0] == 0.0f) { x = m[1] * dg.u + m[3]; y = m[2] * dg.v + m[4]; return; }   // uvMapping
   x = dot(dg.p, mk(m[1], m[2], m[3])) + m[7];                                      // planarMapping
   y = dot(dg.p, mk(m[4], m[5], m[6])) + m[8];
}

// Texture.hs:91-108: mod' a b = let a' = a - (a `div` b) * b in if a' < 0 then a' + b else a'; getPixel / getPixelScalar
static inline long modP(long a, long b) {
   long n = a / b; if ((a % b != 0) && ((a < 0) != (b < 0))) n -= 1;   // Haskell `div` rounds towards minus infinity
   long ap = a - n * b;
   return ap < 0 ? ap + b : ap;
}
static inline const float *imagePixelAt(const blingcu_image &im, float u, float v) {
   long px = modP((long)std::floor(u * (float)im.width), im.width);
   long py = modP((long)std::floor((-v) * (float)im.height), im.height);
   return im.data + ((size_t)py * (size_t)im.width + (size_t)px) * (size_t)im.channels;
}

struct TexEnv { const std::vector<blingcu_texture> &tex; const std::vector<blingcu_image> &images; };

static float evalScalarTexture(const TexEnv &env, int id, const DG &dg) {   // MaterialParser.hs:113-154
   const std::vector<blingcu_texture> &tex = env.tex;
   const blingcu_texture &t = tex[id];
   switch (t.kind) {
   case BLINGCU_STEX_IMAGE: { float x, y; mapping2d(t.s.v, dg, x, y); return imagePixelAt(env.images[t.aux], x, y)[0]; }   // fromIntegral x / 255 done by the host
   case BLINGCU_STEX_CONSTANT: return t.f[0];
   case BLINGCU_STEX_SCALE: return t.f[0] + t.f[1] * evalScalarTexture(env, t.child[0], dg);   // scaleTexture a s t dg = a + s * t dg
   case BLINGCU_STEX_PERLIN: { V3 q = transPoint(t.s.v, dg.p); return perlin3d(q.x, q.y, q.z); }
   case BLINGCU_STEX_FBM: return fbm(t.aux, t.f[0], transPoint(t.s.v, dg.p));
   case BLINGCU_STEX_CELLNOISE: return cellNoise(t.aux, transPoint(t.s.v, dg.p));
   case BLINGCU_STEX_CRYSTAL: { float x, y; mapping2d(t.s.v, dg, x, y); return quasiCrystal(t.aux, x, y); }
   default: return 0;
   }
}

// Reflection.hs:347-377
static DG bump(const TexEnv &tex, int d, const DG &dgg, const DG &dgs) {
   const float du = 0.01f, dv = 0.01f;
   DG dgeu = dgs; dgeu.p = dgs.p + scl(du, dgs.dpdu); dgeu.u = dgs.u + du;   // dgN of the shifted copies is never read
   DG dgev = dgs; dgev.p = dgs.p + scl(dv, dgs.dpdv); dgev.v = dgs.v + dv;
   float uDisp = evalScalarTexture(tex, d, dgeu), vDisp = evalScalarTexture(tex, d, dgev), disp = evalScalarTexture(tex, d, dgs);
   float vscale = (vDisp - disp) / dv;
   V3 dpdv = dgs.dpdv + scl(vscale, dgs.n);
   float uscale = (uDisp - disp) / du;
   V3 dpdu = dgs.dpdu + scl(uscale, dgs.n);
   V3 nnp = normalize(cross(dpdu, dpdv));
   V3 nn = (dot(nnp, dgg.n) < 0) ? -nnp : nnp;   // faceForward (Math.hs:365-369)
   DG o = dgs; o.n = nn; o.dpdu = dpdu; o.dpdv = dpdv;
   return o;
}

}  // namespace orc
