"""Host-side Perlin noise / fbm in float32 (Texture.hs:329-414), needed where the reference evaluates a ScalarMap2d while it
BUILDS geometry (heightMap, Primitive/Heightmap.hs:22-49). The stand-in for the Haskell host's own `fbm`; the kernels and the
oracle carry their own statements (bling_b200/csrc/textures.h, oracle/oracle_texture.h), tests/test_textures.py holds all three
against each other."""
import math

import numpy as np

F = np.float32

# noisePerms (Texture.hs:400-414): Ken Perlin's reference permutation, twice over
_L = [151,160,137,91,90,15,131,13,201,95,96,53,194,233,7,225,140,36,103,30,69,142,8,99,37,240,21,10,23,190,6,148,247,120,234,75,0,26,
      197,62,94,252,219,203,117,35,11,32,57,177,33,88,237,149,56,87,174,20,125,136,171,168,68,175,74,165,71,134,139,48,27,166,77,146,
      158,231,83,111,229,122,60,211,133,230,220,105,92,41,55,46,245,40,244,102,143,54,65,25,63,161,1,216,80,73,209,76,132,187,208,89,
      18,169,200,196,135,130,116,188,159,86,164,100,109,198,173,186,3,64,52,217,226,250,124,123,5,202,38,147,118,126,255,82,85,212,207,
      206,59,227,47,16,58,17,182,189,28,42,223,183,170,213,119,248,152,2,44,154,163,70,221,153,101,155,167,43,172,9,129,22,39,253,19,98,
      108,110,79,113,224,232,178,185,112,104,218,246,97,228,251,34,242,193,238,210,144,12,191,179,162,241,81,51,145,235,249,14,239,107,
      49,192,214,31,181,199,106,157,184,84,204,176,115,121,50,45,127,4,150,254,138,236,205,93,222,114,67,29,24,72,243,141,128,195,78,66,
      215,61,156,180]
PERM = _L + _L


def noise_weight(t):
    t3 = F(F(t * t) * t); t4 = F(t3 * t)
    return F(F(F(F(6) * t4) * t) - F(F(15) * t4)) + F(F(10) * t3)


def grad(x, y, z, dx, dy, dz):
    h = PERM[PERM[PERM[x] + y] + z] & 15
    up = dx if (h < 8 or h in (12, 13)) else dy
    vp = dy if (h < 4 or h in (12, 13)) else dz
    u = -up if h & 1 else up
    v = -vp if h & 2 else vp
    return F(u + v)


def lerp(t, a, b): return F(F(F(F(1) - t) * a) + F(t * b))


def perlin3d(x, y, z):
    x, y, z = F(x), F(y), F(z)
    ixp, iyp, izp = math.floor(x), math.floor(y), math.floor(z)
    dx, dy, dz = F(x - F(ixp)), F(y - F(iyp)), F(z - F(izp))
    ix, iy, iz = ixp & 255, iyp & 255, izp & 255
    o = F(1)
    w = {}
    for a in (0, 1):
        for b in (0, 1):
            for c in (0, 1):
                w[a, b, c] = grad(ix + a, iy + b, iz + c, F(dx - o) if a else dx, F(dy - o) if b else dy, F(dz - o) if c else dz)
    wx, wy, wz = noise_weight(dx), noise_weight(dy), noise_weight(dz)
    x00, x10 = lerp(wx, w[0, 0, 0], w[1, 0, 0]), lerp(wx, w[0, 1, 0], w[1, 1, 0])
    x01, x11 = lerp(wx, w[0, 0, 1], w[1, 0, 1]), lerp(wx, w[0, 1, 1], w[1, 1, 1])
    return lerp(wz, lerp(wy, x00, x10), lerp(wy, x01, x11))


def fbm(octaves, omega, p):
    s, l, o = F(0), F(1), F(1)
    for _ in range(octaves):
        s = F(s + F(o * perlin3d(F(p[0] * l), F(p[1] * l), F(p[2] * l))))
        l = F(F(1.99) * l); o = F(F(omega) * o)
    return s
