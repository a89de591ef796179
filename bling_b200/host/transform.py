"""float32 affine transforms following Graphics/Bling/Transform.hs (host side)."""
from __future__ import annotations

import numpy as np

F = np.float32


class Transform:
    """MkTransform matrix inverse (Transform.hs:120-124); row-major 4x4 float32."""
    __slots__ = ("m", "i")

    def __init__(self, m, i):
        self.m = np.asarray(m, F).reshape(4, 4); self.i = np.asarray(i, F).reshape(4, 4)

    def __mul__(self, other):  # a <> b = concatTrans a b: apply a first (Transform.hs:241-244: mul m1 m2 = M2.M1)
        return Transform(_mm(other.m, self.m), _mm(self.i, other.i))

    def inverse(self): return Transform(self.i, self.m)


def _mm(a, b):
    out = np.zeros((4, 4), F)
    for r in range(4):
        for c in range(4):
            s = F(0)
            for k in range(4): s = F(s + F(a[r, k] * b[k, c]))
            out[r, c] = s
    return out


def identity(): return Transform(np.eye(4, dtype=F), np.eye(4, dtype=F))


def invert(m):  # Gauss-Jordan with full pivoting in float32 (Transform.hs:44-84)
    a = np.array(m, F).reshape(4, 4).copy()
    n = 4; ipiv = [0] * n; indxr = [0] * n; indxc = [0] * n
    for i in range(n):
        big = F(-1); irow = icol = 0
        for j in range(n):
            if ipiv[j] != 1:
                for k in range(n):
                    if ipiv[k] == 0 and abs(a[j, k]) >= big:
                        big = abs(a[j, k]); irow, icol = j, k
        ipiv[icol] += 1
        if irow != icol: a[[irow, icol]] = a[[icol, irow]]
        indxr[i], indxc[i] = irow, icol
        piv = F(F(1) / a[icol, icol]); a[icol, icol] = F(1)
        a[icol, :] = (a[icol, :] * piv).astype(F)
        for j in range(n):
            if j != icol:
                save = a[j, icol]; a[j, icol] = F(0)
                a[j, :] = (a[j, :] - (a[icol, :] * save).astype(F)).astype(F)
    for j in range(n - 1, -1, -1):
        if indxr[j] != indxc[j]: a[:, [indxr[j], indxc[j]]] = a[:, [indxc[j], indxr[j]]]
    return a


def from_matrix(m): return Transform(m, invert(m))


def translate(v):
    m = np.eye(4, dtype=F); i = np.eye(4, dtype=F)
    m[:3, 3] = np.asarray(v, F); i[:3, 3] = -np.asarray(v, F)
    return Transform(m, i)


def scale(v):
    v = np.asarray(v, F); m = np.eye(4, dtype=F); i = np.eye(4, dtype=F)
    for k in range(3): m[k, k] = v[k]; i[k, k] = F(F(1) / v[k])
    return Transform(m, i)


def _radians(x): return F(F(F(x) / F(180)) * F(np.pi))  # Math.hs:64 radians x = x / 180 * pi


def rotate(axis, deg):
    r = _radians(deg); s, c = F(np.sin(r)), F(np.cos(r))
    m = np.eye(4, dtype=F)
    if axis == 0: m[1, 1], m[1, 2], m[2, 1], m[2, 2] = c, -s, s, c
    elif axis == 1: m[0, 0], m[0, 2], m[2, 0], m[2, 2] = c, s, -s, c
    else: m[0, 0], m[0, 1], m[1, 0], m[1, 1] = c, -s, s, c
    return Transform(m, m.T.copy())


def _normalize(v):
    v = np.asarray(v, F); l2 = F(v[0] * v[0] + v[1] * v[1] + v[2] * v[2])
    if l2 != 0: return (v * F(F(1) / F(np.sqrt(l2)))).astype(F)
    return np.array([0, 1, 0], F)


def _cross(u, w):
    return np.array([u[1] * w[2] - u[2] * w[1], -(u[0] * w[2] - u[2] * w[0]), u[0] * w[1] - u[1] * w[0]], F)


def look_at(pos, look, up):  # Transform.hs:221-235
    pos, look, up = (np.asarray(x, F) for x in (pos, look, up))
    d = _normalize(look - pos); left = _normalize(_cross(_normalize(up), d)); u = _cross(d, left)
    m = np.eye(4, dtype=F)
    m[:3, 0], m[:3, 1], m[:3, 2], m[:3, 3] = left, u, d, pos
    return from_matrix(m)


def perspective(fov, n, f):  # Transform.hs:203-213
    fov, n, f = F(fov), F(n), F(f)
    it = F(F(1) / F(np.tan(F(_radians(fov) / F(2)))))
    m = np.zeros((4, 4), F)
    m[0, 0] = m[1, 1] = 1; m[2, 2] = F(f / F(f - n)); m[2, 3] = F(F(-f * n) / F(f - n)); m[3, 2] = 1
    return scale([it, it, 1]) * from_matrix(m)


def trans_point(t: Transform, p):  # Transform.hs:246-256
    m = t.m; x, y, z = (F(c) for c in p)
    r = [F(F(F(m[k, 0] * x + m[k, 1] * y) + m[k, 2] * z) + m[k, 3]) for k in range(4)]
    if r[3] == 1: return np.array(r[:3], F)
    return np.array([r[0] / r[3], r[1] / r[3], r[2] / r[3]], F)


def trans_points(t: Transform, ps: np.ndarray) -> np.ndarray:
    """vectorised transPoint with the same left-to-right float32 evaluation order."""
    m = t.m; ps = np.asarray(ps, F)
    x, y, z = ps[:, 0], ps[:, 1], ps[:, 2]
    r = [(((m[k, 0] * x + m[k, 1] * y).astype(F) + m[k, 2] * z).astype(F) + m[k, 3]).astype(F) for k in range(4)]
    w = r[3]
    out = np.stack(r[:3], 1)
    nz = w != 1
    if nz.any(): out[nz] = (out[nz] / w[nz, None]).astype(F)
    return out.astype(F)


def trans_vector(t: Transform, v):
    m = t.m; x, y, z = (F(c) for c in v)
    return np.array([F(F(m[k, 0] * x + m[k, 1] * y) + m[k, 2] * z) for k in range(3)], F)


def trans_normal(t: Transform, n):  # Transform.hs:267-272: transpose of the inverse
    m = t.i; x, y, z = (F(c) for c in n)
    return np.array([F(F(m[0, k] * x + m[1, k] * y) + m[2, k] * z) for k in range(3)], F)
