"""Spectrum construction in float32, following Graphics/Bling/Spectrum.hs (host side; stays Haskell in a real
deployment). Only construction lives here: the per-band arithmetic of the hot path is in the CUDA kernels."""
from __future__ import annotations

from pathlib import Path

import numpy as np

F = np.float32
BANDS = 16
L_START, L_END = F(400), F(700)          # Spectrum.hs:38-44
_T = np.load(Path(__file__).resolve().parent.parent / "data" / "spectral_tables.npz")


def lerp(t, a, b):                        # Math.hs:116-118
    t, a, b = F(t), F(a), F(b)
    return F(F(F(1) - t) * a + t * b)


class RegularSpd:                         # Spectrum.hs:188-191
    def __init__(self, l0, l1, amps):
        self.l0, self.l1, self.a = F(l0), F(l1), np.asarray(amps, F)

    def eval(self, l):                    # :262-271
        l = F(l)
        if l <= self.l0: return self.a[0]
        if l >= self.l1: return self.a[-1]
        n = len(self.a)
        d1 = F(F(1) / F(F(self.l1 - self.l0) / F(n - 1)))
        x = F(F(l - self.l0) * d1)
        b0 = int(np.floor(x)); b1 = min(b0 + 1, n - 1)
        dx = F(x - F(b0))
        return F(F(F(1) - dx) * self.a[b0] + dx * self.a[b1])

    def avg(self, l0, l1):                # :286-295
        l0, l1 = F(l0), F(l1)
        if l1 <= self.l0: return self.a[0]
        if l0 >= self.l1: return self.a[-1]
        n = len(self.a)
        i0 = max(0, min(n, int(np.floor(F(F(n) * F(F(l0 - self.l0) / F(self.l1 - self.l0)))))))
        i1 = max(0, min(n, int(np.floor(F(F(n) * F(F(l1 - self.l0) / F(self.l1 - self.l0)))))))
        sl = self.a[i0:i1 + 1]
        return F(_sum32(sl) / F(len(sl)))


class IrregularSpd:                       # Spectrum.hs:184-187, mkSpd :203-210
    def __init__(self, pairs):
        pairs = sorted(pairs, key=lambda p: p[0])
        self.ls = np.array([p[0] for p in pairs], F)
        self.vs = np.array([p[1] for p in pairs], F)

    def eval(self, l):                    # :246-260
        l = F(l)
        if l <= self.ls[0]: return self.vs[0]
        if l >= self.ls[-1]: return self.vs[-1]
        lo, hi = 0, len(self.ls) - 1
        while True:
            mid = (lo + hi) // 2
            if lo == mid: i = lo; break
            if self.ls[mid] == l: i = mid; break
            if self.ls[mid] < l: lo = mid
            else: hi = mid
        t = F(F(l - self.ls[i]) / F(self.ls[i + 1] - self.ls[i]))
        return lerp(t, self.vs[i], self.vs[i + 1])

    def avg(self, l0, l1):                # :297-304
        l0, l1 = F(l0), F(l1)
        if l1 <= self.ls[0]: return self.vs[0]
        if l0 >= self.ls[-1]: return self.vs[-1]
        ge0 = np.nonzero(self.ls >= l0)[0]; ge1 = np.nonzero(self.ls >= l1)[0]
        i0 = int(ge0[0]) if len(ge0) else 0
        i1 = int(ge1[0]) if len(ge1) else len(self.vs) - 1
        sl = self.vs[i0:i1 + 1]
        return F(_sum32(sl) / F(len(sl)))


class FuncSpd:                            # SpdFunc; avgSpd fallback :306
    def __init__(self, f): self.f = f
    def eval(self, l): return F(self.f(F(l)))
    def avg(self, l0, l1): return F(F(self.eval(l0) + self.eval(l1)) * F(0.5))


def _sum32(a):
    s = F(0)
    for x in a: s = F(s + F(x))
    return s


def from_spd(spd) -> np.ndarray:          # Spectrum.hs:319-326
    out = np.zeros(BANDS, F)
    for i in range(BANDS):
        l0 = lerp(F(i) / F(BANDS), L_START, L_END)
        l1 = lerp(F(i + 1) / F(BANDS), L_START, L_END)
        out[i] = spd.avg(l0, l1)
    return out


CIE_START, CIE_END = 360, 830
cie_x_spd = RegularSpd(CIE_START, CIE_END, _T["cieXValues"])
cie_y_spd = RegularSpd(CIE_START, CIE_END, _T["cieYValues"])
cie_z_spd = RegularSpd(CIE_START, CIE_END, _T["cieZValues"])
CIE_X, CIE_Y, CIE_Z = from_spd(cie_x_spd), from_spd(cie_y_spd), from_spd(cie_z_spd)   # :328-335
CIE_Y_SUM = _sum32(CIE_Y)                                                             # :337-338


def _rgb_func(name): return from_spd(RegularSpd(380, 720, _T[name]))                 # :537-544


REFL = [_rgb_func("rgbRefl" + n) for n in ("Red", "Green", "Blue", "Cyan", "Magenta", "Yellow", "White")]
ILLUM = [_rgb_func("rgbIllum" + n) for n in ("Red", "Green", "Blue", "Cyan", "Magenta", "Yellow", "White")]


def rgb_to_spectrum(base, rgb) -> np.ndarray:   # Spectrum.hs:146-159; base order r g b c m y w
    r, g, b = (F(x) for x in rgb)
    rb, gb, bb, cb, mb, yb, wb = base
    sc = lambda s, f: (s * F(f)).astype(F)
    if r <= g and r <= b:
        return (sc(wb, r) + ((sc(cb, g - r) + sc(bb, b - g)) if g <= b else (sc(cb, b - r) + sc(gb, g - b)))).astype(F)
    if g <= r and g <= b:
        return (sc(wb, g) + ((sc(mb, r - g) + sc(bb, b - r)) if r <= b else (sc(mb, b - g) + sc(rb, r - b)))).astype(F)
    return (sc(wb, b) + ((sc(yb, r - b) + sc(gb, g - r)) if r <= b else (sc(yb, g - b) + sc(rb, r - g)))).astype(F)


def rgb_refl(rgb): return rgb_to_spectrum(REFL, rgb)
def rgb_illum(rgb): return rgb_to_spectrum(ILLUM, rgb)


def s_y(s) -> np.float32:                 # Spectrum.hs:371-373
    return F(_sum32((np.asarray(s, F) * CIE_Y).astype(F)) / CIE_Y_SUM)


def spd_to_xyz(spd):                      # Spectrum.hs:309-317
    ls = range(CIE_START, CIE_END + 1)
    vs = [spd.eval(F(l)) for l in ls]
    yint = _sum32([cie_y_spd.eval(F(l)) for l in ls])
    x = _sum32([F(cie_x_spd.eval(F(l)) * v) for l, v in zip(ls, vs)])
    y = _sum32([F(cie_y_spd.eval(F(l)) * v) for l, v in zip(ls, vs)])
    z = _sum32([F(cie_z_spd.eval(F(l)) * v) for l, v in zip(ls, vs)])
    return F(x / yint), F(y / yint), F(z / yint)


cie_s0 = RegularSpd(300, 830, _T["cieS0"]); cie_s1 = RegularSpd(300, 830, _T["cieS1"]); cie_s2 = RegularSpd(300, 830, _T["cieS2"])
_sxyz = None


def daylight_xyz():                       # s0XYZ s1XYZ s2XYZ, Spectrum.hs:226-233
    global _sxyz
    if _sxyz is None:
        _sxyz = [np.array(spd_to_xyz(s), F) for s in (cie_s0, cie_s1, cie_s2)]
    return _sxyz


def black_body(temp) -> np.ndarray:       # Spectrum.hs:480-495
    temp = F(temp)

    def planck(w):
        wp = F(F(w) * F(1e-9))
        p5 = F(F(1) / F(wp * wp * wp * wp * wp))
        return F(F(F(0.4e-9) * F(F(3.74183e-16) * p5)) / F(np.exp(F(F(1.4388e-2) / F(wp * temp))) - F(1)))
    return from_spd(FuncSpd(planck))


def sun_curves():
    return (RegularSpd(380, 750, _T["solCurve"]),
            IrregularSpd(list(zip(_T["koCurve_l"], _T["koCurve_a"]))),
            IrregularSpd(list(zip(_T["kgCurve_l"], _T["kgCurve_a"]))),
            IrregularSpd(list(zip(_T["kwaCurve_l"], _T["kwaCurve_a"]))))


def fr_approx_eta(r: np.ndarray) -> np.ndarray:
    """frApproxEta (Fresnel.hs:72-74): (1 + sqrt r') / (1 - sqrt r'), r' = clamp 0 0.999 r, in float32."""
    rp = np.sqrt(np.clip(np.asarray(r, F), F(0), F(0.999))).astype(F)
    return ((F(1) + rp) / (F(1) - rp)).astype(F)


def fr_approx_k(r: np.ndarray) -> np.ndarray:
    """frApproxK (Fresnel.hs:76-78): 2 * sqrt (refl / (1 - refl)), refl = clamp 0 0.999 r, in float32."""
    refl = np.clip(np.asarray(r, F), F(0), F(0.999)).astype(F)
    return (np.sqrt((refl / (F(1) - refl)).astype(F)).astype(F) * F(2)).astype(F)
