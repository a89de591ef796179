"""Infinite-area-light construction (host side): radiance map + Dist2D, following Light.hs:72-82,
Montecarlo.hs:34-104, SunSky.hs and IO/Bitmap.hs:22-29, in float32."""
from __future__ import annotations

from pathlib import Path

import numpy as np

from .. import ir as IR
from . import spectra as S
from . import transform as T

F = np.float32


def _cumsum32(rows: np.ndarray) -> np.ndarray:
    """scanl (\\c f -> c + f / n) 0 per row, sequential float32 adds (Montecarlo.hs:45)."""
    n = rows.shape[1]
    out = np.zeros((rows.shape[0], n + 1), F)
    step = (rows / F(n)).astype(F)
    for j in range(n):
        out[:, j + 1] = (out[:, j] + step[:, j]).astype(F)
    return out


def mk_dist1d(rows: np.ndarray):
    """vectorised mkDist1D over rows (Montecarlo.hs:40-50) -> (func, cdf, funcInt)"""
    rows = np.asarray(rows, F)
    n = rows.shape[1]
    i = _cumsum32(rows)
    fi = i[:, -1].copy()
    cdf = np.empty_like(i)
    nz = fi != 0
    cdf[nz] = (i[nz] / fi[nz, None]).astype(F)
    cdf[~nz] = (np.arange(n + 1, dtype=F) / F(n)).astype(F)
    return rows, cdf, fi


def rgb_to_spectrum_illum_vec(rgb: np.ndarray) -> np.ndarray:
    """vectorised rgbToSpectrumIllum (Spectrum.hs:146-159): (...,3) -> (...,16)"""
    rgb = np.asarray(rgb, F); r, g, b = rgb[..., 0:1], rgb[..., 1:2], rgb[..., 2:3]
    rb, gb, bb, cb, mb, yb, wb = (x[None, :] for x in S.ILLUM)
    sh = rgb.shape[:-1] + (16,)
    r, g, b = (np.broadcast_to(x, rgb.shape[:-1] + (1,)).reshape(-1, 1) for x in (r, g, b))
    m = lambda s, f: (s * f).astype(F)
    c1 = (r <= g) & (r <= b); c2 = ~c1 & (g <= r) & (g <= b); c3 = ~c1 & ~c2
    o1 = np.where(g <= b, m(wb, r) + (m(cb, g - r) + m(bb, b - g)), m(wb, r) + (m(cb, b - r) + m(gb, g - b)))
    o2 = np.where(r <= b, m(wb, g) + (m(mb, r - g) + m(bb, b - r)), m(wb, g) + (m(mb, b - g) + m(rb, r - b)))
    o3 = np.where(r <= b, m(wb, b) + (m(yb, r - b) + m(gb, g - r)), m(wb, b) + (m(yb, g - b) + m(rb, r - g)))
    out = np.where(c1, o1, np.where(c2, o2, o3)).astype(F)
    return out.reshape(sh)


def s_y_vec(spec: np.ndarray) -> np.ndarray:
    acc = np.zeros(spec.shape[:-1], F)
    for i in range(16): acc = (acc + (spec[..., i] * S.CIE_Y[i]).astype(F)).astype(F)
    return (acc / S.CIE_Y_SUM).astype(F)


# ------------------------------------------------------------------------------------------- SunSky.hs
def init_sky(east, sdw, turb):
    """mkSunSkyLight / initSky (SunSky.hs:12-65): returns the blingcu_sunsky POD."""
    up = T._normalize([0, 1, 0]); e = T._normalize(east)
    # coordinateSystem' w v (Math.hs:439-444): w' = normalize w; u = normalize (v x w'); v' = w' x u ; LocalCoordinates u v' w'
    w = T._normalize(up); u = T._normalize(T._cross(e, w)); v = T._cross(w, u)
    w2l = lambda x: np.array([F(np.dot(x, u)), F(np.dot(x, v)), F(np.dot(x, w))], F)
    sdw = np.asarray(sdw, F)
    sd = T._normalize(w2l(T._normalize(sdw)))
    st = F(np.arccos(np.clip(sd[2], -1, 1)))
    t = F(turb); st2, st3, t2 = F(st * st), F(st * st * st), F(t * t)
    chi = F(F(F(4) / F(9) - t / F(120)) * F(F(np.pi) - 2 * st))
    k = IR.SunSky()
    IR.set_arr(k.sun_dir, sd); k.sun_theta = float(st)
    IR.set_arr(k.sun_disc_dir, T._normalize(w2l(sdw)))
    IR.set_arr(k.perez_Y, [0.17872 * t - 1.46303, -0.35540 * t + 0.42749, -0.02266 * t + 5.32505, 0.12064 * t - 2.57705, -0.06696 * t + 0.37027])
    IR.set_arr(k.perez_x, [-0.01925 * t - 0.25922, -0.06651 * t + 0.00081, -0.00041 * t + 0.21247, -0.06409 * t - 0.89887, -0.00325 * t + 0.04517])
    IR.set_arr(k.perez_y, [-0.01669 * t - 0.26078, -0.09495 * t + 0.00921, -0.00792 * t + 0.21023, -0.04405 * t - 1.65369, -0.01092 * t + 0.05291])
    k.zenith_Y = float(F(F(F(F(4.04530) * t - F(4.97100)) * F(np.tan(chi)) - F(0.2155) * t + F(2.4192)) * F(1000)))
    k.zenith_x = float(F(F(F(0.00165) * st3 - F(0.00374) * st2 + F(0.00208) * st) * t2 +
                         F(F(-0.02902) * st3 + F(0.06377) * st2 - F(0.03202) * st + F(0.00394)) * t +
                         F(F(0.11693) * st3 - F(0.21196) * st2 + F(0.06052) * st + F(0.25885))))
    k.zenith_y = float(F(F(F(0.00275) * st3 - F(0.00610) * st2 + F(0.00316) * st) * t2 +
                         F(F(-0.04212) * st3 + F(0.08970) * st2 - F(0.04153) * st + F(0.00515)) * t +
                         F(F(0.15346) * st3 - F(0.26756) * st2 + F(0.06669) * st + F(0.26688))))
    s0, s1, s2 = S.daylight_xyz()
    IR.set_arr(k.s0xyz, s0); IR.set_arr(k.s1xyz, s1); IR.set_arr(k.s2xyz, s2)
    IR.set_arr(k.sun_radiance.v, sun_radiance(sd, st, t))
    return k


def sun_radiance(sd, st, turb):               # sunSpectrum' (SunSky.hs:96-126)
    if sd[2] < 0: return np.zeros(16, F)
    sol, ko, kg, kwa = S.sun_curves()
    t = F(st)

    def sf(l):
        l = F(l)
        m = F(F(1) / F(F(np.cos(t)) + F(0.000940) * F(np.power(F(F(1.6386) - t), F(-1.253)))))
        l1k = F(l / F(1000))
        tR = F(np.exp(F(-m * F(0.008735)) * F(np.power(l1k, F(-4.08)))))
        beta = F(F(0.04608365822050) * turb - F(0.04586025928522))
        tA = F(np.exp(F(-m * beta) * F(np.power(l1k, F(-1.3)))))
        tO = F(np.exp(F(-m * ko.eval(l)) * F(0.35)))
        kgv = kg.eval(l)
        tG = F(np.exp(F(F(-1.41) * kgv * m) / F(np.power(F(F(1.0) + F(118.93) * kgv * m), F(0.45)))))
        kw = kwa.eval(l); w = F(2)
        tWA = F(np.exp(F(F(-0.2385) * kw * w * m) / F(np.power(F(F(1) + F(20.07) * kw * w * m), F(0.45)))))
        return F(sol.eval(l) * tR * tA * tO * tG * tWA)
    return S.from_spd(S.FuncSpd(sf))


def _perez_vec(p, sunT, t, g, lvz):
    p = [F(x) for x in p]; sunT = F(sunT); lvz = F(lvz)
    csg = np.cos(g).astype(F); cst = F(np.cos(sunT))
    num = ((1 + p[0] * np.exp((p[1] / np.cos(t)).astype(F)).astype(F)) * (1 + p[2] * np.exp((p[3] * g).astype(F)).astype(F)) + p[4] * csg * csg).astype(F)
    den = F(F(F(1) + p[0] * F(np.exp(p[1]))) * F(F(1) + p[2] * F(np.exp(F(p[3] * sunT)))) + p[4] * cst * cst)
    return (lvz * num / den).astype(F)


def sunsky_eval_vec(k: IR.SunSky, u: np.ndarray, v: np.ndarray) -> np.ndarray:
    """texMapEval of the sunSky map at Cartesian (u,v), vectorised float32 -> (...,16)"""
    u = np.asarray(u, F); v = np.asarray(v, F)
    phi = (u * F(2) * F(np.pi)).astype(F); theta = (v * F(np.pi)).astype(F)
    sint, cost = np.sin(theta).astype(F), np.cos(theta).astype(F)
    d = np.stack([(sint * np.cos(phi).astype(F)).astype(F), (sint * np.sin(phi).astype(F)).astype(F), cost], -1)
    sd = np.array(list(k.sun_dir), F)
    dz = -d[..., 2]
    with np.errstate(all="ignore"):
        th = np.arccos(dz).astype(F)
        gamma = np.arccos(np.clip((d * sd).sum(-1).astype(F), -1, 1)).astype(F)
        x = _perez_vec(list(k.perez_x), k.sun_theta, th, gamma, k.zenith_x)
        y = _perez_vec(list(k.perez_y), k.sun_theta, th, gamma, k.zenith_y)
        Y = (_perez_vec(list(k.perez_Y), k.sun_theta, th, gamma, k.zenith_Y) * F(1e-4)).astype(F)
        den = (F(0.0241) + F(0.2562) * x - F(0.7341) * y).astype(F)
        m1 = ((F(-1.3515) - F(1.7703) * x + F(5.9114) * y) / den).astype(F)
        m2 = ((F(0.03) - F(31.4424) * x + F(30.0717) * y) / den).astype(F)
        s0, s1, s2 = (np.array(list(a), F) for a in (k.s0xyz, k.s1xyz, k.s2xyz))
        cx, cy, cz = (s0[i] + m1 * s1[i] + m2 * s2[i] for i in range(3))
        X = (cx * Y / cy).astype(F); Z = (cz * Y / cy).astype(F)
        r = (F(3.240479) * X - F(1.537150) * Y - F(0.498535) * Z).astype(F)
        g = (F(-0.969256) * X + F(1.875991) * Y + F(0.041556) * Z).astype(F)
        b = (F(0.055648) * X - F(0.204043) * Y + F(1.057311) * Z).astype(F)
        sky = rgb_to_spectrum_illum_vec(np.stack([r, g, b], -1))
    sky = np.where((dz < F(1e-4))[..., None], F(0), sky)
    sdd = np.array(list(k.sun_disc_dir), F) * np.array([1, 1, -1], F)
    dd = (d * sdd).sum(-1).astype(F)
    stm = F(np.sqrt(max(F(0), F(F(1) - F(F(6.955e5) / F(1.496e8))))))
    sun = np.where((dd > stm)[..., None], np.array(list(k.sun_radiance.v), F), F(0))
    return (sky + sun).astype(F)


def synthetic_hdr(w=1024, h=512) -> np.ndarray:
    """SURVEY.md §8(d) cfg 4b: the reference's HDR is a missing blob, so a synthetic RGBF map stands in."""
    y, x = np.mgrid[0:h, 0:w]
    theta = (np.pi * (y + 0.5) / h); phi = (2 * np.pi * (x + 0.5) / w)
    L = 0.2 + 40.0 * np.exp(-((theta - 0.9) ** 2 + (phi - 1.2) ** 2) / 0.01)
    return (L[..., None] * np.array([1.0, 0.9, 0.7])).astype(F)


def make_envmap(env, w2l: T.Transform, base: Path, env_files: dict):
    e = IR.EnvMap(); a = IR.EnvArrays()
    IR.set_arr(e.w2l, w2l.m); IR.set_arr(e.l2w, w2l.i)
    if env[0] == "constant":
        e.kind = IR.ENV_CONSTANT; e.nu = e.nv = 1; IR.set_arr(e.s.v, env[1])
        lum = np.array([[S.s_y(env[1])]], F)
    elif env[0] == "file":
        rgb = env_files.get(env[1]) if env_files else None
        if rgb is None: rgb = env_files.get("*") if env_files else None
        if rgb is None: raise FileNotFoundError(f"environment map {env[1]} (pass env_files={{name: rgb array}})")
        rgb = np.ascontiguousarray(rgb, F); h, w = rgb.shape[:2]
        e.kind = IR.ENV_RGBTABLE; e.nu, e.nv = w, h; a.rgb = rgb
        # dist evaluated at (x/sx, y/sy) through the flipped nearest lookup (Light.hs:76-82, IO/Bitmap.hs:25-29)
        uu = (np.arange(w, dtype=F) / F(w)).astype(F); vv = (np.arange(h, dtype=F) / F(h)).astype(F)
        xi = np.clip(np.floor(((F(1) - uu) * F(w)).astype(F)).astype(np.int64), 0, w - 1)
        yi = np.clip(np.floor(((F(1) - vv) * F(h)).astype(F)).astype(np.int64), 0, h - 1)
        lum = s_y_vec(rgb_to_spectrum_illum_vec(rgb[yi][:, xi]))
    else:
        e.kind = IR.ENV_SUNSKY; e.nu, e.nv = 640, 480
        e.sky = init_sky(env[1], env[2], env[3])
        uu = (np.arange(640, dtype=F) / F(640)).astype(F); vv = (np.arange(480, dtype=F) / F(480)).astype(F)
        U, V = np.meshgrid(uu, vv)
        lum = s_y_vec(sunsky_eval_vec(e.sky, U, V))
    func, cdf, fi = mk_dist1d(lum)                     # conditional (Montecarlo.hs:81-84)
    mfunc, mcdf, mfi = mk_dist1d(fi[None, :])          # marginal over funcInt of the rows
    a.cond_func, a.cond_cdf, a.cond_int = func, cdf, fi
    a.marg_func, a.marg_cdf = mfunc[0], mcdf[0]; e.marg_int = float(mfi[0])
    return e, a
