"""A THIRD, independent statement of reference functions the oracle and the kernels both restate ("every test
compares CUDA with oracle/, a restatement written by the same builder from the same reading of the Haskell"): pure Python / numpy
float32, written from the Haskell text alone (file:line cited at each function), held against oracle/ through oracle_debug_eval.
The kernels are held against the oracle per sample elsewhere (tests/test_host_and_emu.py, tests/test_gpu_parity.py), so a
misreading shared by oracle and kernels shows up here. tests/test_textures.py does the same for the textures.

Arithmetic: numpy float32 scalars, same operation order as the Haskell. Where only + - * / sqrt occur the comparison is
bit-exact; where pow / sin / cos / acos occur (numpy's and glibc's differ in the last place) it is 4e-6 relative."""
import numpy as np
import pytest

from oracle.oracle_py import debug_eval

F = np.float32
PI = F(np.pi)
INV_PI, INV_TWOPI, TWO_PI = F(1) / PI, F(1) / (F(2) * PI), F(2) * PI
TOL = 4e-6


def v3(x, y, z): return np.array([x, y, z], F)
def dot(a, b): return F(F(F(a[0] * b[0]) + F(a[1] * b[1])) + F(a[2] * b[2]))
def sq_len(a): return dot(a, a)
def normalize(a):                                   # Math.hs:358-362
    if sq_len(a) != 0:
        il = F(1) / np.sqrt(sq_len(a)); return v3(a[0] * il, a[1] * il, a[2] * il)
    return v3(0, 1, 0)
def abs_dot(a, b): return F(abs(dot(a, b)))
def cos_theta(v): return F(v[2])
def abs_cos_theta(v): return F(abs(v[2]))
def same_hemisphere(a, b): return F(a[2] * b[2]) > 0        # Reflection.hs:83-84
def spec(c): return np.full(16, c, F)
def close(a, b, tol=TOL):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.all(np.abs(a - b) <= tol * np.maximum(1e-30, np.maximum(np.abs(a), np.abs(b))))


# ------------------------------------------------------------------------------------------------ Fresnel.hs
def fr_dielectric(etai, etat, cosi):               # Fresnel.hs:31-55 (Q4: etai is not used for cost, indices are not swapped)
    etai, etat, cosi = F(etai), F(etat), F(cosi)
    c = max(F(0), F(F(1) - F(cosi * cosi)))
    costp = F(c / F(etat * etat)) if cosi > 0 else F(c * F(etat * etat))
    cl = min(max(costp, F(0)), F(1))                # clamp cost' 0 1
    cost = np.sqrt(F(F(1) - cl))
    acosi = F(abs(cosi))
    eta = spec(etat) / spec(etai)                   # frDiel: frDiel' cosi (sConst cost) (etat / etai)
    r_parl_p = eta * acosi                          # sScale eta cosi
    r_parl = (spec(cost) - r_parl_p) / (spec(cost) + r_parl_p)
    r_perp_p = eta * spec(cost)
    r_perp = (spec(acosi) - r_perp_p) / (spec(acosi) + r_perp_p)
    return ((r_parl * r_parl + r_perp * r_perp) * F(0.5)).astype(F)


def fr_conductor(eta, k, cosi):                    # Fresnel.hs:58-70
    eta, k, acosi = eta.astype(F), k.astype(F), F(abs(F(cosi)))
    ec2 = eta * F(F(2) * acosi)
    tmp_f = eta * eta + k * k
    tmp = tmp_f * F(acosi * acosi)
    a2 = spec(F(acosi * acosi))
    r_per2 = (tmp_f - ec2 + a2) / (tmp_f + ec2 + a2)
    r_par2 = (tmp - ec2 + spec(1)) / (tmp + ec2 + spec(1))
    return ((r_per2 + r_par2) / spec(2)).astype(F)


@pytest.mark.parametrize("seed", range(3))
def test_fresnel_terms(seed):
    rng = np.random.default_rng(seed)
    for _ in range(200):
        etai, etat, cosi = F(rng.uniform(1, 1.2)), F(rng.uniform(1.0, 2.4)), F(rng.uniform(-1, 1))
        assert np.array_equal(fr_dielectric(etai, etat, cosi), debug_eval(1, [etai, etat, cosi], 16)), (etai, etat, cosi)
        eta, k = rng.uniform(0.1, 3, 16).astype(F), rng.uniform(0.5, 5, 16).astype(F)
        assert np.array_equal(fr_conductor(eta, k, cosi), debug_eval(2, [eta, k, cosi], 16))
    # the quirk itself: with cosi < 0 the transmitted cosine is computed with etat^2 MULTIPLIED (no index swap)
    assert not np.allclose(fr_dielectric(1, 1.5, -0.6), fr_dielectric(1.5, 1, 0.6))
    assert abs(float(fr_dielectric(1, 1.5, 1.0)[0]) - 0.04) < 1e-6          # normal incidence on glass: ((n - 1) / (n + 1))^2


# ------------------------------------------------------------------------------------------------ Reflection/Microfacet.hs
def mf_g(wo, wi, wh):                              # Microfacet.hs:113-120
    n_wh, n_wo, n_wi, wo_wh = abs_cos_theta(wh), abs_cos_theta(wo), abs_cos_theta(wi), abs_dot(wo, wh)
    return min(F(1), min(F(F(F(F(2) * n_wh) * n_wo) / wo_wh), F(F(F(F(2) * n_wh) * n_wi) / wo_wh)))
def blinn_d(e, wh): return F(F(F(e + F(2)) * INV_TWOPI) * F(abs_cos_theta(wh) ** F(e)))        # :194-195
def blinn_pdf(e, wh): return F(F(F(e + F(1)) * F(abs_cos_theta(wh) ** F(e))) * INV_TWOPI)      # :146-147
def spherical_direction(sint, cost, phi): return v3(sint * np.cos(F(phi)), sint * np.sin(F(phi)), cost)   # Math.hs:141-143


def blinn_sample(e, u1, u2):                       # Microfacet.hs:175-182
    cost = F(F(u1) ** F(F(1) / F(e + F(1))))
    sint = np.sqrt(max(F(0), F(F(1) - F(cost * cost))))
    phi = F(F(F(u2) * F(2)) * PI)
    wh = spherical_direction(sint, cost, phi)
    f = F(F(cost ** F(e)) * INV_TWOPI)
    return wh, F(F(e + F(2)) * f), F(F(e + F(1)) * f)


def microfacet_eval(e, r, fr, wo, wi):             # Microfacet.hs:20-33: `e wo wi`
    costo, costi = abs_cos_theta(wo), abs_cos_theta(wi)
    if costi == 0 or costo == 0: return spec(0)
    whp = (wi + wo).astype(F)
    if whp[0] == 0 and whp[1] == 0 and whp[2] == 0: return spec(0)
    wh = normalize(whp)
    if cos_theta(wh) < 0: return spec(0)
    costh = dot(wi, wh)
    x = F(F(blinn_d(e, wh) * mf_g(wo, wi, wh)) / F(F(4) * costi))
    return ((r * fr(costh)) * x).astype(F)


def microfacet_pdf(e, wo, wi):                     # :35-41
    whp = (wo + wi).astype(F)
    if sq_len(whp) == 0: return F(0)
    wh = normalize(whp)
    if cos_theta(wh) < 0: return F(0)
    return F(blinn_pdf(e, wh) / F(F(4) * abs_dot(wo, wh)))


def microfacet_sample(e, r, fr, wo, u1, u2, adj=False):   # :43-54
    whp, d, pdf = blinn_sample(e, u1, u2)
    wh = (-whp).astype(F) if cos_theta(whp) < 0 else whp
    wi = (wh * F(F(2) * dot(wo, wh)) - wo).astype(F)       # (2 * wo `dot` wh) *# wh - wo
    cost_h = dot(wo, wh)
    if not same_hemisphere(wo, wi): return spec(0), wo, F(0)
    fact = F(F(F(d * F(abs(cost_h))) / pdf) * mf_g(wo, wi, wh))
    fp = r * fr(cost_h)
    f = fp * F(fact / abs_cos_theta(wo)) if adj else fp * F(fact / abs_cos_theta(wi))      # Q7: sample divides by |cos theta_i|
    return f.astype(F), wi, F(pdf / F(F(4) * F(abs(cost_h))))


def rand_dir(rng, up=None):
    v = normalize(rng.normal(size=3).astype(F))
    if up is not None and (v[2] > 0) != up: v = v3(v[0], v[1], -v[2])
    return v


@pytest.mark.parametrize("seed", range(3))
def test_blinn_microfacet(seed):
    rng = np.random.default_rng(10 + seed)
    fr = lambda c: fr_dielectric(1, 1.5, c)
    for _ in range(150):
        e = F(rng.choice([0.5, 3, 20, 200, 5000]))
        wo, wi = rand_dir(rng, True), rand_dir(rng, bool(rng.integers(0, 2)))
        wh = normalize((wo + wi).astype(F))
        o3 = debug_eval(3, [e, wh, wo, wi], 3)
        assert close(blinn_d(e, wh), o3[0]) and close(blinn_pdf(e, wh), o3[1]) and mf_g(wo, wi, wh) == o3[2]
        r = rng.uniform(0.05, 1, 16).astype(F); u1, u2 = F(rng.uniform(0.01, 0.99)), F(rng.uniform())
        o4 = debug_eval(4, [e, r, wo, wi, u1, u2], 37)
        assert close(microfacet_eval(e, r, fr, wo, wi), o4[:16], 2e-5), (e, wo, wi)
        assert close(microfacet_pdf(e, wo, wi), o4[16], 2e-5)
        f, ws, pdf = microfacet_sample(e, r, fr, wo, u1, u2)
        assert close(pdf, o4[36], 5e-5) and np.allclose(ws, o4[33:36], atol=2e-6) and close(f, o4[17:33], 1e-4), (e, u1, u2)
    # Q7 as written: eval (after the caller's flip) divides by 4 cos(theta_o), the sample weight by cos(theta_i) -- not consistent
    wo = normalize(v3(0.3, 0.1, 0.9)); f, ws, pdf = microfacet_sample(F(50), spec(1), fr, wo, F(0.5), F(0.3))
    ev = microfacet_eval(F(50), spec(1), fr, ws, wo)        # non-adjoint eval is called with the arguments flipped (Reflection.hs:310)
    assert pdf > 0 and not close(f * pdf, ev, 1e-2)


# ------------------------------------------------------------------------------------------------ Montecarlo.hs / Diffuse.hs
def concentric_sample_disk(u1, u2):                # Montecarlo.hs:164-181
    sx, sy = F(F(F(u1) * F(2)) - F(1)), F(F(F(u2) * F(2)) - F(1))
    if sx == 0 and sy == 0: return F(0), F(0)
    if sx >= -sy:
        if sx > sy: r, tp = (sx, F(sy / sx)) if sy > 0 else (sx, F(F(8) + F(sy / sx)))
        else: r, tp = sy, F(F(2) - F(sx / sy))
    elif sx <= sy: r, tp = F(-sx), F(F(4) - F(sy / F(-sx)))
    else: r, tp = F(-sy), F(F(6) + F(sx / F(-sy)))
    theta = F(F(tp * PI) / F(4))
    return F(r * np.cos(theta)), F(r * np.sin(theta))


def cosine_sample_hemisphere(u1, u2):              # :147-150
    x, y = concentric_sample_disk(u1, u2)
    return v3(x, y, np.sqrt(max(F(0), F(F(F(1) - F(x * x)) - F(y * y)))))


def cos_pdf(wo, wi): return F(INV_PI * abs_cos_theta(wi)) if same_hemisphere(wo, wi) else F(0)      # Diffuse.hs:9-12
def lambert_eval(r, wo, wi): return (r * F(INV_PI * abs_cos_theta(wo))).astype(F)                    # Diffuse.hs:24-26: e wo _


def lambert_sample(r, wo, u1, u2):                 # Diffuse.hs:14-22 (adj = False)
    wi = cosine_sample_hemisphere(u1, u2)
    if wo[2] < 0: wi = v3(wi[0], wi[1], -wi[2])     # toSameHemisphere
    if same_hemisphere(wo, wi): return r, wi, cos_pdf(wo, wi)
    return spec(0), wo, F(0)


# ------------------------------------------------------------------------------------------------ Reflection.hs: the BSDF
REFL, TRANS, DIFFUSE, GLOSSY, SPECULAR = 1, 2, 4, 8, 16     # Reflection.hs:101-107


class Plastic:
    """[Lambertian kd, Microfacet (Blinn e) (frDielectric 1 1.5) ks] (Material.hs:76-86) in the frame (sn, tn, nn) with geometric normal ng"""
    def __init__(self, kd, ks, e, sn, tn, nn, ng):
        self.kd, self.ks, self.e, self.cs, self.ng = kd, ks, e, (sn, tn, nn), ng
        self.types = [REFL | DIFFUSE, REFL | GLOSSY]
    def to_local(self, v): return v3(dot(v, self.cs[0]), dot(v, self.cs[1]), dot(v, self.cs[2]))     # Math.hs:452-454
    def to_world(self, v):                                                                           # Math.hs:456-462
        s, t, n = self.cs
        return v3(F(F(s[0] * v[0]) + F(t[0] * v[1])) + F(n[0] * v[2]), F(F(s[1] * v[0]) + F(t[1] * v[1])) + F(n[1] * v[2]),
                  F(F(s[2] * v[0]) + F(t[2] * v[1])) + F(n[2] * v[2]))
    def bx_eval(self, i, a, b): return lambert_eval(self.kd, a, b) if i == 0 else microfacet_eval(self.e, self.ks, lambda c: fr_dielectric(1, 1.5, c), a, b)
    def bx_pdf(self, i, wo, wi): return cos_pdf(wo, wi) if i == 0 else microfacet_pdf(self.e, wo, wi)
    def bx_sample(self, i, wo, u1, u2):
        return lambert_sample(self.kd, wo, u1, u2) if i == 0 else microfacet_sample(self.e, self.ks, lambda c: fr_dielectric(1, 1.5, c), wo, u1, u2)

    def eval(self, wo_w, wi_w):                    # evalBsdf False (Reflection.hs:318-332)
        cos_wo = dot(wo_w, self.ng)
        side = F(dot(wi_w, self.ng) / cos_wo)
        if side == 0 or abs(cos_wo) < F(1e-5): return spec(0)
        want = TRANS if side < 0 else REFL
        wo, wi = self.to_local(wo_w), self.to_local(wi_w)
        f = spec(0)
        for i in range(2):
            if self.types[i] & want: f = (f + self.bx_eval(i, wi, wo)).astype(F)      # eval b = flip (bxdfEval b): arguments flipped
        return f

    def pdf(self, wo_w, wi_w):                     # bsdfPdf (:251-257, Q6: all components, unfiltered)
        wo, wi = self.to_local(wo_w), self.to_local(wi_w)
        return F(F(self.bx_pdf(0, wo, wi) + self.bx_pdf(1, wo, wi)) / F(2))

    def sample(self, wo_w, u_comp, u1, u2):        # sampleBsdf'' False bxdfAll (:278-316)
        empty = (REFL | DIFFUSE, F(0), spec(0), v3(0, 1, 0))
        wo = self.to_local(wo_w)
        cntf = F(2); inv_cnt = F(1) / cntf
        s_num = max(0, min(1, int(np.floor(F(F(u_comp) * cntf)))))
        f_sample, wi, pdfp = self.bx_sample(s_num, wo, u1, u2)
        wi_w = self.to_world(wi)
        side = F(dot(wi_w, self.ng) / dot(wo_w, self.ng))
        if pdfp == 0 or side == 0: return empty
        want = TRANS if side < 0 else REFL
        if not (self.types[s_num] & want): return empty
        other = 1 - s_num
        pdf = F(F(pdfp + self.bx_pdf(other, wo, wi)) * inv_cnt)
        f_others = self.bx_eval(other, wi, wo) if self.types[other] & want else spec(0)
        f_sum = ((f_sample * pdfp + f_others) * F(F(1) / pdf)).astype(F)
        return self.types[s_num], pdf, f_sum, wi_w


def _frame(rng):
    nn = rand_dir(rng)
    a = rand_dir(rng)
    sn = normalize((a - nn * dot(a, nn)).astype(F))
    tn = np.cross(nn, sn).astype(F)
    return sn, tn, nn


@pytest.mark.parametrize("seed", range(3))
def test_multi_component_bsdf_weights(seed):
    """sampleBsdf'': component choice floor(uComp * n), pdf = (pdf' + sum of the others' pdfs) / n, f = (f_s pdf' + sum of the others'
    evals, flipped) / pdf; evalBsdf's geometric-normal side test; bsdfPdf averaging over ALL components"""
    rng = np.random.default_rng(20 + seed)
    n_ok = 0
    for _ in range(120):
        sn, tn, nn = _frame(rng)
        ng = normalize((nn + rng.normal(size=3).astype(F) * F(0.2)).astype(F))
        kd, ks, e = rng.uniform(0.05, 0.9, 16).astype(F), rng.uniform(0.05, 0.9, 16).astype(F), F(rng.choice([2, 30, 400]))
        b = Plastic(kd, ks, e, sn, tn, nn, ng)
        wo_w, wi_w = rand_dir(rng), rand_dir(rng)
        u = [F(rng.uniform()), F(rng.uniform(0.01, 0.99)), F(rng.uniform())]
        o = debug_eval(5, [kd, ks, e, sn, tn, nn, ng, wo_w, wi_w] + u, 38)
        assert close(b.eval(wo_w, wi_w), o[:16], 5e-5), "evalBsdf"
        assert close(b.pdf(wo_w, wi_w), o[16], 5e-5), "bsdfPdf"
        t, pdf, f, ws = b.sample(wo_w, *u)
        assert t == int(o[17]) and close(pdf, o[18], 1e-4) and np.allclose(ws, o[35:38], atol=3e-6) and close(f, o[19:35], 2e-4), "sampleBsdf"
        n_ok += pdf > 0
    assert n_ok > 30


# ------------------------------------------------------------------------------------------------ Shape.hs: sampling lights' shapes
def lerp(t, a, b): return F(F(F(F(1) - t) * a) + F(t * b))       # Math.hs:116-118


def quad_sample(sx, sy, u1, u2): return v3(lerp(u1, -sx, sx), lerp(u2, -sy, sy), 0), v3(0, 0, -1)      # Shape.hs:404-406 (Q2: normal -z)


def quad_intersect(sx, sy, o, d, tmin, tmax):      # Shape.hs:156-171
    if abs(d[2]) < F(1e-7): return None
    t = F(F(-o[2]) / d[2])
    if t < tmin or t > tmax: return None
    p = (o + d * t).astype(F)
    if abs(p[0]) > sx or abs(p[1]) > sy: return None
    return t, v3(0, 0, 1)                           # the hit normal is +z: normalize (dpdu x dpdv) = (sx,0,0) x (0,sy,0)


def general_pdf(area, hit, p, wi):                 # Shape.hs:346-350
    if hit is None: return F(0)
    t, n = hit
    ph = (p + wi * t).astype(F)
    pd = F(sq_len((p - ph).astype(F)) / F(abs_dot(n, (-wi).astype(F)) * area))
    return F(0) if np.isinf(pd) else pd


def uniform_sample_cone(cs, cos_max, u1, u2):      # Montecarlo.hs:133-145
    x, y, z = cs
    ct = lerp(u1, cos_max, F(1)); st = np.sqrt(F(F(1) - F(ct * ct))); phi = F(F(u2) * TWO_PI)
    return (x * F(np.cos(phi) * st) + y * F(np.sin(phi) * st) + z * ct).astype(F)


def coordinate_system(v):                          # Math.hs:424-437
    if abs(v[0]) > abs(v[1]):
        il = F(1) / np.sqrt(F(F(v[0] * v[0]) + F(v[2] * v[2]))); v2 = v3(-v[2] * il, 0, v[0] * il)
    else:
        il = F(1) / np.sqrt(F(F(v[1] * v[1]) + F(v[2] * v[2]))); v2 = v3(0, v[2] * il, -v[1] * il)
    c = v3(F(v[1] * v2[2]) - F(v[2] * v2[1]), -(F(v[0] * v2[2]) - F(v[2] * v2[0])), F(v[0] * v2[1]) - F(v[1] * v2[0]))   # Math.hs:345-348
    return v2, c, v


def test_area_light_shape_sampling():
    rng = np.random.default_rng(31)
    for _ in range(200):
        sx, sy = F(rng.uniform(0.2, 3)), F(rng.uniform(0.2, 3))
        p = v3(rng.uniform(-4, 4), rng.uniform(-4, 4), rng.choice([-1, 1]) * rng.uniform(0.3, 5))
        u1, u2 = F(rng.uniform()), F(rng.uniform())
        ps, ns = quad_sample(sx, sy, u1, u2)
        wi = normalize((ps - p).astype(F))
        o = debug_eval(6, [3, sx, sy, 0, 0, 0, 0, p, u1, u2, wi], 7)
        assert np.array_equal(ps, o[:3]) and np.array_equal(ns, o[3:6])
        pdf = general_pdf(F(F(F(4) * sx) * sy), quad_intersect(sx, sy, p, wi, F(1e-3), F(np.inf)), p, wi)
        assert close(pdf, o[6], 1e-5), (pdf, o[6])
    for _ in range(200):                            # sphere seen from outside: cone sampling, uniformConePdf (Shape.hs:333-344, 364-374)
        r = F(rng.uniform(0.3, 2))
        p = (rand_dir(rng) * F(r * rng.uniform(1.2, 6))).astype(F)
        u1, u2 = F(rng.uniform()), F(rng.uniform())
        dn = normalize((-p).astype(F))
        cos_max = np.sqrt(max(F(0), F(F(1) - F(F(r * r) / sq_len(p)))))
        d = uniform_sample_cone(coordinate_system(dn), cos_max, u1, u2)
        o = debug_eval(6, [4, r, 0, 0, 0, 0, 0, p, u1, u2, d], 7)
        pdf = F(0) if cos_max >= 1 else F(F(1) / F(TWO_PI * F(F(1) - cos_max)))
        assert close(pdf, o[6], 1e-5)
        ps = o[:3].astype(F)                        # the sampled point lies on the sphere, on the cone around -p, with an outward normal
        assert abs(float(np.sqrt(sq_len(ps))) - float(r)) < 2e-4 * float(r) and np.allclose(o[3:6], ps / np.sqrt(sq_len(ps)), atol=2e-5)
        to_ps = normalize((ps - p).astype(F))
        assert float(dot(to_ps, dn)) >= float(cos_max) - 2e-4
        assert np.allclose(to_ps, d, atol=5e-4)
