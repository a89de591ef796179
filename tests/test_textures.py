"""Textures that compute (second slice): scalar textures, blend / gradient, bump mapping.

Three statements of Texture.hs:129-414 are held against each other:
  * `py_*` below: a small pure-Python one (numpy float32 scalars, Python big integers for Haskell's 64-bit Int),
  * the oracle (oracle/oracle_texture.h),
  * the kernel bodies (bling_b200/csrc/textures.h) -- through the CPU emulator here, through the C ABI on the GPU.
"""
import math

import numpy as np
import pytest

from bling_b200 import ir as IR
from bling_b200.api import BlingCuError
from oracle.oracle_py import Oracle
from tests.conftest import has_gpu, load_scene, small
from tests.emu.emu_py import EmuContext

F = np.float32


# ------------------------------------------------------------------------------------------------ pure-Python statement
# perlin / fbm: the HOST's float32 statement (bling_b200/host/noise.py, what the stand-in loader builds height maps with);
# cellNoise: below, with Python big integers standing for Haskell's 64-bit Int
from bling_b200.host.noise import fbm as py_fbm, perlin3d as py_perlin  # noqa: E402


def _wrap64(v):   # Haskell Int arithmetic wraps at 64 bits (two's complement)
    v &= (1 << 64) - 1
    return v - (1 << 64) if v >> 63 else v


def py_cell_noise(kind, p):
    p = [F(c) for c in p]
    lcg = lambda x: (1103515245 * x + 12345) % 4294967296
    def hsh(x, y, z):
        v = _wrap64(x * 73856093) ^ _wrap64(y * 19349663) ^ _wrap64(z * 83492791)
        return abs(v) % 4294967296
    def prob(v):
        for i, t in enumerate((393325350, 1022645910, 1861739990, 2700834071, 3372109335, 3819626178, 4075350088, 4203212043)):
            if v < t: return i + 1
        return 9
    def dist(q):
        d = [F(a - b) for a, b in zip(p, q)]
        if kind == 0: return F(np.sqrt(F(F(F(d[0] * d[0]) + F(d[1] * d[1])) + F(d[2] * d[2]))))
        if kind == 1: return F(F(F(d[0] * d[0]) + F(d[1] * d[1])) + F(d[2] * d[2]))
        if kind == 2: return F(F(abs(d[0]) + abs(d[1])) + abs(d[2]))
        return max(abs(d[0]), abs(d[1]), abs(d[2]))
    o = [math.floor(c) for c in p]
    best = F(np.inf)
    for x in (-1, 0, 1):
        for y in (-1, 0, 1):
            for z in (-1, 0, 1):
                c = (x + o[0], y + o[1], z + o[2])
                u = lcg(hsh(*c))
                for _ in range(prob(u)):
                    u1 = lcg(u); u2 = lcg(u1); u3 = lcg(u2)
                    q = [F(F(c[k]) + F(F(uu) / F(4294967296.0))) for k, uu in enumerate((u1, u2, u3))]
                    best = min(best, dist(q)); u = u3
    return best


# ------------------------------------------------------------------------------------------------ fixtures
def _texture_scene():
    return small(load_scene("textures"), 48, 30, 2, 2)


def _find(sc, kind, pred=lambda t: True):
    return [i for i, t in enumerate(sc.textures) if t.kind == kind and pred(t)]


def _points(n, seed, span=6.0):
    rng = np.random.default_rng(seed)
    p = ((rng.random((n, 3)) - 0.5) * 2 * span).astype(np.float32)
    p[: n // 8] = np.round(p[: n // 8])          # lattice points: cell borders of perlin / cellNoise
    uv = rng.random((n, 2)).astype(np.float32)
    return p, uv


def _mapped(t, p):   # identityMapping3d: transPoint w2t (all mappings of the scene are affine: w == 1)
    m = np.array(list(t.s.v), np.float32).reshape(4, 4)
    out = []
    for q in p:
        out.append([F(F(F(F(m[r, 0] * q[0]) + F(m[r, 1] * q[1])) + F(m[r, 2] * q[2])) + m[r, 3]) for r in range(3)])
    return out


# ------------------------------------------------------------------------------------------------ oracle vs pure Python
def test_oracle_perlin_fbm_match_python_statement():
    sc = _texture_scene(); o = Oracle(sc)
    p, uv = _points(200, 1)
    for tid in _find(sc, IR.STEX_PERLIN)[:2]:
        got = o.eval_texture(tid, p, uv)[:, 0]
        ref = np.array([py_perlin(*q) for q in _mapped(sc.textures[tid], p)], np.float32)
        assert np.array_equal(got, ref)
    for tid in _find(sc, IR.STEX_FBM)[:2]:
        t = sc.textures[tid]
        got = o.eval_texture(tid, p, uv)[:, 0]
        ref = np.array([py_fbm(t.aux, t.f[0], q) for q in _mapped(t, p)], np.float32)
        assert np.array_equal(got, ref)


def test_oracle_cell_noise_matches_python_statement():
    sc = _texture_scene(); o = Oracle(sc)
    p, uv = _points(120, 2)
    p[:10] *= 1000          # far cells: hash products overflow 32 bits, negative coordinates exercise `abs`
    kinds = set()
    for tid in _find(sc, IR.STEX_CELLNOISE):
        t = sc.textures[tid]
        if t.aux in kinds: continue
        kinds.add(t.aux)
        got = o.eval_texture(tid, p, uv)[:, 0]
        ref = np.array([py_cell_noise(t.aux, q) for q in _mapped(t, p)], np.float32)
        assert np.array_equal(got, ref), t.aux
    assert kinds == {0, 1, 2, 3}


def test_perlin_is_zero_on_the_lattice_and_bounded():
    sc = _texture_scene(); o = Oracle(sc)
    tid = _find(sc, IR.STEX_PERLIN)[0]
    m = np.array(list(sc.textures[tid].s.v), np.float32).reshape(4, 4)
    lattice = np.array([[i, j, k] for i in (-3, 0, 2) for j in (-1, 4) for k in (0, 7)], np.float32)
    p = (lattice - m[:3, 3]) / np.diag(m)[:3]                      # texture-space integers (scale + translate mappings)
    assert np.abs(o.eval_texture(tid, p, np.zeros((len(p), 2)))[:, 0]).max() < 1e-5
    q, uv = _points(2000, 3)
    assert np.abs(o.eval_texture(tid, q, uv)[:, 0]).max() <= 1.5


def test_gradient_and_blend_limits():
    sc = _texture_scene(); o = Oracle(sc)
    p, uv = _points(500, 4)
    for tid in _find(sc, IR.TEX_GRADIENT):
        t = sc.textures[tid]
        f = o.eval_texture(t.aux, p, uv)[:, 0]
        steps = [sc.textures[t.child[0] + k] for k in range(t.child[1])]
        pos = [s.f[0] for s in steps]
        assert pos == sorted(pos)
        got = o.eval_texture(tid, p, uv)
        lo, hi = f <= pos[0], f >= pos[-1]
        assert np.array_equal(got[lo], np.tile(np.array(list(steps[0].s.v), np.float32), (lo.sum(), 1)))
        assert np.array_equal(got[hi], np.tile(np.array(list(steps[-1].s.v), np.float32), (hi.sum(), 1)))
        cols = np.array([list(s.s.v) for s in steps], np.float32)
        assert (got <= cols.max(0) + 1e-6).all() and (got >= cols.min(0) - 1e-6).all()
    for tid in _find(sc, IR.TEX_BLEND):
        t = sc.textures[tid]
        x = o.eval_texture(t.aux, p, uv)[:, 0]
        a, b, got = o.eval_texture(t.child[0], p, uv), o.eval_texture(t.child[1], p, uv), o.eval_texture(tid, p, uv)
        assert np.array_equal(got[x <= 0], a[x <= 0]) and np.array_equal(got[x >= 1], b[x >= 1])
        mid = (x > 0) & (x < 1)
        ref = a[mid] * (1 - x[mid])[:, None] + b[mid] * x[mid][:, None]
        assert np.array_equal(got[mid], ref.astype(np.float32))


# ------------------------------------------------------------------------------------------------ kernel bodies vs oracle
def _all_textures_match(ctx, sc, exact):
    o = Oracle(sc)
    p, uv = _points(3000, 5)
    for tid, t in enumerate(sc.textures):
        if t.kind == IR.TEX_CONSTANT and t.f[0] != 0: continue     # gradient steps are not textures of their own
        a, b = o.eval_texture(tid, p, uv), ctx.eval_texture(tid, p, uv)
        if exact or t.kind not in (IR.STEX_CRYSTAL, IR.TEX_BLEND):
            assert np.array_equal(a, b), (tid, t.kind)
        else:
            # quasiCrystal runs cos / sin: CUDA's libm may differ from glibc in the last place, and `wrap` folds the sum
            # at integers, so compare off the fold
            bad = np.abs(a - b).max(1) > 2e-4
            assert bad.mean() < 2e-3, (tid, t.kind, bad.mean())


def test_emulated_textures_match_oracle_exactly():
    sc = _texture_scene()
    e = EmuContext(); e.upload_scene(sc)
    _all_textures_match(e, sc, exact=True)
    e.close()


def test_upload_rejects_malformed_texture_trees():
    import copy
    base = _texture_scene()

    def upload(mutate):
        sc = copy.copy(base)
        sc.textures = [IR.Texture.from_buffer_copy(t) for t in base.textures]
        sc.materials = [IR.Material.from_buffer_copy(m) for m in base.materials]
        mutate(sc)
        e = EmuContext()
        try:
            with pytest.raises(BlingCuError) as ei: e.upload_scene(sc)
            assert ei.value.code == 1
        finally:
            e.close()

    blend = _find(base, IR.TEX_BLEND)[0]; grad = _find(base, IR.TEX_GRADIENT)[0]; scale = _find(base, IR.STEX_SCALE)[0]
    upload(lambda sc: setattr(sc.textures[blend], "aux", 0))                       # blend factor is a spectrum texture
    upload(lambda sc: setattr(sc.textures[grad], "aux", len(sc.textures)))         # out of range
    upload(lambda sc: sc.textures[grad].child.__setitem__(1, 0))                   # no steps
    upload(lambda sc: setattr(sc.textures[scale], "kind", 99))                     # unknown kind
    upload(lambda sc: sc.textures[scale].child.__setitem__(0, scale))              # scale chain that never ends
    upload(lambda sc: setattr(sc.materials[1], "bump", blend + 1))                 # bump must be a scalar texture
    upload(lambda sc: sc.materials[1].ftex.__setitem__(0, len(sc.textures) + 1))   # out of range

    def deep(sc):   # blend nested three deep
        ids = [blend]
        for _ in range(2):
            t = IR.Texture.from_buffer_copy(sc.textures[blend]); t.child[0] = ids[-1]
            sc.textures.append(t); ids.append(len(sc.textures) - 1)
        sc.materials[0].tex[0] = ids[-1]
    upload(deep)


def test_bump_mapping_tilts_the_shading_normal_only():
    """bump d dgg dgs (Reflection.hs:347-377): a constant displacement changes nothing; fbm changes the radiance of a
    diffuse ground but never the sample positions or the traversal."""
    import copy
    base = _texture_scene()
    o = Oracle(base)
    x0, x1, y0, y1 = o.sample_extent()
    rng = np.random.default_rng(9)
    n = 600
    px, py, s = rng.integers(x0, x1 + 1, n), rng.integers(y0, y1 + 1, n), rng.integers(0, base.spp, n)
    L0, xy0 = o.render_samples(1, 5, px, py, s)
    flat = copy.copy(base)
    flat.textures = [IR.Texture.from_buffer_copy(t) for t in base.textures]
    flat.materials = [IR.Material.from_buffer_copy(m) for m in base.materials]
    c = IR.Texture(); c.kind = IR.STEX_CONSTANT; c.f[0] = 0.37
    flat.textures.append(c)
    for m in flat.materials:
        if m.bump: m.bump = len(flat.textures)
    of = Oracle(flat); L1, xy1 = of.render_samples(1, 5, px, py, s)
    nobump = copy.copy(flat); nobump.materials = [IR.Material.from_buffer_copy(m) for m in flat.materials]
    for m in nobump.materials: m.bump = 0
    on = Oracle(nobump); L2, xy2 = on.render_samples(1, 5, px, py, s)
    assert np.array_equal(xy0, xy1) and np.array_equal(xy1, xy2)
    assert np.allclose(L1, L2, rtol=2e-3, atol=1e-5)           # constant displacement == no bump (up to the 0.01 differences)
    assert np.abs(L0 - L2).max() > 1e-2                          # the noise displacement does change the shading


# ------------------------------------------------------------------------------------------------ GPU, through the C ABI
@pytest.mark.gpu
def test_gpu_textures_match_oracle():
    from bling_b200.api import Context
    sc = _texture_scene()
    ctx = Context(0); ctx.upload_scene(sc)
    _all_textures_match(ctx, sc, exact=False)
    ctx.close()


def test_more_textured_materials_than_slots_share_the_general_kernel():
    """Every material whose textures compute gets its own shade queue (22 of them); beyond that, queues are shared and a queue
    that mixes material kinds is shaded by the any-kind kernel. 24 extra textured materials in front of the real ones force
    both: results must not change."""
    import copy
    base = _texture_scene()
    sc = copy.copy(base)
    sc.materials = [IR.Material.from_buffer_copy(m) for m in base.materials]
    sc.shapes = [IR.Shape.from_buffer_copy(s) for s in base.shapes]
    stex = _find(base, IR.STEX_FBM)[0]
    extra = []
    for k in range(24):
        m = IR.Material(); m.kind = (IR.MAT_PLASTIC, IR.MAT_MIRROR, IR.MAT_METAL)[k % 3]
        m.tex[0] = 0; m.tex[1] = 0; m.tex[2] = -1; m.tex3 = -1; m.f[0] = 0.1; m.bump = stex + 1
        extra.append(m)
    sc.materials = extra + sc.materials
    for s in sc.shapes: s.material += 24
    sc.tri_material = base.tri_material + 24
    o = Oracle(base)
    x0, x1, y0, y1 = o.sample_extent()
    rng = np.random.default_rng(11)
    n = 1500
    px, py, s = rng.integers(x0, x1 + 1, n), rng.integers(y0, y1 + 1, n), rng.integers(0, base.spp, n)
    Lo, _ = o.render_samples(1, 5, px, py, s)
    for scene in (base, sc):
        e = EmuContext(); e.upload_scene(scene)
        Le, _ = e.render_samples(1, 5, px, py, s); e.close()
        rel = np.abs(Lo - Le).max(1) / (np.abs(Lo).max(1) + 1e-6)
        assert (rel < 1e-4).mean() > 0.999, rel.max()
