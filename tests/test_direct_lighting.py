"""Direct-lighting integrator (Integrator/DirectLighting.hs:23-58) on the wavefront kernels.

Per-sample / film / converged parity of the two direct-lighting fixtures (`direct`, `blackbody-emission`) runs with every
other scene in test_host_and_emu.py (emulator) and test_gpu_parity.py (GPU); this file holds what is specific to the
integrator: the branch tree (spawned slots, head-room retry) and its structural properties.
"""
import copy
from pathlib import Path

import numpy as np
import pytest

from bling_b200 import ir as IR
from bling_b200.api import BlingCuError
from oracle.oracle_py import Oracle
from tests.conftest import load_scene, small
from tests.emu.emu_py import EmuContext


def _glass_pixels(sc, n, seed):
    """camera samples whose primary ray hits the big glass sphere of the `direct` scene: every vertex there spawns a branch"""
    o = Oracle(sc)
    x0, x1, y0, y1 = o.sample_extent()
    rng = np.random.default_rng(seed)
    px, py, s = rng.integers(x0, x1 + 1, 8 * n), rng.integers(y0, y1 + 1, 8 * n), rng.integers(0, sc.spp, 8 * n)
    deep = copy.copy(sc); deep.max_depth = 1          # depth-1 radiance is black on glass only where nothing is lit directly
    shallow, _ = Oracle(deep).render_samples(1, 9, px, py, s)
    full, _ = o.render_samples(1, 9, px, py, s)
    branchy = np.abs(full - shallow).max(1) > 1e-3     # radiance that arrived through specular continuations
    idx = np.nonzero(branchy)[0][:n]
    return px[idx], py[idx], s[idx]


def _check_branch_tree(make_ctx):
    sc = small(load_scene("direct"), 90, 60, 4, 4)
    px, py, s = _glass_pixels(sc, 400, 3)
    assert len(px) >= 200
    o = Oracle(sc)
    Lo, xyo = o.render_samples(1, 9, px, py, s)
    c = make_ctx(); c.upload_scene(sc)
    # all samples branch: 2 slots per camera sample (the initial head-room) cannot hold a depth-5 tree of glass vertices,
    # so this call goes through the overflow -> double -> re-run path at least once
    Le, xye = c.render_samples(1, 9, px, py, s)
    st = c.stats()
    c.close()
    assert np.array_equal(xyo, xye)
    rel = np.abs(Lo - Le).max(1) / (np.abs(Lo).max(1) + 1e-6)
    assert (rel < 2e-4).mean() > 0.995, rel.max()
    assert st["rays_extension"] > 2 * len(px)          # more continuation rays than camera samples: the tree did branch
    return st


def test_emulated_branch_tree_and_headroom_retry():
    _check_branch_tree(EmuContext)


def test_direct_lighting_of_diffuse_scene_ignores_max_depth():
    """no specular component -> `cont` never continues (DirectLighting.hs:47-58): maxDepth must not matter"""
    base = small(load_scene("direct"), 60, 40, 2, 2)
    base.materials = [IR.Material.from_buffer_copy(m) for m in base.materials]
    for m in base.materials:
        if m.kind in (IR.MAT_GLASS, IR.MAT_MIRROR, IR.MAT_SHINYMETAL): m.kind = IR.MAT_MATTE
    out = []
    for md in (1, 5):
        sc = copy.copy(base); sc.max_depth = md
        o = Oracle(sc)
        x0, x1, y0, y1 = o.sample_extent()
        rng = np.random.default_rng(4)
        px, py, s = rng.integers(x0, x1 + 1, 500), rng.integers(y0, y1 + 1, 500), rng.integers(0, 4, 500)
        out.append(o.render_samples(1, 2, px, py, s)[0])
        e = EmuContext(); e.upload_scene(sc)
        Le = e.render_samples(1, 2, px, py, s)[0]; e.close()
        assert np.allclose(Le, out[-1], rtol=1e-5, atol=1e-7)
    assert np.array_equal(out[0], out[1])


def test_direct_lighting_misses_are_black_and_emitters_use_wo():
    """directLighting returns black on a miss (no environment radiance, :27) and adds `intLe int wo` (:45): a directly viewed
    emitter is lit where its normal faces the camera -- the opposite of the path integrator's Q1."""
    sc = small(load_scene("direct"), 90, 60, 4, 4)
    path = copy.copy(sc); path.integrator_kind = IR.INTEGRATOR_PATH; path.sample_depth = 3
    od, op = Oracle(sc), Oracle(path)
    x0, x1, y0, y1 = od.sample_extent()
    xs, ys = np.meshgrid(np.arange(x0, x1 + 1), np.arange(y0, y1 + 1))
    px, py = xs.ravel(), ys.ravel(); s = np.zeros_like(px)
    Ld, _ = od.render_samples(1, 1, px, py, s); Lp, _ = op.render_samples(1, 1, px, py, s)
    sky = (Ld.max(1) == 0) & (Lp.max(1) > 0)          # primary rays that leave the scene: env light in the path integrator only
    assert sky.sum() > 50
    bright = Ld.max(1) > 10                             # the emitter sphere (radiance 18..25), seen from outside
    assert bright.sum() >= 3 and (Lp[bright].max(1) < Ld[bright].max(1)).all()


def test_upload_rejects_unknown_integrator():
    sc = small(load_scene("direct"), 30, 20, 2, 2)
    sc.integrator_kind = 7
    e = EmuContext()
    with pytest.raises(BlingCuError) as ei: e.upload_scene(sc)
    assert ei.value.code == 1
    e.close()


@pytest.mark.skipif(not Path("/root/reference/examples/blackbody-emission.bling").exists(), reason="needs the reference checkout")
def test_loader_reads_the_reference_direct_lighting_example():
    from bling_b200.host.loader import load_scene as parse
    ir = parse("/root/reference/examples/blackbody-emission.bling")
    assert ir.integrator_kind == IR.INTEGRATOR_DIRECT and ir.max_depth == 5 and (ir.nu, ir.nv) == (3, 3)
    fx = load_scene("blackbody-emission")
    assert fx.integrator_kind == IR.INTEGRATOR_DIRECT and len(fx.lights) == len(ir.lights) == 15


@pytest.mark.gpu
def test_gpu_branch_tree_and_headroom_retry():
    from bling_b200.api import Context
    _check_branch_tree(lambda: Context(0))


# ------------------------------------------------------------------------------------------------ `debug normals`
def _normals_scene():
    sc = small(load_scene("textures"), 72, 44, 2, 2)      # bump-mapped objects: the SHADING normal is what is shown
    sc.integrator_kind = IR.INTEGRATOR_NORMALS
    return sc


def _check_normal_map(make_ctx, exact):
    sc = _normals_scene()
    o = Oracle(sc)
    x0, x1, y0, y1 = o.sample_extent()
    xs, ys = np.meshgrid(np.arange(x0, x1 + 1), np.arange(y0, y1 + 1))
    px, py = xs.ravel(), ys.ravel(); s = np.ones_like(px)
    Lo, xyo = o.render_samples(1, 1, px, py, s)
    c = make_ctx(); c.upload_scene(sc)
    Lc, xyc = c.render_samples(1, 1, px, py, s)
    c.render_pass(1, 5); film = c.read_film(); c.close()
    assert np.array_equal(xyo, xyc)
    if exact: assert np.array_equal(Lo, Lc)
    else: assert (np.abs(Lo - Lc).max(1) < 1e-4).mean() > 0.999
    assert (Lo.max(1) == 0).sum() > 20 and (Lo.max(1) > 0).sum() > 1000      # black where the ray leaves the scene
    o.render_pass(1, 5, threads=4)
    assert np.abs(film - o.read_film()).mean() <= 1e-5 * np.abs(film).mean()
    # rgbToSpectrumRefl((1 + n) / 2): a flat, un-bumped ground would be one colour; the fbm bump makes it vary
    assert np.unique(np.round(Lo[Lo.max(1) > 0], 3), axis=0).shape[0] > 100


def test_emulated_normal_map_integrator_matches_oracle():
    """mkNormalMap (Integrator/Debug.hs:23-33): rgbToSpectrumRefl ((1 + bsdfShadingNormal) / 2) of the first hit, black on a miss"""
    _check_normal_map(EmuContext, exact=True)


@pytest.mark.gpu
def test_gpu_normal_map_integrator_matches_oracle():
    from bling_b200.api import Context
    _check_normal_map(lambda: Context(0), exact=False)
