"""GPU leg of the randomised differential test (tests/test_fuzz_scenes.py holds the generator and the CPU-emulator leg): random
scenes of the whole supported grammar through the C ABI on cuda:0 against the oracle, per sample. The seeds are the ones whose
scenes produce NON-FINITE weights (second class: the reference turns them into NaN / inf samples that addSample drops,
Image.hs:253-256) plus two ordinary ones. CUDA's libm (sinf / cosf / powf) differs from glibc's in the last bits, which matters
exactly in these ill-conditioned scenes, so the bounds are statistical here and exact on the emulator."""
import numpy as np
import pytest

from bling_b200.host.loader import load_scene as parse
from oracle.oracle_py import Oracle
from tests.conftest import compare_hits, random_rays
from tests.test_fuzz_scenes import random_scene_text


@pytest.mark.gpu
@pytest.mark.parametrize("seed", [3, 7, 12, 13, 20, 38])
def test_gpu_random_scene_matches_oracle(seed, tmp_path):
    from bling_b200.api import Context
    f = tmp_path / f"fuzz{seed}.bling"; f.write_text(random_scene_text(seed))
    try:
        sc = parse(f)
    except NotImplementedError as ex:
        pytest.skip(str(ex))
    o = Oracle(sc); c = Context(0); c.upload_scene(sc)
    rays = random_rays(sc, 4000, seed)
    hg, hb = c.trace_nearest(rays), o.trace_nearest(rays, mode="brute")
    ties, bad = compare_hits(hg, hb)
    assert bad == 0 and ties <= 8, (ties, bad)
    same = (hg["prim"] == hb["prim"]) & (hb["prim"] >= 0)
    assert np.array_equal(hg["t"][same], hb["t"][same])                      # primitive hits are bit-identical (-fmad=false)
    x0, x1, y0, y1 = o.sample_extent()
    rng = np.random.default_rng(2000 + seed)
    n = 20000
    px, py, s = rng.integers(x0, x1 + 1, n), rng.integers(y0, y1 + 1, n), rng.integers(0, sc.spp, n)
    Lo, xyo = o.render_samples(1, 17 + seed, px, py, s)
    Lg, xyg = c.render_samples(1, 17 + seed, px, py, s)
    c.close()
    assert np.array_equal(xyo, xyg)
    fo, fg = np.isfinite(Lo).all(1), np.isfinite(Lg).all(1)
    assert (fo == fg).mean() > 0.995, (seed, int((~fo).sum()), int((~fg).sum()))
    assert abs(int((~fo).sum()) - int((~fg).sum())) <= 0.15 * (~fo).sum() + 5, (seed, int((~fo).sum()), int((~fg).sum()))
    ok = fo & fg
    rel = np.abs(Lo[ok] - Lg[ok]).max(1) / (np.abs(Lo[ok]).max(1) + 1e-6)
    assert (rel < 1e-2).mean() > 0.95, (seed, float((rel < 1e-2).mean()))      # named scenes: > 0.995 at 1e-3 (test_gpu_parity.py)


@pytest.mark.gpu
def test_gpu_flat_triangles_beside_smooth_ones():
    """GPU leg of tests/test_edge_cases.py::test_flat_triangles_beside_smooth_ones (nine-zero normals = flat triangle)."""
    import copy
    from bling_b200 import ir as IR
    from bling_b200.api import Context
    from tests.conftest import load_scene, small
    base = small(load_scene("smooth"), 48, 36, 2, 2); base.integrator_kind = IR.INTEGRATOR_NORMALS
    base.refl_basis = load_scene("textures").refl_basis
    nt = len(base.tri_verts)
    flat = copy.copy(base); flat.tri_normals = None
    zeros = copy.copy(base); zeros.tri_normals = np.zeros((nt, 9), np.float32)
    half = copy.copy(base); half.tri_normals = np.array(base.tri_normals, np.float32).reshape(nt, 9).copy(); half.tri_normals[::2] = 0
    x0, x1, y0, y1 = Oracle(base).sample_extent()
    xs, ys = np.meshgrid(np.arange(x0, x1 + 1), np.arange(y0, y1 + 1)); px, py = xs.ravel(), ys.ravel(); s = np.zeros_like(px)
    L = {}
    for k, sc in dict(smooth=base, flat=flat, zeros=zeros, half=half).items():
        c = Context(0); c.upload_scene(sc); L[k], _ = c.render_samples(1, 2, px, py, s); c.close()
        Lo, _ = Oracle(sc).render_samples(1, 2, px, py, s)
        assert (np.abs(L[k] - Lo).max(1) < 1e-4).mean() > 0.999, k
    assert np.array_equal(L["zeros"], L["flat"]) and not np.array_equal(L["smooth"], L["flat"])
    as_smooth, as_flat = (L["half"] == L["smooth"]).all(1), (L["half"] == L["flat"]).all(1)
    assert (as_smooth | as_flat).all() and (as_smooth & ~as_flat).sum() > 20 and (as_flat & ~as_smooth).sum() > 20


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["cornell-box", "pool"])
def test_gpu_russian_roulette_beyond_depth_seven(name):
    """GPU leg of tests/test_edge_cases.py::test_russian_roulette_beyond_depth_seven (Path.hs:68-72 at maxDepth 14)."""
    from bling_b200.api import Context
    from tests.conftest import load_scene, small
    sc = small(load_scene(name), 40, 30, 2, 2); sc.max_depth = 14
    o = Oracle(sc); c = Context(0); c.upload_scene(sc)
    x0, x1, y0, y1 = o.sample_extent()
    rng = np.random.default_rng(11)
    n = 6000
    px, py, s = rng.integers(x0, x1 + 1, n), rng.integers(y0, y1 + 1, n), rng.integers(0, 4, n)
    Lo, _ = o.render_samples(1, 9, px, py, s); Lg, _ = c.render_samples(1, 9, px, py, s)
    c.close()
    rel = np.abs(Lo - Lg).max(1) / (np.abs(Lo).max(1) + 1e-6)
    # libm differences move a few paths across discontinuities, more of them along 14 vertices than along 5 (0.995 there)
    assert (rel < 1e-3).mean() > 0.98, (name, float((rel < 1e-3).mean()))
    cap = np.percentile(Lo, 99.5)
    assert abs(np.minimum(Lg, cap).mean() / np.minimum(Lo, cap).mean() - 1) < 1e-2
