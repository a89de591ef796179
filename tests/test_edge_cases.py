"""Edge cases of the hot path, kernel bodies (CPU emulator) against the oracle: degenerate and duplicated geometry, extreme
ray ranges, far-away geometry, tiny films under wide filters, non-square strata (Q9), depth-0 integrators."""
import copy

import numpy as np
import pytest

from bling_b200 import ir as IR
from oracle.oracle_py import Oracle
from tests.conftest import compare_hits, load_scene, random_rays, small
from tests.emu.emu_py import EmuContext


def _tri_scene(verts, like="cornell-box"):
    """the camera / film / lights of a fixture with its triangles replaced by `verts` (n, 9); shapes dropped"""
    base = small(load_scene(like), 24, 18, 2, 2)
    sc = copy.copy(base)
    sc.tri_verts = np.asarray(verts, np.float32).reshape(-1, 9)
    n = len(sc.tri_verts)
    sc.tri_uvs = np.tile(np.array([0, 0, 1, 0, 1, 1], np.float32), (n, 1))
    sc.tri_material = np.zeros(n, np.int32); sc.tri_normals = None; sc.tri_prim_id = None; sc.tri_prim_id_base = 0
    sc.shapes = []
    sc.lights = [l for l in base.lights if l.kind != IR.LIGHT_AREA]
    return sc


def _traversal_parity(sc, rays):
    o = Oracle(sc); e = EmuContext(); e.upload_scene(sc)
    ho, hb = o.trace_nearest(rays, mode="kd"), o.trace_nearest(rays, mode="brute")   # the reference's kd-tree, and brute force
    assert compare_hits(ho, hb)[1] == 0
    he = e.trace_nearest(rays)
    oo, oe = o.trace_occluded(rays, mode="brute"), e.trace_occluded(rays)
    e.close()
    ties, bad = compare_hits(he, hb)
    assert bad == 0, (ties, bad)
    hit = hb["prim"] >= 0
    assert np.array_equal(he["t"][hit & (he["prim"] == hb["prim"])], hb["t"][hit & (he["prim"] == hb["prim"])])   # bit-exact t
    assert np.array_equal(oe.astype(bool), oo.astype(bool))
    return ho, he, ties


def test_degenerate_and_duplicate_triangles():
    rng = np.random.default_rng(3)
    good = (rng.random((40, 3, 3)) * 4 - 2).astype(np.float32)
    point = np.repeat(rng.random((5, 1, 3)).astype(np.float32), 3, axis=1)               # all three vertices equal
    a = rng.random((5, 3)).astype(np.float32); d = rng.random((5, 3)).astype(np.float32)
    line = np.stack([a, a + d, a + 2 * d], 1)                                             # collinear
    dup = np.concatenate([good[:6], good[:6]])                                            # exact duplicates: t-ties
    sc = _tri_scene(np.concatenate([good, point, line, dup]).reshape(-1, 9))
    rays = random_rays(sc, 4000, 1)
    ho, he, ties = _traversal_parity(sc, rays)
    n_good = 40
    hit = he["prim"] >= 0
    assert hit.sum() > 200
    assert not np.isin(he["prim"][hit], np.arange(n_good, n_good + 10)).any()             # zero-area triangles are never hit


def test_ray_ranges_that_exclude_everything():
    sc = small(load_scene("zoo"), 24, 18, 2, 2)
    rays = random_rays(sc, 3000, 2)
    e = EmuContext(); e.upload_scene(sc); o = Oracle(sc)
    full = e.trace_nearest(rays)
    for name, tmin, tmax in (("inverted", 5.0, 1.0), ("empty", 0.0, 0.0), ("behind the far hit", 1e7, np.inf)):
        r = rays.copy(); r["tmin"] = tmin; r["tmax"] = tmax
        he, hb = e.trace_nearest(r), o.trace_nearest(r, mode="brute")
        assert (he["prim"] < 0).all() and (hb["prim"] < 0).all(), name
        assert not e.trace_occluded(r).any() and not o.trace_occluded(r).any(), name
    # a range that ends exactly at the hit keeps it (t > tmax rejects, t == tmax does not: TriangleMesh.hs:181, Shape.hs)
    hit = full["prim"] >= 0
    r = rays[hit].copy(); r["tmin"] = 0; r["tmax"] = full["t"][hit]
    he, hb = e.trace_nearest(r), o.trace_nearest(r, mode="brute")
    ties, bad = compare_hits(he, hb)
    assert bad == 0 and (he["prim"] >= 0).mean() > 0.95
    e.close()


def test_far_away_and_tiny_geometry_keep_bit_exact_hits():
    rng = np.random.default_rng(5)
    base = (rng.random((60, 3, 3)) * 2 - 1).astype(np.float32)
    for offset, scale in ((1.0e6, 1.0), (0.0, 1.0e-4), (-3.0e5, 50.0)):
        sc = _tri_scene((base * scale + offset).reshape(-1, 9))
        rays = np.zeros(2000, IR.RAY_DTYPE)
        c = rng.random((2000, 3)).astype(np.float32) * 2 - 1
        d = rng.normal(size=(2000, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
        rays["o"] = (c * scale * 3 + offset).astype(np.float32); rays["d"] = d.astype(np.float32)
        rays["tmin"] = 0; rays["tmax"] = np.inf
        _traversal_parity(sc, rays)


@pytest.mark.parametrize("w,h,filt", [(1, 1, "wide"), (3, 2, "box"), (2, 5, "wide")])
def test_tiny_films(w, h, filt):
    """films smaller than the filter footprint: the sample extent is mostly apron (Image.hs:162-168), every sample reaches
    every pixel, tile images clip on all sides (Q10)"""
    sc = small(load_scene("cornell-box" if filt == "wide" else "ducky"), w, h, 3, 3)
    if filt == "box": sc.filter_w = sc.filter_h = 0.5; sc.filter_table = np.ones(256, np.float32)
    o = Oracle(sc); e = EmuContext(); e.upload_scene(sc)
    assert o.sample_extent() == e.sample_extent()
    o.render_pass(1, 3, threads=2); e.render_pass(1, 3)
    fo, fe = o.read_film(), e.read_film(); e.close()
    assert fo.shape == (h, w, 4)
    assert np.allclose(fo[..., 0], fe[..., 0], rtol=1e-5, atol=1e-6)                       # filter weights: same sample positions
    assert np.allclose(fo, fe, rtol=2e-3, atol=1e-5)


def test_non_square_strata_follow_q9():
    """Sampling.hs:168-171 derives BOTH jitter cells from `quotRem i nu` with du = 1 / nu, dv = 1 / nv: with nu != nv the image
    samples do not tile the pixel. Reproduced, not fixed: positions must equal the oracle's exactly."""
    sc = small(load_scene("glass-torus"), 12, 9, 3, 5)
    o = Oracle(sc); e = EmuContext(); e.upload_scene(sc)
    x0, x1, y0, y1 = o.sample_extent()
    px = np.repeat(np.arange(x0, x1 + 1), 15); py = np.full_like(px, 2); s = np.tile(np.arange(15), x1 - x0 + 1)
    Lo, xyo = o.render_samples(1, 4, px, py, s); Le, xye = e.render_samples(1, 4, px, py, s); e.close()
    assert np.array_equal(xyo, xye)
    fx = xyo[:, 0] - px                                                                    # offsets inside the pixel
    assert fx.min() >= 0 and fx.max() <= 1.0 + 1e-6                                        # x: cells of width 1/3 cover [0, 1]
    fy = xyo[:, 1] - py
    assert fy.min() >= 0 and fy.max() <= 1.0 + 1e-6
    rel = np.abs(Lo - Le).max(1) / (np.abs(Lo).max(1) + 1e-6)
    assert (rel < 1e-4).mean() > 0.99


@pytest.mark.parametrize("name", ["sun-sky", "direct"])
def test_depth_zero_and_one(name):
    """maxDepth 0: the path integrator returns at the first hit (`depth == md`, Path.hs:51) and only adds the environment on a
    miss. Direct lighting has no maxDepth 0: `cont` compares d + 1 with md (DirectLighting.hs:47-49), so 0 would never stop;
    upload rejects it. maxDepth 1 = one shaded vertex, no continuation."""
    from bling_b200.api import BlingCuError
    for md in (0, 1):
        sc = small(load_scene(name), 30, 20, 2, 2); sc.max_depth = md; sc.sample_depth = min(sc.sample_depth, 1)
        if sc.integrator_kind == IR.INTEGRATOR_DIRECT and md == 0:
            e = EmuContext()
            with pytest.raises(BlingCuError): e.upload_scene(sc)
            e.close(); continue
        o = Oracle(sc); e = EmuContext(); e.upload_scene(sc)
        x0, x1, y0, y1 = o.sample_extent()
        rng = np.random.default_rng(8)
        px, py, s = rng.integers(x0, x1 + 1, 800), rng.integers(y0, y1 + 1, 800), rng.integers(0, 4, 800)
        Lo, _ = o.render_samples(1, 6, px, py, s); Le, _ = e.render_samples(1, 6, px, py, s)
        st = e.stats(); e.close()
        rel = np.abs(Lo - Le).max(1) / (np.abs(Lo).max(1) + 1e-6)
        assert (rel < 1e-4).mean() > 0.995, (name, md, rel.max())
        if sc.integrator_kind == IR.INTEGRATOR_PATH and md == 0:
            assert st["rays_shadow"] == 0 and st["rays_extension"] == 0                    # nothing is shaded at depth == md


@pytest.mark.parametrize("name", ["cornell-box", "cornell-box-specular", "pool"])
def test_russian_roulette_beyond_depth_seven(name):
    """Path.hs:68-72: from depth 8 on a path continues with probability min(0.75, Y(t)) and its throughput is divided by it. No
    config reaches that branch with its own maxDepth (5); `examples/pool.bling` has maxDepth 10. Here maxDepth is 14, so thousands
    of paths in the closed box pass the roulette several times; sample dimensions beyond sampleDepth wrap as in the reference."""
    sc = small(load_scene(name), 40, 30, 2, 2); sc.max_depth = 14
    sc8 = copy.copy(sc); sc8.max_depth = 8
    o = Oracle(sc); e = EmuContext(); e.upload_scene(sc)
    x0, x1, y0, y1 = o.sample_extent()
    rng = np.random.default_rng(11)
    n = 6000
    px, py, s = rng.integers(x0, x1 + 1, n), rng.integers(y0, y1 + 1, n), rng.integers(0, 4, n)
    Lo, _ = o.render_samples(1, 9, px, py, s); Le, _ = e.render_samples(1, 9, px, py, s)
    L8, _ = Oracle(sc8).render_samples(1, 9, px, py, s)
    e.close()
    deep = np.abs(Lo - L8).max(1) > 0                       # samples that vertices at depth >= 8 contributed to
    assert deep.sum() > (300 if name != "pool" else 3), (name, int(deep.sum()))
    rel = np.abs(Lo - Le).max(1) / (np.abs(Lo).max(1) + 1e-6)
    assert rel.max() < 1e-5, (name, float(rel.max()))


def test_flat_triangles_beside_smooth_ones():
    """blingcu.h `tri_normals`: nine zeros mark a triangle of a mesh WITHOUT normals (TriangleMesh.hs:39-60 `Nothing`) in a scene
    that also holds smooth meshes (examples/cornell-box-underwater.bling: Cornell walls beside a height-map water surface). Shown by
    the `debug normals` integrator: zeroing every normal gives exactly the flat-shaded scene, zeroing half of them changes only
    samples that the all-smooth and the all-flat scene disagree on, and kernels and oracle agree bit for bit."""
    base = small(load_scene("smooth"), 48, 36, 2, 2); base.integrator_kind = IR.INTEGRATOR_NORMALS
    base.refl_basis = load_scene("textures").refl_basis          # rgbToSpectrumRefl basis (the older fixture carries none)
    nt = len(base.tri_verts)
    assert base.tri_normals is not None and nt > 10
    flat = copy.copy(base); flat.tri_normals = None
    zeros = copy.copy(base); zeros.tri_normals = np.zeros((nt, 9), np.float32)
    half = copy.copy(base); half.tri_normals = np.array(base.tri_normals, np.float32).reshape(nt, 9).copy(); half.tri_normals[::2] = 0
    o = Oracle(base); x0, x1, y0, y1 = o.sample_extent()
    xs, ys = np.meshgrid(np.arange(x0, x1 + 1), np.arange(y0, y1 + 1)); px, py = xs.ravel(), ys.ravel(); s = np.zeros_like(px)
    L = {}
    for k, sc in dict(smooth=base, flat=flat, zeros=zeros, half=half).items():
        e = EmuContext(); e.upload_scene(sc); L[k], _ = e.render_samples(1, 2, px, py, s); e.close()
        Lo, _ = Oracle(sc).render_samples(1, 2, px, py, s)
        assert np.array_equal(L[k], Lo), k
    assert np.array_equal(L["zeros"], L["flat"]) and not np.array_equal(L["smooth"], L["flat"])
    as_smooth, as_flat = (L["half"] == L["smooth"]).all(1), (L["half"] == L["flat"]).all(1)
    assert (as_smooth | as_flat).all() and (as_smooth & ~as_flat).sum() > 20 and (as_flat & ~as_smooth).sum() > 20


@pytest.mark.parametrize("seed", range(16))
def test_random_triangle_distributions_traversal_parity(seed):
    """the quantised BVH4 against brute force on triangle sets it could get wrong: 1 .. 3000 triangles at scales 1e-3 .. 1e4,
    small ones, slivers 2000 times longer than wide, exactly coplanar ones (flat boxes), big overlapping ones. Primitive ids agree
    except t-ties (coplanar / overlapping sets have many), t bit for bit, occlusion exactly. (120 such sets scanned: none bad.)"""
    r = np.random.default_rng(seed)
    n = int(r.choice([1, 2, 3, 7, 50, 400, 3000]))
    scale = 10.0 ** r.uniform(-3, 4)
    kind = seed % 4
    c = r.uniform(-1, 1, (n, 1, 3)) * scale
    if kind == 0: v = c + r.normal(size=(n, 3, 3)) * scale * 0.05
    elif kind == 1: v = c + r.normal(size=(n, 3, 3)) * scale * np.array([2.0, 0.001, 0.001])
    elif kind == 2: v = c * np.array([1, 1, 0.0]) + r.normal(size=(n, 3, 3)) * scale * np.array([0.2, 0.2, 0])
    else: v = c * 0.01 + r.normal(size=(n, 3, 3)) * scale * 0.5
    sc = _tri_scene(v.reshape(n, 9).astype(np.float32))
    _traversal_parity(sc, random_rays(sc, 2500, seed))


@pytest.mark.parametrize("seed", range(8))
@pytest.mark.parametrize("collapse", [(0.0, 2), (0.5, 3)])
def test_rays_through_the_extremes_of_the_boxes(seed, collapse):
    extreme_rays_case(seed, collapse, EmuContext)


def extreme_rays_case(seed, collapse, make_ctx):
    """The child boxes are 8-bit boxes moved out by 1/64 of a cell (bvh_build.cpp, bvh.h: the folded 2^15 of the dequantisation is good
    to 1/512 of a cell). A box that came out too small by a hair would lose exactly the hits at its faces -- so aim rays at the
    vertices and edges of the triangles, i.e. at the extremes of the leaf boxes, from all around, at scales 1e-3 .. 1e4 and offsets up
    to 1e3 scales away from the origin (coarse float grid), for the greedy and the optimal collapse: the quantised tree must find
    every hit brute force finds, t bit for bit."""
    r = np.random.default_rng(1000 + seed)
    n = int(r.choice([40, 400, 5000]))
    scale = 10.0 ** r.uniform(-3, 4)
    off = r.uniform(-1, 1, 3) * scale * (10.0 ** r.uniform(0, 3))
    c = off + r.uniform(-1, 1, (n, 1, 3)) * scale
    v = (c + r.normal(size=(n, 3, 3)) * scale * (0.02 if seed % 2 else 0.2)).astype(np.float32)
    sc = _tri_scene(v.reshape(n, 9))
    m = 3000
    tri = r.integers(0, n, m)
    w = r.dirichlet([0.3, 0.3, 0.3], m); w[: m // 2] = np.eye(3)[r.integers(0, 3, m // 2)]      # half exactly at a vertex, half near an edge
    target = (v[tri].astype(np.float64) * w[:, :, None]).sum(1)
    d = r.normal(size=(m, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    dist = scale * 10.0 ** r.uniform(-1, 1.5, (m, 1))
    rays = np.zeros(m, dtype=random_rays(sc, 1, 0).dtype)
    rays["o"] = (target - d * dist).astype(np.float32); rays["d"] = d.astype(np.float32)
    rays["tmin"] = 0.0; rays["tmax"] = np.inf
    o = Oracle(sc); e = make_ctx(); e.set_option("bvh_collapse_cp", collapse[0]); e.set_option("bvh_leaf", collapse[1]); e.upload_scene(sc)
    hb = o.trace_nearest(rays, mode="brute"); he = e.trace_nearest(rays)
    oo, oe = o.trace_occluded(rays, mode="brute"), e.trace_occluded(rays)
    e.close()
    ties, bad = compare_hits(he, hb)
    assert bad == 0, (ties, bad)
    same = (hb["prim"] >= 0) & (he["prim"] == hb["prim"])
    assert same.sum() > m // 4 and np.array_equal(he["t"][same], hb["t"][same])
    assert np.array_equal(oe.astype(bool), oo.astype(bool))
