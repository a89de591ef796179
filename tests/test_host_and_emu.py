"""CPU suite, part 2: the C-ABI library loads and exports every declared symbol (no compute without a GPU),
the kernel BODIES (bling_b200/csrc/bodies.h) driven by the CPU emulator agree with the oracle, the host-side
renderer logic and its 2-rank (gloo) sharding."""
import ctypes
import os
import re
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

from bling_b200 import api, ir as IR
from bling_b200.renderer import CudaRenderer, PassDone, RenderJob, shard_range
from oracle.oracle_py import Oracle
from tests.conftest import ALL_SCENES, EMU_ONLY, ROOT, SCENES, camera_rays, compare_hits, has_gpu, load_scene, random_rays, small
from tests.emu.emu_py import EmuContext


def test_abi_exports_every_declared_symbol():
    import __graft_entry__ as g
    so = g.build_cuda()
    hdr = (ROOT / "include" / "blingcu.h").read_text()
    declared = sorted(set(re.findall(r"\b(blingcu_[a-z_]+)\s*\(", hdr)))
    assert len(declared) == len(api.SYMBOLS) == 38
    L = ctypes.CDLL(str(so))
    for name in declared:
        assert hasattr(L, name), name
    assert sorted("blingcu_" + s for s in api.SYMBOLS) == declared


@pytest.mark.skipif(has_gpu(), reason="checks the no-GPU failure mode")
def test_ctypes_mirror_matches_the_header_layout():
    """bling_b200/ir.py mirrors include/blingcu.h by hand: hold every struct's size and every field's offset against what the
    C compiler computes from the header (tools/abi_layout.py; the same table generates haskell/Layout.hs)."""
    sys.path.insert(0, str(ROOT / "tools"))
    import abi_layout
    lay = abi_layout.layout()
    mirror = {"blingcu_spectrum": IR.Spectrum, "blingcu_shape": IR.Shape, "blingcu_texture": IR.Texture, "blingcu_image": IR.ImageC,
              "blingcu_material": IR.Material, "blingcu_light": IR.Light, "blingcu_sunsky": IR.SunSky, "blingcu_envmap": IR.EnvMap,
              "blingcu_camera": IR.Camera, "blingcu_scene": IR.SceneC, "blingcu_ray": IR.Ray, "blingcu_hit": IR.Hit, "blingcu_stats": IR.Stats,
              "blingcu_kdnode": IR.KdNode}
    assert IR.KDNODE_DTYPE.itemsize == lay["blingcu_kdnode"]["sizeof"] and all(IR.KDNODE_DTYPE.fields[k][1] == v for k, v in lay["blingcu_kdnode"].items() if k != "sizeof")
    assert set(lay) == set(mirror)
    import ctypes as C
    for name, ty in mirror.items():
        assert C.sizeof(ty) == lay[name]["sizeof"], name
        fields = {f[0]: getattr(ty, f[0]).offset for f in ty._fields_}
        assert fields == {k: v for k, v in lay[name].items() if k != "sizeof"}, name
    hs = (ROOT / "haskell" / "Layout.hs").read_text()
    assert hs == abi_layout.haskell(lay), "haskell/Layout.hs is stale: python tools/abi_layout.py --haskell"
    # GHC is absent, so at least: every layout name the Haskell marshalling code uses exists in the generated table
    defined = set(re.findall(r"^(\w+) ::", hs, flags=re.M))
    for f in ("SceneIR.hs", "Cuda.hs"):
        used = set(re.findall(r"\b(off[A-Z]\w+|sizeOf[A-Z]\w+)\b", (ROOT / "haskell" / f).read_text()))
        assert used <= defined, (f, sorted(used - defined))


def test_product_fails_loudly_without_gpu():
    """no CPU fallback: creating a context without a device is an error, not a silent CPU path."""
    with pytest.raises(api.BlingCuError) as e:
        api.Context(0)
    assert e.value.code == 3 and "no CPU fallback" in str(e.value)


def test_ir_roundtrip(tmp_path):
    sc = load_scene("environment")
    sc.save(tmp_path / "x.npz")
    sc2 = IR.SceneIR.load(tmp_path / "x.npz")
    assert np.array_equal(sc.env_arrays[0].rgb, sc2.env_arrays[0].rgb)
    assert bytes(sc.shapes[3]) == bytes(sc2.shapes[3]) and sc2.max_depth == sc.max_depth
    a, _ = sc.to_c(); b, _ = sc2.to_c()
    assert a.n_shapes == b.n_shapes == 7 and a.width == 1920


@pytest.mark.parametrize("name", ALL_SCENES + EMU_ONLY)
def test_emulated_traversal_matches_oracle(name):
    """parity (a) for the traversal BODY + BVH builder: prim id exact except measured t-ties, t within 1e-5."""
    sc = load_scene(name)
    o = Oracle(sc, kdtree=(name == "ducky")); e = EmuContext(); e.upload_scene(sc)
    n = 3000
    rays = np.concatenate([random_rays(sc, n, 3), camera_rays(None, sc, n, 4)])
    ref = o.trace_nearest(rays, "kd" if name == "ducky" else "brute")
    got = e.trace_nearest(rays)
    ties, bad = compare_hits(got, ref)
    assert bad == 0 and ties <= 0.01 * len(rays), (ties, bad)
    hit = ref["prim"] >= 0
    assert np.array_equal(got["t"][hit & (got["prim"] == ref["prim"])], ref["t"][hit & (got["prim"] == ref["prim"])])   # bit-exact t
    occ_ref = o.trace_occluded(rays, "kd" if name == "ducky" else "brute")
    assert (e.trace_occluded(rays) != occ_ref).mean() < 2e-3
    _, nodes, prims = e.trace_stats(rays[:500])
    assert nodes.max() > 0 and prims.sum() > 0


@pytest.mark.parametrize("name", ["cornell-box", "zoo", "ducky", "glass-torus"])
@pytest.mark.parametrize("cp,leaf", [(0.5, 3), (1.0, 1), (0.25, 8)])
def test_optimal_collapse_gives_the_same_hits(name, cp, leaf):
    """The SAH-optimal collapse (bvh_build.cpp::Collapse, option bvh_collapse_cp) changes the tree, never the answer: the
    nearest hit is global, so prim / t / b1 / b2 and the occlusion flags equal those of the greedy tree bit for
    bit (exact t-ties aside), and every leaf holds at most `bvh_leaf` items."""
    sc = load_scene(name)
    a = EmuContext(); a.set_option("bvh_collapse_cp", 0); a.set_option("bvh_leaf", 2); a.upload_scene(sc)    # round 1's greedy collapse
    b = EmuContext(); b.set_option("bvh_collapse_cp", cp); b.set_option("bvh_leaf", leaf); b.upload_scene(sc)
    rays = np.concatenate([random_rays(sc, 3000, 5), camera_rays(None, sc, 3000, 6)])
    ha, hb = a.trace_nearest(rays), b.trace_nearest(rays)
    ties, bad = compare_hits(hb, ha)
    assert bad == 0 and ties <= 0.01 * len(rays), (ties, bad)
    same = ha["prim"] == hb["prim"]
    for f in ("t", "b1", "b2"):
        assert np.array_equal(ha[f][same], hb[f][same])
    assert np.array_equal(a.trace_occluded(rays), b.trace_occluded(rays))
    _, _, pa = a.trace_stats(rays[:1500]); _, nb_, pb = b.trace_stats(rays[:1500])
    assert nb_.max() > 0
    if leaf == 1 and sc.n_prims > 4:
        assert pb.sum() <= pa.sum()      # single-item leaves never test more primitives than the two-item leaves of the greedy tree
    with pytest.raises(Exception):
        b.set_option("bvh_collapse_cp", -1)


@pytest.mark.parametrize("name", ALL_SCENES + EMU_ONLY)
def test_emulated_path_samples_match_oracle(name):
    """per-sample radiance of the wavefront bodies == oracle's recursive nextVertex on the same sampler SPEC."""
    sc = small(load_scene(name), 40, 30, 4, 4)
    o = Oracle(sc); e = EmuContext(); e.upload_scene(sc)
    x0, x1, y0, y1 = o.sample_extent()
    rng = np.random.default_rng(5)
    n = 1500
    px, py, s = rng.integers(x0, x1 + 1, n), rng.integers(y0, y1 + 1, n), rng.integers(0, 16, n)
    Lo, xyo = o.render_samples(2, 77, px, py, s)
    Le, xye = e.render_samples(2, 77, px, py, s)
    assert np.array_equal(xyo, xye)
    rel = np.abs(Lo - Le).max(1) / (np.abs(Lo).max(1) + 1e-6)
    assert (rel < 1e-4).mean() > 0.999, rel.max()


def test_sampler_non_power_of_two_strata():
    """3x5 strata: the general (division / modulo) path of the sampler SPEC, the configs only reach the power-of-two one."""
    sc = small(load_scene("cornell-box"), 24, 18, 3, 5)
    o = Oracle(sc); e = EmuContext(); e.upload_scene(sc)
    x0, x1, y0, y1 = o.sample_extent()
    rng = np.random.default_rng(9)
    n = 600
    px, py, s = rng.integers(x0, x1 + 1, n), rng.integers(y0, y1 + 1, n), rng.integers(0, 15, n)
    Lo, xyo = o.render_samples(1, 5, px, py, s)
    Le, xye = e.render_samples(1, 5, px, py, s)
    assert np.array_equal(xyo, xye)
    rel = np.abs(Lo - Le).max(1) / (np.abs(Lo).max(1) + 1e-6)
    assert (rel < 1e-4).mean() > 0.995


def test_emulated_film_matches_oracle_tiles():
    """the atomic-free film gather reproduces per-tile addSample + addTile (Q10) to float rounding."""
    for name, wh in (("cornell-box", (37, 29)), ("sun-sky", (33, 20)), ("ducky", (30, 18))):
        sc = small(load_scene(name), *wh, 2, 2)
        o = Oracle(sc); e = EmuContext(); e.upload_scene(sc)
        o.render_pass(1, 21, threads=4); e.render_pass(1, 21)
        fo, fe = o.read_film(), e.read_film()
        assert np.abs(fo - fe).max() <= 2e-5 * max(1.0, np.abs(fo).max()), name
        so, se = o.stats(), e.stats()
        for k in ("samples", "rays_camera", "rays_shadow", "dropped_samples"):
            assert so[k] == se[k], (name, k)
        assert so["rays_extension"] == se["rays_extension"] + se["rays_ext_culled"], name
        if name == "ducky":      # maxDepth 5: the last bounce of every surviving diffuse path is culled
            assert se["rays_ext_culled"] > 0
        # the product does not trace BSDF-MIS rays that cannot reach the chosen light; traced + culled == reference count
        assert so["rays_mis"] == se["rays_mis"] + se["rays_mis_culled"], name
        if name == "cornell-box":
            assert se["rays_mis_culled"] > se["rays_mis"]


@pytest.mark.parametrize("name", ["glass-torus", "direct", "textures"])
def test_slices_and_batches_compose(name):
    """render_pass == union of slices == any batch size (linearity of the film in the sample set) -- also for the
    direct-lighting integrator, whose batches carry spawned branch slots (the sharding unit of SURVEY §8e is the slice)."""
    sc = small(load_scene(name), 32, 24, 4, 4)
    a = EmuContext(); a.upload_scene(sc); a.render_pass(3, 9)
    b = EmuContext(); b.set_option("batch_samples", 3000); b.upload_scene(sc)
    b.render_slice(3, 9, 0, 5); b.render_slice(3, 9, 5, 16)
    fa, fb = a.read_film(), b.read_film()
    assert np.abs(fa - fb).max() <= 1e-5 * np.abs(fa).max()
    with pytest.raises(api.BlingCuError):
        b.render_slice(3, 9, 4, 17)
    b.clear_film(); assert b.read_film().max() == 0
    b.film_add_host(fa); assert np.array_equal(b.read_film(), fa)


def test_fused_and_separate_nee_resolve_are_bit_identical():
    """option "fuse_resolve": the any-hit query adds an unoccluded shadow ray's pending contribution itself (small scenes, the
    default below 2^20 primitives) or a separate resolve launch does (large scenes): same arithmetic, same order."""
    sc = small(load_scene("zoo"), 40, 28, 4, 4)
    films = []
    for mode in (0, 1, -1):
        e = EmuContext(); e.set_option("fuse_resolve", mode); e.upload_scene(sc); e.render_pass(2, 13)
        films.append((e.read_film(), e.stats()["kernel_launches"])); e.close()
    assert np.array_equal(films[0][0], films[1][0]) and np.array_equal(films[1][0], films[2][0])
    assert films[0][1] > films[1][1] == films[2][1]          # one launch per bounce less when fused; small scene: fused by default


def test_edge_scenes_empty_and_unlit():
    """no primitives (every ray escapes to the environment) and no lights (black film, no NEE rays)."""
    from tests.conftest import stripped
    base = small(load_scene("envcam"), 24, 12, 2, 2)
    for sc, lit in ((stripped(base), True), (stripped(base, prims=False, lights=True), False)):
        o = Oracle(sc); e = EmuContext(); e.upload_scene(sc)
        o.render_pass(1, 3, threads=2); e.render_pass(1, 3)
        fo, fe = o.read_film(), e.read_film()
        assert np.isfinite(fe).all() and np.abs(fo - fe).max() <= 2e-5 * max(1.0, np.abs(fo).max())
        assert (fe[..., 1:].sum() > 0) == lit
        so, se = o.stats(), e.stats()
        assert so["samples"] == se["samples"] and se["rays_shadow"] == so["rays_shadow"]
        rays = random_rays(load_scene("envcam"), 100, 1)
        if not len(sc.shapes):
            assert (e.trace_nearest(rays)["prim"] == -1).all() and not e.trace_occluded(rays).any()


def test_state_follows_the_integrator_on_a_reused_context():
    """ADVICE r1 (high), emulator leg: path scene first, then a direct-lighting scene that needs fewer slots."""
    e = EmuContext()
    e.upload_scene(small(load_scene("cornell-box"), 40, 40, 2, 2)); e.render_pass(1, 3)
    dl = small(load_scene("direct"), 24, 16, 2, 2)
    e.upload_scene(dl); e.render_pass(1, 5)
    fe = e.read_film(); e.close()
    e2 = EmuContext(); e2.upload_scene(dl); e2.render_pass(1, 5)
    assert np.array_equal(fe, e2.read_film())            # same film as a fresh context
    e2.close()


def test_null_uvs_default_and_null_tables_are_rejected():
    """tri_uvs == NULL means the default (0,0,1,0,1,1) of TriangleMesh.hs:119-120; null geometry tables
    with a non-zero count are BLINGCU_EINVAL instead of a crash; batch_samples is range-checked."""
    import ctypes as C
    sc = small(load_scene("cornell-box"), 24, 24, 2, 2)
    e = EmuContext()
    c, keep = sc.to_c()
    c.tri_uvs = None
    e._chk(e._f("upload_scene")(e._h, C.byref(c))); e.scene = sc
    e.render_pass(1, 2); f0 = e.read_film()
    sc2 = small(load_scene("cornell-box"), 24, 24, 2, 2)
    sc2.tri_uvs = np.tile(np.array([0, 0, 1, 0, 1, 1], np.float32), (len(sc2.tri_material), 1))
    e.upload_scene(sc2); e.render_pass(1, 2)
    assert np.array_equal(f0, e.read_film())
    c, keep = sc.to_c()
    c.tri_verts = None
    assert e._f("upload_scene")(e._h, C.byref(c)) == 1
    for bad in (0.5, 1e12, float("nan")):
        with pytest.raises(api.BlingCuError):
            e.set_option("batch_samples", bad)
    e.close()


def _check_reference_kdtree(make_ctx, sc, nrays):
    """the oracle's SAH kd-tree (KdTree.hs restated), flattened, uploaded as an alternative accelerator input and
    walked by kdtree.h exactly as `traverse` does: every hit field AND the per-ray dbgTraverse counters equal the oracle's."""
    o = Oracle(sc, kdtree=True)
    nodes, leaf, root, bounds = o.kdtree_flat()
    c = make_ctx(); c.upload_scene(sc); c.upload_kdtree(nodes, leaf, root, bounds)
    rays = np.concatenate([random_rays(sc, nrays, 17), camera_rays(None, sc, nrays, 18)])
    want, wn, wi = o.trace_kd_stats(rays)
    got, gn, gi = c.trace_kdtree(rays)
    for f in ("t", "prim", "b1", "b2"):
        assert np.array_equal(got[f], want[f]), f
    assert np.array_equal(gn, wn) and np.array_equal(gi, wi)          # TraversalStats, ray by ray
    assert (want["prim"] >= 0).mean() > 0.05 and wn.max() > 3
    bvh = c.trace_nearest(rays)                                       # and the product's own accelerator finds the same hits
    ties, bad = compare_hits(bvh, want)
    assert bad == 0
    return c, nodes, leaf, root, bounds


@pytest.mark.parametrize("name", ["cornell-box", "zoo", "ducky"])
def test_emulated_reference_kdtree_traversal(name):
    c, nodes, leaf, root, bounds = _check_reference_kdtree(EmuContext, load_scene(name), 3000 if name == "ducky" else 6000)
    # malformed trees are rejected, not walked
    bad = nodes.copy(); inner = np.flatnonzero(bad["left"] >= 0)
    if len(inner):
        bad["right"][inner[0]] = len(bad) + 5
        with pytest.raises(api.BlingCuError):
            c.upload_kdtree(bad, leaf, root, bounds)
        loop = nodes.copy(); loop["left"][inner[0]] = root
        with pytest.raises(api.BlingCuError):
            c.upload_kdtree(loop, leaf, root, bounds)
    with pytest.raises(api.BlingCuError):
        c.upload_kdtree(nodes, leaf + 10_000_000, root, bounds)
    with pytest.raises(api.BlingCuError):
        c.trace_kdtree(np.zeros(1, IR.RAY_DTYPE))                     # the failed uploads dropped the tree
    c.close()
    fresh = EmuContext()
    with pytest.raises(api.BlingCuError):
        fresh.upload_kdtree(nodes, leaf, root, bounds)                # no scene yet
    fresh.close()


def test_host_array_and_out_buffers():
    """blingcu_host_alloc / host_free and caller-provided output buffers through the binding (emulator: plain malloc)"""
    sc = small(load_scene("cornell-box"), 16, 16, 1, 1)
    e = EmuContext(); e.upload_scene(sc)
    rays = random_rays(sc, 257, 3)
    pr = e.host_array(len(rays), IR.RAY_DTYPE); pr[:] = rays
    ph = e.host_array(len(rays), IR.HIT_DTYPE)
    assert e.trace_nearest(pr, out=ph) is ph and np.array_equal(ph, e.trace_nearest(rays))
    assert len(e.host_array(0, np.uint8)) == 0
    e.close()


def test_api_error_paths():
    e = EmuContext()
    with pytest.raises(api.BlingCuError) as ex:
        e.render_pass(1, 1)
    assert ex.value.code == 4
    sc = small(load_scene("cornell-box"), 16, 16, 1, 1)
    bad = IR.SceneIR.load(ROOT / "tests" / "golden" / "scenes" / "cornell-box.npz")
    bad.tri_material = bad.tri_material.copy(); bad.tri_material[0] = 99
    with pytest.raises(api.BlingCuError) as ex:
        e.upload_scene(bad)
    assert ex.value.code == 1 and "material" in str(ex.value)
    e.upload_scene(sc)
    assert len(e.trace_nearest(np.zeros(0, IR.RAY_DTYPE))) == 0
    with pytest.raises(api.BlingCuError):
        e.set_option("no_such_option", 1)


def test_renderer_progressive_loop():
    """prender semantics: PassDone per pass with the accumulated film; returning False stops (Rendering.hs:137-138)."""
    sc = small(load_scene("cornell-box"), 24, 24, 2, 2)
    r = CudaRenderer(context_cls=EmuContext, seed=5)
    seen = []

    def report(p):
        if isinstance(p, PassDone):
            seen.append((p.pass_num, float(p.final_img[..., 0].sum())))
            return p.pass_num < 3
        return True
    r.render(RenderJob(sc), report)
    assert [p for p, _ in seen] == [1, 2, 3]
    assert seen[0][1] < seen[1][1] < seen[2][1]            # filter weights accumulate over passes
    assert abs(seen[2][1] / seen[0][1] - 3.0) < 2e-2


def test_in_process_group_reduce_sums_the_films():
    """blingcu_comm_init_all + blingcu_reduce_film_group (one process driving several contexts, the Haskell host's mode): film_sum
    is the sum of the private films on every rank (root < 0) or on the root only, the private films stay untouched, and the
    sample shards of the contexts compose to the single-context pass. Emulator leg: the host logic of include/blingcu.h."""
    from bling_b200.renderer import MultiDeviceRenderer
    sc = small(load_scene("cornell-box"), 40, 30, 4, 4)
    one = EmuContext(); one.upload_scene(sc); one.render_pass(1, 0x5EED); full = one.read_film(); one.close()
    r = MultiDeviceRenderer([0, 1, 2], seed=0x5EED, context_cls=EmuContext)
    imgs = []
    r.render(RenderJob(sc), lambda p: (imgs.append(p.final_img.copy()) or len(imgs) < 2) if isinstance(p, PassDone) else True)
    assert np.abs(imgs[0] - full).max() <= 1e-5 * np.abs(full).max()                 # shards 0-5, 5-10, 10-16 compose
    privates = [c.read_film() for c in r.ctxs]
    assert np.abs(sum(privates) - imgs[1]).max() <= 1e-5 * np.abs(imgs[1]).max()     # film_sum == sum of the private films
    assert all(p[..., 0].sum() < imgs[1][..., 0].sum() for p in privates)
    EmuContext.reduce_film_group(r.ctxs, root=-1)                                     # all-reduce: every rank gets the sum
    sums = [c.read_film_sum() for c in r.ctxs]
    assert all(np.array_equal(s_, sums[0]) for s_ in sums[1:])
    with pytest.raises(api.BlingCuError) as ex:                                       # a lone reduce on a 3-rank communicator
        r.ctxs[0].reduce_film()
    assert ex.value.code == 5
    with pytest.raises(api.BlingCuError):
        EmuContext.reduce_film_group(r.ctxs, root=7)
    r.close()
    solo = EmuContext(); solo.upload_scene(sc); solo.render_pass(1, 3)
    with pytest.raises(api.BlingCuError):
        solo.read_film_sum()                                                          # nothing reduced yet
    solo.comm_init(0, 1); solo.reduce_film()
    assert np.array_equal(solo.read_film_sum(), solo.read_film())                     # one rank: film_sum is a copy
    assert len(EmuContext.comm_unique_id()) == api.COMM_ID_BYTES
    solo.close()


def test_shard_range_partitions():
    for spp in (1, 4, 64, 1024):
        for world in (1, 2, 3, 8):
            r = [shard_range(spp, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == spp and all(a[1] == b[0] for a, b in zip(r, r[1:]))


_WORKER = r"""
import os, sys
sys.path.insert(0, {root!r})
import numpy as np, torch.distributed as dist
from bling_b200.renderer import CudaRenderer, RenderJob, PassDone
from tests.emu.emu_py import EmuContext
from tests.conftest import load_scene, small
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=2)
sc = small(load_scene("cornell-box"), 24, 20, 4, 4)
r = CudaRenderer(context_cls=EmuContext, seed=31)
out = []
r.render(RenderJob(sc), lambda p: (out.append(p.final_img) or False) if isinstance(p, PassDone) else True)
if dist.get_rank() == 0: np.save({out!r}, out[0])
dist.destroy_process_group()
"""


def test_two_rank_sample_sharding_gloo(tmp_path):
    """world_size 2: each rank renders half of the pass' sample indices, one all-reduce sums the films; the result
    equals the single-process pass (same sampler keys) to float rounding."""
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    out = str(tmp_path / "film.npy")
    script = tmp_path / "w.py"
    script.write_text(_WORKER.format(root=str(ROOT), port=port, out=out))
    procs = [subprocess.Popen([sys.executable, str(script), str(k)], env=dict(os.environ, OMP_NUM_THREADS="1")) for k in range(2)]
    for p in procs:
        assert p.wait(timeout=240) == 0
    film2 = np.load(out)
    sc = small(load_scene("cornell-box"), 24, 20, 4, 4)
    e = EmuContext(); e.upload_scene(sc); e.render_pass(1, 31)
    film1 = e.read_film()
    assert np.abs(film1 - film2).max() <= 1e-5 * np.abs(film1).max()


def test_loader_bezier_and_heightmap_meshes(tmp_path):
    """host side of SURVEY §8(f)2 "mesh shading normals": tesselateBezier (Primitive/Bezier.hs:80-105) and heightMap
    (Primitive/Heightmap.hs:22-49) become world-space triangles with per-vertex normals and uvs."""
    from bling_b200.host.loader import load_scene as parse
    ctrl = [(x, (x * y) % 3 - 1.0, y) for y in range(4) for x in range(4)]      # 4 x 4 control points, row i = 12 floats
    txt = ("filter box\nimageSize 16 12\n"
           "renderer { sampler sampled { sampler { stratified 2 2 } integrator { path maxDepth 3 sampleDepth 1 } } }\n"
           "camera { perspective fov 40 lensRadius 0 focalDistance 5 }\n"
           "light { infinite { } l { constant rgbI 1 1 1 } }\n"
           "transform { translate 10 0 0 }\n"
           "prim { bezier subdivs 4 p { " + ", ".join(f"{c:g}" for p in ctrl for c in p) + " } }\n"
           "newTransform { }\n"
           "prim { heightMap 5 4 { scale 2 { fbm 0.5 octaves 2 omega 0.5 } } { scale 3 1 2 } }\n")
    f = tmp_path / "m.bling"; f.write_text(txt)
    sc = parse(f)
    nb, nh = 4 * 4 * 2, (5 - 1) * (4 - 1) * 2
    assert len(sc.tri_verts) == nb + nh and sc.tri_normals is not None and sc.tri_normals.shape == (nb + nh, 9)
    # prims are prepended block by block (RenderJob.hs:49-52): the height map (parsed last) comes first
    hm, bz = sc.tri_verts[:nh].reshape(-1, 3), sc.tri_verts[nh:].reshape(-1, 3)
    assert np.allclose(hm[:, [0, 2]].min(0), [0, 0]) and np.allclose(hm[:, [0, 2]].max(0), [3, 2])    # [0,1]^2 scaled by (3, ., 2)
    assert np.abs(hm[:, 1]).max() <= 2 * 1.5                                                         # |2 * fbm| stays small
    # a Bezier patch interpolates its corner control points; the primitive was translated by +10 in x
    corners = np.array([ctrl[0], ctrl[3], ctrl[12], ctrl[15]], np.float32) + np.array([10, 0, 0], np.float32)
    for c in corners: assert np.abs(bz - c).sum(1).min() < 1e-5
    assert np.array_equal(sc.tri_uvs[nh], np.array([0, 0, 0.25, 0, 0, 0.25], np.float32))            # (v00, v10, v01), step 1/4
    n = sc.tri_normals[:nh].reshape(-1, 3)
    assert np.allclose(np.linalg.norm(n * np.array([3, 1, 2], np.float32), axis=1), 1, atol=1e-4)    # unit normals through transNormal (scale 3 1 2)
    o = Oracle(sc); e = EmuContext(); e.upload_scene(sc)
    rays = random_rays(sc, 1500, 4)
    ties, bad = compare_hits(e.trace_nearest(rays), o.trace_nearest(rays, mode="brute"))
    assert bad == 0
    e.close()
