"""Randomised differential test: scenes drawn from the whole grammar this path supports (every shape, material, texture, light,
integrator, sampler and filter kind, random transforms) are written as `.bling` text, flattened by the stand-in loader and run
through the kernel bodies (CPU emulator) and the oracle; per-sample radiance must agree. Deterministic seeds.
(This kind of test found the Box `intersects` / `intersect` quirk of DESIGN §4d.)"""
import numpy as np
import pytest

from bling_b200.host.loader import load_scene as parse
from oracle.oracle_py import Oracle
from tests.conftest import compare_hits, random_rays
from tests.emu.emu_py import EmuContext


def _rgb(r, kind="rgbR", lo=0.05, hi=0.95):
    return f"{kind} {r.uniform(lo, hi):.3f} {r.uniform(lo, hi):.3f} {r.uniform(lo, hi):.3f}"


def _map3d(r):
    return f"map {{ identity {{ scale {r.uniform(0.5, 4):.2f} {r.uniform(0.5, 4):.2f} {r.uniform(0.5, 4):.2f} translate {r.uniform(-1, 1):.2f} {r.uniform(-1, 1):.2f} 0 }} }}"


def _map2d(r):
    if r.random() < 0.5: return f"map {{ uv {r.uniform(1, 8):.2f} {r.uniform(1, 8):.2f} {r.random():.2f} {r.random():.2f} }}"
    return f"map {{ planar {r.uniform(0.2, 2):.2f} 0 0 0 {r.uniform(0.2, 2):.2f} {r.uniform(-1, 1):.2f} {r.random():.2f} {r.random():.2f} }}"


def _scalar(r, depth=0):
    k = r.integers(0, 6 if depth < 2 else 5)
    if k == 0: return f"constant {r.uniform(0.01, 0.6):.3f}"
    if k == 1: return f"perlin {_map3d(r)}"
    if k == 2: return f"fbm octaves {r.integers(1, 4)} omega {r.uniform(0.3, 0.7):.2f} {_map3d(r)}"
    if k == 3: return f"cellNoise {r.choice(['euclidian', 'euclidian2', 'manhattan', 'chebyshev'])} {_map3d(r)}"
    if k == 4: return f"crystal octaves {r.integers(2, 8)} {_map2d(r)}"
    return f"scale {r.uniform(0, 0.3):.3f} {r.uniform(0.05, 0.5):.3f} tex {{ {_scalar(r, depth + 1)} }}"


def _spectrum_tex(r, depth=0):
    k = r.integers(0, 5 if depth < 2 else 1)
    if k == 0: return f"constant {_rgb(r)}"
    if k == 1: return f"graphPaper {r.uniform(0.02, 0.2):.2f} {_map2d(r)} tex1 {{ {_spectrum_tex(r, depth + 1)} }} tex2 {{ {_spectrum_tex(r, depth + 1)} }}"
    if k == 2: return f"checker {r.uniform(0.5, 3):.2f} {r.uniform(0.5, 3):.2f} {r.uniform(0.5, 3):.2f} tex1 {{ {_spectrum_tex(r, depth + 1)} }} tex2 {{ {_spectrum_tex(r, depth + 1)} }}"
    if k == 3: return f"blend tex1 {{ {_spectrum_tex(r, 2)} }} tex2 {{ {_spectrum_tex(r, 2)} }} f {{ {_scalar(r)} }}"
    steps = ", ".join(f"{p:.2f} {_rgb(r)}" for p in sorted(r.uniform(-0.2, 1.2, r.integers(1, 4))))
    return f"gradient f {{ {_scalar(r)} }} steps {{ {steps} }}"


def _small(r): return f"constant {r.uniform(0.005, 0.3):.4f}" if r.random() < 0.6 else _scalar(r)


def _material(r):
    k = r.integers(0, 9)
    bump = f"bumpMap bump {{ scale 0 {r.uniform(0.05, 0.3):.2f} tex {{ {_scalar(r)} }} }} " if r.random() < 0.3 else ""
    if k == 0: body = f"matte kd {{ {_spectrum_tex(r)} }} sigma {{ {_small(r) if r.random() < 0.5 else 'constant 0'} }}"
    elif k == 1: body = f"glass ior {{ constant {r.uniform(1.1, 1.8):.2f} }} kr {{ constant {_rgb(r, lo=0.7, hi=1)} }} kt {{ {_spectrum_tex(r, 2)} }}"
    elif k == 2: body = f"mirror kr {{ {_spectrum_tex(r)} }}"
    elif k == 3: body = f"plastic kd {{ {_spectrum_tex(r)} }} ks {{ constant {_rgb(r)} }} rough {{ {_small(r)} }}"
    elif k == 4: body = f"metal eta {{ constant {_rgb(r, lo=0.2, hi=2)} }} k {{ constant {_rgb(r, lo=1, hi=4)} }} rough {{ {_small(r)} }}"
    elif k == 5: body = f"shinyMetal kr {{ constant {_rgb(r)} }} ks {{ constant {_rgb(r)} }} rough {{ {_small(r)} }}"
    elif k == 6: body = f"transMatte kr {{ constant {_rgb(r)} }} kt {{ {_spectrum_tex(r, 1)} }} ks {{ constant {r.choice([0, 0.4])} }}"
    elif k == 7: body = (f"substrate kd {{ {_spectrum_tex(r)} }} ks {{ constant {_rgb(r, hi=0.3)} }} ka {{ constant {_rgb(r)} }} "
                         f"urough {{ {_small(r)} }} vrough {{ {_small(r)} }} depth {{ constant {r.choice([0, 0.3])} }}")
    else: body = "blackbody"
    return f"material {{ {bump}{body} }}"


def _shape(r):
    k = r.integers(0, 5)
    if k == 0: return f"box pmin {-r.uniform(0.3, 1):.2f} {-r.uniform(0.3, 1):.2f} {-r.uniform(0.3, 1):.2f} pmax {r.uniform(0.3, 1):.2f} {r.uniform(0.3, 1):.2f} {r.uniform(0.3, 1):.2f}"
    if k == 1: return f"cylinder radius {r.uniform(0.3, 0.9):.2f} zmin {-r.uniform(0.2, 1):.2f} zmax {r.uniform(0.2, 1):.2f} phiMax {r.choice([360, 270, 180])}"
    if k == 2: return f"disk height {r.uniform(-0.3, 0.3):.2f} radius {r.uniform(0.5, 1.2):.2f} innerRadius {r.choice([0, 0.2])} phiMax {r.choice([360, 200])}"
    if k == 3: return f"quad {r.uniform(0.4, 1.5):.2f} {r.uniform(0.4, 1.5):.2f}"
    return f"sphere radius {r.uniform(0.4, 1.1):.2f}"


def _transform(r, spread=3.0):
    return (f"newTransform {{ rotateX {r.uniform(-180, 180):.1f} rotateY {r.uniform(-180, 180):.1f} scale {r.uniform(0.6, 1.5):.2f} {r.uniform(0.6, 1.5):.2f} {r.uniform(0.6, 1.5):.2f} "
            f"translate {r.uniform(-spread, spread):.2f} {r.uniform(0.3, 2.5):.2f} {r.uniform(-spread, spread):.2f} }}")


def random_scene_text(seed):
    r = np.random.default_rng(seed)
    md = int(r.integers(1, 6))
    integ = f"directLighting maxDepth {md}" if r.random() < 0.3 else f"path maxDepth {md} sampleDepth {r.integers(0, 4)}"
    smp = f"stratified {r.integers(1, 4)} {r.integers(1, 4)}" if r.random() < 0.7 else f"random {r.integers(1, 9)}"
    filt = r.choice(["box", "triangle 2 2", "mitchell 2 2 0.333333 0.333333", "gauss 2 2 2", "sinc 3 3 3"])
    out = [f"filter {filt}", "imageSize 40 30",
           f"renderer {{ sampler sampled {{ sampler {{ {smp} }} integrator {{ {integ} }} }} }}",
           f"transform {{ lookAt {{ pos {r.uniform(-2, 2):.2f} {r.uniform(2, 5):.2f} {-r.uniform(7, 10):.2f} look 0 1 0 up 0 1 0 }} }}",
           f"camera {{ perspective fov {r.uniform(35, 60):.1f} lensRadius {r.choice([0, 0, 0.1])} focalDistance 9 }}", "newTransform { }"]
    if r.random() < 0.8:
        env = f"constant {_rgb(r, 'rgbI', 0.1, 1.0)}"
        # seeds from 100 on (a separate stream, so that the scenes of the first hundred seeds stay what they were): half of the
        # infinite lights are the Preetham sky of SunSky.hs with a random sun direction and turbidity
        r2 = np.random.default_rng(50_000 + seed)
        if seed >= 100 and r2.random() < 0.5:
            env = f"sunSky east {r2.uniform(-1, 1):.2f} 0 {r2.uniform(0.2, 1):.2f} sunDir {r2.uniform(-1, 1):.2f} {r2.uniform(0.05, 1):.2f} {r2.uniform(-1, 1):.2f} turbidity {r2.uniform(2, 12):.1f}"
        out.append(f"light {{ infinite {{ rotateX -90 }} l {{ {env} }} }}")
    if r.random() < 0.5: out.append(f"light {{ point intensity {_rgb(r, 'rgbI', 5, 40)} position {r.uniform(-4, 4):.2f} {r.uniform(3, 6):.2f} {r.uniform(-4, 4):.2f} }}")
    if r.random() < 0.5: out.append(f"light {{ directional intensity {_rgb(r, 'rgbI', 0.5, 3)} normal {r.uniform(-1, 1):.2f} 1 {r.uniform(-1, 1):.2f} }}")
    out += [_material(r), "newTransform { rotateX -90 }", "prim { shape { quad 9 9 } }"]                      # a ground
    for _ in range(int(r.integers(3, 8))):
        out.append(_material(r))
        emit = r.random() < 0.25
        if emit: out.append(f"emission {{ {_rgb(r, 'rgbI', 3, 25)} }}")
        out += [_transform(r), f"prim {{ shape {{ {_shape(r)} }} }}"]
        if emit: out.append("emission { none }")
    if r.random() < 0.7:                                                                                       # a small mesh
        vs = r.uniform(-1, 1, (5, 3))
        out += [_material(r), _transform(r), "prim { mesh vertexCount 5 faceCount 3 " + " ".join(f"v {a:.3f} {b:.3f} {c:.3f}" for a, b, c in vs) + " f 0 1 2 f 1 2 3 4 f 0 2 4 }"]
    return "\n".join(out) + "\n"


@pytest.mark.parametrize("seed", range(140))
def test_random_scene_bodies_match_oracle(seed, tmp_path):
    f = tmp_path / f"fuzz{seed}.bling"; f.write_text(random_scene_text(seed))
    try:
        sc = parse(f)
    except NotImplementedError as ex:            # e.g. shinyMetal over a computing texture: the loader says so, nothing to compare
        pytest.skip(str(ex))
    o = Oracle(sc); e = EmuContext(); e.upload_scene(sc)
    # parity (a) on random rays: primitive id exact except measured t-ties, t bit-exact, occlusion exact
    rays = random_rays(sc, 1200, seed)
    he, hb = e.trace_nearest(rays), o.trace_nearest(rays, mode="brute")
    ties, bad = compare_hits(he, hb)
    assert bad == 0 and ties <= 3, (ties, bad)
    same = (he["prim"] == hb["prim"]) & (hb["prim"] >= 0)
    assert np.array_equal(he["t"][same], hb["t"][same])
    assert np.array_equal(e.trace_occluded(rays).astype(bool), o.trace_occluded(rays, mode="brute").astype(bool))
    x0, x1, y0, y1 = o.sample_extent()
    rng = np.random.default_rng(1000 + seed)
    n = 700
    px, py, s = rng.integers(x0, x1 + 1, n), rng.integers(y0, y1 + 1, n), rng.integers(0, sc.spp, n)
    Lo, xyo = o.render_samples(1, 17 + seed, px, py, s)
    Le, xye = e.render_samples(1, 17 + seed, px, py, s)
    # the film too (every filter kind, tile clipping) -- where the scene is well conditioned: random glossy parameters let some
    # scenes reach |L| ~ 1e26 with both signs (the reference's sample / eval asymmetries, Q7), and then the ORDER of the film sum
    # (tiles in the oracle, per-pixel gather in the kernels) decides the digits
    o.render_pass(1, 3, threads=2); e.render_pass(1, 3)
    fo, fe = o.read_film(), e.read_film()
    # A NON-FINITE BSDF weight / MIS weight times a black `le` is NaN in the reference and the sample is dropped
    # (Image.hs:253-256): the kernels trace such BSDF-MIS rays instead of culling them and poison L where the throughput
    # is not finite, so the dropped samples agree exactly -- in every scene, also the ill-conditioned ones.
    assert o.stats()["dropped_samples"] == e.stats()["dropped_samples"]
    if np.isfinite(Lo).all() and np.abs(Lo).max() < 1e3:
        assert np.allclose(fo[..., 0], fe[..., 0], rtol=1e-4, atol=1e-6)
        assert np.array_equal(np.isfinite(fo), np.isfinite(fe))
        both = np.isfinite(fo).all(-1)
        assert both.mean() > 0.97 and np.allclose(fo[both], fe[both], rtol=5e-3, atol=1e-4 * max(1.0, float(np.abs(fo[both]).max())))
    e.close()
    assert np.array_equal(xyo, xye)
    ok = np.isfinite(Lo).all(1) & np.isfinite(Le).all(1)
    assert np.array_equal(np.isfinite(Lo).all(1), np.isfinite(Le).all(1))   # the same samples are lost to NaN / inf
    rel = np.abs(Lo[ok] - Le[ok]).max(1) / (np.abs(Lo[ok]).max(1) + 1e-6)
    # same libm on both sides: only the order of sums differs (measured over 400 random scenes x 2000 samples: max 6.3e-7)
    assert rel.max() < 1e-5, (seed, float(rel.max()))


def _anisotropic_shapes_text(seed, decades):
    """1 .. 8 analytic shapes under rotations about all three axes and per-axis scales over +-`decades` powers of ten"""
    r = np.random.default_rng(seed)
    out = ["filter box", "imageSize 16 12", "renderer { sampler sampled { sampler { stratified 1 1 } integrator { path maxDepth 2 sampleDepth 1 } } }",
           "transform { lookAt { pos 0 3 -9 look 0 1 0 up 0 1 0 } }", "camera { perspective fov 50 lensRadius 0 focalDistance 9 }", "newTransform { }",
           "light { point intensity rgbI 10 10 10 position 0 5 0 }", "material { matte kd { constant rgbR 0.5 0.5 0.5 } sigma { constant 0 } }"]
    for _ in range(int(r.integers(1, 9))):
        sx, sy, sz = (10.0 ** r.uniform(-decades, decades) for _ in range(3))
        spread = 10.0 ** r.uniform(-2, 3)
        out.append(f"newTransform {{ rotateX {r.uniform(-180, 180):.1f} rotateY {r.uniform(-180, 180):.1f} rotateZ {r.uniform(-180, 180):.1f} scale {sx:.6f} {sy:.6f} {sz:.6f} "
                   f"translate {r.uniform(-spread, spread):.5f} {r.uniform(-spread, spread):.5f} {r.uniform(-spread, spread):.5f} }}")
        out.append(f"prim {{ shape {{ {_shape(r)} }} }}")
    return "\n".join(out) + "\n"


@pytest.mark.parametrize("seed", range(12))
@pytest.mark.parametrize("decades", [1.0, 1.5])
def test_anisotropic_transforms_traversal_parity(seed, decades, tmp_path):
    """Shapes are intersected in object space (`transRay w2o`, Geometry.hs:26-27) but bounded by the eight corners of the object box
    under `o2w` (`transBox`, Transform.hs:281-292): where `w2o . o2w` is not the identity in f32, hits fall outside the world box
    and the REFERENCE's kd-tree loses them. Measured on 150 such scenes per setting: per-axis scales within +-1 decade -> kd-tree,
    BVH and brute force agree exactly; +-1.5 decades -> kd-tree and BVH lose the same hits (BVH == kd-tree exactly, 3 rays of
    450 000 differ from brute force); from +-2 decades on which of the lost hits is still found depends on the tree (32 rays of
    450 000 differ between the two). So: exact against brute force at 1, exact against the kd-tree restatement at 1.5."""
    f = tmp_path / "aniso.bling"; f.write_text(_anisotropic_shapes_text(seed, decades))
    sc = parse(f)
    rays = random_rays(sc, 3000, seed)
    o = Oracle(sc); e = EmuContext(); e.upload_scene(sc)
    want = o.trace_nearest(rays, mode="brute" if decades <= 1.0 else "kd")
    he = e.trace_nearest(rays); oe = e.trace_occluded(rays)
    oo = o.trace_occluded(rays, mode="brute" if decades <= 1.0 else "kd")
    e.close()
    ties, bad = compare_hits(he, want)
    assert bad == 0 and ties <= 3, (ties, bad)
    same = (he["prim"] == want["prim"]) & (want["prim"] >= 0)
    assert np.array_equal(he["t"][same], want["t"][same])
    assert np.array_equal(oe.astype(bool), oo.astype(bool))
