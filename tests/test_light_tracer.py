"""the light tracer (Renderer/LightTracer.hs) on the kernel bodies (CPU emulator here, the B200 in the -m gpu leg)
against the oracle's restatement: every splat of every photon (pixel position, depth, XYZ), the splat buffer, the ray counts."""
import numpy as np
import pytest

from bling_b200 import api, ir as IR
from bling_b200.host.loader import with_light_tracer_camera
from oracle.oracle_py import Oracle
from tests.conftest import load_scene, small
from tests.emu.emu_py import EmuContext

SCENES = ["cornell-box", "zoo", "glass-torus", "sun-sky", "specular", "extras"]


def _scene(name, w=64, h=48):
    return with_light_tracer_camera(small(load_scene(name), w, h, 2, 2))


def _check(make_ctx, name, n=6000, tol=2e-5, exact_counts=True):
    sc = _scene(name)
    o = Oracle(sc, kdtree=False)
    want = o.light_trace(3, 99, 1000, n, records=True)
    c = make_ctx(); c.upload_scene(sc)
    got = c.light_trace_records(3, 99, 1000, n)
    so, sg = o.stats(), c.stats()
    assert len(want) > 50, (name, len(want))
    if exact_counts:
        assert len(got) == len(want), (name, len(got), len(want))
        assert np.array_equal(got[:, :2], want[:, :2])                               # (photon, depth) of every splat
        assert np.allclose(got[:, 2:4], want[:, 2:4], rtol=1e-5, atol=1e-3)          # raster position
        scale = np.abs(want[:, 4:]).max(1, keepdims=True) + 1e-30
        assert (np.abs(got[:, 4:] - want[:, 4:]) / scale).max() < tol, name          # XYZ
        for k in ("photons", "rays_light", "rays_connect", "splats"):
            assert sg[k] == so[k], (name, k, sg[k], so[k])
    else:   # CUDA libm moves a few paths across discontinuities: compare what the paths have in common, and the totals
        key = lambda r: {(int(a), int(b)): i for i, (a, b) in enumerate(r[:, :2])}
        kg, kw = key(got), key(want)
        common = sorted(set(kg) & set(kw))
        assert len(common) > 0.995 * max(len(kg), len(kw)), (name, len(common), len(kg), len(kw))
        g, w_ = got[[kg[k] for k in common]], want[[kw[k] for k in common]]
        scale = np.abs(w_[:, 4:]).max(1, keepdims=True) + 1e-30
        rel = (np.abs(g[:, 4:] - w_[:, 4:]) / scale).max(1)
        # measured on the B200 (tools/gpu_r02_g.sh, _h.sh): a path that libm nudges at one glossy vertex lands elsewhere at every
        # later one, so only the FIRST vertex of a light path is held record by record (worst scene 0.3 %: sun-sky, the sinf /
        # cosf / powf of the Perez model); over all depths up to 1.0 % of the records differ (extras: anisotropic microfacets),
        # and the splatted energy as a whole must still agree
        first = g[:, 1] == 0
        assert first.sum() > 100 and (rel[first] > 1e-3).mean() < 6e-3, (name, (rel[first] > 1e-3).mean())
        assert (rel > 1e-3).mean() < 3e-2, (name, (rel > 1e-3).mean())
        assert abs(got[:, 4:].sum() / want[:, 4:].sum() - 1) < 2e-2, (name, got[:, 4:].sum() / want[:, 4:].sum())
        assert abs(sg["rays_light"] / so["rays_light"] - 1) < 5e-3 and abs(sg["rays_connect"] / so["rays_connect"] - 1) < 5e-3
    fs, fo = c.read_splat(), o.read_splat()
    assert fs.shape == (sc.height, sc.width, 3) and fo.sum() > 0
    if exact_counts: assert np.abs(fs - fo).max() <= tol * np.abs(fo).max(), name
    else: assert abs(fs.sum() / fo.sum() - 1) < 2e-2, name      # single pixels can move with a nudged path; the image energy cannot
    # the filtered film stays empty (Image.hs:123-129: the light tracer only splats), clear_film clears the splats too
    assert c.read_film().max() == 0
    c.clear_film(); assert c.read_splat().max() == 0
    # photons [first, first + n) compose: two halves == the whole, bit for bit (deterministic splat order within a call,
    # and a pixel's sum over calls is in call order)
    c.light_trace(3, 99, 1000, n)
    whole = c.read_splat(); c.clear_film()
    c.light_trace(3, 99, 1000, n // 2); c.light_trace(3, 99, 1000 + n // 2, n - n // 2)
    halves = c.read_splat()
    assert np.abs(whole - halves).max() <= 1e-5 * np.abs(whole).max()
    c.clear_film(); c.light_trace(3, 99, 1000, n)
    assert np.array_equal(c.read_splat(), whole)                                     # same call, same bits
    c.close(); o.close()


@pytest.mark.parametrize("name", SCENES)
def test_emulated_light_tracer_matches_oracle(name):
    _check(EmuContext, name)


def test_light_tracer_needs_a_sampling_camera():
    sc = small(load_scene("cornell-box"), 32, 24, 2, 2)          # world2raster / pixel_area not set
    e = EmuContext(); e.upload_scene(sc)
    with pytest.raises(api.BlingCuError) as ex:
        e.light_trace(1, 1, 0, 10)
    assert ex.value.code == 1 and "sampleCam" in str(ex.value)
    env = small(load_scene("envcam"), 32, 16, 2, 2)
    with pytest.raises(ValueError):
        with_light_tracer_camera(env)                                # the reference `error`s for the environment camera too
    e.close()


def test_light_tracer_agrees_with_the_path_tracer_in_expectation():
    """two different estimators of the same image (an end-to-end check no shared misreading of a single function can fake): the
    splat buffer / photons must approach the path tracer's film where both estimate the same measurement -- the Cornell box,
    compared over the lower half of the image (floor and walls: never the lamp itself, where Q1 / Q2 make the two differ)"""
    base = small(load_scene("cornell-box"), 40, 30, 4, 4)
    sc = with_light_tracer_camera(base)
    o = Oracle(sc, kdtree=False)
    n = 400_000
    o.light_trace(1, 5, 0, n)
    lt = o.read_splat() / n                                           # splat weight 1 / (n * ppp) (LightTracer.hs:48)
    for p in range(1, 9):
        o.render_pass(p, 77, threads=8)
    film = o.read_film()
    from bling_b200 import image
    pt = image.film_xyz(film)
    # the Cornell quad light: NEE lights the room in both (Q2 makes the BSDF-sampled term vanish in the path tracer, and the
    # light tracer's emission side follows the sampled normal (0,0,-1) mapped to the world: the lit side), so the INDIRECT image
    # agrees; compare the mean luminance over the lower half of the image (floor and walls, never the lamp itself)
    a, b = lt[sc.height // 2:, :, 1].mean(), pt[sc.height // 2:, :, 1].mean()
    assert a > 0 and b > 0 and abs(a / b - 1) < 0.03, (a, b)
    o.close()


def test_light_tracer_renderer_reports_passes_with_splat_weights():
    """the Renderer instance (LightTracer.hs:39-51): PassDone n img (1 / (n * ppp)); getPixel's splat term makes the image"""
    from bling_b200 import image
    from bling_b200.renderer import LightTracerRenderer, PassDone, RenderJob
    sc = _scene("cornell-box", 32, 24)
    r = LightTracerRenderer(5000, seed=11, context_cls=EmuContext)
    seen = []
    r.render(RenderJob(sc), lambda p: (seen.append(p) or len(seen) < 3) if isinstance(p, PassDone) else True)
    assert [p.pass_num for p in seen] == [1, 2, 3] and [p.splat_weight for p in seen] == [1 / 5000, 1 / 10000, 1 / 15000]
    imgs = [image.image_xyz(*p.final_img, p.splat_weight) for p in seen]
    assert imgs[0][..., 1].mean() > 0 and abs(imgs[2][..., 1].mean() / imgs[0][..., 1].mean() - 1) < 0.2      # the estimate, not the sum, is reported
    assert seen[2].final_img[1].sum() > 2.5 * seen[0].final_img[1].sum()                                        # the splats accumulate
    r.close()


def test_loader_reads_the_light_renderer_block(tmp_path):
    """`renderer { light passPhotons n }` (IO/RendererParser.hs:28-30), last renderer block wins (:53-54): the stand-in loader hands
    the scene on with the camera fields sampleCam needs and `renderer_for` picks the light tracer"""
    from bling_b200.host.loader import load_scene
    from bling_b200.renderer import CudaRenderer, LightTracerRenderer, PassDone, RenderJob, renderer_for
    text = """
filter box
renderer { sampler sampled { sampler { stratified 2 2 } integrator { path maxDepth 5 sampleDepth 3 } } }
renderer { light passPhotons 700 }
imageSize 24 16
transform { lookAt { pos 0 3 -8 look 0 0 0 up 0 1 0 } }
camera { perspective fov 40 lensRadius 0 focalDistance 10 }
newTransform { }
material { matte kd { constant rgbR 0.6 0.6 0.6 } sigma { constant 0 } }
prim { mesh vertexCount 4 faceCount 1 v -3 0 -3 v -3 0 3 v 3 0 3 v 3 0 -3 f 0 1 2 3 }
light { point intensity rgbI 30 30 30 position 0 2 0 }
"""
    f = tmp_path / "lt.bling"; f.write_text(text)
    sc = load_scene(f)
    assert sc.pass_photons == 700 and sc.camera.pixel_area > 0 and any(abs(x) > 0 for x in sc.camera.world2raster)
    r = renderer_for(sc, context_cls=EmuContext)
    assert isinstance(r, LightTracerRenderer) and r.ppp == 700
    seen = []
    r.render(RenderJob(sc), lambda p: (seen.append(p) or False) if isinstance(p, PassDone) else True)
    assert seen and seen[0].splat_weight == 1 / 700
    r.close()
    f.write_text(text.replace("renderer { light passPhotons 700 }", ""))
    sc2 = load_scene(f)
    assert sc2.pass_photons == 0
    r2 = renderer_for(sc2, context_cls=EmuContext); assert isinstance(r2, CudaRenderer); r2.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name", SCENES)
def test_gpu_light_tracer_matches_oracle(name):
    _check(lambda: api.Context(0), name, n=20000, exact_counts=False)
