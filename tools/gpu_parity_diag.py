#!/usr/bin/env python
"""Measures what the GPU parity tests tolerate, so that the tolerances in tests/test_gpu_parity.py are the
measured ones: per scene, the fraction of per-sample radiances that differ from the oracle by more than 1e-5 / 1e-3 relative
(CUDA libm vs glibc: sinf / cosf / powf / acosf / atan2f move a few paths across discontinuities), the ratio of the means with and
without clipping, and the any-hit flags that differ, classified by what the rays hit."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from bling_b200 import api  # noqa: E402
from oracle.oracle_py import Oracle  # noqa: E402
from tests.conftest import ALL_SCENES, camera_rays, load_scene, random_rays, small  # noqa: E402

ctx = api.Context(0)
for name in ALL_SCENES:
    sc0 = load_scene(name)
    sc = small(sc0, 64, 48, 4, 4)
    ctx.upload_scene(sc)
    o = Oracle(sc)
    x0, x1, y0, y1 = ctx.sample_extent()
    rng = np.random.default_rng(5)
    n = 20000
    px, py, s = rng.integers(x0, x1 + 1, n), rng.integers(y0, y1 + 1, n), rng.integers(0, 16, n)
    Lo, _ = o.render_samples(2, 77, px, py, s)
    Lg, _ = ctx.render_samples(2, 77, px, py, s)
    rel = np.abs(Lo - Lg).max(1) / (np.abs(Lo).max(1) + 1e-6)
    cap = np.percentile(Lo, 99.5)
    ctx.upload_scene(sc0)
    o2 = Oracle(sc0, kdtree=(name == "ducky")); mode = "kd" if name == "ducky" else "brute"
    rays = np.concatenate([random_rays(sc0, 20000, 3), camera_rays(None, sc0, 20000, 4)])
    a, b = ctx.trace_occluded(rays), o2.trace_occluded(rays, mode)
    bad = np.flatnonzero(a != b)
    shape_prims = {int(sh.prim_id) for sh in sc0.shapes}
    hg, ho = ctx.trace_nearest(rays[bad]), o2.trace_nearest(rays[bad], mode)
    on_shape = sum(1 for k in range(len(bad)) if int(hg["prim"][k]) in shape_prims or int(ho["prim"][k]) in shape_prims)
    print(f"{name:20s} samples rel>1e-5 {100 * (rel > 1e-5).mean():6.3f} %  rel>1e-3 {100 * (rel > 1e-3).mean():6.3f} %  mean ratio {Lg.mean() / Lo.mean():.5f} "
          f"clipped {np.minimum(Lg, cap).mean() / np.minimum(Lo, cap).mean():.5f} | any-hit flags differing {len(bad)} of {len(rays)}, on analytic shapes {on_shape}", flush=True)
    o.close(); o2.close()
ctx.close()
