#!/usr/bin/env python
"""Converged ORACLE renders (>= 4096 spp) of every config scene at reduced resolution, committed under
tests/golden/renders/ and used by tests/test_gpu_parity.py::test_converged_render_vs_golden (parity check (b)).
NOTE: these are outputs of the CPU restatement, not of GHC-built bling (absent: parity unpinned).
The rel-MSE bound is 3x the rel-MSE between two independent 2048-spp halves of the oracle render (the noise floor).

    python tools/make_golden_renders.py [scene ...]
    python tools/make_golden_renders.py --quarter [scene ...]     # SURVEY §8(d) "converged parity": 1/4 linear CONFIG size, 4096 spp
                                                                  # -> tests/golden/renders_quarter/ (hours of CPU: run it niced)
"""
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from bling_b200 import image, ir as IR  # noqa: E402
from bling_b200.host.loader import resized  # noqa: E402
from oracle.oracle_py import Oracle  # noqa: E402

W, H, NU, NV, PASSES, SEED = 80, 60, 4, 4, 256, 0xB11D6
OUT = ROOT / "tests" / "golden" / "renders"
NOISY = {"blackbody-emission"}   # fixtures whose per-channel mean carries a measured tolerance instead of 0.5 %


def rel_mse(a, b):
    return float(np.mean((a - b) ** 2 / (b ** 2 + 1e-3)))


QUARTER = {"cornell-box": (128, 128), "glass-torus": (256, 256), "specular": (256, 256), "ducky": (480, 270), "sun-sky": (480, 270),
           "environment": (480, 270)}   # BASELINE.json configs[0..3] at 1/4 linear size


def main(names, quarter=False):
    global W, H, NU, NV, PASSES, OUT
    if quarter:
        OUT = ROOT / "tests" / "golden" / "renders_quarter"; NU, NV, PASSES = 8, 8, 64
    OUT.mkdir(parents=True, exist_ok=True)
    for name in names:
        if quarter: W, H = QUARTER[name]
        sc = resized(IR.SceneIR.load(ROOT / "tests" / "golden" / "scenes" / f"{name}.npz"), W, H, NU, NV)
        t = time.time()
        halves = []
        for half in range(2):
            o = Oracle(sc)
            for p in range(1 + half * PASSES // 2, 1 + (half + 1) * PASSES // 2):
                o.render_pass(p, SEED, threads=os.cpu_count())
            halves.append(o.read_film())
        full = halves[0] + halves[1]
        xa, xb, xf = (image.film_xyz(f) for f in (halves[0], halves[1], full))
        floor = rel_mse(xa, xb)
        extra = {}
        if name in NOISY:
            # the image MEAN of this scene does not converge to 0.5 % at 4096 spp (a few tiny, very bright emitters): measure
            # the standard deviation of the mean from 8 independent 512-spp renders and let the test allow 4 sigma of the
            # difference of two 4096-spp means instead
            ms = []
            for k in range(8):
                o = Oracle(sc)
                for p in range(1, 1 + PASSES // 8): o.render_pass(p, SEED + 1000 + k, threads=os.cpu_count())
                ms.append(image.film_xyz(o.read_film()).mean((0, 1)))
            ms = np.array(ms); sigma_full = (ms.std(0, ddof=1) / ms.mean(0)).max() / np.sqrt(8.0)
            extra["mean_tol"] = max(5e-3, 4 * np.sqrt(2.0) * float(sigma_full))
        np.savez_compressed(OUT / f"{name}.npz", xyz=xf.astype(np.float32), cfg=np.array([W, H, NU, NV, PASSES]), relmse_halves=floor,
                            relmse_bound=3 * floor, seed=SEED, **extra)
        print(f"{name}: {time.time() - t:.0f}s, rel-MSE between 2048-spp halves {floor:.3e}, mean XYZ {xf.mean((0, 1))} {extra}", flush=True)


if __name__ == "__main__":
    args = [x for x in sys.argv[1:] if x != "--quarter"]
    main(args or ["cornell-box", "glass-torus", "specular", "ducky", "sun-sky", "environment"], quarter="--quarter" in sys.argv)
