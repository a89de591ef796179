#!/usr/bin/env python
"""A/B of two builds of the library on named scenes at config size: pass time, per-kernel-class time and film equality.
    python tools/ab_libs.py bling_b200/libblingcu.so bling_b200/libblingcu_fuse.so [scene ...]
(the second library comes from __graft_entry__.build_variant(name, defines))."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from bling_b200 import api, ir as IR  # noqa: E402

libs = [Path(sys.argv[1]), Path(sys.argv[2])]
names = sys.argv[3:] or ["cornell-box", "ducky", "sun-sky"]
for name in names:
    sc = IR.SceneIR.load(ROOT / "tests" / "golden" / "scenes" / f"{name}.npz")
    films, rows = [], []
    for lib in libs:
        C = type("Ctx", (api.Context,), {"_lib_path": lib})
        c = C(0); c.upload_scene(sc)
        ex = c.sample_extent(); npx = (ex[1] - ex[0] + 1) * (ex[3] - ex[2] + 1)
        k = max(1, min(sc.spp // 2, int(48e6 // npx)))
        c.render_slice(1, 1, 0, k); c.synchronize(); c.reset_stats(); c.set_option("profile_kernels", 1)
        c.render_slice(1, 1, k, 2 * k); c.synchronize()
        kt = c.kernel_times(); st = c.stats()
        films.append(c.read_film())
        rows.append(f"  {lib.name:24s} {st['samples'] / st['last_pass_ms'] / 1e3:7.1f} Msamples/s  pass {st['last_pass_ms']:7.2f} ms  launches {st['kernel_launches']:4d}  " +
                    "  ".join(f"{k_}:{v[0]:.2f}" for k_, v in kt.items()))
        c.close()
    d = float(np.abs(films[0] - films[1]).max())
    print(name, "| max film difference between the two builds:", d, flush=True)
    for r in rows: print(r, flush=True)
