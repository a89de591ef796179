#!/usr/bin/env python
"""Per-kernel-class device time of one timed slice of each named scene (BASELINE.json configs[0..3]) at config size.
A name may carry a size, e.g. `textures@1920x1080x8` = that fixture at 1920x1080 with an 8x8 stratified sampler."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from bling_b200 import api, ir as IR
from bling_b200.host.loader import resized

import os
if os.environ.get("BLINGCU_LIB"):      # A/B: another build of the library (__graft_entry__.build_variant)
    class _Ctx(api.Context):
        _lib_path = Path(os.environ["BLINGCU_LIB"])
    api.Context = _Ctx
names = sys.argv[1:] or ["cornell-box", "glass-torus", "specular", "ducky", "sun-sky", "environment"]
for name in names:
    name, _, size = name.partition("@")
    sc = IR.SceneIR.load(ROOT / "tests" / "golden" / "scenes" / f"{name}.npz")
    if size:
        w, h, n = (int(x) for x in size.split("x"))
        sc = resized(sc, w, h, n, n)
    c = api.Context(0)
    for kv in filter(None, os.environ.get("BLINGCU_OPTIONS", "").split(",")):      # A/B: options set before the upload (builder knobs)
        k_, _, v_ = kv.partition("="); c.set_option(k_, float(v_))
    c.upload_scene(sc)
    ex = c.sample_extent(); npx = (ex[1] - ex[0] + 1) * (ex[3] - ex[2] + 1)
    k = max(1, min(sc.spp // 2, int(48e6 // npx)))
    c.render_slice(1, 1, 0, k); c.synchronize(); c.reset_stats(); c.set_option("profile_kernels", 1)
    c.render_slice(1, 1, k, 2 * k); c.synchronize()
    kt = c.kernel_times(); st = c.stats()
    tot = sum(v[0] for v in kt.values())
    print(f"{name:12s} {st['samples'] / st['last_pass_ms'] / 1e3:7.1f} Msamples/s  pass {st['last_pass_ms']:7.1f} ms  launches {st['kernel_launches']:4d}  " +
          "  ".join(f"{k_}:{100 * v[0] / tot:4.1f}%" for k_, v in kt.items()), flush=True)
    c.close()
