# Round 2, GPU call V: the bidirectional integrator (bidir.h) against the oracle, and what it renders per second.
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout -k 10 900 python -m pytest tests/test_bidir.py -m gpu -x -q ) > gpurun_out/v_pytest_bidir.log 2>&1
tail -15 gpurun_out/v_pytest_bidir.log
( timeout -k 10 600 python - <<'PY'
import copy, time, numpy as np
from bling_b200 import api, ir as IR
from tests.conftest import load_scene
for name in ("cornell-box", "ducky", "sun-sky"):
    sc = copy.copy(load_scene(name)); sc.integrator_kind = IR.INTEGRATOR_BIDIR; sc.max_depth = 5; sc.sample_depth = 3
    c = api.Context(0); c.upload_scene(sc)
    c.render_slice(1, 7, 0, 1); c.read_film()                      # warm-up
    c.reset_stats(); t0 = time.perf_counter()
    c.render_slice(2, 7, 0, min(sc.spp, 4)); f = c.read_film(); dt = time.perf_counter() - t0
    st = c.stats(); c.close()
    rays = st["rays_camera"] + st["rays_extension"] + st["rays_mis"] + st["rays_shadow"]
    print(f"{name:12s} {sc.width}x{sc.height} bidir 5/3: {st['samples'] / dt / 1e6:7.2f} Msamples/s, {rays / dt / 1e6:8.1f} Mrays/s, {st['kernel_launches']} launches, {rays / st['samples']:.1f} rays/sample, finite film {np.isfinite(f).all()}")
PY
) > gpurun_out/v_bidir_rate.log 2>&1
cat gpurun_out/v_bidir_rate.log
