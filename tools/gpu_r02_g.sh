# Round 2, GPU call G: light tracer on the B200, the suite, the bench with the record layout of the path-state spectra
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout -k 10 600 python -m pytest tests/test_light_tracer.py -m gpu -x -q ) > gpurun_out/g_pytest_lt.log 2>&1
tail -15 gpurun_out/g_pytest_lt.log
( time timeout -k 10 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/g_pytest_gpu.log 2>&1
tail -8 gpurun_out/g_pytest_gpu.log
( timeout -k 10 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline ) > gpurun_out/g_bench.log 2> gpurun_out/g_bench.err
python - <<PY
import json
d = json.loads(open("gpurun_out/g_bench.log").read().strip().splitlines()[-1])
print(d["value"], d["unit"], "e2e", d["e2e"]["value"], {k: round(x, 1) for k, x in d["roofline"]["kernel_ms_by_class"].items()})
print({k: round(v["msamples_per_s"]) for k, v in d["scenes"].items()})
PY
python - <<PY
import sys, time
sys.path.insert(0, ".")
from bling_b200 import api
from bling_b200.host.loader import with_light_tracer_camera
from tests.conftest import load_scene
for name in ("cornell-box", "ducky"):
    sc = with_light_tracer_camera(load_scene(name))
    c = api.Context(0); c.upload_scene(sc)
    n = 4_000_000
    c.light_trace(1, 1, 0, n); c.synchronize(); c.reset_stats()
    t0 = time.perf_counter(); c.light_trace(2, 1, 0, n); c.synchronize(); dt = time.perf_counter() - t0
    st = c.stats()
    print(f"light tracer {name} {sc.width}x{sc.height}: {n / dt / 1e6:.1f} Mphotons/s, {(st['rays_light'] + st['rays_connect']) / dt / 1e6:.0f} Mrays/s, splats {st['splats']}, launches {st['kernel_launches']}")
    c.close()
PY
