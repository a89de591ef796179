# Round 2, GPU call H: after the FUSE template split of the any-hit kernel -- named scenes per class, the suite, the bench
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout -k 10 300 python tools/scene_breakdown.py ) > gpurun_out/h_scenes.log 2>&1
cat gpurun_out/h_scenes.log
( time timeout -k 10 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/h_pytest_gpu.log 2>&1
tail -6 gpurun_out/h_pytest_gpu.log
( timeout -k 10 600 python bench.py --steps 5 --warmup 3 ) > gpurun_out/h_bench.log 2> gpurun_out/h_bench.err
python - <<PY
import json
d = json.loads(open("gpurun_out/h_bench.log").read().strip().splitlines()[-1])
print(d["value"], d["unit"], "e2e", d["e2e"]["value"], {k: round(x, 1) for k, x in d["roofline"]["kernel_ms_by_class"].items()})
print({k: round(v["msamples_per_s"]) for k, v in d["scenes"].items()})
PY
