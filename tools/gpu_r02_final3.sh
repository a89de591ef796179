# Round 2, the last single-GPU call: GPU suite, smoke() and the default bench line on the code as committed (shade kernels at the per-kind
# occupancy of profiles/r02_tree_and_requests.md section 8).
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout -k 10 1800 python -m pytest tests -m gpu -x -q ) > gpurun_out/i_pytest_gpu.log 2>&1
tail -4 gpurun_out/i_pytest_gpu.log
( timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" ) > gpurun_out/i_smoke.log 2>&1
tail -2 gpurun_out/i_smoke.log
( timeout -k 10 900 python bench.py ) > gpurun_out/i_bench_default.json 2> gpurun_out/i_bench_default.err
python - <<PY
import json
d = json.loads(open("gpurun_out/i_bench_default.json").read().strip().splitlines()[-1])
print(d["value"], d["unit"], "e2e", d["e2e"]["value"], "e2e_trace", d["e2e_trace"]["value"], d["e2e_trace"]["pageable"], {k: round(x, 1) for k, x in d["roofline"]["kernel_ms_by_class"].items()})
print("roofline", d["roofline"]["frac"], d["roofline_other"]["frac"], "cpu", d["cpu_baseline"]["value"], d["clocks"])
print({k: round(v["msamples_per_s"]) for k, v in d["scenes"].items()})
PY
bash tools/gpu_ncu_scenes.sh > gpurun_out/i_ncu_scenes.log 2>&1
