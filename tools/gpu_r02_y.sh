# Round 2, GPU call Y: the whole -m gpu suite, smoke() and the default bench line on the final tree (after the bidirectional integrator).
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout -k 10 1800 python -m pytest tests -m gpu -x -q ) > gpurun_out/y_pytest_gpu.log 2>&1
tail -6 gpurun_out/y_pytest_gpu.log
( timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" ) > gpurun_out/y_smoke.log 2>&1
tail -2 gpurun_out/y_smoke.log
( timeout -k 10 900 python bench.py ) > gpurun_out/y_bench_default.json 2> gpurun_out/y_bench_default.err
python - <<PY
import json
d = json.loads(open("gpurun_out/y_bench_default.json").read().strip().splitlines()[-1])
print(d["value"], d["unit"], "e2e", d["e2e"]["value"], "e2e_trace", d["e2e_trace"]["value"], d["e2e_trace"]["pageable"], {k: round(x, 1) for k, x in d["roofline"]["kernel_ms_by_class"].items()})
print("roofline", d["roofline"]["frac"], d["roofline"]["traffic"], d["roofline"].get("issue"), d["roofline_other"]["frac"], "cpu", d["cpu_baseline"]["value"], d["clocks"])
print({k: round(v["msamples_per_s"]) for k, v in d["scenes"].items()})
PY
( timeout -k 10 900 python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/y_bench_reference.json 2> gpurun_out/y_bench_reference.err
tail -c 300 gpurun_out/y_bench_reference.json
