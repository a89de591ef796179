# Round 2, GPU call D: the default bench line (roofline against the L1 data pipe, counters of the product kernels, pipelined host
# batches), the launch list, DRAM traffic per ray over every traversal launch of one step, the per-scene ncu triple, the suite.
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout -k 10 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "pipelined or counters_of or film_reduction or variants_agree" ) > gpurun_out/d_pytest_new.log 2>&1
tail -5 gpurun_out/d_pytest_new.log
( time timeout -k 10 900 python bench.py ) > gpurun_out/d_bench_default.log 2> gpurun_out/d_bench_default.err
tail -1 gpurun_out/d_bench_default.log > gpurun_out/d_bench_default.json
tail -3 gpurun_out/d_bench_default.err
python - <<PY
import json
d = json.loads(open("gpurun_out/d_bench_default.json").read())
print(d["value"], d["unit"], "e2e", d["e2e"]["value"], "e2e_trace", d["e2e_trace"], "cpu", d["cpu_baseline"]["value"])
print({k: d["roofline"][k] for k in ("kernel", "bound", "achieved", "peak", "frac", "nodes_per_ray", "prims_per_ray", "mrays_per_s_in_kernel", "share_of_step")})
print({k: d["roofline_other"][k] for k in ("kernel", "bound", "achieved", "peak", "frac", "nodes_per_ray", "prims_per_ray", "mrays_per_s_in_kernel", "share_of_step")})
print({k: (round(v["msamples_per_s"]), round(v.get("cpu_baseline", {}).get("value", 0), 3)) for k, v in d["scenes"].items()})
PY
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 1 --warmup 1 --no-scenes --no-cpu-baseline --no-e2e > gpurun_out/d_ncu_launches.log 2>&1
timeout -k 10 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:kTrace -c 200 --csv --log-file gpurun_out/r02_traffic.csv python bench.py --steps 1 --warmup 0 --no-scenes --no-cpu-baseline --no-e2e > gpurun_out/d_ncu_traffic.log 2>&1
tail -1 gpurun_out/d_ncu_traffic.log > gpurun_out/r02_traffic_run.json
bash tools/gpu_ncu_scenes.sh > gpurun_out/d_scenes.log 2>&1
( time timeout -k 10 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/d_pytest_gpu.log 2>&1
tail -5 gpurun_out/d_pytest_gpu.log
ls -la gpurun_out | tail -30
