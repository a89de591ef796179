# Round 2, GPU call S: sampler key / sample index / depth of a slot as ONE 16-byte record (libblingcu.so) vs three arrays
# (libblingcu_head.so, the previous commit): named scenes per class, cfg 5, films must be identical.
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout -k 10 600 python tools/ab_libs.py bling_b200/libblingcu_head.so bling_b200/libblingcu.so ) > gpurun_out/s_ab.log 2>&1
cat gpurun_out/s_ab.log
for lib in libblingcu_head.so libblingcu.so libblingcu_head.so libblingcu.so; do
  ( timeout -k 10 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-scenes --lib bling_b200/$lib ) > gpurun_out/s_bench_$lib.json 2> gpurun_out/s_bench_$lib.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/s_bench_$lib.json").read().strip().splitlines()[-1])
    print("$lib:", d["value"], d["unit"], {k: round(x, 1) for k, x in d["roofline"]["kernel_ms_by_class"].items()})
except Exception as e:
    print("$lib: no line", e)
PY
done
( time timeout -k 10 1800 python -m pytest tests -m gpu -x -q ) > gpurun_out/s_pytest_gpu.log 2>&1
tail -6 gpurun_out/s_pytest_gpu.log
