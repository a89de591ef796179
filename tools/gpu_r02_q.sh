# Round 2, GPU call Q: refill threshold 4 (product) vs 6, 8, 12 idle lanes.
# Build HERE first: g.build_variant('rf6', ['TR_REFILL=6']) ...
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for lib in libblingcu.so libblingcu_rf6.so libblingcu_rf8.so libblingcu_rf12.so libblingcu.so; do
  ( timeout -k 10 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-scenes --lib bling_b200/$lib ) > gpurun_out/q_bench_$lib.json 2> gpurun_out/q_bench_$lib.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/q_bench_$lib.json").read().strip().splitlines()[-1])
    print("$lib:", d["value"], d["unit"], {k: round(x, 1) for k, x in d["roofline"]["kernel_ms_by_class"].items()})
except Exception as e:
    print("$lib: no line", e)
PY
done
