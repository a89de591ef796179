# Round 2, GPU call P: L1 prefetch of the next trip's node right after the pop (pf4: one sector, pf12: both).
# Build HERE first: g.build_variant('pf4', ['TQ_PREFETCH=4']); g.build_variant('pf12', ['TQ_PREFETCH=12'])
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for lib in libblingcu.so libblingcu_pf4.so libblingcu_pf12.so libblingcu.so; do
  ( timeout -k 10 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-scenes --lib bling_b200/$lib ) > gpurun_out/p_bench_$lib.json 2> gpurun_out/p_bench_$lib.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/p_bench_$lib.json").read().strip().splitlines()[-1])
    print("$lib:", d["value"], d["unit"], {k: round(x, 1) for k, x in d["roofline"]["kernel_ms_by_class"].items()})
except Exception as e:
    print("$lib: no line", e)
PY
done
