# Round 2, GPU call Z9: L1 policies on the traversal loads -- leaf items bypass L1 (BL_L1_POLICY bit 0: LDG.E.NA), nodes evict_last (bit 1: LDG.E.EL).
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for lib in libblingcu.so libblingcu_l1p1.so libblingcu_l1p2.so libblingcu_l1p3.so; do
  ( timeout -k 10 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-scenes --lib bling_b200/$lib ) > gpurun_out/z9_bench_$lib.json 2> gpurun_out/z9_bench_$lib.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/z9_bench_$lib.json").read().strip().splitlines()[-1])
    print("$lib:", round(d["value"], 2), d["unit"], {k: round(x, 1) for k, x in d["roofline"]["kernel_ms_by_class"].items()})
except Exception as e:
    print("$lib: no line", e)
PY
done
for lib in libblingcu.so libblingcu_l1p3.so; do
  ( BLINGCU_LIB=bling_b200/$lib timeout -k 10 300 python tools/scene_breakdown.py cornell-box ducky ) 2>&1 | tail -2
done
