# Round 2, GPU call J: any-hit child order. libblingcu.so enters the child the ray stays in longest, libblingcu_slot.so (the
# previous commit, built from a stash) the first hit child in slot order.
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout -k 10 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "variants_agree or soup_traversal or counters or any" ) > gpurun_out/j_pytest_new.log 2>&1
tail -5 gpurun_out/j_pytest_new.log
for lib in libblingcu.so libblingcu_slot.so libblingcu.so libblingcu_slot.so; do
  ( timeout -k 10 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-scenes --lib bling_b200/$lib ) > gpurun_out/j_bench_$lib.json 2> gpurun_out/j_bench_$lib.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/j_bench_$lib.json").read().strip().splitlines()[-1])
    print("$lib:", d["value"], d["unit"], {k: round(x, 1) for k, x in d["roofline"]["kernel_ms_by_class"].items()}, d.get("stats", {}))
except Exception as e:
    print("$lib: no line", e)
PY
done
for lib in libblingcu.so libblingcu_slot.so; do
  ( BLINGCU_LIB=bling_b200/$lib timeout -k 10 300 python tools/scene_breakdown.py ) > gpurun_out/j_scenes_$lib.log 2>&1
  cat gpurun_out/j_scenes_$lib.log
done
