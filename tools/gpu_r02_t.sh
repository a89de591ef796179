# Round 2, GPU call T: L2 eviction policies on the traversal loads: leaf items evict_first (l2p1), nodes evict_last (l2p2), both (l2p3).
# Build HERE first: g.build_variant('l2p1', ['BL_L2_POLICY=1']) ...
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for lib in libblingcu.so libblingcu_l2p1.so libblingcu_l2p2.so libblingcu_l2p3.so libblingcu.so; do
  ( timeout -k 10 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-scenes --lib bling_b200/$lib ) > gpurun_out/t_bench_$lib.json 2> gpurun_out/t_bench_$lib.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/t_bench_$lib.json").read().strip().splitlines()[-1])
    print("$lib:", d["value"], d["unit"], {k: round(x, 1) for k, x in d["roofline"]["kernel_ms_by_class"].items()})
except Exception as e:
    print("$lib: no line", e)
PY
done
for lib in libblingcu.so libblingcu_l2p3.so; do
  ( BLINGCU_LIB=bling_b200/$lib timeout -k 10 300 python tools/scene_breakdown.py ) > gpurun_out/t_scenes_$lib.log 2>&1
  cat gpurun_out/t_scenes_$lib.log
done
