# Round 2, GPU call Z3: cross-item prefetch of the hit-shading kernels (SH_PREFETCH, cuda_backend.cu::kRunQueueShade) against the
# plain loop (libblingcu_nopf.so), and the leaf-queue flush threshold / shared stack depth re-tuned on the new tree.
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for lib in libblingcu.so libblingcu_nopf.so libblingcu_f24.so libblingcu_f16.so libblingcu_ss12.so libblingcu.so; do
  ( timeout -k 10 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-scenes --lib bling_b200/$lib ) > gpurun_out/z3_bench_$lib.json 2> gpurun_out/z3_bench_$lib.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/z3_bench_$lib.json").read().strip().splitlines()[-1])
    print("$lib:", round(d["value"], 2), d["unit"], {k: round(x, 1) for k, x in d["roofline"]["kernel_ms_by_class"].items()})
except Exception as e:
    print("$lib: no line", e)
PY
done
for lib in libblingcu.so libblingcu_nopf.so; do
  ( BLINGCU_LIB=bling_b200/$lib timeout -k 10 300 python tools/scene_breakdown.py ) > gpurun_out/z3_scenes_$lib.log 2>&1
  cat gpurun_out/z3_scenes_$lib.log
done
( timeout -k 10 600 python -m pytest tests -m gpu -x -q -k "film or samples or fuzz or direct or bidir" ) > gpurun_out/z3_pytest.log 2>&1
tail -3 gpurun_out/z3_pytest.log
