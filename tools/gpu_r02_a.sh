# Round 2, GPU call A: does the warp-level leaf queue (trace_variant 2 / 3, trace_warpq.cuh) give the same hits, and what does
# it buy? Build HERE first:  python -c "import __graft_entry__ as g; g.build(); g.build_variant('refill2', ['TR_REFILL=2']); g.build_variant('refill1f24', ['TR_REFILL=1','TQ_FLUSH=24'])"
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > gpurun_out/smi.txt
# 1. the new kernels first, under a short timeout (a hang must not take the box)
( time timeout -k 10 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "variants_agree or soup_traversal or state_follows" ) > gpurun_out/a_pytest_new.log 2>&1
tail -15 gpurun_out/a_pytest_new.log
# 2. kernel-only A/B on the 10 M soup (4 M bounce-like + 4 M primary rays)
( timeout -k 10 600 python tools/trace_bench.py --variants 1 2 3 --lib bling_b200/libblingcu.so bling_b200/libblingcu_refill2.so bling_b200/libblingcu_refill1f24.so ) > gpurun_out/a_trace_bench.log 2>&1
cat gpurun_out/a_trace_bench.log
# 3. the whole pipeline on cfg 5
for v in 1 2 3; do
  ( timeout -k 10 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-scenes --option trace_variant=$v ) > gpurun_out/a_bench_v$v.json 2> gpurun_out/a_bench_v$v.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/a_bench_v$v.json").read().strip().splitlines()[-1])
    print("variant $v:", d["value"], d["unit"], d.get("mrays_per_s"), {k: round(x, 1) for k, x in d["roofline"]["kernel_ms_by_class"].items()})
except Exception as e:
    print("variant $v: no line", e)
PY
done
# 4. the whole GPU suite
( time timeout -k 10 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/a_pytest_gpu.log 2>&1
tail -8 gpurun_out/a_pytest_gpu.log
