# Round 2, GPU call M: the stack inside the warp's region with an address as stack pointer, the slot in shared memory (no local
# memory reloads in the loop).
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout -k 10 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "variants_agree or soup_traversal or counters or any or deep or stack" ) > gpurun_out/m_pytest_new.log 2>&1
tail -5 gpurun_out/m_pytest_new.log
for i in 1 2; do
  ( timeout -k 10 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-scenes ) > gpurun_out/m_bench_$i.json 2> gpurun_out/m_bench_$i.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/m_bench_$i.json").read().strip().splitlines()[-1])
    print("run $i:", d["value"], d["unit"], {k: round(x, 1) for k, x in d["roofline"]["kernel_ms_by_class"].items()})
except Exception as e:
    print("run $i: no line", e)
PY
done
( timeout -k 10 300 python tools/scene_breakdown.py ) > gpurun_out/m_scenes.log 2>&1
cat gpurun_out/m_scenes.log
