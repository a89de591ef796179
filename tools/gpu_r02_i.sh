# Round 2, GPU call I: rays as one 32-byte record; nearest-hit kernel compiled for 7 CTAs per SM (72 registers) vs 8.
# Build HERE first: python -c "import __graft_entry__ as g; g.build(); g.build_variant('near7', ['TQ_NEAR_BLOCKS=7'])"
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout -k 10 600 python -m pytest tests/test_gpu_parity.py tests/test_light_tracer.py -m gpu -x -q -k "variants_agree or soup_traversal or pipelined or light_tracer or kdtree" ) > gpurun_out/i_pytest_new.log 2>&1
tail -5 gpurun_out/i_pytest_new.log
( timeout -k 10 300 python tools/scene_breakdown.py ) > gpurun_out/i_scenes.log 2>&1
cat gpurun_out/i_scenes.log
for lib in libblingcu.so libblingcu_near7.so; do
  ( timeout -k 10 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-scenes --lib bling_b200/$lib ) > gpurun_out/i_bench_$lib.json 2> gpurun_out/i_bench_$lib.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/i_bench_$lib.json").read().strip().splitlines()[-1])
    print("$lib:", d["value"], d["unit"], {k: round(x, 1) for k, x in d["roofline"]["kernel_ms_by_class"].items()})
except Exception as e:
    print("$lib: no line", e)
PY
done
( time timeout -k 10 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/i_pytest_gpu.log 2>&1
tail -6 gpurun_out/i_pytest_gpu.log
