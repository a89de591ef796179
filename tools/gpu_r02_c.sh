# Round 2, GPU call C: the hand-tuned warp-queue kernels (64-byte leaf items, no o/d per ray, shared-memory ray copies, PTX
# shared addressing). Build HERE first:
#   python -c "import __graft_entry__ as g; g.build(); g.build_variant('mb7', ['TR_MINBLOCKS=7']); g.build_variant('item48', ['BL_ITEM_F4=3'])"
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout -k 10 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "variants_agree or soup_traversal or state_follows or film_reduction" ) > gpurun_out/c_pytest_new.log 2>&1
tail -5 gpurun_out/c_pytest_new.log
( timeout -k 10 600 python tools/trace_bench.py --variants 1 3 --lib bling_b200/libblingcu.so bling_b200/libblingcu_mb7.so bling_b200/libblingcu_item48.so ) > gpurun_out/c_trace_bench.log 2>&1
cat gpurun_out/c_trace_bench.log
for leaf in 3 4; do
  ( timeout -k 10 300 python tools/trace_bench.py --variants 3 --option bvh_leaf=$leaf ) > gpurun_out/c_trace_bench_leaf$leaf.log 2>&1
  cat gpurun_out/c_trace_bench_leaf$leaf.log
done
for lib in libblingcu.so libblingcu_mb7.so; do
  ( timeout -k 10 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-scenes --lib bling_b200/$lib ) > gpurun_out/c_bench_$lib.json 2> gpurun_out/c_bench_$lib.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/c_bench_$lib.json").read().strip().splitlines()[-1])
    print("$lib:", d["value"], d["unit"], d.get("mrays_per_s"), {k: round(x, 1) for k, x in d["roofline"]["kernel_ms_by_class"].items()})
except Exception as e:
    print("$lib: no line", e)
PY
done
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:kTrace -s 4 -c 3 -f -o /tmp/prof_c \
   python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-scenes > gpurun_out/c_ncu.log 2>&1
tail -2 gpurun_out/c_ncu.log
ncu -i /tmp/prof_c.ncu-rep --page raw --csv > gpurun_out/r02_trace_c.raw.csv
ncu -i /tmp/prof_c.ncu-rep --page source --csv --kernel-id :::1 2>/dev/null | gzip > gpurun_out/r02_trace_c.source.csv.gz
( time timeout -k 10 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/c_pytest_gpu.log 2>&1
tail -5 gpurun_out/c_pytest_gpu.log
