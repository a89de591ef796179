# Round 2, GPU call K: why do 12 % fewer node visits (any-hit, longest-overlap-first) buy only 1.2 % of kernel time? ncu --set full
# of the same any-hit launches with both builds (libblingcu_slot.so = first hit child in slot order), plus the A/B of the cheaper
# predicate form of the choice.
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout -k 10 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "variants_agree or soup_traversal or counters or any" ) > gpurun_out/k_pytest_new.log 2>&1
tail -5 gpurun_out/k_pytest_new.log
for lib in libblingcu.so libblingcu_slot.so libblingcu.so libblingcu_slot.so; do
  ( timeout -k 10 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-scenes --lib bling_b200/$lib ) > gpurun_out/k_bench_$lib.json 2> gpurun_out/k_bench_$lib.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/k_bench_$lib.json").read().strip().splitlines()[-1])
    print("$lib:", d["value"], d["unit"], {k: round(x, 1) for k, x in d["roofline"]["kernel_ms_by_class"].items()}, d["roofline"].get("nodes_per_ray"))
except Exception as e:
    print("$lib: no line", e)
PY
done
for lib in libblingcu.so libblingcu_slot.so; do
  timeout -k 10 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:kTraceWarpQ<\(bool\)1' -s 2 -c 2 -f -o /tmp/prof_k_$lib \
     python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-scenes --lib bling_b200/$lib > gpurun_out/k_ncu_$lib.log 2>&1
  tail -2 gpurun_out/k_ncu_$lib.log
  ncu -i /tmp/prof_k_$lib.ncu-rep --page raw --csv > gpurun_out/r02_any_$lib.raw.csv
  ncu -i /tmp/prof_k_$lib.ncu-rep --page source --csv --kernel-id :::1 2>/dev/null | gzip > gpurun_out/r02_any_$lib.source.csv.gz
done
# the nearest-hit kernel of the product build, for the per-instruction profile
timeout -k 10 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:kTraceWarpQ<\(bool\)0' -s 2 -c 1 -f -o /tmp/prof_k_near \
   python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-scenes > gpurun_out/k_ncu_near.log 2>&1
ncu -i /tmp/prof_k_near.ncu-rep --page raw --csv > gpurun_out/r02_near.raw.csv
ncu -i /tmp/prof_k_near.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/r02_near.source.csv.gz
ls -la gpurun_out
