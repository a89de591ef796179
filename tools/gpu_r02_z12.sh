# Round 2, GPU call Z12: ONLY the plastic shade kernel compiled for five CTAs per SM (SH_KINDS5=0x08: 96 registers, 112 bytes of spills).
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for lib in libblingcu.so; do
  ( timeout -k 10 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-scenes --lib bling_b200/$lib ) > gpurun_out/z12_bench_$lib.json 2> gpurun_out/z12_bench_$lib.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/z12_bench_$lib.json").read().strip().splitlines()[-1])
    print("$lib:", round(d["value"], 2), d["unit"], {k: round(x, 1) for k, x in d["roofline"]["kernel_ms_by_class"].items()})
except Exception as e:
    print("$lib: no line", e)
PY
done
python tools/ab_libs.py bling_b200/libblingcu_k5p.so bling_b200/libblingcu.so ducky sun-sky environment > gpurun_out/z12_ab.log 2>&1
cat gpurun_out/z12_ab.log
