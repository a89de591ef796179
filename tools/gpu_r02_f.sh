# Round 2, GPU call F: path-state layout A/B prepared in round 1 and never measured (bodies.h::spec4At, BL_SPEC_AOS: one 64-byte record
# per slot and spectrum instead of four float4 planes). Build HERE first: python -c "import __graft_entry__ as g; g.build(); g.build_variant('aos', ['BL_SPEC_AOS'])"
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout -k 10 600 python tools/ab_libs.py bling_b200/libblingcu.so bling_b200/libblingcu_aos.so cornell-box glass-torus specular ducky sun-sky environment ) > gpurun_out/f_ab_aos.log 2>&1
cat gpurun_out/f_ab_aos.log
for lib in libblingcu.so libblingcu_aos.so; do
  ( timeout -k 10 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-scenes --lib bling_b200/$lib ) > gpurun_out/f_bench_$lib.json 2> gpurun_out/f_bench_$lib.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/f_bench_$lib.json").read().strip().splitlines()[-1])
    print("$lib:", d["value"], d["unit"], {k: round(x, 1) for k, x in d["roofline"]["kernel_ms_by_class"].items()})
except Exception as e:
    print("$lib: no line", e)
PY
done
