# Round 2, GPU call Z5: libblingcu_tri64.so = spectrum records as two 32-byte accesses (BL_SPEC_V8) AND the triangle shading geometry as one
# 64-byte record (shading.h::DScene::tri_p) against libblingcu.so = the product before both.
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for lib in libblingcu_tri64.so libblingcu.so; do
  ( timeout -k 10 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-scenes --lib bling_b200/$lib ) > gpurun_out/z5_bench_$lib.json 2> gpurun_out/z5_bench_$lib.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/z5_bench_$lib.json").read().strip().splitlines()[-1])
    print("$lib:", round(d["value"], 2), d["unit"], {k: round(x, 1) for k, x in d["roofline"]["kernel_ms_by_class"].items()})
except Exception as e:
    print("$lib: no line", e)
PY
done
python tools/ab_libs.py bling_b200/libblingcu_tri64.so bling_b200/libblingcu.so cornell-box glass-torus specular ducky sun-sky environment > gpurun_out/z5_ab.log 2>&1
cat gpurun_out/z5_ab.log
cp bling_b200/libblingcu_tri64.so bling_b200/libblingcu.so   # scratch copy on the box: the suite below runs on the new layout
( timeout -k 10 600 python -m pytest tests -m gpu -x -q -k "film or samples or fuzz or direct or bidir or light" ) > gpurun_out/z5_pytest.log 2>&1
tail -3 gpurun_out/z5_pytest.log
