# Round 2, GPU call Z8: the vector reads of Z7 again with the occupancy confound removed -- material kinds whose shade kernels need
# 114-135 registers compiled for four CTAs per SM (SH_KINDS4=0x9F). vec4 = every table read vectorised; vecb = only the per-lane
# divergent ones (BxDF spectra, rgb -> spectrum basis), warp-uniform reads (light / environment spectra, CIE curves) scalar again.
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for lib in libblingcu_vecb.so libblingcu_vec4.so libblingcu.so; do
  ( timeout -k 10 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-scenes --lib bling_b200/$lib ) > gpurun_out/z8_bench_$lib.json 2> gpurun_out/z8_bench_$lib.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/z8_bench_$lib.json").read().strip().splitlines()[-1])
    print("$lib:", round(d["value"], 2), d["unit"], {k: round(x, 1) for k, x in d["roofline"]["kernel_ms_by_class"].items()})
except Exception as e:
    print("$lib: no line", e)
PY
done
python tools/ab_libs.py bling_b200/libblingcu_vecb.so bling_b200/libblingcu.so cornell-box glass-torus specular ducky sun-sky environment > gpurun_out/z8_ab_vecb.log 2>&1
cat gpurun_out/z8_ab_vecb.log
python tools/ab_libs.py bling_b200/libblingcu_vec4.so bling_b200/libblingcu.so cornell-box glass-torus specular ducky sun-sky environment > gpurun_out/z8_ab_vec4.log 2>&1
grep -v "libblingcu.so " gpurun_out/z8_ab_vec4.log
cp bling_b200/libblingcu_vecb.so bling_b200/libblingcu.so   # scratch copy on the box: the suite below runs on the new code
( timeout -k 10 600 python -m pytest tests -m gpu -x -q -k "film or samples or fuzz or direct or bidir or light" ) > gpurun_out/z8_pytest.log 2>&1
tail -3 gpurun_out/z8_pytest.log
