# Round 2, GPU call Z7: libblingcu_vec.so = table spectra (BxDF reflectance / eta / k, rgb -> spectrum basis, CIE curves) read four bands per 16-byte
# load (hd.h::specQuarter) against libblingcu.so = band-by-band scalar loads (70 % of the L1 requests of the matte shade kernel: Z6).
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for lib in libblingcu_vec.so libblingcu.so; do
  ( timeout -k 10 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-scenes --lib bling_b200/$lib ) > gpurun_out/z7_bench_$lib.json 2> gpurun_out/z7_bench_$lib.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/z7_bench_$lib.json").read().strip().splitlines()[-1])
    print("$lib:", round(d["value"], 2), d["unit"], {k: round(x, 1) for k, x in d["roofline"]["kernel_ms_by_class"].items()})
except Exception as e:
    print("$lib: no line", e)
PY
done
python tools/ab_libs.py bling_b200/libblingcu_vec.so bling_b200/libblingcu.so cornell-box glass-torus specular ducky sun-sky environment > gpurun_out/z7_ab.log 2>&1
cat gpurun_out/z7_ab.log
cp bling_b200/libblingcu_vec.so bling_b200/libblingcu.so   # scratch copy on the box: the suite below runs on the new layout
( timeout -k 10 600 python -m pytest tests -m gpu -x -q -k "film or samples or fuzz or direct or bidir or light" ) > gpurun_out/z7_pytest.log 2>&1
tail -3 gpurun_out/z7_pytest.log
