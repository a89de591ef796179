# Round 2, GPU call Z10: the plastic / translucent-matte shade kernels compiled for four CTAs per SM (SH_KINDS4=0x88: 127 / 125 registers,
# no spills) against three (162 / 160 registers).
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python tools/ab_libs.py bling_b200/libblingcu_k4.so bling_b200/libblingcu.so ducky sun-sky environment specular > gpurun_out/z10_ab.log 2>&1
cat gpurun_out/z10_ab.log
cp bling_b200/libblingcu_k4.so bling_b200/libblingcu.so   # scratch copy on the box
( timeout -k 10 600 python -m pytest tests -m gpu -x -q -k "film or samples or fuzz" ) > gpurun_out/z10_pytest.log 2>&1
tail -3 gpurun_out/z10_pytest.log
