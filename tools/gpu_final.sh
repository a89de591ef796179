set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 200 python -m pytest tests -m gpu -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -4 gpurun_out/pytest_gpu.log
( timeout 100 python tools/ab_libs.py bling_b200/libblingcu.so bling_b200/libblingcu_fuse.so cornell-box ducky sun-sky ) > gpurun_out/ab_fuse.log 2>&1
cat gpurun_out/ab_fuse.log
