set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 120 python -m pytest tests -m gpu -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -4 gpurun_out/pytest_gpu.log
( timeout 60 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-scenes ) > gpurun_out/bench_lite.log 2> gpurun_out/bench_lite.err
tail -1 gpurun_out/bench_lite.log | cut -c1-330
