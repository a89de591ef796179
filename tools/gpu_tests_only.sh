set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 300 python -m pytest tests -m gpu -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -6 gpurun_out/pytest_gpu.log
( timeout 120 python tools/scene_breakdown.py textures@1920x1080x8 direct@1920x1080x8 extras@1920x1080x8 ) > gpurun_out/breakdown_textures.log 2>&1
cat gpurun_out/breakdown_textures.log
