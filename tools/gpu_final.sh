cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 28 python -m pytest tests -m gpu -q -x -k "fused_and_separate or (film_matches and (cornell or zoo or textures or direct))" > gpurun_out/pytest_gpu_switch.log 2>&1
tail -3 gpurun_out/pytest_gpu_switch.log
