set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -8 gpurun_out/pytest_gpu.log
bash tools/gpu_ncu_texshade.sh
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:DlShadeBody" -c 2 -f -o gpurun_out/prof_dlshade_r01 python tools/scene_breakdown.py direct@1920x1080x8 > gpurun_out/ncu_dlshade.log 2>&1
tail -3 gpurun_out/ncu_dlshade.log
