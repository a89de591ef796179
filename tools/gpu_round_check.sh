set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > gpurun_out/smi.txt
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -5 gpurun_out/pytest_gpu.log
( time timeout 600 python bench.py ) > gpurun_out/bench_default.log 2> gpurun_out/bench_default.err
tail -1 gpurun_out/bench_default.log > gpurun_out/bench_default.json
( timeout 300 python tools/scene_breakdown.py textures@1920x1080x8 extras@1920x1080x8 zoo@1920x1080x8 ) > gpurun_out/breakdown_textures.log 2>&1
cat gpurun_out/breakdown_textures.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ShadeHitBodyILi9 -c 2 -f -o gpurun_out/prof_texshade_r01 python tools/scene_breakdown.py textures@1920x1080x8 > gpurun_out/ncu_texshade.log 2>&1
tail -3 gpurun_out/ncu_texshade.log
( timeout 200 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
