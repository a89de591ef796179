set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:ShadeHitBody<.int.0>" -s 2 -c 2 -f -o gpurun_out/prof_shade_cornell_r01 python tools/scene_breakdown.py cornell-box > gpurun_out/ncu_shade_cornell.log 2>&1
tail -3 gpurun_out/ncu_shade_cornell.log
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:ShadeHitBody<.int.0>" -s 1 -c 2 -f -o gpurun_out/prof_shade_cfg5_r01 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-scenes > gpurun_out/ncu_shade_cfg5.log 2>&1
tail -3 gpurun_out/ncu_shade_cfg5.log
ls -la gpurun_out
