set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:ShadeHitBody<.int.16>" -c 3 -f -o gpurun_out/prof_texshade_r01 python tools/scene_breakdown.py textures@1920x1080x8 > gpurun_out/ncu_texshade.log 2>&1
tail -3 gpurun_out/ncu_texshade.log
ls -la gpurun_out
