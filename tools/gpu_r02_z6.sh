# Round 2, GPU call Z6: ncu source-level capture of the hit-shading kernels (which instructions make the requests), cfg 5 and two named scenes.
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout -k 10 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:ShadeHitBody -s 1 -c 1 -f -o /tmp/prof_shade5 \
   python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --no-scenes > gpurun_out/z6_ncu5.log 2>&1
ncu -i /tmp/prof_shade5.ncu-rep --page raw --csv > gpurun_out/z6_shade_cfg5.raw.csv
ncu -i /tmp/prof_shade5.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/z6_shade_cfg5.source.csv.gz
for sc in sun-sky environment; do
  timeout -k 10 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:ShadeHitBody -s 3 -c 1 -f -o /tmp/prof_$sc \
     python tools/scene_breakdown.py $sc > gpurun_out/z6_ncu_$sc.log 2>&1
  ncu -i /tmp/prof_$sc.ncu-rep --page raw --csv > gpurun_out/z6_shade_$sc.raw.csv
  ncu -i /tmp/prof_$sc.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/z6_shade_$sc.source.csv.gz
done
ls -la gpurun_out | grep z6
