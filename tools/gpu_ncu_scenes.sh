set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
M=gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct,smsp__thread_inst_executed_per_inst_executed.ratio,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,l1tex__t_sector_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active
for s in cornell-box glass-torus specular ducky sun-sky environment; do
  timeout 400 ncu --metrics $M --clock-control none --kernel-name-base demangled -c 600 --csv --log-file gpurun_out/scene_metrics_$s.csv python tools/scene_breakdown.py $s > gpurun_out/scene_metrics_$s.log 2>&1
  tail -1 gpurun_out/scene_metrics_$s.log
done
ls -la gpurun_out | tail -8
