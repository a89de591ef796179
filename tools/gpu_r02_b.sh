# Round 2, GPU call B: ncu --set full of the traversal kernels, variant 1 (majority vote) vs variant 3 (warp-level leaf queue).
# The .ncu-rep files are converted to CSV on the box (raw page; source page of the nearest-hit kernel) and removed: two full
# reports exceed the 64 MiB that come back.
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for v in 1 3; do
  timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:kTrace -s 4 -c 3 -f -o /tmp/prof_trace_r02_v$v \
     python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-scenes --option trace_variant=$v > gpurun_out/b_ncu_v$v.log 2>&1
  tail -3 gpurun_out/b_ncu_v$v.log
  ncu -i /tmp/prof_trace_r02_v$v.ncu-rep --page raw --csv > gpurun_out/r02_trace_v$v.raw.csv
  ncu -i /tmp/prof_trace_r02_v$v.ncu-rep --page source --csv --kernel-id :::1 2>/dev/null | gzip > gpurun_out/r02_trace_v$v.source.csv.gz
  ls -la /tmp/prof_trace_r02_v$v.ncu-rep
done
ls -la gpurun_out/
