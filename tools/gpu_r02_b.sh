# Round 2, GPU call B: ncu --set full of the traversal kernels, variant 1 (majority vote) vs variant 3 (warp-level leaf queue)
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for v in 1 3; do
  timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:kTrace -s 4 -c 4 -f -o gpurun_out/prof_trace_r02_v$v \
     python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-scenes --option trace_variant=$v > gpurun_out/b_ncu_v$v.log 2>&1
  tail -3 gpurun_out/b_ncu_v$v.log
done
ls -la gpurun_out/*.ncu-rep
