# Round 2, GPU call N: ncu --set full with the source page of both product traversal kernels (after the stack restructure).
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout -k 10 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:kTraceWarpQ<\(bool\)1' -s 3 -c 1 -f -o /tmp/prof_n_any \
   python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-scenes > gpurun_out/n_ncu_any.log 2>&1
ncu -i /tmp/prof_n_any.ncu-rep --page raw --csv > gpurun_out/r02_any_n.raw.csv
ncu -i /tmp/prof_n_any.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/r02_any_n.source.csv.gz
timeout -k 10 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:kTraceWarpQ<\(bool\)0' -s 2 -c 1 -f -o /tmp/prof_n_near \
   python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-scenes > gpurun_out/n_ncu_near.log 2>&1
ncu -i /tmp/prof_n_near.ncu-rep --page raw --csv > gpurun_out/r02_near_n.raw.csv
ncu -i /tmp/prof_n_near.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/r02_near_n.source.csv.gz
ls -la gpurun_out | tail -8
