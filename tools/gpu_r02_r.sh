# Round 2, GPU call R: upper bound of ray binning: the same 4 M incoherent rays in random order and sorted by origin cell / octant.
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout -k 10 600 python tools/trace_bench.py --variants 3 --sorted ) > gpurun_out/r_trace_bench.log 2>&1
cat gpurun_out/r_trace_bench.log
( timeout -k 10 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-scenes ) > gpurun_out/r_bench.json 2> gpurun_out/r_bench.err
tail -c 600 gpurun_out/r_bench.json
