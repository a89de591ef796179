# Round 2, GPU call W: compute-sanitizer over the kernels this round wrote or rewrote: the warp-queue traversal kernels (shared-memory
# stack, ring, per-lane records: memcheck + racecheck + synccheck), the light tracer's splat sort, the bidirectional integrator, the
# kd-tree walk and the pipelined host batches.
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
SEL='(film_matches and (zoo or cornell)) or variants_agree or reference_kdtree or (light_tracer and (cornell or zoo)) or (bidir and (zoo or film)) or pipelined'
( timeout -k 10 900 compute-sanitizer --tool memcheck python -m pytest tests -m gpu -x -q -k "$SEL" ) > gpurun_out/w_memcheck.log 2>&1; tail -5 gpurun_out/w_memcheck.log
( timeout -k 10 900 compute-sanitizer --tool racecheck python -m pytest tests -m gpu -x -q -k "(film_matches and zoo) or variants_agree or (bidir and film)" ) > gpurun_out/w_racecheck.log 2>&1; tail -5 gpurun_out/w_racecheck.log
( timeout -k 10 600 compute-sanitizer --tool synccheck python -m pytest tests -m gpu -x -q -k "(film_matches and zoo) or variants_agree" ) > gpurun_out/w_synccheck.log 2>&1; tail -5 gpurun_out/w_synccheck.log
( timeout -k 10 600 compute-sanitizer --tool initcheck python -m pytest tests -m gpu -x -q -k "(film_matches and zoo) or (bidir and film) or (light_tracer and cornell)" ) > gpurun_out/w_initcheck.log 2>&1; tail -5 gpurun_out/w_initcheck.log
grep -c "ERROR SUMMARY" gpurun_out/w_*.log
grep -h "ERROR SUMMARY\|RACECHECK SUMMARY" gpurun_out/w_*.log | sort | uniq -c
