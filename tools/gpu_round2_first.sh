# First GPU call of round 2: confirm what round 1 could only check on the kernel-body emulator, then the A/B that is
# prepared but unmeasured. Before calling gpurun, build the variant HERE (it travels with the snapshot):
#   python -c "import __graft_entry__ as g; g.build(); g.build_variant('aos', ['BL_SPEC_AOS'])"
#   gpurun --timeout 1500 -- 'bash tools/gpu_round2_first.sh'
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > gpurun_out/smi.txt
# 1. the whole GPU suite: non-finite weights (test_zz_gpu_fuzz.py), nine-zero normals, Russian roulette, and everything of round 1
( time timeout 600 python -m pytest tests -m gpu -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -6 gpurun_out/pytest_gpu.log
( timeout 100 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
# 2. the headline bench: against round 1's 183 Msamples/s / 1118 Mrays/s (two more idle launches per bounce since then)
( time timeout 400 python bench.py ) > gpurun_out/bench_default.log 2> gpurun_out/bench_default.err
tail -1 gpurun_out/bench_default.log > gpurun_out/bench_default.json
# 3. launch list of the same command (share of each kernel; the idle nearest-hit MIS launches should be ~µs)
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 1 --warmup 1 --no-scenes --no-cpu-baseline --no-e2e > gpurun_out/ncu_bench.log 2>&1
# 4. path-state layout A/B (bodies.h::spec4At, BL_SPEC_AOS: one 64-byte record per slot instead of four float4 planes)
if [ -f bling_b200/libblingcu_aos.so ]; then
  ( timeout 300 python tools/ab_libs.py bling_b200/libblingcu.so bling_b200/libblingcu_aos.so cornell-box glass-torus ducky sun-sky environment ) > gpurun_out/ab_aos.log 2>&1
  cat gpurun_out/ab_aos.log
fi
