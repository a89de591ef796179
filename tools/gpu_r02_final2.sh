# Round 2, last GPU call on the final code (tree of profiles/r02_tree_and_requests.md, wide spectrum / triangle records): the whole -m gpu
# suite, smoke(), the default bench line, the reference-arm line, a memcheck pass, and the ncu passes the committed profiles/r02_*
# summaries are regenerated from (tools/make_r02_profiles.py, tools/ncu_scene_table.py).
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout -k 10 1800 python -m pytest tests -m gpu -x -q ) > gpurun_out/g_pytest_gpu.log 2>&1
tail -6 gpurun_out/g_pytest_gpu.log
( timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" ) > gpurun_out/g_smoke.log 2>&1
tail -2 gpurun_out/g_smoke.log
( timeout -k 10 900 python bench.py ) > gpurun_out/g_bench_default.json 2> gpurun_out/g_bench_default.err
python - <<PY
import json
d = json.loads(open("gpurun_out/g_bench_default.json").read().strip().splitlines()[-1])
print(d["value"], d["unit"], "e2e", d["e2e"]["value"], "e2e_trace", d["e2e_trace"]["value"], d["e2e_trace"]["pageable"], {k: round(x, 1) for k, x in d["roofline"]["kernel_ms_by_class"].items()})
print("roofline", d["roofline"]["frac"], d["roofline_other"]["frac"], "cpu", d["cpu_baseline"]["value"], d["clocks"])
print({k: round(v["msamples_per_s"]) for k, v in d["scenes"].items()})
PY
( timeout -k 10 900 python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/g_bench_reference.json 2> gpurun_out/g_bench_reference.err
tail -c 400 gpurun_out/g_bench_reference.json
# launch list and DRAM traffic of one step
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -c 800 --csv --log-file gpurun_out/r02_launches.csv \
   python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-scenes > gpurun_out/g_ncu_launches.log 2>&1
timeout -k 10 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:kTrace --csv --log-file gpurun_out/r02_traffic.csv \
   python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --no-scenes > gpurun_out/r02_traffic.json 2> gpurun_out/g_ncu_traffic.err
tail -c 300 gpurun_out/r02_traffic.json
# the three launches every version of the traversal kernels was captured on (profiles/r02_trace_warpq.md)
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:kTrace -s 4 -c 3 -f -o /tmp/prof_final \
   python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-scenes > gpurun_out/g_ncu_full.log 2>&1
ncu -i /tmp/prof_final.ncu-rep --page raw --csv > gpurun_out/r02_trace_final2.raw.csv
bash tools/gpu_ncu_scenes.sh > gpurun_out/g_ncu_scenes.log 2>&1
SEL='(film_matches and (textures or direct or zoo or ducky)) or branch_tree or gpu_textures or nearest_and_any'
( timeout 600 compute-sanitizer --tool memcheck python -m pytest tests -m gpu -x -q -k "$SEL" ) > gpurun_out/g_memcheck.log 2>&1; tail -4 gpurun_out/g_memcheck.log
ls -la gpurun_out | tail -30
du -sh gpurun_out
