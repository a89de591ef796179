# Round 2, GPU call Z2: the new tree (SAH-optimal collapse, cp 0.5 / leaves <= 3; child boxes on the 2^15 grid without the extra
# cell) -- the GPU suite, cfg 5 against the greedy collapse and neighbouring cp values, the named scenes both ways.
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout -k 10 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/z2_pytest_gpu.log 2>&1
tail -4 gpurun_out/z2_pytest_gpu.log
run() {  # name, options...
  name=$1; shift
  opts=""; for o in "$@"; do opts="$opts --option $o"; done
  ( timeout -k 10 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-scenes $opts ) > gpurun_out/z2_bench_$name.json 2> gpurun_out/z2_bench_$name.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/z2_bench_$name.json").read().strip().splitlines()[-1])
    r, o = d["roofline"], d["roofline_other"]
    print("$name:", round(d["value"], 2), d["unit"], {k: round(x, 1) for k, x in r["kernel_ms_by_class"].items()}, "bvh", d["bvh"],
          "any n/p", r.get("nodes_per_ray"), r.get("prims_per_ray"), "near n/p", o.get("nodes_per_ray"), o.get("prims_per_ray"))
except Exception as e:
    print("$name: no line", e)
PY
}
run default
run greedy bvh_collapse_cp=0 bvh_leaf=2
run cp045l3 bvh_collapse_cp=0.45
run cp05l4 bvh_leaf=4
run cp06l3 bvh_collapse_cp=0.6
run cp04l2 bvh_collapse_cp=0.4 bvh_leaf=2
( timeout -k 10 300 python tools/scene_breakdown.py ) > gpurun_out/z2_scenes_default.log 2>&1
cat gpurun_out/z2_scenes_default.log
( BLINGCU_OPTIONS="bvh_collapse_cp=0,bvh_leaf=2" timeout -k 10 300 python tools/scene_breakdown.py ) > gpurun_out/z2_scenes_greedy.log 2>&1
cat gpurun_out/z2_scenes_greedy.log
