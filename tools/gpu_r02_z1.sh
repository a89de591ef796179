# Round 2, GPU call Z1: SAH-optimal collapse (bvh_build.cpp::Collapse, option bvh_collapse_cp) against the greedy collapse on cfg 5.
# tools/travsim.cpp says: same node visits, 23-46 % fewer primitive tests, node count -8 .. +25 % depending on cp.
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
run() {  # name, options...
  name=$1; shift
  opts=""; for o in "$@"; do opts="$opts --option $o"; done
  ( timeout -k 10 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-scenes $opts ) > gpurun_out/z1_bench_$name.json 2> gpurun_out/z1_bench_$name.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/z1_bench_$name.json").read().strip().splitlines()[-1])
    r, o = d["roofline"], d["roofline_other"]
    print("$name:", round(d["value"], 2), d["unit"], {k: round(x, 1) for k, x in r["kernel_ms_by_class"].items()}, "bvh", d["bvh"],
          "any n/p", r.get("nodes_per_ray"), r.get("prims_per_ray"), "near n/p", o.get("nodes_per_ray"), o.get("prims_per_ray"), "upload_s", d.get("e2e", {}).get("scene_upload_s"))
except Exception as e:
    print("$name: no line", e)
PY
}
run base
run leaf1 bvh_leaf=1
run cp04l2 bvh_collapse_cp=0.4 bvh_leaf=2
run cp05l3 bvh_collapse_cp=0.5 bvh_leaf=3
run cp06l3 bvh_collapse_cp=0.6 bvh_leaf=3
run cp07l2 bvh_collapse_cp=0.7 bvh_leaf=2
run cp10l1 bvh_collapse_cp=1.0 bvh_leaf=1
run base2
