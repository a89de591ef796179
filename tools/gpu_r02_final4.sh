# Round 2: the bench's per-scene table once more (a one-shot 653 Msamples/s for sun-sky in gpu_r02_final3.sh against 921-947 in every A/B run).
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout -k 10 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e ) > gpurun_out/j_bench.json 2> gpurun_out/j_bench.err
python - <<PY
import json
d = json.loads(open("gpurun_out/j_bench.json").read().strip().splitlines()[-1])
print(d["value"], {k: round(v["msamples_per_s"]) for k, v in d["scenes"].items()})
PY
python tools/scene_breakdown.py sun-sky sun-sky 2>&1 | cut -c1-120
