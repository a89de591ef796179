# Round 2, GPU call X (gpurun --gpus 8): the library's NCCL film reduction at 4 and 8 ranks (torchrun, one process per GPU), and
# one process driving 8 contexts (comm_init_all / reduce_film_group: the Haskell host's shape).
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L | wc -l
for n in 8 4; do
  ( timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 3 --warmup 3 --no-scenes --no-cpu-baseline ) > gpurun_out/x_bench_n$n.json 2> gpurun_out/x_bench_n$n.err
  tail -3 gpurun_out/x_bench_n$n.err
done
( timeout -k 10 300 python - <<'PY'
import numpy as np, time
from bling_b200 import api
from bling_b200.renderer import MultiDeviceRenderer, RenderJob, PassDone
from tests.conftest import load_scene, small
sc = small(load_scene("cornell-box"), 256, 256, 4, 4)
r = MultiDeviceRenderer(list(range(8)), seed=5)
seen = []
r.render(RenderJob(sc), lambda p: (seen.append(p) or len(seen) < 2) if isinstance(p, PassDone) else True)
one = api.Context(0); one.upload_scene(sc); one.render_pass(1, 5); ref = one.read_film(); one.close()
f = seen[0].final_img
print("8 contexts in one process: film vs one device max rel diff", float(np.abs(f - ref).max() / np.abs(ref).max()))
r.close()
PY
) > gpurun_out/x_multi8.log 2>&1
tail -3 gpurun_out/x_multi8.log
python - <<PY
import json
for n in (4, 8):
    try:
        d = json.loads(open(f"gpurun_out/x_bench_n{n}.json").read().strip().splitlines()[-1])
        print(n, d["value"], d["unit"], "e2e", d["e2e"]["value"], "ms/step", d["ms_per_step"])
    except Exception as e:
        print(n, "no line", e)
PY
