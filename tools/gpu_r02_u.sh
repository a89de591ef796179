# Round 2, GPU call U (gpurun --gpus 2): the two-device tests and the N = 1 / N = 2 bench lines of the final code on the same box.
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout -k 10 600 python -m pytest tests -m gpu -x -q -k "two_devices or film_reduction or comm" ) > gpurun_out/u_pytest_2gpu.log 2>&1
tail -4 gpurun_out/u_pytest_2gpu.log
( timeout -k 10 600 python bench.py --steps 5 --warmup 3 --no-scenes --no-cpu-baseline ) > gpurun_out/u_bench_n1.json 2> gpurun_out/u_bench_n1.err
( timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 --no-scenes --no-cpu-baseline ) > gpurun_out/u_bench_n2.json 2> gpurun_out/u_bench_n2.err
( timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 ) > gpurun_out/u_bench_ref_n2.json 2> gpurun_out/u_bench_ref_n2.err
python - <<PY
import json
for f in ("u_bench_n1", "u_bench_n2", "u_bench_ref_n2"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d.get("n_gpus"), d.get("value"), d.get("unit"), "e2e", (d.get("e2e") or {}).get("value"), d.get("impl"))
    except Exception as e:
        print(f, "no line", e)
PY
