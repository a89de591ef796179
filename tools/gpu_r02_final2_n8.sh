# Round 2, last multi-GPU call (gpurun --gpus 8) on the final code: bench lines at N = 8, 4, 2 through the library's NCCL film reduction,
# and the two-device test of the GPU suite.
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L | wc -l
for n in 8 4 2; do
  ( timeout -k 10 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 3 --warmup 3 --no-scenes --no-cpu-baseline ) > gpurun_out/h_bench_n$n.json 2> gpurun_out/h_bench_n$n.err
  tail -2 gpurun_out/h_bench_n$n.err
done
( timeout -k 10 200 python bench.py --steps 3 --warmup 3 --no-scenes --no-cpu-baseline ) > gpurun_out/h_bench_n1.json 2> gpurun_out/h_bench_n1.err
( timeout -k 10 300 python -m pytest tests -m gpu -x -q -k "two_devices or comm or reduce" ) > gpurun_out/h_pytest_2gpu.log 2>&1
tail -3 gpurun_out/h_pytest_2gpu.log
python - <<PY
import json
for n in (1, 2, 4, 8):
    try:
        d = json.loads(open(f"gpurun_out/h_bench_n{n}.json").read().strip().splitlines()[-1])
        print(n, round(d["value"], 1), d["unit"], "e2e", round(d["e2e"]["value"], 1), "ms/step", round(d["ms_per_step"], 1))
    except Exception as e:
        print(n, "no line", e)
PY
