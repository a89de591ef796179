# Round 2, GPU call E (gpurun --gpus 2): the library's own NCCL film reduction -- one process driving two contexts
# (comm_init_all / reduce_film_group) and one process per GPU under torchrun (comm_unique_id / comm_init / reduce_film) --
# then the parity diagnostics and the quarter-config-size converged renders on one GPU.
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L
( time timeout -k 10 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "two_devices or film_reduction" ) > gpurun_out/e_pytest_comm.log 2>&1
tail -5 gpurun_out/e_pytest_comm.log
( timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 --no-scenes ) > gpurun_out/e_bench_n2.log 2> gpurun_out/e_bench_n2.err
tail -1 gpurun_out/e_bench_n2.log > gpurun_out/e_bench_n2.json
tail -5 gpurun_out/e_bench_n2.err
python - <<PY
import json
d = json.loads(open("gpurun_out/e_bench_n2.json").read())
print("N=2:", d["value"], d["unit"], "e2e", d["e2e"]["value"], "ms/step", d["ms_per_step"])
PY
( timeout -k 10 300 python bench.py --steps 5 --warmup 3 --no-scenes --no-cpu-baseline ) > gpurun_out/e_bench_n1.log 2> gpurun_out/e_bench_n1.err
python - <<PY
import json
d = json.loads(open("gpurun_out/e_bench_n1.log").read().strip().splitlines()[-1])
print("N=1:", d["value"], d["unit"], "e2e", d["e2e"]["value"], "ms/step", d["ms_per_step"])
PY
( timeout -k 10 600 python tools/gpu_parity_diag.py ) > gpurun_out/e_parity_diag.log 2>&1
cat gpurun_out/e_parity_diag.log
( time timeout -k 10 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "quarter_config or reference_kdtree" ) > gpurun_out/e_pytest_quarter.log 2>&1
tail -5 gpurun_out/e_pytest_quarter.log
