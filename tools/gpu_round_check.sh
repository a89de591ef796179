set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > gpurun_out/smi.txt
( time timeout 600 python -m pytest tests -m gpu -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -6 gpurun_out/pytest_gpu.log
( timeout 200 python tools/scene_breakdown.py textures@1920x1080x8 direct@1920x1080x8 ) > gpurun_out/breakdown_textures.log 2>&1
cat gpurun_out/breakdown_textures.log
( time timeout 400 python bench.py ) > gpurun_out/bench_default.log 2> gpurun_out/bench_default.err
tail -1 gpurun_out/bench_default.log > gpurun_out/bench_default.json
python -c "
import json; b=json.load(open('gpurun_out/bench_default.json')); print(b['value'], b['mrays_per_s'], b['e2e']['value'], b['roofline']['frac'], {k: round(v['msamples_per_s']) for k, v in b['scenes'].items()})"
( timeout 100 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
