set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
SEL='(film_matches and (textures or direct or zoo)) or branch_tree or gpu_textures'
( timeout 500 compute-sanitizer --tool memcheck python -m pytest tests -m gpu -x -q -k "$SEL" ) > gpurun_out/san_memcheck.log 2>&1; tail -4 gpurun_out/san_memcheck.log
( timeout 500 compute-sanitizer --tool racecheck python -m pytest tests -m gpu -x -q -k "film_matches and (textures or direct)" ) > gpurun_out/san_racecheck.log 2>&1; tail -4 gpurun_out/san_racecheck.log
( timeout 400 compute-sanitizer --tool initcheck python -m pytest tests -m gpu -x -q -k "film_matches and (textures or direct)" ) > gpurun_out/san_initcheck.log 2>&1; tail -4 gpurun_out/san_initcheck.log
