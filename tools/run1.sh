set -x
cd $GRAFT_REPO_ROOT
nvidia-smi -L
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25
timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --no-aux > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 3000 gpurun_out/bench_n1.json; tail -5 gpurun_out/bench_n1.err
