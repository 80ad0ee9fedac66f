set -x
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 900 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 6000 gpurun_out/bench_n1.json; tail -5 gpurun_out/bench_n1.err
timeout 600 python bench.py --workload q3 --steps 3 --warmup 2 > gpurun_out/bench_q3_n1.json 2> gpurun_out/bench_q3_n1.err; tail -c 3000 gpurun_out/bench_q3_n1.json; tail -5 gpurun_out/bench_q3_n1.err
timeout 600 python bench.py --workload c5 --steps 3 --warmup 2 > gpurun_out/bench_c5_n1.json 2> gpurun_out/bench_c5_n1.err; tail -c 3000 gpurun_out/bench_c5_n1.json; tail -5 gpurun_out/bench_c5_n1.err
