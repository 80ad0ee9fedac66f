set -x
cd $GRAFT_REPO_ROOT
nvidia-smi -L
nvidia-smi topo -m | head -20
export GSQL_MG_SECTIONS=push,join
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25
timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --no-aux > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 3000 gpurun_out/bench_n1.json; tail -5 gpurun_out/bench_n1.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; tail -c 3000 gpurun_out/bench_n2.json; tail -30 gpurun_out/bench_n2.err
