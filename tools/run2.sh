set -x
cd $GRAFT_REPO_ROOT
nvidia-smi -L
nvidia-smi topo -m | head -12
timeout 900 python -m pytest tests/test_multigpu.py -m gpu -x -q 2>&1 | tail -25
for S in 4; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-e2e --no-cpu --slabs $S > gpurun_out/bench_n2_s$S.json 2> gpurun_out/bench_n2_s$S.err; tail -c 2500 gpurun_out/bench_n2_s$S.json; tail -30 gpurun_out/bench_n2_s$S.err
done
