set -x
cd $GRAFT_REPO_ROOT
for S in 4 8; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-e2e --no-cpu --slabs $S > gpurun_out/bench_n2_s$S.json 2> gpurun_out/bench_n2_s$S.err; python - <<PY
import json
d=json.loads(open("gpurun_out/bench_n2_s$S.json").read().strip().splitlines()[-1])
print("SLABS $S ms_per_step", d["ms_per_step"], "value", d["value"]/1e9)
for k,v in d["roofline"]["per_kernel_ms_per_step"].items(): print("   ", k, round(v,3))
PY
tail -3 gpurun_out/bench_n2_s$S.err
done
GSQL_XCHG_PUSH_CTAS_PER_SM=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-e2e --no-cpu --slabs 4 > gpurun_out/bench_n2_c1.json 2> gpurun_out/bench_n2_c1.err; python - <<PY
import json
d=json.loads(open("gpurun_out/bench_n2_c1.json").read().strip().splitlines()[-1])
print("CTAS_PER_SM 1 ms_per_step", d["ms_per_step"], "value", d["value"]/1e9)
for k,v in d["roofline"]["per_kernel_ms_per_step"].items(): print("   ", k, round(v,3))
PY
