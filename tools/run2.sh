set -x
cd $GRAFT_REPO_ROOT
run() {
  tag=$1; shift
  env "$@" timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-e2e --no-cpu --slabs $SL > gpurun_out/bench_n2_$tag.json 2> gpurun_out/bench_n2_$tag.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_n2_$tag.json").read().strip().splitlines()[-1])
    print("CFG $tag slabs $SL ms_per_step", round(d["ms_per_step"],2), "value", round(d["value"]/1e9,2), {k:round(v,2) for k,v in d["roofline"]["per_kernel_ms_per_step"].items() if v>0.5})
except Exception as e:
    print("CFG $tag failed", e); print(open("gpurun_out/bench_n2_$tag.err").read()[-1500:])
PY
}
SL=4 run c8s4 GSQL_XCHG_PUSH_CTAS_PER_SM=8
SL=4 run c4s4 GSQL_XCHG_PUSH_CTAS_PER_SM=4
SL=4 run c2s4 GSQL_XCHG_PUSH_CTAS_PER_SM=2
SL=8 run c4s8 GSQL_XCHG_PUSH_CTAS_PER_SM=4
SL=2 run c4s2 GSQL_XCHG_PUSH_CTAS_PER_SM=4
SL=1 run c8s1 GSQL_XCHG_PUSH_CTAS_PER_SM=8
