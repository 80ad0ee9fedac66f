set -x
cd $GRAFT_REPO_ROOT
nvidia-smi -L | wc -l
GSQL_TEST_GPUS=8 timeout 900 python -m pytest tests/test_multigpu.py -m gpu -q 2>&1 | tail -4
run() {
  tag=$1; n=$2; wl=$3; shift 3
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --workload $wl --steps 4 --warmup 3 --no-e2e --no-cpu --slabs $SL > gpurun_out/scale_${tag}.json 2> gpurun_out/scale_${tag}.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/scale_${tag}.json").read().strip().splitlines()[-1])
    print("CFG ${tag} n=$n wl=$wl slabs=$SL ms_per_step", round(d["ms_per_step"],2), "value", round(d["value"]/1e9,2), {k:round(v,2) for k,v in d["roofline"]["per_kernel_ms_per_step"].items() if v>0.5}, d.get("parity",{}).get("match"))
except Exception as e:
    print("CFG ${tag} failed", e); print(open("gpurun_out/scale_${tag}.err").read()[-1500:])
PY
}
SL=1 run c2_n8_s1 8 c2
SL=2 run c2_n8_s2 8 c2
SL=4 run c2_n8_s4 8 c2
SL=1 run c2_n4_s1 4 c2
SL=4 run q3_n8_s4 8 q3
SL=1 run q3_n8_s1 8 q3
SL=4 run c5_n8_s4 8 c5
SL=1 run c5_n8_s1 8 c5
