set -x
cd $GRAFT_REPO_ROOT
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2_full.json 2> gpurun_out/bench_n2_full.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_n2_full.json").read().strip().splitlines()[-1])
    print("N2 ms_per_step", round(d["ms_per_step"],2), "value", round(d["value"]/1e9,2))
    print("e2e", d.get("e2e"))
    for k in ("pipeline_q3","pipeline_c5"):
        e=d["roofline"].get(k,{}); print(k, e.get("ms_per_step"), e.get("rows_per_s"), e.get("parity",{}).get("match"))
except Exception as e:
    print("failed", e)
PY
grep -v "^\[W\|^W0\|^\*\*\*\|OMP_NUM" gpurun_out/bench_n2_full.err | tail -25
