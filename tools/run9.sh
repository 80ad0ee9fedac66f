set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_pipelines_gpu.py -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench.py --workload q3 --steps 3 --warmup 3 > gpurun_out/bench_q3_n1.json 2> gpurun_out/bench_q3_n1.err; python - <<PY
import json
d=json.loads(open("gpurun_out/bench_q3_n1.json").read().strip().splitlines()[-1])
print("q3 ms", d["ms_per_step"], "value", d["value"]/1e9, d["parity"]["match"])
print({a:round(b,3) for a,b in d["roofline"]["per_kernel_ms_per_step"].items() if b>0.2})
PY
tail -2 gpurun_out/bench_q3_n1.err
