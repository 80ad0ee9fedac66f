set -x
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_pipelines_gpu.py tests/test_multigpu.py -m gpu -x -q -k "agg or scan or q3 or two_phase or push or across" 2>&1 | tail -8
timeout 300 python tools/pushbench.py 2>&1 | tail -3
timeout 900 python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; python - <<PY
import json
d=json.loads(open("gpurun_out/bench_n1.json").read().strip().splitlines()[-1])
print("C2 ms", d["ms_per_step"], "frac", d["roofline"]["frac"])
for k in ("agg_q1","agg_c5","agg_c1"):
    e=d["roofline"][k]; print(k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in e.items() if a in ("ms_wall","kernel_ms","frac","kernel","per_kernel_ms","error")})
PY
tail -3 gpurun_out/bench_n1.err
timeout 600 python bench.py --workload q3 --steps 3 --warmup 2 > gpurun_out/bench_q3_n1.json 2> gpurun_out/bench_q3_n1.err; python - <<PY
import json
d=json.loads(open("gpurun_out/bench_q3_n1.json").read().strip().splitlines()[-1])
print("Q3 ms", d["ms_per_step"], "value", d["value"]/1e9, d["parity"]["match"])
print({a:round(b,3) for a,b in d["roofline"]["per_kernel_ms_per_step"].items()})
PY
tail -3 gpurun_out/bench_q3_n1.err
timeout 600 python bench.py --workload c5 --steps 3 --warmup 2 > gpurun_out/bench_c5_n1.json 2> gpurun_out/bench_c5_n1.err; python - <<PY
import json
d=json.loads(open("gpurun_out/bench_c5_n1.json").read().strip().splitlines()[-1])
print("C5 ms", d["ms_per_step"], "value", d["value"]/1e9, d["parity"]["match"])
print({a:round(b,3) for a,b in d["roofline"]["per_kernel_ms_per_step"].items()})
PY
tail -3 gpurun_out/bench_c5_n1.err
