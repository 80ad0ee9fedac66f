set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; python - <<PY
import json
d=json.loads(open("gpurun_out/bench_n1.json").read().strip().splitlines()[-1])
print("C2 ms", d["ms_per_step"], "value", d["value"]/1e9, "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"]/1e9, "cpu", d["cpu_baseline"]["value"]/1e9, d["clocks"])
for k in ("agg_q1","agg_c5","agg_c1"):
    e=d["roofline"][k]; print(k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in e.items() if a in ("ms_wall","kernel_ms","frac","kernel","per_kernel_ms","error")})
for k in ("pipeline_q3","pipeline_c5"):
    e=d["roofline"].get(k,{}); print(k, e.get("ms_per_step"), e.get("rows_per_s"), e.get("parity",{}).get("match"))
PY
tail -3 gpurun_out/bench_n1.err
