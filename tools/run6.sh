set -x
cd $GRAFT_REPO_ROOT
for D in 0 1; do
GSQL_JOIN_SCATTER_DIRECT=$D timeout 600 python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu --no-aux > gpurun_out/bench_direct$D.json 2> gpurun_out/bench_direct$D.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_direct$D.json").read().strip().splitlines()[-1])
print("DIRECT=$D ms", round(d["ms_per_step"],2), "frac", round(d["roofline"]["frac"],4), {k:round(v,2) for k,v in d["roofline"]["per_kernel_ms_per_step"].items()})
PY
tail -2 gpurun_out/bench_direct$D.err
done
