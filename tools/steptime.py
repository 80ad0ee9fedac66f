"""Host-side timeline of one C2 step (device-resident): wall time of each C-ABI call with a sync after it."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from galaxysql_b200 import api, native as N, synth
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
dev = torch.device("cuda", 0)
ctx = api.Context(0)
nb, npr = int(1e8 * scale), int(1e9 * scale)
build, probe = synth.c2_tables_t(nb, npr, dev)
torch.cuda.synchronize()
types = [N.T_INT64, N.T_INT32, N.T_INT32]
tt = {N.T_INT64: torch.int64, N.T_INT32: torch.int32}
out = [(torch.empty(npr, dtype=tt[t], device=dev), None) for t in types + types]
b = lambda cols: [(c, None) for c in cols]
prof = len(sys.argv) > 2
for it in range(4):
    if it == 2 and prof:
        ctx.profile(True)
    t = [time.perf_counter()]
    j = api.HashJoin(ctx, N.JOIN_INNER, types, types, [0], [0], expected_build_rows=nb); ctx.sync(); t.append(time.perf_counter())
    j.build_consume(b(build)); ctx.sync(); t.append(time.perf_counter())
    j.build_finish(); ctx.sync(); t.append(time.perf_counter())
    n = j.probe_into(b(probe), out, npr); ctx.sync(); t.append(time.perf_counter())
    j.close(); ctx.sync(); t.append(time.perf_counter())
    if prof and it == 3:
        print(ctx.profile_dump())
    print(f"iter {it}: create {1e3*(t[1]-t[0]):.2f} consume {1e3*(t[2]-t[1]):.2f} finish {1e3*(t[3]-t[2]):.2f} probe {1e3*(t[4]-t[3]):.2f} close {1e3*(t[5]-t[4]):.2f} ms  rows {n}")
