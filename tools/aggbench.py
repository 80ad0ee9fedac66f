"""Group-by throughput on the BASELINE shapes (device-resident inputs): C1, C3-like (Q1), C5 per-GPU share."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from galaxysql_b200 import api, native as N, synth

dev = torch.device("cuda", 0)
ctx = api.Context(0)
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0


def run(name, cols, types, groups, aggs, expected, bytes_per_row, reps=3):
    rows = cols[0].numel()
    best = None
    for i in range(reps + 1):
        if i == reps:
            ctx.profile(True); ctx.profile_reset()
        a = api.HashAgg(ctx, types, groups, aggs, expected)
        ctx.sync(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        a.consume([(c, None) for c in cols])
        n = a.finish()
        ctx.sync()
        dt = time.perf_counter() - t0
        a.close()
        if i > 0 and (best is None or dt < best):
            best = dt
    print({k: round(v[1], 3) for k, v in ctx.profile_dump().items()}); ctx.profile(False)
    print(f"{name}: {rows} rows, {n} groups: {best*1e3:.2f} ms  {rows/best/1e9:.2f} Grows/s  {rows*bytes_per_row/best/1e9:.0f} GB/s algorithmic", flush=True)


# C1: 1M-row INT column, k in [0, 65536), COUNT(*)
k = synth.rand_i64_t(1_000_000, 1, dev, post=lambda b: synth._u64_mod(b, 65536).to(torch.int32))
run("C1", [k], [N.T_INT32], [0], [(N.AGG_COUNT_STAR, [])], 65535, 4)

# C5 share of one GPU: 500M rows, 6.25M distinct BIGINT keys, SUM(double)
n5 = int(500_000_000 * scale)
k5 = synth.rand_i64_t(n5, 2, dev, post=lambda b: synth._u64_mod(b, 6_250_000) * 7919 + 13)
v5 = synth.rand_i64_t(n5, 3, dev, post=lambda b: (synth._lsr(b, 11).to(torch.float64) / float(1 << 53)))
run("C5/8", [k5, v5], [N.T_INT64, N.T_FP64], [0], [(N.AGG_SUM, [1])], 6_250_000, 16)
del k5, v5

# C3 (Q1 shape): 600M rows: 2 INT keys (3 x 2 values), 4 fp64 measures, INT shipdate; 8 accumulators over plain columns
n3 = int(600_037_902 * scale)
flag = synth.rand_i64_t(n3, 4, dev, post=lambda b: synth._u64_mod(b, 3).to(torch.int32))
status = synth.rand_i64_t(n3, 5, dev, post=lambda b: synth._u64_mod(b, 2).to(torch.int32))
qty = synth.rand_i64_t(n3, 6, dev, post=lambda b: (synth._u64_mod(b, 50) + 1).to(torch.float64))
price = synth.rand_i64_t(n3, 7, dev, post=lambda b: (synth._u64_mod(b, 10_410_000) + 90_000).to(torch.float64) / 100.0)
disc = synth.rand_i64_t(n3, 8, dev, post=lambda b: synth._u64_mod(b, 11).to(torch.float64) / 100.0)
tax = synth.rand_i64_t(n3, 9, dev, post=lambda b: synth._u64_mod(b, 9).to(torch.float64) / 100.0)
ship = synth.rand_i64_t(n3, 10, dev, post=lambda b: (synth._u64_mod(b, 2526) + 8036).to(torch.int32))
aggs = [(N.AGG_SUM, [2]), (N.AGG_SUM, [3]), (N.AGG_SUM, [4]), (N.AGG_SUM, [5]), (N.AGG_AVG, [2]), (N.AGG_AVG, [3]), (N.AGG_AVG, [4]), (N.AGG_COUNT_STAR, [])]
run("C3-plain", [flag, status, qty, price, disc, tax, ship], [0, 0, 2, 2, 2, 2, 0], [0, 1], aggs, 8, 44, reps=2)
