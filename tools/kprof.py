"""Small driver for `ncu`: runs each kernel of interest once or twice on realistic shapes (one GPU).
KPROF=agg_c5,agg_q3,push,scan,agg_q1 selects."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from galaxysql_b200 import api, native as N, synth  # noqa: E402

which = os.environ.get("KPROF", "agg_c5,agg_q3,push,scan,agg_q1").split(",")
dev = torch.device("cuda", 0)
ctx = api.Context(0)
E = api.E
if "agg_c5" in which:
    n, keys = 250_000_000, 6_250_000
    k = synth.rand_i64_t(n, 11, dev, post=lambda b: synth._u64_mod(b, keys) * 8 + 3)
    v = synth.rand_i64_t(n, 12, dev, post=lambda b: synth._lsr(b, 11).to(torch.float64) / float(1 << 53))
    a = api.HashAgg(ctx, [N.T_INT64, N.T_FP64], [0], [(N.AGG_SUM, [1])], keys)
    a.consume([(k, None), (v, None)])
    a.consume([(k, None), (v, None)])   # second batch: every group exists
    print("c5 groups", a.finish())
    a.close()
    del k, v
if "agg_q3" in which:
    n = 3_000_000
    k0 = synth.rand_i64_t(n, 13, dev, post=lambda b: synth._u64_mod(b, 1_200_000) * 8 + 1)
    k1 = synth.rand_i64_t(n, 14, dev, post=lambda b: (synth._u64_mod(b, 1000) + 8035).to(torch.int32))
    k2 = torch.zeros(n, dtype=torch.int32, device=dev)
    v = synth.rand_i64_t(n, 15, dev, post=lambda b: synth._lsr(b, 11).to(torch.float64) / float(1 << 53))
    a = api.HashAgg(ctx, [N.T_INT64, N.T_FP64, N.T_INT64, N.T_INT32, N.T_INT32], [0, 3, 4], [(N.AGG_SUM, [1])], 5_600_000)
    a.consume([(k0, None), (v, None), (k0, None), (k1, None), (k2, None)])
    print("q3 groups", a.finish())
    a.close()
if "push" in which:
    n = 250_000_000
    k = synth.rand_i64_t(n, 3, dev, post=lambda b: synth._u64_mod(b, 100_000_000))
    p1 = synth.rand_i64_t(n, 4, dev, post=synth._top31)
    p2 = synth.rand_i64_t(n, 5, dev, post=synth._top31)
    x = api.Exchange(ctx, [N.T_INT64, N.T_INT32, N.T_INT32], [0], 1)
    x.open_p2p(n + 1024)
    x.push([(k, None), (p1, None), (p2, None)], 1)
    x.push_wait()
    x.close()
    del k, p1, p2
if "scan" in which:
    n = 100_000_000
    ok = synth.rand_i64_t(n, 23, dev, post=lambda b: synth._u64_mod(b, 1 << 30))
    pr = synth.rand_i64_t(n, 24, dev, post=lambda b: (synth._u64_mod(b, 10_410_000) + 90_000).to(torch.float64) / 100.0)
    di = synth.rand_i64_t(n, 25, dev, post=lambda b: synth._u64_mod(b, 11).to(torch.float64) / 100.0)
    sd = synth.rand_i64_t(n, 26, dev, post=lambda b: (synth._u64_mod(b, 2557) + 8035).to(torch.int32))
    s = api.Scan(ctx, [N.T_INT64, N.T_FP64, N.T_FP64, N.T_INT32], [E.col(0), E.col(1) * (1.0 - E.col(2))], filter=E.col(3) > 9204)
    out = s.apply([(ok, None), (pr, None), (di, None), (sd, None)], nullable_out=False)
    print("scan rows", out[0][0].shape[0])
    s.close()
if "agg_q1" in which:
    n3 = 150_000_000
    flag = synth.rand_i64_t(n3, 4, dev, post=lambda b: synth._u64_mod(b, 3).to(torch.int32))
    status = synth.rand_i64_t(n3, 5, dev, post=lambda b: synth._u64_mod(b, 2).to(torch.int32))
    qty = synth.rand_i64_t(n3, 6, dev, post=lambda b: (synth._u64_mod(b, 50) + 1).to(torch.float64))
    price = synth.rand_i64_t(n3, 7, dev, post=lambda b: (synth._u64_mod(b, 10_410_000) + 90_000).to(torch.float64) / 100.0)
    disc = synth.rand_i64_t(n3, 8, dev, post=lambda b: synth._u64_mod(b, 11).to(torch.float64) / 100.0)
    tax = synth.rand_i64_t(n3, 9, dev, post=lambda b: synth._u64_mod(b, 9).to(torch.float64) / 100.0)
    ship = synth.rand_i64_t(n3, 10, dev, post=lambda b: (synth._u64_mod(b, 2526) + 8036).to(torch.int32))
    aggs = [(N.AGG_SUM, [2]), (N.AGG_SUM, [3]), (N.AGG_SUM, [7]), (N.AGG_SUM, [8]), (N.AGG_AVG, [2]), (N.AGG_AVG, [3]),
            (N.AGG_AVG, [4]), (N.AGG_COUNT_STAR, [])]
    a = api.HashAgg(ctx, [0, 0, 2, 2, 2, 2, 0], [0, 1], aggs, 8, derived=[(N.EXPR_MUL_1MINUS, 3, 4, 0), (N.EXPR_MUL_1MINUS_1PLUS, 3, 4, 5)],
                    row_filter=(6, N.CMP_LE, 10471))
    a.consume([(c, None) for c in (flag, status, qty, price, disc, tax, ship)])
    print("q1 groups", a.finish())
    a.close()
ctx.sync()
