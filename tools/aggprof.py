import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from galaxysql_b200 import api, native as N, synth
dev = torch.device("cuda", 0); ctx = api.Context(0)
n3 = 60_000_000
flag = synth.rand_i64_t(n3, 4, dev, post=lambda b: synth._u64_mod(b, 3).to(torch.int32))
status = synth.rand_i64_t(n3, 5, dev, post=lambda b: synth._u64_mod(b, 2).to(torch.int32))
qty = synth.rand_i64_t(n3, 6, dev, post=lambda b: (synth._u64_mod(b, 50) + 1).to(torch.float64))
price = synth.rand_i64_t(n3, 7, dev, post=lambda b: (synth._u64_mod(b, 10_410_000) + 90_000).to(torch.float64) / 100.0)
disc = synth.rand_i64_t(n3, 8, dev, post=lambda b: synth._u64_mod(b, 11).to(torch.float64) / 100.0)
tax = synth.rand_i64_t(n3, 9, dev, post=lambda b: synth._u64_mod(b, 9).to(torch.float64) / 100.0)
ship = synth.rand_i64_t(n3, 10, dev, post=lambda b: (synth._u64_mod(b, 2526) + 8036).to(torch.int32))
aggs = [(N.AGG_SUM, [2]), (N.AGG_SUM, [3]), (N.AGG_SUM, [4]), (N.AGG_SUM, [5]), (N.AGG_AVG, [2]), (N.AGG_AVG, [3]), (N.AGG_AVG, [4]), (N.AGG_COUNT_STAR, [])]
for i in range(2):
    a = api.HashAgg(ctx, [0, 0, 2, 2, 2, 2, 0], [0, 1], aggs, 8)
    a.consume([(c, None) for c in [flag, status, qty, price, disc, tax, ship]])
    print(a.finish()); a.close()
