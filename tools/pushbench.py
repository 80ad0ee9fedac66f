"""Stand-alone timing of the split-and-push exchange on ONE GPU (every destination is local memory: kernel efficiency
without NVLink): C2's probe side, 1 B rows x (BIGINT, INT, INT), for several fake rank counts is not possible with one
rank — so R = 1 here; the N-GPU numbers come from bench.py --gpus N.  Prints the per-kernel profile."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from galaxysql_b200 import api, native as N, synth  # noqa: E402

n = int(float(os.environ.get("ROWS", "1e9")))
dev = torch.device("cuda", 0)
ctx = api.Context(0)
k = synth.rand_i64_t(n, 3, dev, post=lambda b: synth._u64_mod(b, 100_000_000))
p1 = synth.rand_i64_t(n, 4, dev, post=synth._top31)
p2 = synth.rand_i64_t(n, 5, dev, post=synth._top31)
x = api.Exchange(ctx, [N.T_INT64, N.T_INT32, N.T_INT32], [0], 1)
x.open_p2p(n + 1024)
for slabs in (1, 4):
    for rep in range(3):
        ctx.profile(True)
        ctx.profile_reset()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        x.push([(k, None), (p1, None), (p2, None)], slabs)
        x.push_wait()
        dt = time.perf_counter() - t0
        prof = ctx.profile_dump()
        ctx.profile(False)
    print(f"push {n} rows, {slabs} slabs: wall {dt * 1e3:.2f} ms", {a: round(b[1], 3) for a, b in prof.items()}, flush=True)
x.close()
