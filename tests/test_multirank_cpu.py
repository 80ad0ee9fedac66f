"""world_size-2 gloo tests (CPU) of the N > 1 path's host logic.  The layout arithmetic under test is the PRODUCT's:
gsql_xchg_plan_layout, the host function gsql_xchg_push itself calls (exported by libgsql_gpu.so, no GPU needed).  The
ranks exchange real per-slab count matrices over gloo, place their segments where the plan says, and the result must
tile every receive buffer exactly once, slab by slab; joining the shuffled shards must equal the global join.
Destination ids come from the oracle (tests may use it); the GPU data path (split-and-push kernels over peer memory)
is covered by the -m gpu tests, tests/test_multigpu.py and bench.py --gpus N."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from galaxysql_b200 import api
from oracle import oracle as orc
from tests import kat_util as ku

TILE = 2048  # rows per slab are a multiple of the push kernel's tile (xchg.cu: PUSH_TILE)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _slab_rows(n, nslabs):
    per = -(-max(n, 1) // nslabs)
    return -(-per // TILE) * TILE


def _push(rank, world, cols, key_col, nslabs):
    """What gsql_xchg_push does, with gloo carrying the bytes: returns (received columns, slab_rows)."""
    n = len(cols[0])
    pid = orc.partition_ids(orc.hash_rows([(cols[key_col], None)]), world)
    sr = _slab_rows(n, nslabs)
    counts = np.zeros((nslabs, world), dtype=np.int64)
    for i in range(nslabs):
        counts[i] = np.bincount(pid[i * sr:(i + 1) * sr], minlength=world)
    gathered = [torch.zeros(nslabs * world, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(gathered, torch.from_numpy(counts.reshape(-1)))
    matrix = torch.stack(gathered).numpy().reshape(world, nslabs, world)
    send_base, recv_base, slab_rows, worst = api.plan_layout(world, nslabs, rank, matrix)
    total = int(slab_rows.sum())
    assert worst >= total
    out = [np.full(total, -1, dtype=c.dtype) for c in cols]
    written = np.zeros(total, dtype=np.int32)
    for i in range(nslabs):
        lo, hi = i * sr, min(n, (i + 1) * sr)
        for peer in range(world):
            sel = np.nonzero(pid[lo:hi] == peer)[0] + lo if hi > lo else np.zeros(0, dtype=np.int64)
            ns, nr = int(counts[i, peer]), int(matrix[peer, i, rank])
            assert len(sel) == ns
            ro = int(recv_base[i, peer])
            if peer == rank:
                assert int(send_base[i, peer]) == ro      # my own segment: where I send is where I receive
                for c, o in zip(cols, out):
                    o[ro:ro + nr] = c[sel]
                written[ro:ro + nr] += 1
                continue
            reqs, bufs = [], []
            # the sender's idea of the segment start must be the receiver's: exchange and compare it
            reqs.append(dist.isend(torch.tensor([int(send_base[i, peer])]), peer))
            theirs = torch.zeros(1, dtype=torch.int64)
            reqs.append(dist.irecv(theirs, peer))
            for c, o in zip(cols, out):
                if ns:
                    reqs.append(dist.isend(torch.from_numpy(np.ascontiguousarray(c[sel])), peer))
                if nr:
                    t = torch.empty(nr, dtype=torch.from_numpy(c[:1]).dtype)
                    reqs.append(dist.irecv(t, peer))
                    bufs.append((o, t))
            for q_ in reqs:
                q_.wait()
            assert int(theirs.item()) == ro
            for o, t in bufs:
                o[ro:ro + nr] = t.numpy()
            written[ro:ro + nr] += 1
    assert (written == 1).all()                            # the segments tile the receive buffer exactly once
    return out, slab_rows


def _worker(rank, world, port, nb, npr, nslabs, out_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        nb_r, npr_r = nb + 13 * rank, npr + 37 * rank        # ranks hold different row counts
        bkey = (np.argsort(ku.rand_u64(nb_r, 100 + rank)).astype(np.int64)) * world + rank   # disjoint key shares
        bpay = (ku.rand_u64(nb_r, 200 + rank) % np.uint64(1000)).astype(np.int64)
        pkey = (ku.rand_u64(npr_r, 300 + rank) % np.uint64(nb * world)).astype(np.int64)
        ppay = (np.arange(npr_r) + rank * 1_000_000).astype(np.int64)
        (bk, bp), _ = _push(rank, world, [bkey, bpay], 0, 1)
        (pk, pp), slab_rows = _push(rank, world, [pkey, ppay], 0, nslabs)
        assert (orc.partition_ids(orc.hash_rows([(pk, None)]), world) == rank).all()   # every received key is mine
        # slab-major layout: inside a slab the sources follow each other, so a slab is one contiguous batch
        off = 0
        for i in range(nslabs):
            src = pp[off:off + int(slab_rows[i])] // 1_000_000
            assert (np.diff(src) >= 0).all()
            off += int(slab_rows[i])
        spec = orc.JoinSpec(orc.JOIN_INNER, [0], [0], [orc.T_INT64])
        local = orc.hash_join(spec, [(pk, None), (pp, None)], [(bk, None), (bp, None)])
        out_q.put((rank, [c[0] for c in local], bkey, bpay, pkey, ppay))
    finally:
        dist.destroy_process_group()


def test_two_rank_pushed_join_equals_global_join():
    world, nb, npr, nslabs = 2, 3000, 9000, 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, nb, npr, nslabs, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort(key=lambda x: x[0])
    bkey = np.concatenate([r[2] for r in res]); bpay = np.concatenate([r[3] for r in res])
    pkey = np.concatenate([r[4] for r in res]); ppay = np.concatenate([r[5] for r in res])
    spec = orc.JoinSpec(orc.JOIN_INNER, [0], [0], [orc.T_INT64])
    glob = orc.hash_join(spec, [(pkey, None), (ppay, None)], [(bkey, None), (bpay, None)])
    got = ku.rows_multiset([(np.concatenate([r[1][c] for r in res]), None) for c in range(4)])
    assert got == ku.rows_multiset(glob)


def test_plan_layout_known_answer():
    m = np.arange(8).reshape(2, 2, 2) + 1                 # m[src][slab][dst]
    s0, r0, rows0, w0 = api.plan_layout(2, 2, 0, m)
    s1, r1, rows1, w1 = api.plan_layout(2, 2, 1, m)
    assert w0 == w1 == 20 and rows0.tolist() == [6, 10] and rows1.tolist() == [8, 12]
    assert s0.tolist() == [[0, 0], [6, 8]] and r0.tolist() == [[0, 1], [6, 9]]
    assert s1.tolist() == [[1, 2], [9, 12]] and r1.tolist() == [[0, 2], [8, 12]]
    # what rank a sends to rank b starts where rank b expects rank a's rows
    for slab in range(2):
        assert s0[slab][1] == r1[slab][0] and s1[slab][0] == r0[slab][1]
