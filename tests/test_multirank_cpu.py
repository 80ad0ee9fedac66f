"""world_size-2 gloo test (CPU) of the N > 1 path's host logic: count matrix -> exchange plan -> AllToAllv layout, and
that joining the shuffled shards equals the global join.  Partition ids come from the oracle (tests may use it); the
GPU data path (partition kernel + NCCL) is covered by the -m gpu tests and bench.py --gpus N."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from galaxysql_b200 import shuffle
from oracle import oracle as orc
from tests import kat_util as ku


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, nb, npr, out_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        perm = np.argsort(ku.rand_u64(nb, 100 + rank)).astype(np.int64)
        bkey = shuffle.weak_scaling_build_keys(perm, rank, world)
        bpay = (ku.rand_u64(nb, 200 + rank) % np.uint64(1000)).astype(np.int32)
        pkey = (ku.rand_u64(npr, 300 + rank) % np.uint64(shuffle.global_key_space(nb, world))).astype(np.int64)
        ppay = (np.arange(npr) + rank * npr).astype(np.int32)

        def shuffle_side(key, pay):
            # what gsql_xchg_all_to_all does: partition ids (ExecUtils.partition), contiguous per-destination segments,
            # AllGather of counts, AllToAllv
            pid = orc.partition_ids(orc.hash_rows([(key, None)]), world)
            order = np.argsort(pid, kind="stable")
            counts = np.bincount(pid, minlength=world).astype(np.int64)
            gathered = [torch.zeros(world, dtype=torch.int64) for _ in range(world)]
            dist.all_gather(gathered, torch.from_numpy(counts))
            matrix = torch.stack(gathered).numpy()
            plan = shuffle.exchange_plan(matrix, rank)
            assert plan.send_counts.tolist() == counts.tolist()
            outs = []
            for col in (key[order], pay[order]):
                t = torch.from_numpy(np.ascontiguousarray(col))
                send = list(torch.split(t, plan.send_counts.tolist()))
                recv = [torch.empty(int(n), dtype=t.dtype) for n in plan.recv_counts]
                # grouped point-to-point, like the ncclSend/ncclRecv group of the native path (gloo has no alltoall)
                reqs = []
                for peer in range(world):
                    if peer == rank:
                        recv[peer].copy_(send[peer])
                    else:
                        reqs.append(dist.isend(send[peer].contiguous(), peer))
                        reqs.append(dist.irecv(recv[peer], peer))
                for q_ in reqs:
                    q_.wait()
                buf = torch.empty(plan.recv_total, dtype=t.dtype)
                for src in range(world):
                    buf[plan.recv_offsets[src]: plan.recv_offsets[src] + plan.recv_counts[src]] = recv[src]
                outs.append(buf.numpy())
            # every received key belongs to this rank
            assert (orc.partition_ids(orc.hash_rows([(outs[0], None)]), world) == rank).all()
            return outs

        bk, bp = shuffle_side(bkey, bpay)
        pk, pp = shuffle_side(pkey, ppay)
        spec = orc.JoinSpec(orc.JOIN_INNER, [0], [0], [orc.T_INT64])
        local = orc.hash_join(spec, [(pk, None), (pp, None)], [(bk, None), (bp, None)])
        out_q.put((rank, [c[0] for c in local], bkey, bpay, pkey, ppay))
    finally:
        dist.destroy_process_group()


def test_two_rank_shuffled_join_equals_global_join():
    world, nb, npr = 2, 3000, 8000
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, nb, npr, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort(key=lambda x: x[0])
    bkey = np.concatenate([r[2] for r in res]); bpay = np.concatenate([r[3] for r in res])
    pkey = np.concatenate([r[4] for r in res]); ppay = np.concatenate([r[5] for r in res])
    assert len(np.unique(bkey)) == world * nb                       # disjoint cover of the global key space
    spec = orc.JoinSpec(orc.JOIN_INNER, [0], [0], [orc.T_INT64])
    glob = orc.hash_join(spec, [(pkey, None), (ppay, None)], [(bkey, None), (bpay, None)])
    got = ku.rows_multiset([(np.concatenate([r[1][c] for r in res]), None) for c in range(4)])
    assert got == ku.rows_multiset(glob)
    assert sum(len(r[1][0]) for r in res) == world * npr            # every probe row matches exactly once


def test_exchange_plan_offsets():
    m = np.array([[3, 1, 0], [2, 2, 5], [0, 4, 1]])
    p = shuffle.exchange_plan(m, 1)
    assert p.send_counts.tolist() == [2, 2, 5] and p.send_offsets.tolist() == [0, 2, 4]
    assert p.recv_counts.tolist() == [1, 2, 4] and p.recv_offsets.tolist() == [0, 1, 3] and p.recv_total == 7
    assert shuffle.worst_case_capacity(1000, 1) == 1000 and shuffle.worst_case_capacity(1000, 8) > 1000


def _slab_worker(rank, world, port, n, slabs, out_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        nloc = n + 37 * rank                                      # ranks hold different row counts
        key = (ku.rand_u64(nloc, 500 + rank) % np.uint64(10_000)).astype(np.int64)
        pay = (np.arange(nloc) + rank * 1_000_000).astype(np.int64)
        pid = orc.partition_ids(orc.hash_rows([(key, None)]), world)
        sr = shuffle.slab_rows(nloc, slabs)
        counts = np.zeros((slabs, world), dtype=np.int64)
        for i in range(slabs):
            counts[i] = np.bincount(pid[i * sr:(i + 1) * sr], minlength=world)
        gathered = [torch.zeros(slabs * world, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(gathered, torch.from_numpy(counts.reshape(-1)))
        tensor = torch.stack(gathered).numpy().reshape(world, slabs, world)
        plan = shuffle.slab_exchange_plan(tensor, rank, nloc)
        assert plan.slab_rows == sr and (plan.send_counts == counts).all()
        # staging: every slab partitioned on its own (what k_xchg_scatter does for the slab's row range)
        skey, spay = np.zeros(slabs * sr, dtype=np.int64), np.zeros(slabs * sr, dtype=np.int64)
        for i in range(slabs):
            lo, hi = i * sr, min(nloc, (i + 1) * sr)
            if hi <= lo:
                continue
            order = np.argsort(pid[lo:hi], kind="stable")
            skey[lo:hi], spay[lo:hi] = key[lo:hi][order], pay[lo:hi][order]
        rkey, rpay = np.full(plan.recv_total, -1, dtype=np.int64), np.full(plan.recv_total, -1, dtype=np.int64)
        for i in range(slabs):                                    # slab after slab, like the stripe streams
            reqs, bufs = [], []
            for peer in range(world):
                ns, nr = int(plan.send_counts[i, peer]), int(plan.recv_counts[i, peer])
                so, ro = int(plan.send_offsets[i, peer]), int(plan.recv_offsets[i, peer])
                if peer == rank:
                    rkey[ro:ro + nr], rpay[ro:ro + nr] = skey[so:so + ns], spay[so:so + ns]
                    continue
                for src, dst in ((skey, rkey), (spay, rpay)):
                    if ns:
                        reqs.append(dist.isend(torch.from_numpy(np.ascontiguousarray(src[so:so + ns])), peer))
                    if nr:
                        t = torch.empty(nr, dtype=torch.int64)
                        reqs.append(dist.irecv(t, peer))
                        bufs.append((dst, ro, nr, t))
            for q_ in reqs:
                q_.wait()
            for dst, ro, nr, t in bufs:
                dst[ro:ro + nr] = t.numpy()
        assert (rkey >= 0).all()                                  # every receive row written exactly once (no gaps)
        assert (orc.partition_ids(orc.hash_rows([(rkey, None)]), world) == rank).all()
        # per-source contiguity: rows of source s occupy one block, in slab order
        src_of = rpay // 1_000_000
        assert (np.diff(src_of) >= 0).all()
        out_q.put((rank, rkey, rpay, key, pay))
    finally:
        dist.destroy_process_group()


def test_two_rank_slabbed_exchange_plan_delivers_every_row_once():
    """Host mirror of the opt-in slabbed AllToAllv (xchg.cu: all_to_all_slabbed): 3 slabs (the last one short), ranks with
    different row counts; the union of what the ranks receive is the union of what they sent, routed by ExecUtils.partition."""
    world, n, slabs = 2, 1500, 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_slab_worker, args=(r, world, port, n, slabs, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    sent = ku.rows_multiset([(np.concatenate([r[3] for r in res]), None), (np.concatenate([r[4] for r in res]), None)])
    got = ku.rows_multiset([(np.concatenate([r[1] for r in res]), None), (np.concatenate([r[2] for r in res]), None)])
    assert got == sent


def test_slab_plan_offsets():
    t = np.zeros((2, 2, 2), dtype=np.int64)       # [src][slab][dst]
    t[0] = [[3, 1], [2, 2]]
    t[1] = [[0, 4], [5, 1]]
    p = shuffle.slab_exchange_plan(t, 1, 8)
    assert p.slab_rows == 256
    assert p.send_offsets.tolist() == [[0, 0], [256, 261]] and p.send_counts.tolist() == [[0, 4], [5, 1]]
    assert p.recv_counts.tolist() == [[1, 4], [2, 1]]         # [slab][src]
    assert p.recv_offsets.tolist() == [[0, 3], [1, 7]] and p.recv_total == 8
