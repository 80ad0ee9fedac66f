"""Host-side bookkeeping of the multi-GPU hash-partition exchange (one process per GPU).

The data path itself is native (gsql_xchg_all_to_all: partition kernel + NCCL AllToAllv); this module holds the pure
host logic around it — the R x R count matrix -> per-peer send/recv offsets, capacity planning, the weak-scaling shard
layout bench.py uses — so that it can be exercised with world_size 2 on CPU (gloo) without a GPU.
Mirrors PartitionedOutputBuffer / ExchangeClient bookkeeping (EX/mpp/execution/buffer/PartitionedOutputBuffer.java:138-175).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Sequence

import numpy as np


@dataclass
class ExchangePlan:
    send_offsets: np.ndarray  # [R] start row of the segment for each destination inside the partitioned send buffer
    send_counts: np.ndarray   # [R]
    recv_offsets: np.ndarray  # [R] start row of each source's rows inside the receive buffer
    recv_counts: np.ndarray   # [R]
    recv_total: int


def exchange_plan(count_matrix: np.ndarray, rank: int) -> ExchangePlan:
    """count_matrix[src, dst] = rows src sends to dst (the AllGather of every rank's partition counts)."""
    m = np.asarray(count_matrix, dtype=np.int64)
    assert m.ndim == 2 and m.shape[0] == m.shape[1]
    send = m[rank, :].copy()
    recv = m[:, rank].copy()
    return ExchangePlan(np.concatenate([[0], np.cumsum(send)[:-1]]), send, np.concatenate([[0], np.cumsum(recv)[:-1]]), recv,
                        int(recv.sum()))


@dataclass
class SlabPlan:
    slab_rows: int            # rows per slab of the local batch (multiple of 256; the last slabs may be short or empty)
    send_offsets: np.ndarray  # [S, R] start row of segment (slab, dst) inside the partitioned send staging
    send_counts: np.ndarray   # [S, R]
    recv_offsets: np.ndarray  # [S, R] start row of segment (slab, src) inside the receive buffer
    recv_counts: np.ndarray   # [S, R]
    recv_total: int


def slab_rows(rows: int, slabs: int, block: int = 256) -> int:
    """Rows per slab as all_to_all_slabbed cuts the batch: ceil(rows / slabs) rounded up to the partition kernel's block."""
    per = -(-max(rows, 1) // slabs)
    return -(-per // block) * block


def slab_exchange_plan(count_tensor: np.ndarray, rank: int, rows: int) -> SlabPlan:
    """count_tensor[src, slab, dst] = rows of src's slab routed to dst (the AllGather of every rank's S x R counts).

    Mirrors xchg.cu: all_to_all_slabbed (GSQL_XCHG_SLABS).  Send staging: slab i occupies rows [i * slab_rows, ...) of the
    staging buffer, its destinations contiguous inside.  Receive buffer: per SOURCE contiguous (same layout as the
    sequential path), the source's slabs one after the other — so a consumer cannot tell the two paths apart."""
    m = np.asarray(count_tensor, dtype=np.int64)
    assert m.ndim == 3 and m.shape[0] == m.shape[2]
    R, S = m.shape[0], m.shape[1]
    sr = slab_rows(rows, S)
    send = m[rank]                                   # [S, R]
    soff = np.zeros((S, R), dtype=np.int64)
    for i in range(S):
        soff[i] = i * sr + np.concatenate([[0], np.cumsum(send[i])[:-1]])
    recv = m[:, :, rank].T.copy()                    # [S, R(src)]
    roff = np.zeros((S, R), dtype=np.int64)
    base = 0
    for src in range(R):
        acc = base
        for i in range(S):
            roff[i, src] = acc
            acc += int(recv[i, src])
        base = acc
    return SlabPlan(sr, soff, send.copy(), roff, recv, int(recv.sum()))


def worst_case_capacity(rows_per_rank: int, world: int, slack: float = 0.02, pad: int = 1_000_000) -> int:
    """Receive-buffer rows for a uniform hash shuffle of `rows_per_rank` rows per rank (binomial spread is ~ sqrt(n))."""
    return rows_per_rank if world == 1 else int(rows_per_rank * (1.0 + slack)) + pad


def weak_scaling_build_keys(perm_local: np.ndarray, rank: int, world: int) -> np.ndarray:
    """Rank r owns keys {r, r+W, r+2W, ...} of the global key space [0, W * n): a disjoint cover for any W."""
    return perm_local * world + rank


def global_key_space(build_rows_per_rank: int, world: int) -> int:
    return build_rows_per_rank * world
