"""MPP plan fragments composed from the GPU operators: the shapes polardbx-optimizer emits around the hot path.

* ShuffledJoin  — hash/hash-distributed join: both inputs repartitioned on the join key, joined locally
                  (MppHashJoinConvertRule hash distribution; RuleUtils.ensureKeyDataTypeDistribution).
* TwoPhaseAgg   — partial HashAgg -> hash exchange on the group keys -> final HashAgg, or raw-row shuffle + one HashAgg,
                  chosen like MppHashAggConvertRule.tryConvertToPartialAgg:125-147 (selectivity / bucket thresholds);
                  aggregate calls are split like CBOPushAggRule.splitAgg:236-330.
* Q3Pipeline    — TPC-H Q3's MPP plan (MppTpchPlan100gTest.yml:124-135): broadcast of the filtered customer keys,
                  customer x orders, exchange on the order key, x lineitem, group-by with SUM(price*(1-discount)).

Everything that computes runs in libgsql_gpu.so (exchange pushes over NVLink peer memory, joins, aggregations, the
vectorised filter/project); this module only sequences the calls — it is the host-side plan fragment, the role the
reference's LocalExecutionPlanner / PlanFragmenter play.  torch is used for device buffers and process-group plumbing.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

from . import api, native as N

try:
    import torch
except Exception:  # pragma: no cover
    torch = None


def _world(ctx) -> Tuple[int, int]:
    """(nranks, rank) of the context's communicator (1, 0 without gsql_comm_init)."""
    return ctx.nranks, ctx.rank


def _alloc(ctx, types: Sequence[int], rows: int, nullable: Sequence[bool]):
    return api._alloc_out(ctx, types, rows, N.MEM_DEVICE, list(nullable))


class ShuffledJoin:
    """Repartition both sides on the equi-join key with the push exchange, join locally.  The probe side travels in
    `nslabs` slabs: slab k is probed on the context stream while slab k+1 is still crossing NVLink."""

    def __init__(self, ctx: api.Context, join_type: int, outer_types: Sequence[int], inner_types: Sequence[int],
                 outer_keys: Sequence[int], inner_keys: Sequence[int], build_capacity: int, probe_capacity: int,
                 nslabs: int = 4, key_types: Optional[Sequence[int]] = None, outer_nullable: Sequence[int] = (),
                 inner_nullable: Sequence[int] = ()):
        self.ctx, self.join_type = ctx, join_type
        self.outer_types, self.inner_types = list(outer_types), list(inner_types)
        self.outer_keys, self.inner_keys = list(outer_keys), list(inner_keys)
        self.key_types = list(key_types) if key_types else [outer_types[k] for k in outer_keys]
        self.nslabs = nslabs
        world = ctx.nranks
        # partition channels = the join keys converted to the unified key type (RuleUtils.ensureKeyDataTypeDistribution)
        self.xb = api.Exchange(ctx, inner_types, inner_keys, world, key_types=self.key_types)
        self.xb.open_p2p(build_capacity, nullable=inner_nullable)
        self.xp = api.Exchange(ctx, outer_types, outer_keys, world, key_types=self.key_types)
        self.xp.open_p2p(probe_capacity, nullable=outer_nullable)
        self.probe_capacity = probe_capacity
        self.last_info = None
        self.last_rows = 0
        self.out_cols = None

    def close(self):
        for x in (self.xb, self.xp):
            x.close()

    def run(self, probe_cols, build_cols, out_cols=None, out_capacity: Optional[int] = None):
        """Collective.  Returns the joined rows this rank produced as device columns (trimmed views of out_cols)."""
        ctx = self.ctx
        self.xb.push(build_cols, 1)
        slab_rows = self.xp.push(probe_cols, self.nslabs)   # on the wire while the table is being built
        build = self.xb.recv(-1)
        j = api.HashJoin(ctx, self.join_type, self.outer_types, self.inner_types, self.outer_keys, self.inner_keys,
                         key_types=self.key_types, expected_build_rows=int(build[0][0].numel()))
        try:
            j.build_consume_ref(build)
            j.build_finish()
            total_in = sum(slab_rows)
            if out_cols is None:
                # every operator of this family emits at most max(1, matches) rows per probe row only for unique build keys;
                # the caller passes buffers for anything else
                cap = out_capacity if out_capacity is not None else max(total_in, 1)
                outer = self.join_type in (N.JOIN_LEFT, N.JOIN_RIGHT)
                out_cols = _alloc(ctx, j.out_types, cap, [outer] * len(j.out_types))
            else:
                cap = out_capacity if out_capacity is not None else int(out_cols[0][0].numel())
            off = 0
            for k in range(self.nslabs):
                if slab_rows[k] == 0:
                    continue
                view = self.xp.recv(k)
                dst = [(d[off:], None if nl is None else nl[off:]) for d, nl in out_cols]
                off += j.probe_into(view, dst, cap - off)
            self.last_info = j.info()
        finally:
            j.close()
        self.last_rows = off
        self.out_cols = out_cols
        return [(d[:off], None if nl is None else nl[:off]) for d, nl in out_cols]
