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


# ------------------------------------------------------------------------------------------------ two-phase aggregation
def split_agg_calls(nkeys: int, aggs: Sequence[Tuple[int, Sequence[int]]], in_types: Sequence[int]):
    """CBOPushAggRule.splitAgg:236-330 for the aggregate kinds of the GPU path: -> (partial calls, final calls) where the
    final calls address the partial result's columns (group keys first).  None when a call cannot be split here
    (SUM over integers yields DECIMAL partials, which are output-only on the GPU path)."""
    partial, final = [], []
    for kind, cols in aggs:
        at = nkeys + len(partial)
        if kind in (N.AGG_COUNT_STAR, N.AGG_COUNT):
            partial.append((kind, list(cols)))
            final.append((N.AGG_SUM0, [at]))                      # COUNT -> SUM0 of the partial counts (:239-255)
        elif kind == N.AGG_AVG:
            partial.append((N.AGG_SUM, list(cols)))               # partial_sum, partial_count (:256-296)
            partial.append((N.AGG_COUNT, list(cols)))
            final.append((N.AGG_AVG_MERGE, [at, at + 1]))         # global_sum / global_count (:297-310)
        elif kind == N.AGG_SUM:
            if in_types[cols[0]] != N.T_FP64:
                return None
            partial.append((kind, list(cols)))
            final.append((kind, [at]))
        elif kind in (N.AGG_MIN, N.AGG_MAX, N.AGG_SUM0):
            partial.append((kind, list(cols)))
            final.append((kind, [at]))
        else:
            return None
    return partial, final


class TwoPhaseAgg:
    """GROUP BY across ranks.  mode 'partial': local HashAgg -> push the partial rows on the group keys -> final HashAgg
    (HashAgg.isPartial); mode 'shuffle': push the raw rows on the group keys -> one HashAgg.  Default: the reference's
    choice (MppHashAggConvertRule.tryConvertToPartialAgg:125-147: partial only when groups <= rows *
    PARTIAL_AGG_SELECTIVITY_THRESHOLD (0.2) and groups <= PARTIAL_AGG_BUCKET_THRESHOLD (64))."""

    def __init__(self, ctx: api.Context, input_types: Sequence[int], groups: Sequence[int],
                 aggs: Sequence[Tuple[int, Sequence[int]]], expected_groups: int, capacity: int, mode: Optional[str] = None,
                 nslabs: int = 4, expected_rows: Optional[int] = None, nullable: Sequence[int] = (),
                 selectivity_threshold: float = 0.2, bucket_threshold: int = 64):
        self.ctx = ctx
        self.input_types, self.groups, self.aggs = list(input_types), list(groups), [(k, list(c)) for k, c in aggs]
        self.expected_groups, self.nslabs = expected_groups, nslabs
        world = ctx.nranks
        split = split_agg_calls(len(groups), self.aggs, self.input_types)
        if mode is None:
            rows = expected_rows if expected_rows is not None else capacity
            mode = "partial" if (split is not None and rows * selectivity_threshold >= expected_groups and expected_groups <= bucket_threshold) else "shuffle"
        if mode == "partial" and split is None:
            raise ValueError("these aggregate calls cannot be split into partial + final on the GPU path")
        self.mode = mode
        if mode == "shuffle":
            self.x = api.Exchange(ctx, self.input_types, self.groups, world)
            self.x.open_p2p(capacity, nullable=nullable)
        else:
            self.partial_calls, self.final_calls = split
            # schema of the partial result: group keys, then one column per partial call
            probe = api.HashAgg(ctx, self.input_types, self.groups, self.partial_calls, expected_groups)
            self.partial_types = list(probe.out_types)
            probe.close()
            nk = len(groups)
            self.x = api.Exchange(ctx, self.partial_types, list(range(nk)), world)
            self.x.open_p2p(capacity, nullable=list(range(len(self.partial_types))))   # agg results always carry a mask

    def close(self):
        self.x.close()

    def run(self, cols):
        """Collective.  -> this rank's share of the final groups: (group keys || aggregate values) as device columns."""
        ctx = self.ctx
        if self.mode == "shuffle":
            slab_rows = self.x.push(cols, self.nslabs)
            a = api.HashAgg(ctx, self.input_types, self.groups, self.aggs, max(self.expected_groups // max(ctx.nranks, 1), 1024))
            try:
                for k in range(self.nslabs):
                    if slab_rows[k]:
                        a.consume(self.x.recv(k))        # slab k is aggregated while slab k+1 crosses NVLink
                return a.result(N.MEM_DEVICE)
            finally:
                a.close()
        part = api.HashAgg(ctx, self.input_types, self.groups, self.partial_calls, self.expected_groups)
        try:
            part.consume(cols)
            partial_rows = part.result(N.MEM_DEVICE)
        finally:
            part.close()
        self.x.push(partial_rows, 1)
        recv = self.x.recv(-1)
        fin = api.HashAgg(ctx, self.partial_types, list(range(len(self.groups))), self.final_calls,
                          max(self.expected_groups // max(ctx.nranks, 1), 1024))
        try:
            fin.consume(recv)
            return fin.result(N.MEM_DEVICE)
        finally:
            fin.close()


# ------------------------------------------------------------------------------------------------ TPC-H Q3
Q3_DATE = 9204  # 1995-03-15 as days since 1970-01-01
Q3_SEGMENT = 1  # dictionary code of 'BUILDING'
CUSTOMER_TYPES = [N.T_INT64, N.T_INT32]                       # c_custkey, c_mktsegment
ORDERS_TYPES = [N.T_INT64, N.T_INT64, N.T_INT32, N.T_INT32]   # o_orderkey, o_custkey, o_orderdate, o_shippriority
LINEITEM_TYPES = [N.T_INT64, N.T_FP64, N.T_FP64, N.T_INT32]   # l_orderkey, l_extendedprice, l_discount, l_shipdate


class Q3Pipeline:
    """TPC-H Q3 as the MPP plan of MppTpchPlan100gTest.yml:124-135, tables round-robin over the ranks:

        customer --filter(segment)--> c_custkey --exchange(broadcast)--> build J1
        orders   --filter(o_orderdate < D)--> probe J1 on o_custkey
                 --project(o_orderkey, o_orderdate, o_shippriority)--exchange(hash o_orderkey)--> build J2
        lineitem --filter(l_shipdate > D), project(l_orderkey, price*(1-discount))--exchange(hash l_orderkey, slabs)-->
                 probe J2 --> HashAgg(group l_orderkey, o_orderdate, o_shippriority; SUM(revenue))

    The plan's last exchange (hash[group keys]) moves nothing here: the rows are already distributed on l_orderkey, which
    is one of the group keys, so no group spans two ranks.  Sort / limit sit above the hot path (not built)."""

    def __init__(self, ctx: api.Context, customer_capacity: int, orders_capacity: int, lineitem_capacity: int, nslabs: int = 4,
                 expected_groups: int = 1 << 20):
        E = api.E
        self.ctx, self.nslabs, self.expected_groups = ctx, nslabs, expected_groups
        world = ctx.nranks
        self.scan_c = api.Scan(ctx, CUSTOMER_TYPES, [E.col(0)], filter=E.col(1).eq(Q3_SEGMENT))
        self.scan_o = api.Scan(ctx, ORDERS_TYPES, [E.col(0), E.col(1), E.col(2), E.col(3)], filter=E.col(2) < Q3_DATE)
        self.scan_l = api.Scan(ctx, LINEITEM_TYPES, [E.col(0), E.col(1) * (1.0 - E.col(2))], filter=E.col(3) > Q3_DATE)
        self.xc = api.Exchange(ctx, [N.T_INT64], [0], world, mode=N.XCHG_BROADCAST)
        self.xc.open_p2p(customer_capacity)
        self.xo = api.Exchange(ctx, [N.T_INT64, N.T_INT32, N.T_INT32], [0], world)
        self.xo.open_p2p(orders_capacity)
        self.xl = api.Exchange(ctx, [N.T_INT64, N.T_FP64], [0], world)
        self.xl.open_p2p(lineitem_capacity)
        self.stats = {}

    def close(self):
        for o in (self.scan_c, self.scan_o, self.scan_l, self.xc, self.xo, self.xl):
            o.close()

    def run(self, customer, orders, lineitem):
        """Collective.  -> (l_orderkey, o_orderdate, o_shippriority, revenue) groups owned by this rank (device columns)."""
        ctx = self.ctx
        # the lineitem side does not depend on the joins: filter + project it first and put it on the wire, so that it
        # crosses NVLink while the two tables are being built
        li = self.scan_l.apply(lineitem, nullable_out=False)
        li_slabs = self.xl.push(li, self.nslabs)
        ckeys = self.scan_c.apply(customer, nullable_out=False)
        self.xc.push(ckeys, 1)
        j1 = api.HashJoin(ctx, N.JOIN_INNER, ORDERS_TYPES, [N.T_INT64], [1], [0])
        j2 = None
        agg = None
        try:
            j1.build_consume_ref(self.xc.recv(-1))
            j1.build_finish()
            od = self.scan_o.apply(orders, nullable_out=False)
            oj = j1.probe(od, nullable_out=False)                              # orders of BUILDING customers (+ c_custkey)
            self.xo.push([oj[0], oj[2], oj[3]], 1)                             # project: o_orderkey, o_orderdate, o_shippriority
            j2 = api.HashJoin(ctx, N.JOIN_INNER, [N.T_INT64, N.T_FP64], [N.T_INT64, N.T_INT32, N.T_INT32], [0], [0])
            j2.build_consume_ref(self.xo.recv(-1))
            j2.build_finish()
            # join row: l_orderkey, revenue, o_orderkey, o_orderdate, o_shippriority
            agg = api.HashAgg(ctx, [N.T_INT64, N.T_FP64, N.T_INT64, N.T_INT32, N.T_INT32], [0, 3, 4], [(N.AGG_SUM, [1])], self.expected_groups)
            joined = 0
            for k in range(self.nslabs):
                if li_slabs[k] == 0:
                    continue
                rows = j2.probe(self.xl.recv(k), nullable_out=False)
                joined += int(rows[0][0].shape[0])
                if rows[0][0].shape[0]:
                    agg.consume(rows)
            out = agg.result(N.MEM_DEVICE)
            self.stats = {"customer_keys": int(ckeys[0][0].shape[0]), "orders_after_filter": int(od[0][0].shape[0]),
                          "orders_joined": int(oj[0][0].shape[0]), "lineitem_after_filter": int(li[0][0].shape[0]),
                          "lineitem_received": int(sum(li_slabs)), "joined_rows": joined, "groups": int(out[0][0].shape[0]),
                          "j1_fast": int(j1.info().fast_path), "j2_fast": int(j2.info().fast_path)}
            return out
        finally:
            j1.close()
            if j2 is not None:
                j2.close()
            if agg is not None:
                agg.close()
