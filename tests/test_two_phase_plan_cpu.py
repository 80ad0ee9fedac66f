"""CPU checks of the two-phase aggregation plan's host logic (galaxysql_b200/pipelines.py), with the oracle as the engine:
the partial / final call lists produced by split_agg_calls (CBOPushAggRule.splitAgg:236-330), run as partial aggregation
per "rank" -> repartition of the partial rows by group key -> final aggregation, must equal the single-phase aggregation
over all rows.  The final-stage AVG (GSQL_AGG_AVG_MERGE = global_sum / global_count) is restated here with numpy on top of
the oracle's SUM / SUM0 because the oracle — a restatement of the reference's single-phase operators — has no such kind."""
import numpy as np

from galaxysql_b200 import native as N
from galaxysql_b200 import pipelines
from oracle import oracle as orc
from tests import kat_util as ku

_ORC = {N.AGG_COUNT_STAR: orc.AGG_COUNT_STAR, N.AGG_COUNT: orc.AGG_COUNT, N.AGG_SUM: orc.AGG_SUM, N.AGG_AVG: orc.AGG_AVG, N.AGG_MIN: orc.AGG_MIN,
        N.AGG_MAX: orc.AGG_MAX, N.AGG_SUM0: orc.AGG_SUM0}


def test_split_follows_split_agg():
    p, f = pipelines.split_agg_calls(2, [(N.AGG_COUNT_STAR, []), (N.AGG_AVG, [3]), (N.AGG_SUM, [3]), (N.AGG_MAX, [2]), (N.AGG_COUNT, [3])],
                                     [N.T_INT32, N.T_INT64, N.T_INT32, N.T_FP64])
    # partial columns sit after the 2 group keys: 2 = COUNT(*), 3 / 4 = SUM, COUNT of the AVG, 5 = SUM, 6 = MAX, 7 = COUNT
    assert p == [(N.AGG_COUNT_STAR, []), (N.AGG_SUM, [3]), (N.AGG_COUNT, [3]), (N.AGG_SUM, [3]), (N.AGG_MAX, [2]), (N.AGG_COUNT, [3])]
    assert f == [(N.AGG_SUM0, [2]), (N.AGG_AVG_MERGE, [3, 4]), (N.AGG_SUM, [5]), (N.AGG_MAX, [6]), (N.AGG_SUM0, [7])]
    # SUM over integers yields DECIMAL partials (output-only on the GPU path): such a plan is not split
    assert pipelines.split_agg_calls(1, [(N.AGG_SUM, [1])], [N.T_INT64, N.T_INT64]) is None


def _final_with_avg_merge(cols, nkeys, final_calls):
    """Final stage over the concatenated partial rows with the oracle; AVG_MERGE = SUM(partial sums) / SUM0(partial counts),
    NULL when the count is 0 (Avg.java: a group that only saw NULLs)."""
    ocalls, shape = [], []
    for kind, c in final_calls:
        if kind == N.AGG_AVG_MERGE:
            shape.append(("avg", len(ocalls)))
            ocalls += [orc.AggCall(orc.AGG_SUM, [c[0]]), orc.AggCall(orc.AGG_SUM0, [c[1]])]
        else:
            shape.append(("plain", len(ocalls)))
            ocalls.append(orc.AggCall(_ORC[kind], list(c)))
    res = orc.hash_agg(cols, list(range(nkeys)), ocalls, 64)
    out = list(res[:nkeys])
    for what, at in shape:
        if what == "plain":
            out.append(res[nkeys + at])
        else:
            (s, sn), (c, _) = res[nkeys + at], res[nkeys + at + 1]
            cnt = np.asarray(c, dtype=np.int64)
            sn = np.zeros(len(cnt), bool) if sn is None else np.asarray(sn, bool)
            nul = sn | (cnt == 0)
            with np.errstate(divide="ignore", invalid="ignore"):
                out.append((np.where(nul, 0.0, np.asarray(s, dtype=np.float64) / np.maximum(cnt, 1)), nul))
    return out


def test_partial_then_final_equals_single_phase_on_the_oracle():
    n, world = 60_000, 3
    k0 = (ku.rand_u64(n, 1) % np.uint64(37)).astype(np.int32)
    k1 = (ku.rand_u64(n, 2) % np.uint64(5)).astype(np.int64) - 2
    v = ku.with_nulls((ku.rand_u64(n, 3) % np.uint64(100000)).astype(np.float64) / 8.0, 0.2, 4)   # exact binary fractions: sums are order-independent
    w = ku.with_nulls((ku.rand_u64(n, 5) % np.uint64(1000)).astype(np.int32) - 500, 0.1, 6)
    cols = [(k0, None), (k1, None), v, w]
    types = [N.T_INT32, N.T_INT64, N.T_FP64, N.T_INT32]
    aggs = [(N.AGG_COUNT_STAR, []), (N.AGG_AVG, [2]), (N.AGG_SUM, [2]), (N.AGG_MIN, [3]), (N.AGG_MAX, [3]), (N.AGG_COUNT, [2, 3])]
    partial, final = pipelines.split_agg_calls(2, aggs, types)
    single = orc.hash_agg(cols, [0, 1], [orc.AggCall(_ORC[k], list(c)) for k, c in aggs], 64)
    # phase 1: every "rank" aggregates its slice of the rows
    parts = []
    for r in range(world):
        sl = slice(r * n // world, (r + 1) * n // world)
        rc = [(d[sl], None if nl is None else nl[sl]) for d, nl in cols]
        parts.append(orc.hash_agg(rc, [0, 1], [orc.AggCall(_ORC[k], list(c)) for k, c in partial], 64))
    # the exchange: partial rows of all ranks, repartitioned on the group keys; rank by rank the final stage
    allp = [(np.concatenate([np.asarray(p[c][0]) for p in parts]),
             np.concatenate([np.zeros(len(p[c][0]), bool) if p[c][1] is None else np.asarray(p[c][1], bool) for p in parts])) for c in range(len(parts[0]))]
    _, counts = orc.partition_exchange(allp, [0, 1], world)
    dest_cols, _ = orc.partition_exchange(allp, [0, 1], world)
    got_rows = []
    off = 0
    for r in range(world):
        rows = [(np.asarray(d)[off:off + counts[r]], np.asarray(nl)[off:off + counts[r]]) for d, nl in dest_cols]
        off += counts[r]
        if counts[r]:
            got_rows.append(_final_with_avg_merge(rows, 2, final))
    got = [(np.concatenate([np.asarray(g[c][0]) for g in got_rows]),
            np.concatenate([np.zeros(len(g[c][0]), bool) if g[c][1] is None else np.asarray(g[c][1], bool) for g in got_rows])) for c in range(len(got_rows[0]))]
    # a group lives on exactly one rank after the repartition: the union of the ranks' results is the single-phase result
    assert len(got[0][0]) == len(single[0][0])
    assert ku.rows_multiset(got) == ku.rows_multiset(single)


def test_q3_scan_programs_have_the_shape_the_specialised_kernel_recognises():
    """scan.cu's scan_fast_plan matches `COL c, CONST_I64 k, <cmp>` filters and `COL` / `COL a, CONST_F64 1.0, COL b, SUB, MUL`
    outputs.  The Q3 pipeline builds its programs with api.E: if the builder ever emitted another instruction order the
    pipeline would silently fall back to the bytecode kernel (2.3x slower scans) — pin the sequences here."""
    from galaxysql_b200 import api
    E = api.E
    assert (E.col(1) * (1.0 - E.col(2))).ins == [(N.OP_COL, 1, 0), (N.OP_CONST_F64, 0, 1.0), (N.OP_COL, 2, 0), (N.OP_SUB, 0, 0), (N.OP_MUL, 0, 0)]
    assert (E.col(3) > pipelines.Q3_DATE).ins == [(N.OP_COL, 3, 0), (N.OP_CONST_I64, 0, pipelines.Q3_DATE), (N.OP_GT, 0, 0)]
    assert (E.col(2) < pipelines.Q3_DATE).ins == [(N.OP_COL, 2, 0), (N.OP_CONST_I64, 0, pipelines.Q3_DATE), (N.OP_LT, 0, 0)]
    assert E.col(1).eq(pipelines.Q3_SEGMENT).ins == [(N.OP_COL, 1, 0), (N.OP_CONST_I64, 0, pipelines.Q3_SEGMENT), (N.OP_EQ, 0, 0)]
    assert E.col(0).ins == [(N.OP_COL, 0, 0)]
