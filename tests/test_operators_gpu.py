"""Operator-level tests through the mirrored Executor / ConsumerExecutor interface — they read like the reference's
HashJoinTest / HashAggExecTest (MockExec sources, IntegerBlock.of(...), CHUNK_SIZE = 2 to exercise output re-slicing,
order-insensitive row comparison as in BaseExecTest.assertExecResultByRow)."""
from collections import Counter

import numpy as np
import pytest

from tests import kat_util as ku
from tests.golden import reference_kats as kats

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from galaxysql_b200 import operators
    return operators


def assertExecResultByRow(actual_chunks, expected_chunk_rows):
    got = Counter()
    for ch in actual_chunks:
        got.update(ch.rows())
    assert got == Counter(expected_chunk_rows)


def test_testInnerJoin_Simple(ops):
    """HashJoinTest.java:119-166 with the StringBlock payload dictionary-encoded (a=1 .. f=6)."""
    o = ops
    context = o.ExecutionContext(chunk_size=2)          # HashJoinTest.java:73
    outerInput = o.MockExec.builder(o.DataTypes.IntegerType, o.DataTypes.IntegerType) \
        .withChunk(o.Chunk(o.IntegerBlock.of(0, 1, 2, 3), o.IntegerBlock.of(3, 4, 9, 7))) \
        .withChunk(o.Chunk(o.IntegerBlock.of(4, 5, 6, 7), o.IntegerBlock.of(5, 3, 8, 1))).build()
    innerInput = o.MockExec.builder(o.DataTypes.IntegerType, o.DataTypes.IntegerType) \
        .withChunk(o.Chunk(o.IntegerBlock.of(1, 2, 3, 4), o.IntegerBlock.of(1, 2, 3, None))) \
        .withChunk(o.Chunk(o.IntegerBlock.of(5, 6, 7, 8), o.IntegerBlock.of(4, 5, 6, None))).build()
    joinKeys = [o.EquiJoinKey(1, 0, o.DataTypes.IntegerType)]
    exec_ = o.GpuParallelHashJoinExec(outerInput, innerInput, o.JoinRelType.INNER, False, joinKeys, None, None, False, context)
    test = o.SingleExecTest(exec_, innerInput).exec()
    assert all(ch.getPositionCount() <= 2 for ch in test.result())
    assertExecResultByRow(test.result(), list(zip([0, 1, 3, 4, 5, 6, 7], [3, 4, 7, 5, 3, 8, 1], [3, 4, 7, 5, 3, 8, 1],
                                                  [3, None, 6, 4, 3, None, 1])))


_TYPE = {"int": "IntegerType", "str": "IntegerType", "long": "LongType", "double": "DoubleType"}
_BLOCK = {"int": "IntegerBlock", "str": "IntegerBlock", "long": "LongBlock", "double": "DoubleBlock"}


def _mock(o, types, chunks, sdict):
    b = o.MockExec.builder(*[getattr(o.DataTypes, _TYPE[t]) for t in types])
    for ch in chunks:
        blocks = [getattr(o, _BLOCK[t]).of(*[(sdict[v] if isinstance(v, str) else v) for v in col]) for t, col in zip(types, ch)]
        b.withChunk(o.Chunk(*blocks))
    return b.build()


@pytest.mark.parametrize("case", kats.JOIN_KATS, ids=[c["name"] for c in kats.JOIN_KATS])
def test_hash_join_kats_through_operator_interface(ops, case):
    o = ops
    sdict = kats.encode_case_strings(case["outer"], case["inner"], case.get("expect", []), [v for _, v in case.get("cond_ne", [])])
    outerInput = _mock(o, case["outer_types"], case["outer"], sdict)
    innerInput = _mock(o, case["inner_types"], case["inner"], sdict)
    jt = getattr(o.JoinRelType, case["join_type"])
    keys = [o.EquiJoinKey(a, b, getattr(o.DataTypes, _TYPE[t])) for a, b, t in case["keys"]]
    cond = [(c, sdict[v] if isinstance(v, str) else v) for c, v in case.get("cond_ne", [])]
    exec_ = o.GpuParallelHashJoinExec(outerInput, innerInput, jt, case.get("max_one_row", False), keys, cond,
                                      case.get("anti_operands"), False, o.ExecutionContext(chunk_size=2))
    test = o.SingleExecTest(exec_, innerInput)
    if case.get("expect_error"):
        with pytest.raises(o.TddlRuntimeException) as ei:
            test.exec()
        assert ei.value.error_code == o.ErrorCode.ERR_SCALAR_SUBQUERY_RETURN_MORE_THAN_ONE_ROW
        return
    test.exec()
    got = Counter()
    for ch in test.result():
        got.update(ch.rows())
    assert got == ku.expect_multiset(case["expect"], sdict)


def test_testHashAggSimpleCount_and_Sum(ops):
    """HashAggExecTest.java:93-163."""
    o = ops
    chunks = [o.Chunk(o.IntegerBlock.of(0, 1, 2, 3), o.IntegerBlock.of(3, 4, 9, 7)),
              o.Chunk(o.IntegerBlock.of(0, 1, 2, 3), o.IntegerBlock.of(5, 3, 8, 1))]
    types = [o.DataTypes.IntegerType, o.DataTypes.IntegerType]
    exec_ = o.GpuHashAggExec(types, [0], [o.Count([1], False, -1)], None, 1024, o.ExecutionContext(chunk_size=2))
    test = o.SingleExecTest(exec_, None, chunks).exec()
    assertExecResultByRow(test.result(), list(zip([0, 1, 2, 3], [2, 2, 2, 2])))
    exec_ = o.GpuHashAggExec(types, [0], [o.Sum(1, False, o.DataTypes.DecimalType, -1)], None, 1024, o.ExecutionContext())
    test = o.SingleExecTest(exec_, None, chunks).exec()
    assertExecResultByRow(test.result(), list(zip([0, 1, 2, 3], [8, 7, 17, 8])))


def test_partitioning_exchanger_feeds_parallel_aggs(ops):
    """LocalExchange(PARTITION on group keys) -> P HashAggExec consumers (LocalExecutionPlanner.visitHashAgg:1487-1535):
    the union of the consumers' results equals a single aggregation, and no key lands on two consumers."""
    o = ops
    n = 20000
    k = (ku.rand_u64(n, 5) % np.uint64(700)).astype(np.int32)
    v = (ku.rand_u64(n, 6) % np.uint64(50)).astype(np.int64)
    types = [o.DataTypes.IntegerType, o.DataTypes.LongType]
    ctx = o.ExecutionContext(chunk_size=1000, gpu_batch_rows=4096)
    aggs = [o.GpuHashAggExec(types, [0], [o.CountRow(), o.Sum0(1)], None, 64, ctx) for _ in range(4)]
    ex = o.GpuPartitioningExchanger(aggs, types, [0], ctx)
    ex.openConsume()
    for lo in range(0, n, 1000):
        ex.consumeChunk(o.Chunk(o.IntegerBlock(k[lo:lo + 1000]), o.LongBlock(v[lo:lo + 1000])))
    ex.buildConsume()
    seen, got = set(), Counter()
    for a in aggs:
        rows = []
        while True:
            ch = a.nextChunk()
            if ch is None:
                break
            rows.extend(ch.rows())
        keys = {r[0] for r in rows}
        assert not (keys & seen)
        seen |= keys
        got.update(rows)
    exp = Counter()
    for kk in np.unique(k):
        m = k == kk
        exp[(int(kk), int(m.sum()), int(v[m].sum()))] += 1
    assert got == exp
    ex.closeConsume(True)
