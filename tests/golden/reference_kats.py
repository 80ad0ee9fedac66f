"""Known-answer tests ported from the reference's own operator unit tests (the only artefacts in
/root/reference that pin *results* for this hot path — SURVEY.md §8c).

Each case keeps the literal input chunks and expected rows of the Java test it cites; nothing here was
produced by running code.  Conventions:

* a chunk is a list of columns, a column is a Python list, ``None`` = SQL NULL;
* column types: "int" (IntegerBlock), "long" (LongBlock), "double" (DoubleBlock), "str" (StringBlock);
* "str" columns are dictionary-encoded to INT32 codes by :func:`encode_case` before they reach the oracle
  or the GPU (v1 of the path has INT32/INT64/FP64 blocks).  The encoding is injective per case, and join /
  group-by semantics depend only on equality and NULL-ness, so the expected multisets carry over unchanged;
* comparison is an order-insensitive row multiset, exactly like ``BaseExecTest.assertExecResultByRow``
  (polardbx-executor/src/test/java/com/alibaba/polardbx/executor/operator/BaseExecTest.java:78-103, order=false).

HT = polardbx-executor/src/test/java/com/alibaba/polardbx/executor/operator/HashJoinTest.java
AT = .../operator/HashAggExecTest.java ; ST = .../operator/SpilledHashAggExecTest.java
MT = .../operator/util/ChunkRowOpenHashMapTest.java
"""

# ---- shared inputs (the Java tests repeat these literals) ------------------------------------------------
_OUTER_SIMPLE = [
    [[0, 1, 2, 3], [3, 4, 9, 7]],
    [[4, 5, 6, 7], [5, 3, 8, 1]],
]
_INNER_SIMPLE = [
    [[1, 2, 3, 4], ["a", "b", "c", None]],
    [[5, 6, 7, 8], ["d", "e", "f", None]],
]
_OUTER_MULTI = [
    [[0, 1, 2, 3, 4], [1, 1, 2, 2, None], ["a", "b", "a", "b", "a"]],
    [[5, 6, 7, 8, 9], [3, 3, 4, 4, 4], ["a", "b", "a", "b", None]],
]
_INNER_MULTI = [
    [[1, 2, 3, 4], ["a", "a", "a", None], ["A", "B", "C", "D"]],
    [[1, 2, 3, None], ["a", "b", "c", "b"], ["E", "F", "G", "H"]],
]
_OUTER_ANTI = [
    [[0, 1, 2, 3], [3, 4, 9, 7]],
    [[4, 5, 6, 7], [5, 3, 8, None]],
]
_INNER_SEMI = [
    [[1, 2, 3, 4]],
    [[3, 4, 5, 6]],
]
_INNER_SINGLE = [
    [["a", "b", "c", None], [1, 2, 3, 4]],
    [["d", "e", "f", None], [5, 6, 7, 8]],
]
_INNER_SINGLE_DUP = [
    [["a", "b", "c", None], [1, 2, 3, 4]],
    [["d", "e", "f", None], [4, 5, 6, 7]],
]

JOIN_KATS = [
    dict(name="testInnerJoin_Simple", src="HT:119-166",
         outer_types=["int", "int"], outer=_OUTER_SIMPLE, inner_types=["int", "str"], inner=_INNER_SIMPLE,
         join_type="INNER", keys=[(1, 0, "int")],
         expect=[[0, 1, 3, 4, 5, 6, 7], [3, 4, 7, 5, 3, 8, 1], [3, 4, 7, 5, 3, 8, 1],
                 ["c", None, "f", "d", "c", None, "a"]]),
    dict(name="testInnerJoin_MultiKey", src="HT:169-225",
         outer_types=["int", "int", "str"], outer=_OUTER_MULTI, inner_types=["int", "str", "str"], inner=_INNER_MULTI,
         join_type="INNER", keys=[(1, 0, "int"), (2, 1, "str")],
         expect=[[0, 0, 2, 3, 5], [1, 1, 2, 2, 3], ["a", "a", "a", "b", "a"], [1, 1, 2, 2, 3],
                 ["a", "a", "a", "b", "a"], ["E", "A", "B", "F", "C"]]),
    dict(name="testLeftOuterJoin_Simple", src="HT:228-277",
         outer_types=["int", "int"], outer=_OUTER_SIMPLE, inner_types=["int", "str"], inner=_INNER_SIMPLE,
         join_type="LEFT", keys=[(1, 0, "int")],
         expect=[[0, 1, 2, 3, 4, 5, 6, 7], [3, 4, 9, 7, 5, 3, 8, 1], [3, 4, None, 7, 5, 3, 8, 1],
                 ["c", None, None, "f", "d", "c", None, "a"]]),
    dict(name="testLeftOuterJoin_WithCondition", src="HT:280-337",
         outer_types=["int", "int"], outer=_OUTER_SIMPLE, inner_types=["int", "str"], inner=_INNER_SIMPLE,
         join_type="LEFT", keys=[(1, 0, "int")], cond_ne=[(3, "d")],
         expect=[[0, 1, 2, 3, 4, 5, 6, 7], [3, 4, 9, 7, 5, 3, 8, 1], [3, 4, None, 7, None, 3, 8, 1],
                 ["c", None, None, "f", None, "c", None, "a"]]),
    dict(name="testLeftOuterJoin_MultiKey", src="HT:340-391",
         outer_types=["int", "int", "str"], outer=_OUTER_MULTI, inner_types=["int", "str", "str"], inner=_INNER_MULTI,
         join_type="LEFT", keys=[(1, 0, "int"), (2, 1, "str")],
         expect=[[0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9], [1, 1, 1, 2, 2, None, 3, 3, 4, 4, 4],
                 ["a", "a", "b", "a", "b", "a", "a", "b", "a", "b", None],
                 [1, 1, None, 2, 2, None, 3, None, None, None, None],
                 ["a", "a", None, "a", "b", None, "a", None, None, None, None],
                 ["E", "A", None, "B", "F", None, "C", None, None, None, None]]),
    dict(name="testRightOuterJoin_Simple", src="HT:394-443",
         outer_types=["int", "int"], outer=_OUTER_SIMPLE, inner_types=["int", "str"], inner=_INNER_SIMPLE,
         join_type="RIGHT", keys=[(1, 0, "int")],
         expect=[[3, 4, None, 7, 5, 3, 8, 1], ["c", None, None, "f", "d", "c", None, "a"],
                 [0, 1, 2, 3, 4, 5, 6, 7], [3, 4, 9, 7, 5, 3, 8, 1]]),
    dict(name="testRightOuterJoin_MultiKey", src="HT:446-500",
         outer_types=["int", "int", "str"], outer=_OUTER_MULTI, inner_types=["int", "str", "str"], inner=_INNER_MULTI,
         join_type="RIGHT", keys=[(1, 0, "int"), (2, 1, "str")],
         expect=[[1, 1, None, 2, 2, None, 3, None, None, None, None],
                 ["a", "a", None, "a", "b", None, "a", None, None, None, None],
                 ["E", "A", None, "B", "F", None, "C", None, None, None, None],
                 [0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9], [1, 1, 1, 2, 2, None, 3, 3, 4, 4, 4],
                 ["a", "a", "b", "a", "b", "a", "a", "b", "a", "b", None]]),
    dict(name="testSemiJoin", src="HT:503-545",
         outer_types=["int", "int"], outer=_OUTER_SIMPLE, inner_types=["int"], inner=_INNER_SEMI,
         join_type="SEMI", keys=[(1, 0, "int")],
         expect=[[0, 1, 4, 5, 7], [3, 4, 5, 3, 1]]),
    dict(name="testSemiJoin_InnerEmpty", src="HT:548-574",
         outer_types=["int", "int"], outer=[[[0, 1, 2, 3], [3, None, 9, None]]], inner_types=["int"], inner=[],
         join_type="SEMI", keys=[(1, 0, "int")],
         expect=[[], []]),
    dict(name="testAntiJoin_NotExists", src="HT:577-619",
         outer_types=["int", "int"], outer=_OUTER_ANTI, inner_types=["int"], inner=_INNER_SEMI,
         join_type="ANTI", keys=[(1, 0, "int")],
         expect=[[2, 3, 6, 7], [9, 7, 8, None]]),
    dict(name="testAntiJoin_NotIn", src="HT:622-669",
         outer_types=["int", "int"], outer=_OUTER_ANTI, inner_types=["int"], inner=_INNER_SEMI,
         join_type="ANTI", keys=[(1, 0, "int")], anti_operands=[1],
         expect=[[2, 3, 6], [9, 7, 8]]),
    dict(name="testAntiJoin_NotIn_InnerEmpty", src="HT:672-706",
         outer_types=["int", "int"], outer=[[[4, 5, 6, 7], [5, None, 8, None]]], inner_types=["int"], inner=[],
         join_type="ANTI", keys=[(1, 0, "int")], anti_operands=[1],
         expect=[[4, 5, 6, 7], [5, None, 8, None]]),
    dict(name="testAntiJoin_NotIn_InnerContainsNull", src="HT:709-749",
         outer_types=["int", "int"], outer=_OUTER_ANTI, inner_types=["int"],
         inner=[[[1, 2, 3, 4]], [[3, None, 5, 6]]],
         join_type="ANTI", keys=[(1, 0, "int")], anti_operands=[1],
         expect=[[], []]),
    dict(name="testAntiJoin_WithCondition", src="HT:752-803",
         outer_types=["int", "int"], outer=_OUTER_ANTI, inner_types=["int"], inner=_INNER_SEMI,
         join_type="ANTI", keys=[(1, 0, "int")], cond_ne=[(2, 5)],
         expect=[[2, 3, 4, 6, 7], [9, 7, 5, 8, None]]),
    dict(name="testInnerSingleJoin", src="HT:806-852",
         outer_types=["int", "int"], outer=_OUTER_SIMPLE, inner_types=["str", "int"], inner=_INNER_SINGLE,
         join_type="INNER", max_one_row=True, keys=[(1, 1, "int")],
         expect=[[0, 1, 3, 4, 5, 6, 7], [3, 4, 7, 5, 3, 8, 1], ["c", None, "f", "d", "c", None, "a"]]),
    dict(name="testInnerSingleJoin_withError", src="HT:855-905",
         outer_types=["int", "int"], outer=_OUTER_SIMPLE, inner_types=["str", "int"], inner=_INNER_SINGLE_DUP,
         join_type="INNER", max_one_row=True, keys=[(1, 1, "int")],
         expect_error="ERR_SCALAR_SUBQUERY_RETURN_MORE_THAN_ONE_ROW"),
    dict(name="testLeftSingleJoin", src="HT:908-954",
         outer_types=["int", "int"], outer=_OUTER_SIMPLE, inner_types=["str", "int"], inner=_INNER_SINGLE,
         join_type="LEFT", max_one_row=True, keys=[(1, 1, "int")],
         expect=[[0, 1, 2, 3, 4, 5, 6, 7], [3, 4, 9, 7, 5, 3, 8, 1],
                 ["c", None, None, "f", "d", "c", None, "a"]]),
    dict(name="testLeftSingleJoin_WithCondition", src="HT:957-1011",
         outer_types=["int", "int"], outer=_OUTER_SIMPLE, inner_types=["str", "int"], inner=_INNER_SINGLE,
         join_type="LEFT", max_one_row=True, keys=[(1, 1, "int")], cond_ne=[(2, "d")],
         expect=[[0, 1, 2, 3, 4, 5, 6, 7], [3, 4, 9, 7, 5, 3, 8, 1],
                 ["c", None, None, "f", None, "c", None, "a"]]),
    dict(name="testLeftSingleJoin_withError", src="HT:1014-1064",
         outer_types=["int", "int"], outer=_OUTER_SIMPLE, inner_types=["str", "int"], inner=_INNER_SINGLE_DUP,
         join_type="LEFT", max_one_row=True, keys=[(1, 1, "int")],
         expect_error="ERR_SCALAR_SUBQUERY_RETURN_MORE_THAN_ONE_ROW"),
]

# ---- aggregation -----------------------------------------------------------------------------------------
_AGG_IN = [
    [[0, 1, 2, 3], [3, 4, 9, 7]],
    [[0, 1, 2, 3], [5, 3, 8, 1]],
]
_AGG_IN_NULL = [
    [[0, 1, 2, 3], [None, None, None, None]],
    [[0, 1, 2, 3], [None, None, None, None]],
]


def _seq_chunk(length, *initial):
    """SequenceChunkBuilder.createSequenceChunk: column c holds initial[c] + i, i in [0, length)."""
    return [[init + i for i in range(length)] for init in initial]


AGG_KATS = [
    dict(name="testHashAggSimpleCount", src="AT:93-127", types=["int", "int"], chunks=_AGG_IN, groups=[0],
         aggs=[("COUNT", [1])], expected_groups=1024,
         expect=[[0, 1, 2, 3], [2, 2, 2, 2]]),
    dict(name="testHashAggSimpleSum", src="AT:130-163", types=["int", "int"], chunks=_AGG_IN, groups=[0],
         aggs=[("SUM", [1])], expected_groups=1024,
         expect=[[0, 1, 2, 3], [8, 7, 17, 8]]),
    dict(name="testHashAggDoubleNullAvg", src="AT:245-276", types=["int", "double"], chunks=_AGG_IN_NULL, groups=[0],
         aggs=[("AVG", [1])], expected_groups=1024,
         expect=[[0, 1, 2, 3], [None, None, None, None]]),
    dict(name="testHashAggNullSum", src="AT:279-310", types=["int", "int"], chunks=_AGG_IN_NULL, groups=[0],
         aggs=[("SUM", [1])], expected_groups=1024,
         expect=[[0, 1, 2, 3], [None, None, None, None]]),
    dict(name="testHashAggNullCount", src="AT:313-337", types=["int", "int"], chunks=_AGG_IN_NULL, groups=[0],
         aggs=[("COUNT", [1])], expected_groups=1024,
         expect=[[0, 1, 2, 3], [0, 0, 0, 0]]),
    dict(name="testMemoryHashAggr (NULL group key)", src="ST:67-83", types=["int", "int"],
         chunks=[[[None, 1, 2, 3], [1, 2, 3, 4]]], groups=[0], aggs=[("COUNT", [1])], expected_groups=100,
         expect=[[None, 1, 2, 3], [1, 1, 1, 1]]),
    dict(name="testSpillHashAggr (20 chunks x 2 rows)", src="ST:86-110", types=["int", "int"],
         chunks=[[[0, 1], [0, 1]]] * 20, groups=[0], aggs=[("COUNT", [1])], expected_groups=100,
         expect=[[0, 1], [20, 20]]),
    dict(name="testSpillHashAggr2 (20 chunks x 1024 rows)", src="ST:113-140", types=["int", "int"],
         chunks=[[list(range(1024)), list(range(1024))]] * 20, groups=[0], aggs=[("COUNT", [1])], expected_groups=100,
         expect=[list(range(1024)), [20] * 1024]),
    dict(name="testSpillHashAggr3 (string keys)", src="ST:143-171", types=["str", "int"],
         chunks=[[["a%d" % j for j in range(1024)], list(range(1024))]] * 20, groups=[0], aggs=[("COUNT", [1])],
         expected_groups=100,
         expect=[["a%d" % j for j in range(1024)], [20] * 1024]),
]

# testSpillHashAggWithCountV2 (ST:174-225) asserts spilled == in-memory execution on sequence chunks; no literal
# expectation exists, so it is ported as an input shape whose expectation comes from a brute-force group-by
# (tests/test_oracle_golden.py::test_agg_sequence_chunks).
AGG_SEQUENCE_INPUT = dict(
    src="ST:174-225", types=["long", "long"],
    chunks=[_seq_chunk(1024, 10000, 10000), _seq_chunk(1024, 22000, 13000), _seq_chunk(1024, 30000, 13000),
            _seq_chunk(1024, 40000, 15000), _seq_chunk(200, 70, 70), _seq_chunk(1024, 50000, 16000),
            _seq_chunk(1024, 60000, 17000), _seq_chunk(1024, 80000, 19000), _seq_chunk(10240, 80000, 19000)],
    groups=[0], agg_sets=[[("COUNT", [1])], [("SUM", [1])], [("SUM", [1]), ("COUNT", [1])]])

# ---- ChunkRowOpenHashMapTest.test (MT:33-69): exact put / get position vectors -----------------------------
CHUNK_ROW_OPEN_HASH_MAP = dict(
    src="MT:33-69",
    build=[[6, 5, 4, 3, 2, 1, 0, 3, 2, 1, 0, 6, 5, 4], [7, 8, 9, 4, 2, 6, 2, 8, 9, 4, 2, 6, 2, 7]],
    probe=[[6, 4, 0, 4, 4, 1, 0, 3], [7, 9, 2, 9, 7, 4, 0, 8]],
    expected_put=[-1, -1, -1, -1, -1, -1, -1, -1, -1, -1, 6, -1, -1, -1],
    expected_get=[0, 2, 10, 2, 13, 9, -1, 7],
)

# HashJoinTest.java:201-208 documents LIFO chain order for a duplicate build key: the later build row ("E",
# row 4) is emitted before the earlier one ("A", row 0) for probe row 0 — checked in test_oracle_golden.py.


# ---- helpers ----------------------------------------------------------------------------------------------
def encode_case_strings(*chunk_lists_and_expect):
    """Build one injective str->int32 dictionary over every string in the case (sorted for determinism)."""
    seen = set()

    def walk(x):
        if isinstance(x, str):
            seen.add(x)
        elif isinstance(x, (list, tuple)):
            for y in x:
                walk(y)

    for item in chunk_lists_and_expect:
        walk(item)
    return {s: 1000 + i for i, s in enumerate(sorted(seen))}
