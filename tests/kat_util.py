"""Helpers shared by the oracle (CPU) and GPU parity tests: KAT decoding, row multisets, synthetic tables."""
from __future__ import annotations

import os
import sys
from collections import Counter
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import oracle as orc  # noqa: E402  (test infrastructure)
from tests.golden import reference_kats as kats  # noqa: E402

Col = Tuple[np.ndarray, Optional[np.ndarray]]

_NP = {"int": np.int32, "str": np.int32, "long": np.int64, "double": np.float64}
_ORC_T = {"int": orc.T_INT32, "str": orc.T_INT32, "long": orc.T_INT64, "double": orc.T_FP64}
_JT = {"INNER": orc.JOIN_INNER, "LEFT": orc.JOIN_LEFT, "RIGHT": orc.JOIN_RIGHT, "SEMI": orc.JOIN_SEMI,
       "ANTI": orc.JOIN_ANTI}
_AGG = {"COUNT": orc.AGG_COUNT, "COUNT_STAR": orc.AGG_COUNT_STAR, "SUM": orc.AGG_SUM, "AVG": orc.AGG_AVG,
        "MIN": orc.AGG_MIN, "MAX": orc.AGG_MAX, "SUM0": orc.AGG_SUM0}


def column(values: Sequence, typ: str, sdict: Dict[str, int]) -> Col:
    """Python list with None -> (numpy values, nulls or None)."""
    n = len(values)
    data = np.zeros(n, dtype=_NP[typ])
    nulls = np.zeros(n, dtype=bool)
    for i, v in enumerate(values):
        if v is None:
            nulls[i] = True
        else:
            data[i] = sdict[v] if isinstance(v, str) else v
    return data, (nulls if nulls.any() else None)


def chunks_to_cols(chunks: List[List[List]], types: List[str], sdict: Dict[str, int]) -> List[Col]:
    """Concatenate a list of chunks into one column set (a MockExec drained in order)."""
    cols = []
    for c, t in enumerate(types):
        vals: List = []
        for ch in chunks:
            vals.extend(ch[c])
        cols.append(column(vals, t, sdict))
    return cols


def chunks_to_chunk_cols(chunks: List[List[List]], types: List[str], sdict: Dict[str, int]) -> List[List[Col]]:
    """Keep the chunk boundaries: one column set per chunk."""
    return [[column(ch[c], t, sdict) for c, t in enumerate(types)] for ch in chunks]


def rows_multiset(cols: Sequence[Col], float_round: Optional[int] = None) -> Counter:
    """Order-insensitive row multiset (BaseExecTest.assertExecResultByRow, order=false)."""
    n = len(cols[0][0]) if cols else 0
    lists = []
    for data, nulls in cols:
        vals = data.tolist() if hasattr(data, "tolist") else list(data)
        if float_round is not None:
            vals = [round(v, float_round) if isinstance(v, float) else v for v in vals]
        if nulls is not None:
            nl = np.asarray(nulls).astype(bool).tolist()
            vals = [None if isnull else v for v, isnull in zip(vals, nl)]
        lists.append(vals)
    return Counter(zip(*lists)) if lists else Counter({(): n})


def expect_multiset(expect_cols: List[List], sdict: Dict[str, int]) -> Counter:
    enc = [[(sdict[v] if isinstance(v, str) else v) for v in col] for col in expect_cols]
    return Counter(zip(*enc))


def join_case(case: dict):
    """-> (JoinSpec, outer cols, inner cols, expected Counter | None, expect_error)."""
    sdict = kats.encode_case_strings(case["outer"], case["inner"], case.get("expect", []),
                                     [v for _, v in case.get("cond_ne", [])])
    outer = chunks_to_cols(case["outer"], case["outer_types"], sdict)
    inner = chunks_to_cols(case["inner"], case["inner_types"], sdict)
    spec = orc.JoinSpec(
        join_type=_JT[case["join_type"]],
        outer_keys=[k[0] for k in case["keys"]],
        inner_keys=[k[1] for k in case["keys"]],
        key_types=[_ORC_T[k[2]] for k in case["keys"]],
        max_one_row=case.get("max_one_row", False),
        build_outer=case.get("build_outer", False),
        anti_operands=case.get("anti_operands"),
        cond_ne=tuple((c, sdict[v] if isinstance(v, str) else v) for c, v in case.get("cond_ne", [])),
    )
    expect = expect_multiset(case["expect"], sdict) if "expect" in case else None
    return spec, outer, inner, expect, case.get("expect_error")


def agg_case(case: dict):
    sdict = kats.encode_case_strings(case["chunks"], case.get("expect", []))
    cols = chunks_to_cols(case["chunks"], case["types"], sdict)
    aggs = [orc.AggCall(_AGG[k], cols_) for k, cols_ in case["aggs"]]
    expect = expect_multiset(case["expect"], sdict) if "expect" in case else None
    return cols, case["groups"], aggs, case["expected_groups"], expect


def agg_calls(spec_list) -> List["orc.AggCall"]:
    return [orc.AggCall(_AGG[k], c) for k, c in spec_list]


# ---------------------------------------------------------------------------------------------- synthetic tables
def splitmix64(x: np.ndarray) -> np.ndarray:
    """Counter-based generator used for every synthetic table: splitmix64(seed + row_id)."""
    with np.errstate(over="ignore"):
        z = (x.astype(np.uint64) + np.uint64(0x9E3779B97F4A7C15))
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def rand_u64(n: int, seed: int, stream: int = 0) -> np.ndarray:
    return splitmix64(np.arange(n, dtype=np.uint64) + np.uint64(seed) + np.uint64(stream) * np.uint64(0x1000000000))


def with_nulls(data: np.ndarray, frac: float, seed: int) -> Col:
    if frac <= 0:
        return data, None
    r = rand_u64(len(data), seed, stream=99)
    nulls = (r % np.uint64(1000000)) < np.uint64(int(frac * 1000000))
    return data, nulls


def brute_force_join_inner(outer: Sequence[Col], inner: Sequence[Col], okeys: Sequence[int], ikeys: Sequence[int]):
    """Nested-loop-equivalent INNER join through a Python dict (oracle self-check, SURVEY.md §8c last row)."""
    def keyrows(cols, keys):
        out = []
        for r in range(len(cols[0][0])):
            k = []
            for c in keys:
                d, nl = cols[c]
                if nl is not None and nl[r]:
                    k = None
                    break
                k.append(d[r].item())
            out.append(tuple(k) if k is not None else None)
        return out

    ok, ik = keyrows(outer, okeys), keyrows(inner, ikeys)
    index: Dict[tuple, List[int]] = {}
    for r, k in enumerate(ik):
        if k is not None:
            index.setdefault(k, []).append(r)

    def val(cols, c, r):
        d, nl = cols[c]
        return None if (nl is not None and nl[r]) else d[r].item()

    rows = Counter()
    for r, k in enumerate(ok):
        if k is None:
            continue
        for m in index.get(k, ()):
            rows[tuple(val(outer, c, r) for c in range(len(outer))) +
                 tuple(val(inner, c, m) for c in range(len(inner)))] += 1
    return rows
