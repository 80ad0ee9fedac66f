"""Toy-scale TPC-H Q3 tables and the oracle's answer (numpy filters + oracle joins + oracle group-by on the GLOBAL
tables), shared by the single-GPU parity test and the multi-GPU worker."""
import numpy as np

from oracle import oracle as orc
from tests import kat_util as ku


def q3_tables(rank: int, world: int, ncust=4000, nord=30000, nline=110000, seed=700):
    """This rank's round-robin share.  Keys are globally unique (i * world + rank)."""
    s = seed + 1000 * rank
    c_custkey = (np.arange(ncust, dtype=np.int64) * world + rank)
    c_seg = (ku.rand_u64(ncust, s + 1) % np.uint64(5)).astype(np.int32)
    o_orderkey = (np.argsort(ku.rand_u64(nord, s + 2)).astype(np.int64) * world + rank)
    o_custkey = (ku.rand_u64(nord, s + 3) % np.uint64(ncust * world)).astype(np.int64)
    o_orderdate = (8035 + ku.rand_u64(nord, s + 4) % np.uint64(2557)).astype(np.int32)     # 1992-01-01 .. 1998-12-31
    o_ship = np.zeros(nord, dtype=np.int32)
    l_orderkey = (ku.rand_u64(nline, s + 5) % np.uint64(nord * world)).astype(np.int64)
    l_price = (90000 + ku.rand_u64(nline, s + 6) % np.uint64(10_410_000)).astype(np.float64) / 100.0
    l_disc = (ku.rand_u64(nline, s + 7) % np.uint64(11)).astype(np.float64) / 100.0
    l_ship = (8035 + ku.rand_u64(nline, s + 8) % np.uint64(2557)).astype(np.int32)
    n = lambda *cols: [(c, None) for c in cols]
    return n(c_custkey, c_seg), n(o_orderkey, o_custkey, o_orderdate, o_ship), n(l_orderkey, l_price, l_disc, l_ship)


def q3_oracle(customer, orders, lineitem, date=9204, segment=1):
    """Filters in numpy (value semantics only), joins and group-by through the oracle."""
    c = customer[1][0] == segment
    ck = [(customer[0][0][c], None)]
    o = orders[2][0] < date
    od = [(col[0][o], None) for col in orders]
    spec = orc.JoinSpec(orc.JOIN_INNER, [1], [0], [orc.T_INT64])
    oj = orc.hash_join(spec, od, ck)                        # o_orderkey, o_custkey, o_orderdate, o_shippriority, c_custkey
    l = lineitem[3][0] > date
    li = [(lineitem[0][0][l], None), (lineitem[1][0][l] * (1.0 - lineitem[2][0][l]), None)]
    spec2 = orc.JoinSpec(orc.JOIN_INNER, [0], [0], [orc.T_INT64])
    lj = orc.hash_join(spec2, li, [oj[0], oj[2], oj[3]])   # l_orderkey, revenue, o_orderkey, o_orderdate, o_shippriority
    return orc.hash_agg(lj, [0, 3, 4], [orc.AggCall(orc.AGG_SUM, [1])], 1024)


def check_q3_multi(ctx, device, rank, world, gather_cols, dev, host):
    from galaxysql_b200 import pipelines
    from tests import gpu_util as gu
    cust, orders, line = q3_tables(rank, world)
    q3 = pipelines.Q3Pipeline(ctx, customer_capacity=4000 * world + 16, orders_capacity=60000, lineitem_capacity=220000, nslabs=3,
                              expected_groups=4096)
    out = host(q3.run(dev(cust, device), dev(orders, device), dev(line, device)))
    stats = q3.stats
    q3.close()
    gc, go, gl = gather_cols(cust), gather_cols(orders), gather_cols(line)
    exp = q3_oracle(gc, go, gl)
    allout = gather_cols(out)
    if rank == 0:
        assert len(exp[0][0]) > 100, "toy tables too selective to be a test"
        gu.approx_rows_equal(allout, exp, float_cols=[3], key_cols=[0, 1, 2], rtol=1e-6)
        assert stats["j1_fast"] == 1 and stats["j2_fast"] == 1, stats
