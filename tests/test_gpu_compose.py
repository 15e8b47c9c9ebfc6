"""GPU parity of operator chains ABOVE an aggregate — the db-side chain Packet -> Limit -> Sort -> HavingFilter -> (Merge)Agg the
reference builds (src/physical_plan/separate.cpp:241-260, src/exec/exec_node.cpp:347-394): HAVING over aggregate slots, ORDER BY
over aggregate output (values and keys, both directions, NULL placement), LIMIT / OFFSET on top, over a scan, over MERGE_AGG rows,
over a joined aggregate, over a scalar aggregate.  Sort keys are unique where the order is compared row by row."""
import numpy as np
import pytest

from baikaldb_b200 import datagen, plan as P, queries
from baikaldb_b200.column import make_column
from baikaldb_b200.plan import PlanNodeType, PrimitiveType as T
from oracle import oracle
from tests.util import run_both

pytestmark = pytest.mark.gpu

AGGS = [P.agg_expr("count_star", 1, 1), P.agg_expr("sum", 1, 2, None, P.slot_ref(0, 3, T.DOUBLE)), P.agg_expr("avg", 1, 3, 4, P.slot_ref(0, 4, T.DOUBLE))]
TUPLES = {0: [(1, T.INT32), (2, T.INT32), (3, T.DOUBLE), (4, T.DOUBLE)], 1: P.agg_tuple_slots(AGGS, [T.INT64, T.DOUBLE, T.DOUBLE])}
CNT, SUM, AVG, KEY = P.slot_ref(1, 1, T.INT64), P.slot_ref(1, 2, T.DOUBLE), P.slot_ref(1, 3, T.DOUBLE), P.slot_ref(0, 1, T.INT32)


def _agg(merge=False):
    if merge:
        return P.agg(P.scan(0), 1, [KEY], AGGS, merge=True)
    return P.agg(P.where(P.scan(0), P.lt(P.slot_ref(0, 2, T.INT32), P.int_lit(1 << 19))), 1, [KEY], AGGS)


def _having(child, *conj):
    return P.where(child, *conj, node_type=PlanNodeType.HAVING_FILTER_NODE)


@pytest.mark.parametrize("n_groups", [40, 3000])
def test_having_over_aggregate(n_groups):
    cols = datagen.c2_table(0, 200_000, n_groups=n_groups)
    thr = 200_000 // n_groups // 2
    pl = P.Plan(_having(_agg(), P.gt(CNT, P.int_lit(thr)), P.lt(AVG, P.double_lit(25.0))), TUPLES)
    got, _, want = run_both(pl, cols, keys=["0_1"])
    assert 0 < len(got[0]) < n_groups          # the predicate really cuts


@pytest.mark.parametrize("asc", [True, False])
@pytest.mark.parametrize("by", ["sum", "key", "avg_then_key"])
def test_order_by_over_aggregate(asc, by):
    cols = datagen.c2_table(0, 150_000, n_groups=500)
    exprs = {"sum": [SUM], "key": [KEY], "avg_then_key": [P.cast_to_signed(P.divides(AVG, P.double_lit(10.0))), KEY]}[by]
    pl = P.Plan(P.sort(_agg(), exprs, [asc] * len(exprs), tuple_id=1), TUPLES)
    run_both(pl, cols, keys=None)


def test_limit_sort_having_chain_with_offset():
    cols = datagen.c2_table(0, 300_000, n_groups=1000)
    chain = P.limit(P.sort(_having(_agg(), P.ge(CNT, P.int_lit(140))), [SUM], [False], tuple_id=1), 25, 3)
    got, _, _ = run_both(P.Plan(P.packet(chain), TUPLES), cols, keys=None)
    assert len(got[0]) == 25
    # ORDER BY ... LIMIT carried by the SORT node itself (top-n), no HAVING
    run_both(P.Plan(P.sort(_agg(), [SUM], [True], tuple_id=1, limit=10), TUPLES), cols, keys=None)
    # LIMIT over HAVING without a sort: any `limit` of the qualifying groups is a correct answer — compare the counts only
    pl = P.Plan(P.limit(_having(_agg(), P.ge(CNT, P.int_lit(140))), 7), TUPLES)
    want = oracle.execute(pl.serialize(), cols)
    from baikaldb_b200.exec_node import execute
    got, _ = execute(pl, cols, device=0)
    assert len(got[0]) == want.nrows == 7 and all(c >= 140 for c in {c.name: c for c in got}["1_1"].to_list())


def test_db_side_chain_over_merge_agg_rows():
    """Packet -> Limit -> Sort -> Having -> MergeAgg over the stores' partial rows (separate.cpp:249-258)"""
    store_plan = queries.c2_filter_groupby()
    parts = [oracle.execute(store_plan.serialize(), datagen.c2_table(r * 50_000, 50_000, n_groups=300)).columns for r in range(4)]
    rows = []
    for cs in zip(*parts):
        valid = None if all(c.valid is None for c in cs) else np.concatenate([c.valid if c.valid is not None else np.ones(len(c), bool) for c in cs])
        rows.append(make_column(cs[0].tuple_id, cs[0].slot_id, cs[0].prim_type, np.concatenate([c.values for c in cs]), valid))
    chain = P.packet(P.limit(P.sort(_having(_agg(merge=True), P.gt(CNT, P.int_lit(330))), [SUM], [False], tuple_id=1), 40))
    run_both(P.Plan(chain, TUPLES), rows, keys=None)


def test_order_by_with_null_aggregates_and_null_placement():
    rng = np.random.default_rng(31)
    n = 60_000
    cols = [make_column(0, 1, T.INT32, rng.integers(0, 200, n)), make_column(0, 2, T.INT32, rng.integers(0, 1 << 20, n)),
            make_column(0, 3, T.DOUBLE, rng.random(n), (rng.integers(0, 200, n) % 3 != 0) & (rng.random(n) > 0.2)),
            make_column(0, 4, T.DOUBLE, rng.normal(size=n), rng.random(n) > 0.5)]
    cols[2] = make_column(0, 3, T.DOUBLE, cols[2].values, (cols[0].values % 3 != 0) & (rng.random(n) > 0.2))   # a third of the groups: SUM is NULL
    for null_first in (True, False):
        pl = P.Plan(P.sort(_agg(), [SUM, KEY], [True, True], [null_first, True], tuple_id=1), TUPLES)
        run_both(pl, cols, keys=None)


def test_sort_and_having_over_joined_aggregate():
    nd, nf = 5_000, 120_000
    fact, dim = datagen.c3_fact(0, nf, nd), datagen.c3_dim(0, nd, nd, n_groups=300)
    aggs = [P.agg_expr("count_star", 2, 1), P.agg_expr("sum", 2, 2, None, P.slot_ref(0, 2, T.DOUBLE))]
    j = P.join(P.scan(1), P.scan(0), [P.eq(P.slot_ref(1, 1, T.INT32), P.slot_ref(0, 1, T.INT32))])
    a = P.agg(j, 2, [P.slot_ref(1, 2, T.INT32)], aggs)
    h = _having(a, P.gt(P.slot_ref(2, 1, T.INT64), P.int_lit(390)))
    pl = P.Plan(P.limit(P.sort(h, [P.slot_ref(2, 2, T.DOUBLE)], [False], tuple_id=2), 30),
                {0: [(1, T.INT32), (2, T.DOUBLE)], 1: [(1, T.INT32), (2, T.INT32)], 2: P.agg_tuple_slots(aggs, [T.INT64, T.DOUBLE])})
    run_both(pl, dim + fact, keys=None, batches=[dim, fact])
    # LIMIT directly over the joined aggregate (the limit used to be dropped): any 11 groups
    pl2 = P.Plan(P.limit(a, 11), {0: [(1, T.INT32), (2, T.DOUBLE)], 1: [(1, T.INT32), (2, T.INT32)], 2: P.agg_tuple_slots(aggs, [T.INT64, T.DOUBLE])})
    from baikaldb_b200.exec_node import execute
    got, _ = execute(pl2, [dim, fact], device=0)
    assert len(got[0]) == 11


def test_having_over_scalar_aggregate():
    cols = datagen.c1_table(0, 100_000)
    aggs = [P.agg_expr("count_star", 1, 1)]
    a = P.agg(P.where(P.scan(0), P.lt(P.slot_ref(0, 1, T.INT32), P.int_lit(1 << 19))), 1, [], aggs)
    tuples = {0: [(1, T.INT32)], 1: P.agg_tuple_slots(aggs, [T.INT64])}
    got, _, _ = run_both(P.Plan(P.packet(_having(a, P.gt(P.slot_ref(1, 1, T.INT64), P.int_lit(10)))), tuples), cols, keys=[])
    assert len(got[0]) == 1
    got, _, _ = run_both(P.Plan(P.packet(_having(a, P.gt(P.slot_ref(1, 1, T.INT64), P.int_lit(10**9)))), tuples), cols, keys=[])
    assert not got or len(got[0]) == 0


def test_aggregate_over_a_filtered_join():
    """AGG -> FILTER -> JOIN (the store-side chain of `... FROM a JOIN b ON ... WHERE f(a, b) GROUP BY ...`): the filter sees the joined row;
    above an INNER join it joins the residual conditions, above an outer join it is refused"""
    nd, nf = 4_000, 90_000
    fact, dim = datagen.c3_fact(0, nf, nd), datagen.c3_dim(0, nd, nd, n_groups=60)
    aggs = [P.agg_expr("count_star", 2, 1), P.agg_expr("sum", 2, 2, None, P.slot_ref(0, 2, T.DOUBLE))]
    tuples = {0: [(1, T.INT32), (2, T.DOUBLE)], 1: [(1, T.INT32), (2, T.INT32)], 2: P.agg_tuple_slots(aggs, [T.INT64, T.DOUBLE])}
    on = [P.eq(P.slot_ref(1, 1, T.INT32), P.slot_ref(0, 1, T.INT32))]
    where = [P.gt(P.add(P.slot_ref(0, 2, T.DOUBLE), P.slot_ref(1, 2, T.INT32)), P.double_lit(20.25)), P.ne(P.mod(P.slot_ref(1, 1, T.INT32), P.int_lit(7)), P.int_lit(0))]
    a = P.agg(P.where(P.join(P.scan(1), P.scan(0), on), *where), 2, [P.slot_ref(1, 2, T.INT32)], aggs)
    got, _, _ = run_both(P.Plan(a, tuples), dim + fact, keys=["1_2"], batches=[dim, fact])
    assert 0 < len(got[0]) <= 60
    from baikaldb_b200 import _lib
    left = P.agg(P.where(P.join(P.scan(1), P.scan(0), on, join_type=P.JoinType.LEFT_JOIN), *where), 2, [P.slot_ref(1, 2, T.INT32)], aggs)
    with pytest.raises(_lib.BkgpuError) as e:
        _lib.explain(P.Plan(left, tuples).serialize())
    assert e.value.code == _lib.EUNSUPPORTED
