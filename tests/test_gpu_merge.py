"""MERGE_AGG_NODE over rows (separate.cpp:249-258: the db side folds the stores' partial rows; AggFnCall::merge,
agg_fn_call.cpp:719-822): the GPU merger against the oracle's merger on the same partial rows, and merged-of-regions
against the single-pass aggregate."""
import numpy as np
import pytest

from baikaldb_b200 import datagen, queries
from baikaldb_b200 import plan as P
from baikaldb_b200.column import make_column
from baikaldb_b200.plan import PrimitiveType as T
from oracle import oracle
from tests.util import assert_same_rows, run_both


def _concat(parts):
    out = []
    for cols in zip(*parts):
        c0 = cols[0]
        vals = np.concatenate([c.values for c in cols])
        valid = None
        if any(c.valid is not None for c in cols):
            valid = np.concatenate([c.valid if c.valid is not None else np.ones(len(c), bool) for c in cols])
        out.append(make_column(c0.tuple_id, c0.slot_id, c0.prim_type, vals, valid))
    return out


def _regions_partial_rows(n, regions, n_groups):
    plan = queries.c2_filter_groupby()
    step = (n + regions - 1) // regions
    parts = [oracle.execute(plan.serialize(), datagen.c2_table(r0, min(step, n - r0), n_groups=n_groups)).columns for r0 in range(0, n, step)]
    return _concat(parts)


@pytest.mark.gpu
@pytest.mark.parametrize("n,regions,groups", [(60_000, 4, 1000), (5_000, 7, 37)])
def test_merge_rows_matches_oracle_and_single_pass(n, regions, groups):
    rows = _regions_partial_rows(n, regions, groups)
    got, _, _ = run_both(queries.c2_filter_groupby(merge=True), rows, ["0_1"])
    single = oracle.execute(queries.c2_filter_groupby().serialize(), datagen.c2_table(0, n, n_groups=groups))
    assert_same_rows(got, single.columns, ["0_1"], rel=1e-9, abs_tol=1e-9)


def _minmax_plan(merge, grouped=True):
    aggs = [P.agg_expr("count", 1, 1, None, P.slot_ref(0, 2, T.INT64)), P.agg_expr("sum", 1, 2, None, P.slot_ref(0, 2, T.INT64)),
            P.agg_expr("min", 1, 3, None, P.slot_ref(0, 3, T.DOUBLE)), P.agg_expr("max", 1, 4, None, P.slot_ref(0, 2, T.INT64)),
            P.agg_expr("avg", 1, 5, 6, P.slot_ref(0, 2, T.INT64)), P.agg_expr("count_star", 1, 7)]
    tuples = {0: [(1, T.INT32), (2, T.INT64), (3, T.DOUBLE)], 1: P.agg_tuple_slots(aggs, [T.INT64, T.INT64, T.DOUBLE, T.INT64, T.INT64, T.INT64])}
    root = P.agg(P.scan(0), 1, [P.slot_ref(0, 1, T.INT32)] if grouped else [], aggs, merge=merge)
    return P.Plan(P.packet(root), tuples)


def _nully_table(n, seed):
    rng = np.random.default_rng(seed)
    return [make_column(0, 1, T.INT32, rng.integers(0, 9, n), rng.random(n) > 0.1),
            make_column(0, 2, T.INT64, rng.integers(-1000, 1000, n), rng.random(n) > 0.6),
            make_column(0, 3, T.DOUBLE, rng.normal(size=n), rng.random(n) > 0.6)]


@pytest.mark.gpu
@pytest.mark.parametrize("grouped", [True, False])
def test_merge_all_kinds_with_nulls(grouped):
    """COUNT / SUM / MIN / MAX / AVG / COUNT(*) partials with NULL sums, NULL keys and all-NULL groups."""
    parts = [oracle.execute(_minmax_plan(False, grouped).serialize(), _nully_table(40, s)).columns for s in range(5)]
    rows = _concat(parts)
    run_both(_minmax_plan(True, grouped), rows, ["0_1"] if grouped else [])


@pytest.mark.gpu
def test_scalar_merge_of_nothing_and_of_initial_rows():
    """no input rows / only rows in their initial state: the merger returns the blank row (agg_node.cpp:489-522)."""
    empty = oracle.execute(_minmax_plan(False, False).serialize(), [c.__class__(c.tuple_id, c.slot_id, c.prim_type, c.values[:0], None) for c in _nully_table(4, 1)]).columns
    assert len(empty[0]) == 1          # COUNT = 0, everything else NULL
    run_both(_minmax_plan(True, False), empty, [])
    run_both(_minmax_plan(True, False), _concat([empty, empty, empty]), [])
    none = [c.__class__(c.tuple_id, c.slot_id, c.prim_type, c.values[:0], None) for c in empty]
    run_both(_minmax_plan(True, False), none, [])


def _distinct_case(seed=1, n=60_000, nk=9, nx=300):
    """COUNT(DISTINCT x), SUM(DISTINCT x), AVG(DISTINCT x), SUM(v), COUNT(*) GROUP BY k as the planner lays it out (select_planner.cpp:612-700):
    a lower aggregate GROUP BY (k, x) with the ordinary aggregates, and above it a MERGE_AGG GROUP BY k whose *_distinct functions are
    UPDATED from the (k, x) rows (AggFnCall::merge -> update for distinct aggregates, agg_fn_call.cpp:719-727) while the ordinary ones merge"""
    import numpy as np
    from baikaldb_b200 import plan as P
    from baikaldb_b200.column import make_column
    from baikaldb_b200.plan import PrimitiveType as T
    rng = np.random.default_rng(seed)
    k, x, v = rng.integers(0, nk, n), rng.integers(-50, nx, n), rng.random(n)
    xv = rng.random(n) > 0.1
    cols = [make_column(0, 1, T.INT32, k), make_column(0, 2, T.INT32, x, xv), make_column(0, 3, T.DOUBLE, v)]
    low_aggs = [P.agg_expr("sum", 1, 1, None, P.slot_ref(0, 3, T.DOUBLE)), P.agg_expr("count_star", 1, 2)]
    low = P.Plan(P.agg(P.scan(0), 1, [P.slot_ref(0, 1, T.INT32), P.slot_ref(0, 2, T.INT32)], low_aggs),
                 {0: [(1, T.INT32), (2, T.INT32), (3, T.DOUBLE)], 1: P.agg_tuple_slots(low_aggs, [T.DOUBLE, T.INT64])})
    top_aggs = [P.agg_expr("sum", 1, 1, None, P.slot_ref(0, 3, T.DOUBLE)), P.agg_expr("count_star", 1, 2), P.agg_expr("count_distinct", 1, 3, None, P.slot_ref(0, 2, T.INT32)),
                P.agg_expr("sum_distinct", 1, 4, None, P.slot_ref(0, 2, T.INT32)), P.agg_expr("avg_distinct", 1, 5, 6, P.slot_ref(0, 2, T.INT32))]
    top = P.Plan(P.agg(P.scan(0), 1, [P.slot_ref(0, 1, T.INT32)], top_aggs, merge=True),
                 {0: [(1, T.INT32), (2, T.INT32)], 1: [(1, T.DOUBLE), (2, T.INT64), (3, T.INT64), (4, T.INT64), (5, T.DOUBLE), (6, T.STRING)]})
    return (k, x, xv, v), cols, low, top


@pytest.mark.gpu
def test_distinct_aggregates_two_level_plan():
    import numpy as np
    from baikaldb_b200.exec_node import execute
    (k, x, xv, v), cols, low, top = _distinct_case()
    mid, _, _ = run_both(low, cols, keys=["0_1", "0_2"])                 # GPU: GROUP BY (k, x), NULL x is its own group
    got, _, _ = run_both(top, list(mid), keys=["0_1"])                   # GPU: the merger over those rows == oracle
    by = {c.name: c.to_list() for c in got}
    for i, kk in enumerate(by["0_1"]):                                   # ... and == an independent numpy computation
        m = k == kk
        dx = sorted(set(x[m & xv].tolist()))
        assert by["1_3"][i] == len(dx) and by["1_4"][i] == sum(dx) and by["1_2"][i] == int(m.sum())
        assert abs(by["1_5"][i] - sum(dx) / len(dx)) < 1e-9 and abs(by["1_1"][i] - v[m].sum()) < 1e-6 * max(1.0, v[m].sum())
