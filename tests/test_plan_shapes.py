"""Fragment shapes the lowering accepts and refuses (host only: bkgpu_plan_explain parses, type-infers and lowers without a device) —
the composite chains of round 2: operators above an aggregate, filters around joins, joins that return rows, distinct aggregates."""
import pytest

from baikaldb_b200 import _lib
from baikaldb_b200 import plan as P
from baikaldb_b200.plan import PrimitiveType as T

T0 = [(1, T.INT32), (2, T.DOUBLE), (3, T.INT32)]
T1 = [(1, T.INT32), (2, T.INT32)]
ON = [P.eq(P.slot_ref(1, 1, T.INT32), P.slot_ref(0, 1, T.INT32))]
AGGS = [P.agg_expr("count_star", 2, 1), P.agg_expr("sum", 2, 2, None, P.slot_ref(0, 2, T.DOUBLE))]
TUPLES = {0: T0, 1: T1, 2: P.agg_tuple_slots(AGGS, [T.INT64, T.DOUBLE])}


def explain(root, tuples=TUPLES):
    return _lib.explain(P.Plan(root, tuples).serialize())


def refused(root, tuples=TUPLES):
    with pytest.raises(_lib.BkgpuError) as e:
        explain(root, tuples)
    assert e.value.code == _lib.EUNSUPPORTED, e.value
    return str(e.value)


def test_chains_above_an_aggregate():
    a = P.agg(P.where(P.scan(0), P.gt(P.slot_ref(0, 3, T.INT32), P.int_lit(3))), 2, [P.slot_ref(0, 1, T.INT32)], AGGS)
    having = P.where(a, P.gt(P.slot_ref(2, 1, T.INT64), P.int_lit(10)), node_type=P.PlanNodeType.HAVING_FILTER_NODE)
    text = explain(P.packet(P.limit(P.sort(having, [P.slot_ref(2, 2, T.DOUBLE)], [False], tuple_id=2), 10, offset=5)))
    assert text.startswith("kind=1") and "post fragment above the aggregate: kind=3" in text and "limit=15 offset=5" in text
    assert "post fragment above the aggregate: kind=2" in explain(having)                      # HAVING alone: a filter over the groups
    # ORDER BY a column the aggregate does not output
    assert "does not output" in refused(P.sort(a, [P.slot_ref(0, 3, T.INT32)], [True], tuple_id=2))


@pytest.mark.parametrize("jt", ["INNER_JOIN", "LEFT_JOIN", "RIGHT_JOIN", "SEMI_JOIN", "ANTI_SEMI_JOIN"])
def test_aggregate_over_joins(jt):
    ch = (P.scan(0), P.scan(1)) if jt == "RIGHT_JOIN" else (P.scan(1), P.scan(0))
    j = P.join(ch[0], ch[1], ON + [P.gt(P.add(P.slot_ref(0, 3, T.INT32), P.slot_ref(1, 2, T.INT32)), P.int_lit(7))], join_type=getattr(P.JoinType, jt))
    assert explain(P.agg(j, 2, [P.slot_ref(1, 2, T.INT32)], AGGS[:1] if "SEMI" in jt else AGGS)).startswith("kind=4")
    filtered = P.agg(P.where(j, P.lt(P.slot_ref(1, 2, T.INT32), P.int_lit(5))), 2, [P.slot_ref(1, 2, T.INT32)], AGGS[:1] if "SEMI" in jt else AGGS)
    if jt == "INNER_JOIN":
        assert explain(filtered).startswith("kind=4")                                           # AGG -> FILTER -> JOIN: the filter joins the residual conditions
    else:
        assert "filter between the aggregate" in refused(filtered)


def test_joins_that_return_rows():
    j = P.join(P.scan(1), P.scan(0), ON)
    rows = {0: T0, 1: T1}
    text = explain(P.packet(j), rows)
    assert text.startswith("kind=5") and "cols=5" in text
    f = P.where(j, P.ne(P.slot_ref(1, 2, T.INT32), P.int_lit(4)))
    assert "kind=3 keys=1" in explain(P.limit(P.sort(f, [P.slot_ref(0, 2, T.DOUBLE)], [True], tuple_id=0), 9), rows)   # SORT + LIMIT -> a top-k sink
    assert explain(P.join(P.scan(1), P.scan(0), ON, join_type=P.JoinType.LEFT_JOIN), rows).startswith("kind=5")
    assert "returns rows" in refused(P.join(P.scan(1), P.scan(0), ON, join_type=P.JoinType.SEMI_JOIN), rows)
    assert "filter" in refused(P.where(P.join(P.scan(1), P.scan(0), ON, join_type=P.JoinType.LEFT_JOIN), P.ne(P.slot_ref(1, 2, T.INT32), P.int_lit(4))), rows)
    assert "equality" in refused(P.join(P.scan(1), P.scan(0), [P.lt(P.slot_ref(1, 1, T.INT32), P.slot_ref(0, 1, T.INT32))]), rows)


def test_distinct_aggregates_lower_only_as_the_planner_lays_them_out():
    top = [P.agg_expr("count_distinct", 2, 1, None, P.slot_ref(0, 3, T.INT32)), P.agg_expr("sum", 2, 2, None, P.slot_ref(0, 2, T.DOUBLE))]
    tuples = {0: [(1, T.INT32), (3, T.INT32)], 2: [(1, T.INT64), (2, T.DOUBLE)]}
    text = explain(P.agg(P.scan(0), 2, [P.slot_ref(0, 1, T.INT32)], top, merge=True), tuples)   # MERGE_AGG: count_distinct UPDATES (kind 1 = COUNT), sum MERGES
    assert "agg[0] kind=1" in text and "agg[1] kind=2" in text
    with pytest.raises(_lib.BkgpuError):
        explain(P.agg(P.scan(0), 2, [P.slot_ref(0, 1, T.INT32)], [P.agg_expr("multi_count_distinct", 2, 1, None, P.slot_ref(0, 3, T.INT32))], merge=True), tuples)
