"""GPU parity: ORDER BY / top-k (K5) and filter-only fragments (K1) vs the row-engine oracle, through the C ABI.
Row order is compared exactly: TopNSorter is stable by arrival (topn_sorter.h:96-106) and the GPU path carries the
arrival index as the last sort component; the full (LSD radix) sort is stable as well."""
import numpy as np
import pytest

from baikaldb_b200 import datagen, plan as P, queries
from baikaldb_b200.column import make_column
from baikaldb_b200.plan import PrimitiveType as T
from tests.util import run_both

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n", [0, 1, 999, 1000, 1001, 50_000, 3_000_000])
def test_c5_topk_sizes(n):
    """n < k, n == k, the collect-everything path and (3M rows > half the candidate buffer) the sampling path"""
    cols = datagen.c5_table(0, n)
    got, stats, _ = run_both(queries.c5_topk(1000), cols, keys=None)
    assert len(got[0]) == min(n, 1000)


@pytest.mark.parametrize("asc", [True, False])
def test_topk_heavy_duplicates_are_stable(asc):
    rng = np.random.default_rng(21)
    n = 2_500_000
    cols = [make_column(0, 1, T.INT64, rng.integers(0, 7, n)), make_column(0, 2, T.INT32, np.arange(n, dtype=np.int32))]
    run_both(queries.c5_topk(1000, asc=asc), cols, keys=None)


def test_topk_all_keys_equal():
    n = 2_200_000
    cols = [make_column(0, 1, T.INT64, np.full(n, 42)), make_column(0, 2, T.INT32, np.arange(n, dtype=np.int32))]
    got, _, _ = run_both(queries.c5_topk(100), cols, keys=None)
    assert got[1].to_list() == list(range(100))     # earliest arrivals win


@pytest.mark.parametrize("asc,null_first", [(True, True), (True, False), (False, True), (False, False)])
def test_topk_null_keys_placement(asc, null_first):
    rng = np.random.default_rng(5)
    n = 40_000
    cols = [make_column(0, 1, T.INT64, rng.integers(-1000, 1000, n), rng.random(n) > 0.01), make_column(0, 2, T.INT32, np.arange(n, dtype=np.int32))]
    pl = P.Plan(P.sort(P.scan(0), [P.slot_ref(0, 1, T.INT64)], [asc], [null_first], limit=700, tuple_id=0), {0: [(1, T.INT64), (2, T.INT32)]})
    got, _, _ = run_both(pl, cols, keys=None)
    assert (None in got[0].to_list()) == null_first


@pytest.mark.parametrize("ktype,vals", [(T.DOUBLE, lambda r, n: r.normal(size=n) * 1e6), (T.INT32, lambda r, n: r.integers(-2**31, 2**31 - 1, n)),
                                        (T.UINT64, lambda r, n: r.integers(0, 2**63, n).astype(np.uint64) * 2), (T.FLOAT, lambda r, n: r.normal(size=n).astype(np.float32))])
def test_topk_key_types(ktype, vals):
    rng = np.random.default_rng(8)
    n = 100_000
    cols = [make_column(0, 1, ktype, vals(rng, n)), make_column(0, 2, T.INT32, np.arange(n, dtype=np.int32))]
    for asc in (True, False):
        pl = P.Plan(P.sort(P.scan(0), [P.slot_ref(0, 1, ktype)], [asc], limit=333, tuple_id=0), {0: [(1, ktype), (2, T.INT32)]})
        run_both(pl, cols, keys=None)


def test_topk_with_filter_and_expression_key():
    rng = np.random.default_rng(13)
    n = 300_000
    cols = [make_column(0, 1, T.INT64, rng.integers(-10**6, 10**6, n)), make_column(0, 2, T.INT32, rng.integers(0, 100, n), rng.random(n) > 0.1)]
    s1, s2 = P.slot_ref(0, 1, T.INT64), P.slot_ref(0, 2, T.INT32)
    child = P.where(P.scan(0), P.lt(s2, P.int_lit(50)))
    pl = P.Plan(P.sort(child, [P.minus(s1, P.multiplies(s2, P.int_lit(1000)))], [False], limit=500, tuple_id=0), {0: [(1, T.INT64), (2, T.INT32)]})
    run_both(pl, cols, keys=None)


def test_topk_streaming_batches_keep_the_best_rows():
    cols = datagen.c5_table(0, 120_000)
    batches = [[make_column(c.tuple_id, c.slot_id, c.prim_type, c.values[a:b]) for c in cols] for a, b in ((0, 30_000), (30_000, 30_001), (30_001, 120_000))]
    run_both(queries.c5_topk(1000), cols, keys=None, batches=batches)


def test_full_sort_single_and_multi_key():
    rng = np.random.default_rng(17)
    n = 70_000
    cols = [make_column(0, 1, T.INT64, rng.integers(0, 50, n), rng.random(n) > 0.05), make_column(0, 2, T.INT32, rng.integers(-5, 5, n)),
            make_column(0, 3, T.DOUBLE, rng.normal(size=n))]
    tuples = {0: [(1, T.INT64), (2, T.INT32), (3, T.DOUBLE)]}
    run_both(P.Plan(P.sort(P.scan(0), [P.slot_ref(0, 3, T.DOUBLE)], [True], tuple_id=0), tuples), cols, keys=None)
    pl = P.Plan(P.sort(P.scan(0), [P.slot_ref(0, 1, T.INT64), P.slot_ref(0, 2, T.INT32)], [True, False], tuple_id=0), tuples)
    run_both(pl, cols, keys=None)
    pl = P.Plan(P.sort(P.scan(0), [P.slot_ref(0, 2, T.INT32), P.slot_ref(0, 1, T.INT64)], [False, False], [False, True], limit=20_000, tuple_id=0), tuples)
    run_both(pl, cols, keys=None)   # k > 4096: the sort path serves large limits


def test_filter_only_fragment_with_limit_and_offset():
    rng = np.random.default_rng(19)
    n = 200_000
    cols = [make_column(0, 1, T.INT64, rng.integers(0, 100, n)), make_column(0, 2, T.INT32, np.arange(n, dtype=np.int32), rng.random(n) > 0.2)]
    tuples = {0: [(1, T.INT64), (2, T.INT32)]}
    conj = P.and_(P.gt(P.slot_ref(0, 1, T.INT64), P.int_lit(90)), P.not_(P.is_null(P.slot_ref(0, 2, T.INT32))))
    got, stats, want = run_both(P.Plan(P.where(P.scan(0), conj), tuples), cols, keys=None)
    assert stats.rows_filtered == want.rows_filtered
    run_both(P.Plan(P.limit(P.where(P.scan(0), conj), 1000, 17), tuples), cols, keys=None, check_scanned=False)
    run_both(P.Plan(P.where(P.scan(0), conj, limit=123), tuples), cols, keys=None, check_scanned=False)
    batches = [[make_column(c.tuple_id, c.slot_id, c.prim_type, c.values[a:b], None if c.valid is None else c.valid[a:b]) for c in cols]
               for a, b in ((0, 70_000), (70_000, 200_000))]
    run_both(P.Plan(P.limit(P.where(P.scan(0), conj), 5000, 3), tuples), cols, keys=None, batches=batches, check_scanned=False)


@pytest.mark.parametrize("dist", ["full_range", "few_values", "sorted", "reverse"])
def test_full_sort_two_million_rows_is_stable(dist):
    """ORDER BY without LIMIT over 2M rows (~500 radix tiles: the chained look-back, the constant-digit skip for narrow keys):
    keys ascending, equal keys in arrival order (Sorter is a stable sort over MemRowCompare, sorter.cpp:54-114) — checked against
    numpy's stable argsort"""
    rng = np.random.default_rng(len(dist))
    n = 2_000_003
    keys = {"full_range": rng.integers(-(1 << 63), (1 << 63) - 1, n), "few_values": rng.integers(-3, 4, n) * (1 << 40),
            "sorted": np.arange(n, dtype=np.int64) - n // 2, "reverse": np.arange(n, dtype=np.int64)[::-1].copy()}[dist]
    pay = np.arange(n, dtype=np.int32)
    cols = [make_column(0, 1, T.INT64, keys), make_column(0, 2, T.INT32, pay)]
    pl = P.Plan(P.sort(P.scan(0), [P.slot_ref(0, 1, T.INT64)], [True], tuple_id=0), {0: [(1, T.INT64), (2, T.INT32)]})
    from baikaldb_b200.exec_node import execute
    got, stats = execute(pl, cols, device=0)
    order = np.argsort(keys, kind="stable")
    by = {c.name: c for c in got}
    assert np.array_equal(by["0_1"].values, keys[order])
    assert np.array_equal(by["0_2"].values, pay[order])
    assert stats.main_kernel_name.decode() == "radix_sort(k_rs_pass)"
