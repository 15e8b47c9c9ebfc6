"""GPU parity of the warp-private aggregate kernel (csrc/agg_wp.cuh) against the row-engine oracle: key-table inserts and
bucket overflow, same-group collisions inside a warp (skew), groups beyond the per-warp capacity (global path), capacity learned
from an earlier run, integer sums, one / no value column, every comparison operator of the filter."""
import ctypes

import numpy as np
import pytest

from baikaldb_b200 import datagen, plan as P, queries
from baikaldb_b200.column import make_column
from baikaldb_b200.exec_node import ColumnSource, GpuExecNode, RowBatch, RuntimeState
from baikaldb_b200.plan import PrimitiveType as T
from tests.util import assert_same_rows, run_both
from oracle import oracle

pytestmark = pytest.mark.gpu
WP = "k_agg_group_wp"
USE = {"use_wp": 1}   # the kernel is opt-in (measured at par with k_agg_group_lean)


@pytest.mark.parametrize("n_groups", [1, 2, 7, 33, 1000, 1100, 1500, 5000, 40_000])
def test_wp_cardinalities(n_groups):
    """1..40k groups: few groups = every lane of a warp fights for the same count word; 1100+ = more groups than one warp table
    holds on the first run (the rest take the global path); the result never depends on which path a row took"""
    cols = datagen.c2_table(0, 300_000, n_groups=n_groups)
    _, stats, _ = run_both(queries.c2_filter_groupby(), cols, keys=["0_1"], options=USE)
    assert stats.main_kernel_name.decode() == WP


@pytest.mark.parametrize("n_groups", [50, 1000, 3000])
def test_wp_second_run_uses_learned_cardinality(n_groups):
    """prepared-statement reuse: the second run sizes the tables from the first run's group count; same rows both times"""
    cols = datagen.c2_table(0, 200_000, n_groups=n_groups)
    pl = queries.c2_filter_groupby()
    want = oracle.execute(pl.serialize(), cols)
    node, st = GpuExecNode(), RuntimeState(device=0, options=dict(USE))
    node.init(pl)
    node.add_child(ColumnSource([cols]))
    try:
        assert node.open(st) == 0, st.error_msg
        for run in range(3):
            if run:
                node.reset()
                node.push(cols)
                node.finish()
            got, eos, rb = [], False, RowBatch()
            while not eos:
                rc, eos = node.get_next(st, rb)
                assert rc == 0
                got = got or list(rb.columns)
            assert_same_rows(got, want.columns, ["0_1"])
            # (3000 groups x 8 warp tables do not fit 227 KB: once the cardinality is known a CTA-shared-table kernel takes over)
            name = node.stats().main_kernel_name.decode()
            assert name == WP if (n_groups <= 1000 or run == 0) else name in ("k_agg_group_lean", "k_agg_group_lean_fx", "k_agg_group_direct")
    finally:
        node.close(st)


def test_wp_sparse_and_negative_keys():
    rng = np.random.default_rng(5)
    n = 150_000
    keys = rng.choice(np.array([-(1 << 31), -1, 0, 1, (1 << 31) - 1, 123456789, -987654321] + list(rng.integers(-(1 << 31), 1 << 31, 500))), n)
    cols = [make_column(0, 1, T.INT32, keys), make_column(0, 2, T.INT32, rng.integers(0, 1 << 20, n)),
            make_column(0, 3, T.DOUBLE, rng.random(n)), make_column(0, 4, T.DOUBLE, rng.normal(size=n) * 1e3)]
    _, stats, _ = run_both(queries.c2_filter_groupby(), cols, keys=["0_1"], options=USE)
    assert stats.main_kernel_name.decode() == WP


def test_wp_uint32_key_with_all_ones():
    rng = np.random.default_rng(6)
    n = 90_000
    keys = rng.choice(np.array([0xFFFFFFFF, 0xFFFFFFFE, 0xFFFFFFF0, 0, 1, 0x80000000], dtype=np.uint32), n)
    cols = [make_column(0, 1, T.UINT32, keys), make_column(0, 2, T.DOUBLE, rng.random(n))]
    aggs = [P.agg_expr("count_star", 1, 1), P.agg_expr("sum", 1, 2, None, P.slot_ref(0, 2, T.DOUBLE))]
    pl = P.Plan(P.agg(P.scan(0), 1, [P.slot_ref(0, 1, T.UINT32)], aggs), {0: [(1, T.UINT32), (2, T.DOUBLE)], 1: P.agg_tuple_slots(aggs, [T.INT64, T.DOUBLE])})
    _, stats, _ = run_both(pl, cols, keys=["0_1"], options=USE)
    assert stats.main_kernel_name.decode() == WP


@pytest.mark.parametrize("shape", ["count_only", "one_double", "one_int64", "int64_and_double", "two_int64"])
def test_wp_value_shapes(shape):
    rng = np.random.default_rng(len(shape))
    n = 120_001
    cols = [make_column(0, 1, T.INT32, rng.integers(0, 300, n)), make_column(0, 2, T.INT32, rng.integers(0, 100, n)),
            make_column(0, 3, T.DOUBLE, rng.normal(size=n)), make_column(0, 4, T.INT64, rng.integers(-(1 << 61), 1 << 61, n)),
            make_column(0, 5, T.INT64, rng.integers(-5, 5, n))]
    d, i1, i2 = P.slot_ref(0, 3, T.DOUBLE), P.slot_ref(0, 4, T.INT64), P.slot_ref(0, 5, T.INT64)
    aggs = {"count_only": [P.agg_expr("count_star", 1, 1)],
            "one_double": [P.agg_expr("count_star", 1, 1), P.agg_expr("avg", 1, 2, 3, d)],
            "one_int64": [P.agg_expr("sum", 1, 1, None, i1)],
            "int64_and_double": [P.agg_expr("sum", 1, 1, None, i1), P.agg_expr("sum", 1, 2, None, d), P.agg_expr("count_star", 1, 3)],
            "two_int64": [P.agg_expr("sum", 1, 1, None, i1), P.agg_expr("sum", 1, 2, None, i2)]}[shape]
    types = {"count_only": [T.INT64], "one_double": [T.INT64, T.DOUBLE], "one_int64": [T.INT64], "int64_and_double": [T.INT64, T.DOUBLE, T.INT64],
             "two_int64": [T.INT64, T.INT64]}[shape]
    root = P.agg(P.where(P.scan(0), P.lt(P.slot_ref(0, 2, T.INT32), P.int_lit(60))), 1, [P.slot_ref(0, 1, T.INT32)], aggs)
    pl = P.Plan(root, {0: [(1, T.INT32), (2, T.INT32), (3, T.DOUBLE), (4, T.INT64), (5, T.INT64)], 1: P.agg_tuple_slots(aggs, types)})
    _, stats, _ = run_both(pl, cols, keys=["0_1"], options=USE)
    assert stats.main_kernel_name.decode() == WP


@pytest.mark.parametrize("op", ["eq", "ne", "lt", "le", "gt", "ge"])
@pytest.mark.parametrize("c", [-(1 << 31), -7, 0, 41, (1 << 31) - 1])
def test_wp_filter_operators(op, c):
    """every comparison of `int32 column <cmp> constant` at the int32 limits (the kernel folds all six into one unsigned compare)"""
    rng = np.random.default_rng(abs(c) % 1000 + len(op))
    n = 50_000
    f = rng.choice(np.array([-(1 << 31), -(1 << 31) + 1, -8, -7, -6, -1, 0, 1, 40, 41, 42, (1 << 31) - 2, (1 << 31) - 1]), n)
    cols = [make_column(0, 1, T.INT32, rng.integers(0, 64, n)), make_column(0, 2, T.INT32, f), make_column(0, 3, T.DOUBLE, rng.random(n))]
    aggs = [P.agg_expr("count_star", 1, 1), P.agg_expr("sum", 1, 2, None, P.slot_ref(0, 3, T.DOUBLE))]
    pred = getattr(P, op)(P.slot_ref(0, 2, T.INT32), P.int_lit(c))
    root = P.agg(P.where(P.scan(0), pred, P.ge(P.slot_ref(0, 1, T.INT32), P.int_lit(3))), 1, [P.slot_ref(0, 1, T.INT32)], aggs)
    pl = P.Plan(root, {0: [(1, T.INT32), (2, T.INT32), (3, T.DOUBLE)], 1: P.agg_tuple_slots(aggs, [T.INT64, T.DOUBLE])})
    _, stats, _ = run_both(pl, cols, keys=["0_1"], options=USE)
    assert stats.main_kernel_name.decode() == WP
