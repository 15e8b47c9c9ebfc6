"""Pins the oracle (oracle/bk_oracle.c) to the reference's own known-answer tests and fixtures
(SURVEY.md §8c): test/test_expr_value.cpp:456-653 (ExprValue compare / cast) and the Arrow fixture of
test/test_arrow_compute.cpp:51-277, plus cross-checks of the row-engine restatement against the Acero
plan the reference's vectorized engine builds.  CPU only."""
import struct

import numpy as np
import pytest

from baikaldb_b200 import datagen, plan as P, queries
from baikaldb_b200.column import make_column, rows_as_set
from baikaldb_b200.plan import PrimitiveType as T
from oracle import acero_oracle as A
from oracle import oracle


def _bits(fmt, v):
    return struct.unpack("<Q", struct.pack(fmt, v).ljust(8, b"\0"))[0]


def i64(v): return _bits("<q", v)
def u64(v): return _bits("<Q", v & 0xFFFFFFFFFFFFFFFF)
def i32(v): return _bits("<i", v)
def u32(v): return _bits("<I", v & 0xFFFFFFFF)
def f64(v): return _bits("<d", v)


def cmp(ta, ba, tb, bb, diff=False):
    return oracle.lib().bko_ev_compare(int(ta), ba, int(tb), bb, 1 if diff else 0)


# ---- test/test_expr_value.cpp TEST(test_compare, case_all) ----
def test_int64_1_equals_int32_1_after_promotion():      # :498-503
    assert cmp(T.INT64, i64(1), T.INT32, i32(1), diff=True) == 0


def test_int64_ordering():                               # :504-510
    assert cmp(T.INT64, i64(65571188177), T.INT64, i64(72856896263)) < 0


def test_uint64_ordering():                              # :511-524
    assert cmp(T.UINT64, u64(65571188177), T.UINT64, u64(72856896263)) < 0
    assert cmp(T.UINT64, u64(1), T.UINT64, u64(-1)) < 0   # 1 < 0xFFFF...FFFF


def test_int32_compare_reads_int32_member_of_other():    # :525-531  v2 is typed INT64 but only int32_val is set
    assert cmp(T.INT32, i32(2147483610), T.INT64, i32(-2147483610)) > 0


def test_uint32_minus_one_is_large():                    # :532-538
    assert cmp(T.UINT32, u32(-1), T.UINT32, u32(1)) > 0


def test_datetime_compares_as_uint64():                  # :539-548
    a = oracle.lib().bko_ev_cast(int(T.UINT64), u64(9223372036854775800), int(T.DATETIME))
    b = oracle.lib().bko_ev_cast(int(T.UINT64), u64(9223372036854775810), int(T.DATETIME))
    assert cmp(T.DATETIME, a, T.DATETIME, b) < 0


def test_maxvalue_type():                                # :583-603
    MAXV = 24
    assert cmp(MAXV, 0, T.INT32, i32(2147483647)) > 0
    assert cmp(MAXV, 0, MAXV, 0) == 0
    assert cmp(T.INT32, i32(2147483647), MAXV, 0) < 0


def test_int64_double_round_trip():                      # :468-476
    d = oracle.lib().bko_ev_cast(int(T.INT64), i64(123372036854775800), int(T.DOUBLE))
    assert struct.unpack("<d", struct.pack("<Q", d))[0] == float(123372036854775800)
    back = oracle.lib().bko_ev_cast(int(T.DOUBLE), d, int(T.INT64))
    assert struct.unpack("<q", struct.pack("<Q", back))[0] == int(float(123372036854775800))


def test_cast_truncations():
    L = oracle.lib()
    assert L.bko_ev_cast(int(T.INT64), i64(3_000_000_000), int(T.INT32)) & 0xFFFFFFFF == u32(3_000_000_000 - (1 << 32))
    assert L.bko_ev_cast(int(T.DOUBLE), f64(0.9), int(T.INT64)) == 0
    assert L.bko_ev_cast(int(T.DOUBLE), f64(-1.5), int(T.INT64)) == i64(-1)
    assert L.bko_ev_cast(int(T.INT32), i32(-1), int(T.UINT64)) == u64(-1)


def test_nan_compares_equal():                           # expr_value.h:928-933
    nan = f64(float("nan"))
    assert cmp(T.DOUBLE, nan, T.DOUBLE, f64(1.0)) == 0
    assert cmp(T.DOUBLE, f64(1.0), T.DOUBLE, nan) == 0


def test_memcomparable_key_encoding():                   # mut_table_key.h:60-160, key_encoder.h:120-135
    L = oracle.lib()
    buf = bytes(8)

    def enc(t, b):
        out = bytearray(8)
        c = (np.frombuffer(out, dtype=np.uint8)).ctypes
        import ctypes
        raw = ctypes.create_string_buffer(8)
        n = L.bko_key_encode(int(t), b, raw)
        return raw.raw[:n]
    assert enc(T.INT32, i32(0)) == b"\x80\x00\x00\x00"
    assert enc(T.INT32, i32(-1)) == b"\x7f\xff\xff\xff"
    assert enc(T.INT64, i64(1)) == b"\x80" + b"\x00" * 6 + b"\x01"
    assert enc(T.UINT32, u32(258)) == b"\x00\x00\x01\x02"
    assert enc(T.INT32, i32(-5)) < enc(T.INT32, i32(3))       # memcomparable
    assert enc(T.DOUBLE, f64(-2.5)) < enc(T.DOUBLE, f64(-1.0)) < enc(T.DOUBLE, f64(0.0)) < enc(T.DOUBLE, f64(1e300))


# ---- test/test_arrow_compute.cpp:51-277: the 5-row, 14-column fixture and its 6-key hash_sum plan ----
def _arrow_compute_fixture():
    u, q = T.UINT32, T.UINT64
    data = {
        (0, 1, u): [0, 1, 0, 1, 0], (0, 9, u): [841665] * 5,
        (0, 2, q): [134384483009, 100507578, 19687499137, 19687499137, 100507578],
        (0, 3, u): [1381619] * 5, (0, 4, q): [920714] * 5, (0, 5, q): [105320959] * 5,
        (0, 6, T.INT64): [1] * 5, (0, 7, T.INT64): [0] * 5, (0, 8, T.INT64): [0] * 5, (0, 10, u): [1] * 5,
        (1, 1, T.INT64): [2, 5, 6, 1, 8], (1, 2, T.INT64): [0, 0, 0, 0, 2], (1, 3, T.INT64): [0, 0, 0, 0, 7664],
    }
    return [make_column(t, s, pt, v) for (t, s, pt), v in data.items()]


def test_arrow_compute_fixture_six_key_group_by():
    """keys 0_1,0_9,0_2,0_3,0_4,0_5; hash_sum over 1_1,1_2,1_3 (the reference prints the result; the five
    key tuples are distinct, so every sum equals its row's value).  Row oracle == Acero == expected."""
    cols = _arrow_compute_fixture()
    keys = [(0, 1, T.UINT32), (0, 9, T.UINT32), (0, 2, T.UINT64), (0, 3, T.UINT32), (0, 4, T.UINT64), (0, 5, T.UINT64)]
    aggs = [P.agg_expr("sum", 2, i + 1, None, P.slot_ref(1, i + 1, T.INT64)) for i in range(3)]
    root = P.agg(P.scan(0), 2, [P.slot_ref(t, s, pt) for t, s, pt in keys], aggs)
    tuples = {0: [(s, pt) for (t, s, pt) in [(c.tuple_id, c.slot_id, c.prim_type) for c in cols] if t == 0],
              1: [(1, T.INT64), (2, T.INT64), (3, T.INT64)], 2: P.agg_tuple_slots(aggs, [T.INT64] * 3)}
    res = oracle.execute(P.Plan(root, tuples).serialize(), cols)
    got = rows_as_set(res.columns, ["0_1", "0_2"])
    expected = {(0, 134384483009): (2, 0, 0), (1, 100507578): (5, 0, 0), (0, 19687499137): (6, 0, 0),
                (1, 19687499137): (1, 0, 0), (0, 100507578): (8, 2, 7664)}
    names = [c.name for c in res.columns]
    assert set(got) == set(expected)
    for k, sums in expected.items():
        row = got[k]
        assert tuple(row[names.index(n)] for n in ("2_1", "2_2", "2_3")) == sums
    table = A.to_table(cols)
    acero = A.filter_groupby(table, None, ["0_1", "0_9", "0_2", "0_3", "0_4", "0_5"],
                             [("hash_sum", "1_1", "2_1"), ("hash_sum", "1_2", "2_2"), ("hash_sum", "1_3", "2_3")])
    arows = A.table_rows(acero, ["0_1", "0_2"])
    an = acero.column_names
    for k, sums in expected.items():
        assert tuple(arows[k][an.index(n)] for n in ("2_1", "2_2", "2_3")) == sums


# ---- row-engine restatement == Acero plan on the BASELINE configs (small sizes) ----
def test_c1_row_oracle_equals_acero_and_numpy():
    cols = datagen.c1_table(0, 1_000_000)
    res = oracle.execute(queries.c1_count_where().serialize(), cols)
    want = int((cols[0].values < (1 << 19)).sum())
    assert res.columns[0].to_list() == [want]
    assert A.c1_count_where(A.to_table(cols), 1 << 19).column("1_1").to_pylist() == [want]
    assert res.rows_scanned == 1_000_000 and res.rows_filtered == 1_000_000 - want


@pytest.mark.parametrize("k", [1 << 19, 10486])
def test_c2_row_oracle_equals_acero(k):
    cols = datagen.c2_table(0, 300_000)
    res = oracle.execute(queries.c2_filter_groupby(k).serialize(), cols)
    got = rows_as_set(res.columns, ["0_1"])
    names = [c.name for c in res.columns]
    ac = A.c2_filter_groupby(A.to_table(cols), k)
    arows = A.table_rows(ac, ["0_1"])
    an = ac.column_names
    assert set(got) == set(arows)
    for key, row in got.items():
        assert row[names.index("1_1")] == arows[key][an.index("1_1")]                      # COUNT(*) bit-exact
        assert row[names.index("1_2")] == pytest.approx(arows[key][an.index("1_2")], rel=1e-9)   # SUM(double)
        assert row[names.index("1_3")] == pytest.approx(arows[key][an.index("1_3")], rel=1e-9)   # AVG(double)
        blob = row[names.index("1_4")]                                                     # AVG intermediate {sum, count}
        s, c = struct.unpack("<dq", blob)
        assert c == row[names.index("1_1")] and s / c == row[names.index("1_3")]


def test_c3_row_oracle_equals_acero():
    fact, dim = datagen.c3_fact(0, 200_000, 5000), datagen.c3_dim(0, 5000, 5000, n_groups=40)
    assert sorted(dim[0].values.tolist()) == list(range(5000))  # the generator's permutation is a bijection
    res = oracle.execute(queries.c3_join_groupby().serialize(), fact + dim)
    got = rows_as_set(res.columns, ["1_2"])
    names = [c.name for c in res.columns]
    ac = A.c3_join_groupby(A.to_table(fact), A.to_table(dim))
    arows = A.table_rows(ac, ["1_2"])
    an = ac.column_names
    assert set(got) == set(arows)
    for key, row in got.items():
        assert row[names.index("2_1")] == arows[key][an.index("2_1")]
        assert row[names.index("2_2")] == pytest.approx(arows[key][an.index("2_2")], rel=1e-9)


def test_c5_row_oracle_equals_acero_with_ties():
    rng = np.random.default_rng(9)
    n = 50_000
    cols = [make_column(0, 1, T.INT64, rng.integers(0, 2000, n)), make_column(0, 2, T.INT32, np.arange(n))]
    res = oracle.execute(queries.c5_topk(1000).serialize(), cols)
    keys, payload = res.columns[0].to_list(), res.columns[1].to_list()
    order = np.lexsort((np.arange(n), cols[0].values))[:1000]          # stable: ties by arrival index (topn_sorter.h:96-106)
    assert keys == cols[0].values[order].tolist() and payload == order.tolist()
    ac = A.c5_topk(A.to_table(cols), 1000)
    assert ac.column("0_1").to_pylist() == keys                        # Acero's sort is not stable: keys only


def test_empty_input_semantics():
    empty = [make_column(0, 1, T.INT32, np.zeros(0, np.int32))]
    assert oracle.execute(queries.c1_count_where().serialize(), empty).columns[0].to_list() == [0]   # under PACKET: COUNT=0 row
    r = oracle.execute(queries.c2_filter_groupby().serialize(), datagen.c2_table(0, 0))
    assert r.nrows == 0


def test_three_valued_logic_and_null_arithmetic():
    n = 6
    a = make_column(0, 1, T.INT32, [1, 0, 1, 0, 5, 5], [True, True, False, False, True, True])
    b = make_column(0, 2, T.INT64, [0, 0, 0, 0, 0, 2])
    c = make_column(0, 3, T.DOUBLE, np.arange(n, dtype=float))
    d = make_column(0, 4, T.INT32, np.arange(n))
    s1, s2 = P.slot_ref(0, 1, T.INT32), P.slot_ref(0, 2, T.INT64)
    # WHERE (a = 1 OR a IS NULL) AND NOT (b / b > 0)   -- b/0 is NULL, NOT NULL is NULL -> row dropped
    aggs = [P.agg_expr("count_star", 1, 1)]
    conj = [P.or_(P.eq(s1, P.int_lit(1)), P.is_null(s1)), P.not_(P.gt(P.divides(s2, s2), P.int_lit(0)))]
    root = P.agg(P.where(P.scan(0), *conj), 1, [], aggs)
    pl = P.Plan(P.packet(root), {0: [(1, T.INT32), (2, T.INT64), (3, T.DOUBLE), (4, T.INT32)], 1: P.agg_tuple_slots(aggs, [T.INT64])})
    assert oracle.execute(pl.serialize(), [a, b, c, d]).columns[0].to_list() == [0]
    conj2 = [P.or_(P.eq(s1, P.int_lit(1)), P.is_null(s1))]
    root2 = P.agg(P.where(P.scan(0), *conj2), 1, [], aggs)
    pl2 = P.Plan(P.packet(root2), pl.tuples)
    assert oracle.execute(pl2.serialize(), [a, b, c, d]).columns[0].to_list() == [3]


# ---- test/test_expr_value.cpp TEST(type_merge, type_merge) :655-840: return types of case_when / if (complete_common_fn +
#      has_merged_type).  The numeric vectors are on the GPU path; both the oracle's inference and the device lowering
#      (bkgpu_plan_explain, host-only) must give the reference's answer.  (STRING / date-time merges stay outside the path.) ----
TYPE_MERGE_VECTORS = [
    ("case_when", [T.INT8, T.INT64], T.INT64),                       # {STRING, INT8, INT64}        -> INT64   (:667-673)
    ("case_when2", [T.INT8, T.DOUBLE], T.DOUBLE),                    # {STRING, INT8, INT64, DOUBLE} -> DOUBLE  (:676-682; index 2 is a condition)
    ("case_when", [T.INT64, T.UINT64], T.DOUBLE),                    # {STRING, INT64, UINT64}       -> DOUBLE  (:685-691)
    ("if", [T.INT8, T.INT64], T.INT64),                              # if {STRING, INT8, INT64}      -> INT64
    ("if", [T.INT64, T.UINT64], T.DOUBLE),                           # if {STRING, INT64, UINT64}    -> DOUBLE
    ("if", [T.NULL_TYPE, T.INT8], T.INT8),                           # if {STRING, NULL_TYPE, INT8}  -> INT8
]


@pytest.mark.parametrize("fn_name,branch_types,expected", TYPE_MERGE_VECTORS)
def test_type_merge_known_answers(fn_name, branch_types, expected):
    import re
    from baikaldb_b200 import _lib
    slots = {T.INT8: 1, T.INT64: 2, T.UINT64: 3, T.DOUBLE: 4}
    tuple0 = [(s, t) for t, s in slots.items()] + [(9, T.INT32)]
    branch = lambda t: P.null_lit() if t == T.NULL_TYPE else P.slot_ref(0, slots[t], t)
    cond = P.gt(P.slot_ref(0, 9, T.INT32), P.int_lit(0))
    a, b = branch(branch_types[0]), branch(branch_types[1])
    if fn_name == "if":
        e = P.if_(cond, a, b)
    elif fn_name == "case_when2":
        e = P.case_when(cond, a, P.lt(P.slot_ref(0, 2, T.INT64), P.int_lit(5)), b)     # WHEN c THEN a WHEN c2(INT64-typed slot in it) ... ELSE b
    else:
        e = P.case_when(cond, a, b)                                                     # WHEN c THEN a ELSE b
    aggs = [P.agg_expr("min", 1, 1, None, e)]
    pl = P.Plan(P.agg(P.scan(0), 1, [], aggs), {0: sorted(tuple0), 1: []})     # the aggregate slot is left undeclared: its type is the inferred one
    cols = [make_column(0, 1, T.INT8, [1, -2]), make_column(0, 2, T.INT64, [10, 3]), make_column(0, 3, T.UINT64, [7, 8]),
            make_column(0, 4, T.DOUBLE, [0.5, 1.5]), make_column(0, 9, T.INT32, [1, 0])]
    res = oracle.execute(pl.serialize(), cols)
    assert res.columns[0].prim_type == expected                                         # MIN(x) carries x's inferred type
    out_prims = re.findall(r"agg\[0\].*out_prim=(\d+)", _lib.explain(pl.serialize()))
    assert out_prims and int(out_prims[0]) == int(expected)


def test_nullable_min_max_sum_count_row_oracle_equals_acero():
    """NULL keys form their own group, NULL filter operands drop the row, aggregates skip NULL inputs, all-NULL groups give NULL
    (COUNT 0): the row-engine restatement against Acero's hash_min / hash_max / hash_sum / hash_count / hash_mean on the same table
    (the Acero functions the vectorized engine maps them to, src/expr/agg_fn_call.cpp:1365-1392)."""
    import pyarrow.compute as pc
    rng = np.random.default_rng(5)
    n = 20_000
    cols = [make_column(0, 1, T.INT32, rng.integers(0, 15, n), rng.random(n) > 0.1), make_column(0, 2, T.INT32, rng.integers(0, 100, n), rng.random(n) > 0.15),
            make_column(0, 3, T.DOUBLE, rng.normal(size=n) * 100, (rng.random(n) > 0.4) & (rng.integers(0, 15, n) != 3)),
            make_column(0, 4, T.INT64, rng.integers(-1 << 40, 1 << 40, n), rng.random(n) > 0.3)]
    d, i64 = P.slot_ref(0, 3, T.DOUBLE), P.slot_ref(0, 4, T.INT64)
    aggs = [P.agg_expr("count_star", 1, 1), P.agg_expr("min", 1, 2, None, d), P.agg_expr("max", 1, 3, None, d), P.agg_expr("sum", 1, 4, None, d),
            P.agg_expr("count", 1, 5, None, d), P.agg_expr("max", 1, 6, None, i64), P.agg_expr("sum", 1, 7, None, i64), P.agg_expr("avg", 1, 8, 9, i64)]
    root = P.agg(P.where(P.scan(0), P.lt(P.slot_ref(0, 2, T.INT32), P.int_lit(70))), 1, [P.slot_ref(0, 1, T.INT32)], aggs)
    pl = P.Plan(root, {0: [(1, T.INT32), (2, T.INT32), (3, T.DOUBLE), (4, T.INT64)],
                       1: P.agg_tuple_slots(aggs, [T.INT64, T.DOUBLE, T.DOUBLE, T.DOUBLE, T.DOUBLE, T.INT64, T.INT64, T.INT64])})
    row = oracle.execute(pl.serialize(), cols)
    t = A.filter_groupby(A.to_table(cols), pc.field("0_2") < pc.scalar(pa_scalar(70)), ["0_1"],
                         [("hash_count_all", None, "1_1"), ("hash_min", "0_3", "1_2"), ("hash_max", "0_3", "1_3"), ("hash_sum", "0_3", "1_4"),
                          ("hash_count", "0_3", "1_5"), ("hash_max", "0_4", "1_6"), ("hash_sum", "0_4", "1_7"), ("hash_mean", "0_4", "1_8")])
    want = A.table_rows(t, ["0_1"])
    got = rows_as_set(row.columns, ["0_1"])
    names = [c.name for c in row.columns]
    assert set(got) == set(want) and (None,) in got
    tn = t.column_names
    for k in got:
        for nm in ("1_1", "1_2", "1_3", "1_5", "1_6", "1_7"):          # exact: counts, extremes, integer sums
            assert got[k][names.index(nm)] == want[k][tn.index(nm)], (k, nm)
        for nm in ("1_4", "1_8"):                                       # double sum / mean: same adds, different order
            a, b = got[k][names.index(nm)], want[k][tn.index(nm)]
            assert (a is None and b is None) or abs(a - b) <= 1e-9 * max(1.0, abs(b)), (k, nm, a, b)


def pa_scalar(v):
    import pyarrow as pa
    return pa.scalar(v, pa.int32())


def test_multi_key_group_by_with_null_keys_row_oracle_equals_acero():
    """two GROUP BY expressions, NULLs in both: (NULL, x), (x, NULL) and (NULL, NULL) are groups of their own
    (encode_exprs_key's null-flag byte, exec_node.cpp:555-571; Acero groups NULL keys the same way)"""
    rng = np.random.default_rng(21)
    n = 30_000
    cols = [make_column(0, 1, T.INT32, rng.integers(0, 6, n), rng.random(n) > 0.15), make_column(0, 2, T.INT64, rng.integers(-2, 3, n), rng.random(n) > 0.15),
            make_column(0, 3, T.DOUBLE, rng.random(n))]
    aggs = [P.agg_expr("count_star", 1, 1), P.agg_expr("sum", 1, 2, None, P.slot_ref(0, 3, T.DOUBLE))]
    pl = P.Plan(P.agg(P.scan(0), 1, [P.slot_ref(0, 1, T.INT32), P.slot_ref(0, 2, T.INT64)], aggs),
                {0: [(1, T.INT32), (2, T.INT64), (3, T.DOUBLE)], 1: P.agg_tuple_slots(aggs, [T.INT64, T.DOUBLE])})
    row = oracle.execute(pl.serialize(), cols)
    got = rows_as_set(row.columns, ["0_1", "0_2"])
    t = A.filter_groupby(A.to_table(cols), None, ["0_1", "0_2"], [("hash_count_all", None, "1_1"), ("hash_sum", "0_3", "1_2")])
    want = A.table_rows(t, ["0_1", "0_2"])
    assert set(got) == set(want) and (None, None) in got and any(k[0] is None and k[1] is not None for k in got)
    names, tn = [c.name for c in row.columns], t.column_names
    for k in got:
        assert got[k][names.index("1_1")] == want[k][tn.index("1_1")]
        assert got[k][names.index("1_2")] == pytest.approx(want[k][tn.index("1_2")], rel=1e-9)


def test_join_with_duplicate_and_null_keys_row_oracle_equals_acero():
    """inner equi-join with several build rows per key and NULL keys on both sides: every (probe, build) pair counts once, NULL never
    matches (Joiner multimap, joiner.cpp:608-685; Acero hashjoin 'inner')"""
    rng = np.random.default_rng(23)
    nd, nf = 2_000, 40_000
    dim = [make_column(1, 1, T.INT32, rng.integers(0, 500, nd), rng.random(nd) > 0.1), make_column(1, 2, T.INT32, rng.integers(0, 12, nd))]
    fact = [make_column(0, 1, T.INT32, rng.integers(-20, 520, nf), rng.random(nf) > 0.1), make_column(0, 2, T.DOUBLE, rng.random(nf))]
    res = oracle.execute(queries.c3_join_groupby().serialize(), fact + dim)
    got = rows_as_set(res.columns, ["1_2"])
    names = [c.name for c in res.columns]
    ac = A.c3_join_groupby(A.to_table(fact), A.to_table(dim))
    want = A.table_rows(ac, ["1_2"])
    an = ac.column_names
    assert set(got) == set(want)
    pairs = 0
    for key, row in got.items():
        assert row[names.index("2_1")] == want[key][an.index("2_1")]
        assert row[names.index("2_2")] == pytest.approx(want[key][an.index("2_2")], rel=1e-9)
        pairs += row[names.index("2_1")]
    # independent count of matching pairs
    dk = dim[0].values[dim[0].valid]
    fk = fact[0].values[fact[0].valid]
    assert pairs == int(np.bincount(dk, minlength=600)[np.clip(fk, 0, 599)][(fk >= 0) & (fk < 600)].sum())


@pytest.mark.parametrize("asc", [True, False])
def test_order_by_null_placement_row_oracle_equals_acero(asc):
    """ORDER BY with NULL keys: the planner sets is_null_first = is_asc (logical_planner.cpp:4071) — NULLs first ascending, last
    descending (MemRowCompare, mem_row_compare.cpp:18-38); Acero's order_by with the matching null_placement agrees on the keys"""
    rng = np.random.default_rng(29)
    n = 5_000
    cols = [make_column(0, 1, T.INT64, rng.integers(-300, 300, n), rng.random(n) > 0.2), make_column(0, 2, T.INT32, np.arange(n))]
    res = oracle.execute(queries.c5_topk(n, asc=asc).serialize(), cols)
    keys = res.columns[0].to_list()
    ac = A.c5_topk(A.to_table(cols), n, ascending=asc)
    assert keys == ac.column("0_1").to_pylist()
    nn = int((~cols[0].valid).sum())
    assert (keys[:nn] == [None] * nn) if asc else (keys[-nn:] == [None] * nn)


@pytest.mark.parametrize("jt", ["LEFT_JOIN", "RIGHT_JOIN", "SEMI_JOIN", "ANTI_SEMI_JOIN"])
def test_row_oracle_outer_semi_anti_joins_equal_acero(jt):
    """the row restatement of LEFT / RIGHT / SEMI / ANTI_SEMI (join_node.cpp:151-156,1200-1276; joiner.cpp:633-685) against the
    Acero hashjoin the reference's vectorized engine declares for them (join_node.cpp:853-868: LEFT_OUTER / LEFT_SEMI / LEFT_ANTI,
    outer side first) — duplicate keys, NULL keys on both sides, outer rows without a partner"""
    import pyarrow.acero as ac
    from baikaldb_b200 import plan as P
    from baikaldb_b200.column import make_column
    from baikaldb_b200.plan import PrimitiveType as T
    from oracle import acero_oracle as A, oracle
    rng = np.random.default_rng(5)
    nd, nf = 2_500, 40_000
    dim = [make_column(1, 1, T.INT32, rng.integers(0, 1_200, nd), rng.random(nd) > 0.04), make_column(1, 2, T.INT32, rng.integers(0, 12, nd))]
    fact = [make_column(0, 1, T.INT32, rng.integers(400, 1_500, nf), rng.random(nf) > 0.05), make_column(0, 2, T.DOUBLE, rng.random(nf), rng.random(nf) > 0.1)]
    semi = "SEMI" in jt
    aggs = [P.agg_expr("count_star", 2, 1)] + ([] if semi else [P.agg_expr("count", 2, 2, None, P.slot_ref(0, 2, T.DOUBLE)), P.agg_expr("sum", 2, 3, None, P.slot_ref(0, 2, T.DOUBLE))])
    ch = (P.scan(0), P.scan(1)) if jt == "RIGHT_JOIN" else (P.scan(1), P.scan(0))
    j = P.join(ch[0], ch[1], [P.eq(P.slot_ref(1, 1, T.INT32), P.slot_ref(0, 1, T.INT32))], join_type=getattr(P.JoinType, jt))
    pl = P.Plan(P.agg(j, 2, [P.slot_ref(1, 2, T.INT32)], aggs), {0: [(1, T.INT32), (2, T.DOUBLE)], 1: [(1, T.INT32), (2, T.INT32)],
                                                                   2: P.agg_tuple_slots(aggs, [T.INT64] if semi else [T.INT64, T.INT64, T.DOUBLE])})
    r = oracle.execute(pl.serialize(), fact + dim)
    mine = {c.name: c.to_list() for c in r.columns}
    at = {"LEFT_JOIN": "left outer", "RIGHT_JOIN": "left outer", "SEMI_JOIN": "left semi", "ANTI_SEMI_JOIN": "left anti"}[jt]
    jd = ac.Declaration("hashjoin", ac.HashJoinNodeOptions(at, ["1_1"], ["0_1"]),
                        inputs=[ac.Declaration("table_source", ac.TableSourceNodeOptions(A.to_table(dim))), ac.Declaration("table_source", ac.TableSourceNodeOptions(A.to_table(fact)))])
    specs = [([], "hash_count_all", None, "c")] + ([] if semi else [("0_2", "hash_count", None, "n"), ("0_2", "hash_sum", None, "s")])
    t = ac.Declaration("aggregate", ac.AggregateNodeOptions(specs, keys=["1_2"]), inputs=[jd]).to_table()
    want = {k: i for i, k in enumerate(t.column("1_2").to_pylist())}
    assert sorted(mine["1_2"]) == sorted(want)
    for i, k in enumerate(mine["1_2"]):
        assert mine["2_1"][i] == t.column("c")[want[k]].as_py()
        if not semi:
            assert mine["2_2"][i] == t.column("n")[want[k]].as_py()
            a, b = mine["2_3"][i], t.column("s")[want[k]].as_py()
            assert (a is None and b is None) or abs(a - b) <= 1e-9 * max(1.0, abs(b))


@pytest.mark.parametrize("jt", ["INNER_JOIN", "LEFT_JOIN", "RIGHT_JOIN"])
def test_row_oracle_joined_rows_equal_acero(jt):
    """a JOIN that returns its rows: the row restatement's joined rows are Acero's hashjoin rows (inner / left outer), as multisets —
    duplicate keys and NULL keys on both sides"""
    import pyarrow.acero as ac
    from baikaldb_b200 import plan as P
    from baikaldb_b200.column import make_column
    from baikaldb_b200.plan import PrimitiveType as T
    from oracle import acero_oracle as A, oracle
    rng = np.random.default_rng(8)
    nd, nf = 900, 12_000
    dim = [make_column(1, 1, T.INT32, rng.integers(0, 500, nd), rng.random(nd) > 0.04), make_column(1, 2, T.INT32, rng.integers(0, 12, nd))]
    fact = [make_column(0, 1, T.INT32, rng.integers(200, 700, nf), rng.random(nf) > 0.05), make_column(0, 2, T.DOUBLE, rng.random(nf), rng.random(nf) > 0.1)]
    ch = (P.scan(0), P.scan(1)) if jt == "RIGHT_JOIN" else (P.scan(1), P.scan(0))
    j = P.join(ch[0], ch[1], [P.eq(P.slot_ref(1, 1, T.INT32), P.slot_ref(0, 1, T.INT32))], join_type=getattr(P.JoinType, jt))
    r = oracle.execute(P.Plan(P.packet(j), {0: [(1, T.INT32), (2, T.DOUBLE)], 1: [(1, T.INT32), (2, T.INT32)]}).serialize(), fact + dim)
    names = ["0_1", "0_2", "1_1", "1_2"]
    mine = {c.name: c.to_list() for c in r.columns}
    key = lambda t: tuple((0, 0) if v is None else (1, v) for v in t)
    got = sorted(zip(*[mine[n] for n in names]), key=key)
    jd = ac.Declaration("hashjoin", ac.HashJoinNodeOptions("inner" if jt == "INNER_JOIN" else "left outer", ["1_1"], ["0_1"]),
                        inputs=[ac.Declaration("table_source", ac.TableSourceNodeOptions(A.to_table(dim))), ac.Declaration("table_source", ac.TableSourceNodeOptions(A.to_table(fact)))])
    t = jd.to_table()
    want = sorted(zip(*[t.column(n).to_pylist() for n in names]), key=key)
    assert len(got) == len(want) > 1000 and got == want


def test_row_oracle_distinct_aggregates_equal_numpy():
    """COUNT / SUM / AVG (DISTINCT x) through the planner's two-level layout (tests/test_gpu_merge.py::_distinct_case): oracle rows == numpy sets"""
    from oracle import oracle
    from tests.test_gpu_merge import _distinct_case
    (k, x, xv, v), cols, low, top = _distinct_case(seed=3, n=20_000)
    mid = oracle.execute(low.serialize(), cols)
    res = oracle.execute(top.serialize(), list(mid.columns))
    by = {c.name: c.to_list() for c in res.columns}
    assert sorted(by["0_1"]) == sorted(set(k.tolist()))
    for i, kk in enumerate(by["0_1"]):
        m = k == kk
        dx = sorted(set(x[m & xv].tolist()))
        assert by["1_3"][i] == len(dx) and by["1_4"][i] == sum(dx) and by["1_2"][i] == int(m.sum())
        assert abs(by["1_5"][i] - sum(dx) / len(dx)) < 1e-9 and abs(by["1_1"][i] - v[m].sum()) < 1e-9 * max(1.0, v[m].sum())
