"""Named builtins on the path: IF / IFNULL / CASE WHEN, ABS / FLOOR / CEIL / ROUND, CAST AS SIGNED / UNSIGNED /
DOUBLE — src/expr/internal_functions.cpp:52-99,2351-2395,2941-2963, typed by fn_manager.cpp:398-401,466-514 and
has_merged_type (include/common/type_utils.h:502-560).  CPU: the oracle against hand-computed answers that follow the
reference's definitions line by line; GPU: device bytecode == oracle, inside filters, GROUP BY keys and aggregate arguments."""
import math

import numpy as np
import pytest

from baikaldb_b200 import _lib, plan as P
from baikaldb_b200.column import make_column
from baikaldb_b200.plan import PrimitiveType as T
from oracle import oracle

TUPLES0 = [(1, T.INT32), (2, T.DOUBLE), (3, T.INT64), (4, T.UINT64)]


def _table(n=4000, seed=9):
    rng = np.random.default_rng(seed)
    return [make_column(0, 1, T.INT32, rng.integers(-6, 6, n), rng.random(n) > 0.15),
            make_column(0, 2, T.DOUBLE, np.round(rng.normal(scale=20, size=n), 3), rng.random(n) > 0.15),
            make_column(0, 3, T.INT64, rng.integers(-1000, 1000, n), rng.random(n) > 0.15),
            make_column(0, 4, T.UINT64, rng.integers(0, 1 << 40, n, dtype=np.uint64))]


def _c(i): return P.slot_ref(0, i, dict(TUPLES0)[i])


def _group_plan(key_expr, key_type, arg_expr, arg_type, where=None):
    aggs = [P.agg_expr("count_star", 1, 1), P.agg_expr("sum", 1, 2, None, arg_expr), P.agg_expr("min", 1, 3, None, arg_expr), P.agg_expr("count", 1, 4, None, arg_expr)]
    child = P.scan(0) if where is None else P.where(P.scan(0), where)
    root = P.agg(child, 1, [key_expr], aggs)
    sum_t = T.DOUBLE if arg_type in (T.DOUBLE, T.FLOAT) else (T.UINT64 if arg_type == T.UINT64 else T.INT64)
    return P.Plan(root, {0: TUPLES0, 1: [(1, T.INT64), (2, sum_t), (3, arg_type), (4, T.INT64)]})


def _scalar_rows(expr, out_type, cols):
    """evaluate `expr` per row through the oracle: GROUP BY a unique row id, MIN(expr)"""
    n = len(cols[0])
    rid = make_column(0, 9, T.INT64, np.arange(n))
    aggs = [P.agg_expr("min", 1, 1, None, expr)]
    root = P.agg(P.scan(0), 1, [P.slot_ref(0, 9, T.INT64)], aggs)
    pl = P.Plan(root, {0: TUPLES0 + [(9, T.INT64)], 1: [(1, out_type)]})
    res = oracle.execute(pl.serialize(), cols + [rid])
    by = dict(zip(res.columns[0].to_list(), res.columns[1].to_list()))
    return [by[i] for i in range(n)], pl, cols + [rid]


def test_oracle_known_answers():
    cols = [make_column(0, 1, T.INT32, [3, -2, 0, 5], [True, True, True, False]), make_column(0, 2, T.DOUBLE, [2.5, -2.5, 0.125, 7.0], [True, True, False, True]),
            make_column(0, 3, T.INT64, [10, 20, 30, 40]), make_column(0, 4, T.UINT64, [1, 2, 3, 4])]
    a, d, i64 = _c(1), _c(2), _c(3)
    ev = lambda e, t: _scalar_rows(e, t, cols)[0]
    assert ev(P.if_(P.gt(a, P.int_lit(0)), d, i64), T.DOUBLE) == [2.5, 20.0, 30.0, 40.0]           # NULL condition is false; INT64 branch widens to DOUBLE
    assert ev(P.ifnull(d, a), T.DOUBLE) == [2.5, -2.5, 0.0, 7.0]
    assert ev(P.ifnull(a, P.null_lit()), T.INT32) == [3, -2, 0, None]
    assert ev(P.case_when(P.lt(a, P.int_lit(0)), P.int_lit(-1), P.eq(a, P.int_lit(0)), P.int_lit(0), P.int_lit(1)), T.INT64) == [1, -1, 0, 1]
    assert ev(P.case_when(P.lt(a, P.int_lit(0)), i64), T.INT64) == [None, 20, None, None]        # no ELSE -> NULL
    assert ev(P.abs_(a), T.DOUBLE) == [3.0, 2.0, 0.0, None]
    assert ev(P.floor_(d), T.INT64) == [2, -3, None, 7] and ev(P.ceil_(d), T.INT64) == [3, -2, None, 7]
    assert ev(P.round_(d), T.DOUBLE) == [3.0, -3.0, None, 7.0]                                   # half away from zero
    assert ev(P.round_(P.divides(i64, P.int_lit(3)), P.int_lit(2)), T.DOUBLE) == [3.33, 6.67, 10.0, 13.33]
    assert ev(P.cast_to_signed(d), T.INT64) == [2, -2, None, 7]                                  # static_cast truncates
    assert ev(P.cast_to_unsigned(a), T.UINT64) == [3, (1 << 64) - 2, 0, None]
    assert ev(P.cast_to_double(i64), T.DOUBLE) == [10.0, 20.0, 30.0, 40.0]


def test_round_matches_the_references_known_answers():
    """test/test_internal_functions.cpp:34-166 (TEST(round, round)): the twelve (value, decimals) -> result vectors of the reference's own
    unit test, through the oracle's restatement of `round` (internal_functions.cpp) — decimals beyond the double's digits leave the value,
    negative decimals round left of the point, |decimals| past the range gives 0"""
    vectors = [(3.1356, None, 3.0), (3.1356, 0, 3.0), (3.1356, 2, 3.14), (3.1356, 1, 3.1), (123456.1356, 30, 123456.1356), (123456.1356, -1, 123460.0),
               (123456.1356, -3, 123000.0), (123456.1356, -300, 0.0), (-3.1356, 2, -3.14), (-3.1356, 3, -3.136), (-123456.1356, -2, -123500.0),
               (-123456.1356, -30, 0.0)]
    for x, dec, want in vectors:
        cols = [make_column(0, 1, T.INT32, [0]), make_column(0, 2, T.DOUBLE, [x]), make_column(0, 3, T.INT64, [0]), make_column(0, 4, T.UINT64, [0])]
        e = P.round_(_c(2)) if dec is None else P.round_(_c(2), P.int_lit(dec))
        got, pl, _ = _scalar_rows(e, T.DOUBLE, cols)
        assert got == [want], (x, dec, got, want)
        _lib.explain(pl.serialize())   # the library lowers the same fragment (host side: parse, type inference, bytecode)


def test_oracle_known_answers_of_the_remaining_numeric_builtins():
    """sqrt / sign / ln / log / pow / mod / greatest / least / bit_count / pi / trigonometry (internal_functions.cpp:101-350): NULL outside
    the domain, DOUBLE arithmetic on get_numberic<double>() of the arguments"""
    cols = [make_column(0, 1, T.INT32, [4, -9, 0, 5], [True, True, True, False]), make_column(0, 2, T.DOUBLE, [2.0, -0.5, 1.0, 7.0], [True, True, True, True]),
            make_column(0, 3, T.INT64, [10, 20, 30, 40]), make_column(0, 4, T.UINT64, [1, 255, (1 << 64) - 1, 0])]
    a, d, i64, u64 = _c(1), _c(2), _c(3), _c(4)
    ev = lambda e, t: _scalar_rows(e, t, cols)[0]
    assert ev(P.sqrt_(a), T.DOUBLE) == [2.0, None, 0.0, None]                       # negative -> NULL
    assert ev(P.sign_(d), T.INT64) == [1, -1, 1, 1] and ev(P.sign_(a), T.INT64) == [1, -1, 0, None]
    assert ev(P.ln_(d), T.DOUBLE) == [math.log(2.0), None, 0.0, math.log(7.0)]      # <= 0 -> NULL
    assert ev(P.log_(P.int_lit(2), i64), T.DOUBLE) == [math.log(10) / math.log(2), math.log(20) / math.log(2), math.log(30) / math.log(2), math.log(40) / math.log(2)]
    assert ev(P.log_(d, i64), T.DOUBLE)[1:3] == [None, None]                         # base <= 0, base == 1 -> NULL
    assert ev(P.pow_(d, P.int_lit(3)), T.DOUBLE) == [8.0, -0.125, 1.0, 343.0]
    assert ev(P.fmod_(i64, d), T.DOUBLE) == [0.0, 0.0, 0.0, 5.0] and ev(P.fmod_(i64, P.int_lit(0)), T.DOUBLE) == [None] * 4
    assert ev(P.fmod_(P.uminus(i64), P.int_lit(7)), T.DOUBLE) == [-3.0, -6.0, -2.0, -5.0]   # std::fmod keeps the dividend's sign
    assert ev(P.greatest(a, d, P.int_lit(1)), T.DOUBLE) == [4.0, 1.0, 1.0, None] and ev(P.least(a, d, P.int_lit(1)), T.DOUBLE) == [1.0, -9.0, 0.0, None]
    assert ev(P.bit_count(u64), T.INT64) == [1, 8, 64, 0] and ev(P.bit_count(P.uminus(P.int_lit(1))), T.INT64) == [64] * 4
    assert ev(P.pi_(), T.DOUBLE) == [math.pi] * 4
    assert ev(P.trig("asin", d), T.DOUBLE) == [None, math.asin(-0.5), math.asin(1.0), None] and ev(P.trig("cot", a), T.DOUBLE)[2] is None
    got = ev(P.trig("sin", d), T.DOUBLE)
    assert all(abs(g - math.sin(x)) < 1e-15 for g, x in zip(got, [2.0, -0.5, 1.0, 7.0]))


def test_lowering_accepts_and_rejects():
    ok = _group_plan(P.floor_(P.divides(_c(2), P.int_lit(10))), T.INT64, P.if_(P.gt(_c(1), P.int_lit(0)), _c(2), P.int_lit(0)), T.DOUBLE)
    text = _lib.explain(ok.serialize())
    assert "SELECT" in text and "MATH" in text
    with pytest.raises(_lib.BkgpuError):
        _lib.explain(_group_plan(_c(1), T.INT32, P.common("substr", _c(2)), T.DOUBLE).serialize())
    with pytest.raises(_lib.BkgpuError):
        _lib.explain(_group_plan(_c(1), T.INT32, P.round_(_c(2), _c(1)), T.DOUBLE).serialize())        # decimals must be a literal


CASES = [
    ("if_mixed", lambda: (_c(1), T.INT32, P.if_(P.gt(_c(1), P.int_lit(0)), _c(2), _c(3)), T.DOUBLE)),
    ("if_int_uint", lambda: (_c(1), T.INT32, P.if_(P.is_null(_c(2)), _c(3), _c(4)), T.DOUBLE)),       # INT64 vs UINT64 -> DOUBLE
    ("ifnull", lambda: (_c(1), T.INT32, P.ifnull(_c(2), P.double_lit(-1.5)), T.DOUBLE)),
    ("case_else", lambda: (P.case_when(P.lt(_c(1), P.int_lit(-2)), P.int_lit(0), P.lt(_c(1), P.int_lit(2)), P.int_lit(1), P.int_lit(2)), T.INT64, _c(3), T.INT64)),
    ("case_no_else", lambda: (_c(1), T.INT32, P.case_when(P.gt(_c(2), P.double_lit(0.0)), _c(3), P.lt(_c(2), P.double_lit(-10.0)), P.uminus(_c(3))), T.INT64)),
    ("floor_key", lambda: (P.floor_(P.divides(_c(2), P.int_lit(10))), T.INT64, P.abs_(_c(2)), T.DOUBLE)),
    ("ceil_round", lambda: (P.ceil_(_c(2)), T.INT64, P.round_(_c(2), P.int_lit(1)), T.DOUBLE)),
    ("round0_neg", lambda: (P.cast_to_signed(P.round_(_c(2))), T.INT64, P.round_(P.multiplies(_c(2), P.double_lit(0.5))), T.DOUBLE)),
    ("casts", lambda: (P.cast_to_signed(_c(2)), T.INT64, P.cast_to_double(P.cast_to_unsigned(_c(3))), T.DOUBLE)),
    ("sqrt_ln", lambda: (P.sign_(_c(2)), T.INT64, P.add(P.ifnull(P.sqrt_(_c(2)), P.double_lit(-1.0)), P.ifnull(P.ln_(_c(3)), P.double_lit(0.25))), T.DOUBLE)),
    ("pow_log", lambda: (_c(1), T.INT32, P.add(P.pow_(P.divides(_c(2), P.int_lit(10)), P.int_lit(3)), P.ifnull(P.log_(P.int_lit(3), _c(3)), P.double_lit(0.0))), T.DOUBLE)),
    ("fmod_greatest", lambda: (P.cast_to_signed(P.fmod_(_c(3), P.int_lit(5))), T.INT64, P.minus(P.greatest(_c(1), _c(2), P.int_lit(0)), P.least(_c(2), _c(3))), T.DOUBLE)),
    ("bit_count_pi", lambda: (P.bit_count(_c(4)), T.INT64, P.multiplies(P.pi_(), P.bit_count(_c(3))), T.DOUBLE)),
    ("trig", lambda: (_c(1), T.INT32, P.add(P.add(P.trig("sin", _c(2)), P.trig("cos", _c(3))), P.add(P.ifnull(P.trig("asin", P.divides(_c(1), P.int_lit(5))), P.double_lit(9.0)), P.trig("atan", _c(2)))), T.DOUBLE)),
    ("tan_cot_acos", lambda: (_c(1), T.INT32, P.add(P.trig("tan", P.divides(_c(2), P.int_lit(40))), P.add(P.ifnull(P.trig("cot", _c(1)), P.double_lit(0.5)), P.ifnull(P.trig("acos", P.divides(_c(1), P.int_lit(6))), P.double_lit(4.0)))), T.DOUBLE)),
]


@pytest.mark.gpu
@pytest.mark.parametrize("name,build", CASES, ids=[c[0] for c in CASES])
def test_gpu_matches_oracle(name, build):
    from tests.util import run_both
    key, kt, arg, at = build()
    where = P.ifnull(P.gt(P.abs_(_c(2)), P.double_lit(1.0)), P.bool_lit(True)) if name in ("ifnull", "floor_key") else None
    pl = _group_plan(key, kt, arg, at, where)
    key_name = "0_1" if key.node_type == P.ExprNodeType.SLOT_REF else "-1_0"
    run_both(pl, _table(), keys=[key_name], rel=1e-9, abs_tol=1e-9)
