"""STRING columns through order-preserving dictionary codes (baikaldb_b200/dictionary.py): the rewritten fragment — INT32 codes, literal
comparisons as code thresholds — executed by the oracle equals pyarrow's own string kernels on the original strings: filters with every
comparison operator and literals that are / are not in the dictionary, IN, NULLs, GROUP BY a string key with MIN / MAX / COUNT of another
string column, a join on string keys (both sides share one dictionary), ORDER BY a string LIMIT k; unsupported uses are refused."""
import numpy as np
import pyarrow as pa
import pyarrow.compute as pc
import pytest

from baikaldb_b200 import dictionary as D, plan as P
from baikaldb_b200._lib import explain
from baikaldb_b200.column import make_column
from baikaldb_b200.plan import PrimitiveType as T
from oracle import oracle

WORDS = [b"", b"a", b"ab", b"abc", b"b", b"ba", b"k", b"kk", b"m", b"zebra", b"Zebra", b"\xe4\xb8\xad", b"0", b"10", b"9"]


def _strings(rng, n, null_frac=0.1, words=WORDS):
    idx = rng.integers(0, len(words), n)
    ok = rng.random(n) >= null_frac
    return [words[i] if o else None for i, o in zip(idx, ok)]


def _run(plan, string_cols, other_cols):
    enc = D.encode_strings(plan, string_cols)
    explain(enc.plan.serialize())                           # the library lowers the rewritten fragment (host side)
    res = oracle.execute(enc.plan.serialize(), enc.columns + other_cols)
    return enc, enc.decode(res.columns)


def test_filters_with_literals_group_by_string_min_max_count():
    rng = np.random.default_rng(1)
    n = 20_000
    s1, s2, s3 = _strings(rng, n), _strings(rng, n, 0.2), _strings(rng, n, 0.0)
    v = rng.normal(size=n)
    tbl = pa.table({"s1": pa.array(s1, pa.binary()), "s2": pa.array(s2, pa.binary()), "s3": pa.array(s3, pa.binary()), "v": v})
    S = lambda slot: P.slot_ref(0, slot, T.STRING)
    cases = [("ge_present", P.ge(S(3), P.str_lit("k")), pc.greater_equal(tbl["s3"], b"k")),
             ("gt_absent", P.gt(S(3), P.str_lit("c")), pc.greater(tbl["s3"], b"c")),
             ("lt_absent", P.lt(S(3), P.str_lit("aa")), pc.less(tbl["s3"], b"aa")),
             ("le_present", P.le(S(3), P.str_lit("ab")), pc.less_equal(tbl["s3"], b"ab")),
             ("ne_present", P.ne(S(3), P.str_lit("abc")), pc.not_equal(tbl["s3"], b"abc")),
             ("eq_absent", P.eq(S(3), P.str_lit("nope")), pc.equal(tbl["s3"], b"nope")),
             ("lit_left", P.lt(P.str_lit("b"), S(3)), pc.greater(tbl["s3"], b"b")),
             ("in", P.in_(S(3), P.str_lit("zebra"), P.str_lit("k"), P.str_lit("missing")), pc.is_in(tbl["s3"], value_set=pa.array([b"zebra", b"k", b"missing"], pa.binary()))),
             ("nullable_ne_absent", P.ne(S(2), P.str_lit("nope")), pc.not_equal(tbl["s2"], b"nope"))]
    for name, pred, mask in cases:
        aggs = [P.agg_expr("count_star", 1, 1), P.agg_expr("sum", 1, 2, None, P.slot_ref(0, 4, T.DOUBLE)), P.agg_expr("min", 1, 3, None, S(2)),
                P.agg_expr("max", 1, 4, None, S(2)), P.agg_expr("count", 1, 5, None, S(2))]
        plan = P.Plan(P.agg(P.where(P.scan(0), pred), 1, [S(1)], aggs),
                      {0: [(1, T.STRING), (2, T.STRING), (3, T.STRING), (4, T.DOUBLE)], 1: [(1, T.INT64), (2, T.DOUBLE), (3, T.STRING), (4, T.STRING), (5, T.INT64)]})
        scols = [D.StringColumn(0, 1, s1), D.StringColumn(0, 2, s2), D.StringColumn(0, 3, s3)]
        enc, got = _run(plan, scols, [make_column(0, 4, T.DOUBLE, v)])
        by = {c.name: (c.values if isinstance(c, D.StringColumn) else c.to_list()) for c in got}
        want = tbl.filter(pc.fill_null(mask, False)).group_by("s1", use_threads=False).aggregate([([], "count_all"), ("v", "sum"), ("s2", "min"), ("s2", "max"), ("s2", "count")]).to_pydict()
        w = {k: (c, sv, mn, mx, cn) for k, c, sv, mn, mx, cn in zip(want["s1"], want["count_all"], want["v_sum"], want["s2_min"], want["s2_max"], want["s2_count"])}
        g = {k: (c, sv, mn, mx, cn) for k, c, sv, mn, mx, cn in zip(by["0_1"], by["1_1"], by["1_2"], by["1_3"], by["1_4"], by["1_5"])}
        assert set(g) == set(w), (name, set(g) ^ set(w))
        for k in g:
            assert g[k][0] == w[k][0] and g[k][2:] == w[k][2:], (name, k, g[k], w[k])
            assert abs(g[k][1] - w[k][1]) <= 1e-9 * max(1.0, abs(w[k][1])), (name, k)


def test_join_on_string_keys_shares_one_dictionary():
    rng = np.random.default_rng(2)
    names = [f"dim{i:04d}".encode() for i in range(400)] + [b"only_in_dim"]
    dk = list(names); rng.shuffle(dk)
    dattr = rng.integers(0, 12, len(dk))
    fk = [names[i] if o else None for i, o in zip(rng.integers(0, 400, 30_000), rng.random(30_000) > 0.05)] + [b"only_in_fact"] * 7
    fv = rng.normal(size=len(fk))
    aggs = [P.agg_expr("count_star", 2, 1), P.agg_expr("sum", 2, 2, None, P.slot_ref(0, 2, T.DOUBLE))]
    j = P.join(P.scan(1), P.scan(0), [P.eq(P.slot_ref(1, 1, T.STRING), P.slot_ref(0, 1, T.STRING))])
    plan = P.Plan(P.agg(j, 2, [P.slot_ref(1, 2, T.INT32)], aggs), {0: [(1, T.STRING), (2, T.DOUBLE)], 1: [(1, T.STRING), (2, T.INT32)], 2: P.agg_tuple_slots(aggs, [T.INT64, T.DOUBLE])})
    enc = D.encode_strings(plan, [D.StringColumn(0, 1, fk), D.StringColumn(1, 1, dk)])
    assert enc.dictionaries[(0, 1)] is enc.dictionaries[(1, 1)]                     # one comparison domain
    explain(enc.plan.serialize())
    res = oracle.execute(enc.plan.serialize(), [enc.columns[0], make_column(0, 2, T.DOUBLE, fv), enc.columns[1], make_column(1, 2, T.INT32, dattr)])
    got = {k: (c, s) for k, c, s in zip(res.columns[0].to_list(), res.columns[1].to_list(), res.columns[2].to_list())}
    fact = pa.table({"k": pa.array(fk, pa.binary()), "v": fv}); dim = pa.table({"k": pa.array(dk, pa.binary()), "attr": dattr})
    want = fact.join(dim, "k", join_type="inner").group_by("attr", use_threads=False).aggregate([([], "count_all"), ("v", "sum")]).to_pydict()
    w = {k: (c, s) for k, c, s in zip(want["attr"], want["count_all"], want["v_sum"])}
    assert set(got) == set(w)
    for k in got:
        assert got[k][0] == w[k][0] and abs(got[k][1] - w[k][1]) <= 1e-9 * max(1.0, abs(w[k][1]))


@pytest.mark.parametrize("asc", [True, False])
def test_order_by_string_limit(asc):
    rng = np.random.default_rng(3)
    n = 5000
    s = _strings(rng, n, 0.15, [f"{i:03d}".encode() for i in range(300)])
    rid = np.arange(n, dtype=np.int32)
    plan = P.Plan(P.sort(P.scan(0), [P.slot_ref(0, 1, T.STRING)], [asc], limit=40, tuple_id=0), {0: [(1, T.STRING), (2, T.INT32)]})
    enc, got = _run(plan, [D.StringColumn(0, 1, s)], [make_column(0, 2, T.INT32, rid)])
    keys, rows = got[0].values, got[1].to_list()
    # the planner's default: NULLs first when ascending, last when descending; ties keep arrival order (TopNSorter is stable)
    order = sorted(range(n), key=lambda i: ((0, b"") if s[i] is None else (1, s[i]), i)) if asc else \
        sorted(range(n), key=lambda i: ((1, b"") if s[i] is None else (0, bytes(255 - b for b in s[i]) + b"\xff"), i))
    assert rows == [int(rid[i]) for i in order[:40]] and keys == [s[i] for i in order[:40]]


def test_unsupported_uses_of_strings_are_refused():
    s = D.StringColumn(0, 1, [b"a", b"b", None])
    S = P.slot_ref(0, 1, T.STRING)
    tuples = {0: [(1, T.STRING), (2, T.DOUBLE)], 1: [(1, T.DOUBLE)]}
    with pytest.raises(D.Unsupported):   # SUM over strings
        D.encode_strings(P.Plan(P.agg(P.scan(0), 1, [], [P.agg_expr("sum", 1, 1, None, S)]), tuples), [s])
    with pytest.raises(D.Unsupported):   # a string compared with a number
        D.encode_strings(P.Plan(P.agg(P.where(P.scan(0), P.gt(S, P.int_lit(3))), 1, [], [P.agg_expr("count_star", 1, 1)]), tuples), [s])
    with pytest.raises(D.Unsupported):   # a function of a string
        D.encode_strings(P.Plan(P.agg(P.where(P.scan(0), P.eq(P.common("length", S), P.int_lit(1))), 1, [], [P.agg_expr("count_star", 1, 1)]), tuples), [s])


@pytest.mark.gpu
@pytest.mark.skipif(__import__("os").environ.get("BKGPU_UNVERIFIED") != "1", reason="written after round 2's last GPU window: not yet run on a GPU")
def test_gpu_runs_the_rewritten_fragment():
    """the GPU sees only INT32 codes: the rewritten GROUP BY / MIN / MAX fragment through the C ABI equals the oracle, and decodes to pyarrow's answer"""
    from tests.util import run_both
    rng = np.random.default_rng(4)
    n = 200_000
    s1, s2 = _strings(rng, n), _strings(rng, n, 0.2)
    v = rng.normal(size=n)
    S = lambda slot: P.slot_ref(0, slot, T.STRING)
    aggs = [P.agg_expr("count_star", 1, 1), P.agg_expr("sum", 1, 2, None, P.slot_ref(0, 4, T.DOUBLE)), P.agg_expr("min", 1, 3, None, S(2)), P.agg_expr("max", 1, 4, None, S(2))]
    plan = P.Plan(P.agg(P.where(P.scan(0), P.ge(S(2), P.str_lit("b"))), 1, [S(1)], aggs),
                  {0: [(1, T.STRING), (2, T.STRING), (4, T.DOUBLE)], 1: [(1, T.INT64), (2, T.DOUBLE), (3, T.STRING), (4, T.STRING)]})
    enc = D.encode_strings(plan, [D.StringColumn(0, 1, s1), D.StringColumn(0, 2, s2)])
    got, _, _ = run_both(enc.plan, enc.columns + [make_column(0, 4, T.DOUBLE, v)], keys=["0_1"])
    dec = {c.name: (c.values if isinstance(c, D.StringColumn) else c.to_list()) for c in enc.decode(got)}
    tbl = pa.table({"s1": pa.array(s1, pa.binary()), "s2": pa.array(s2, pa.binary()), "v": v})
    want = tbl.filter(pc.fill_null(pc.greater_equal(tbl["s2"], b"b"), False)).group_by("s1", use_threads=False).aggregate([([], "count_all"), ("s2", "min"), ("s2", "max")]).to_pydict()
    w = {k: (c, mn, mx) for k, c, mn, mx in zip(want["s1"], want["count_all"], want["s2_min"], want["s2_max"])}
    g = {k: (c, mn, mx) for k, c, mn, mx in zip(dec["0_1"], dec["1_1"], dec["1_3"], dec["1_4"])}
    assert g == w
