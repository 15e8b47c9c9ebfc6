"""Date/time expressions on the device against the oracle: text / numeric / typed literals folded on the host, calendar conversions
between DATE, DATETIME and TIMESTAMP columns in the bytecode interpreter, IN lists, +/- on the images; and the plans the slice
refuses (TIME as a date, STRING operands) fail when compiled instead of answering differently."""
import numpy as np
import pytest

from baikaldb_b200 import _lib
from baikaldb_b200 import plan as P
from baikaldb_b200.plan import PrimitiveType as T
from oracle import oracle
from tests.dt_plans import TUPLE0, dt_image, fragment, table

SEEDS = list(range(40))


@pytest.mark.parametrize("seed", SEEDS)
def test_datetime_fragment_lowers_and_oracle_runs(seed):
    plan, _ = fragment(seed)
    assert _lib.explain(plan.serialize()).startswith("kind=")
    res = oracle.execute(plan.serialize(), table(400, seed))
    assert res.columns is not None


def _count_where(pred):
    aggs = [P.agg_expr("count_star", 1, 1)]
    return P.Plan(P.agg(P.where(P.scan(0), pred), 1, [], aggs), {0: TUPLE0, 1: [(1, T.INT64)]})


def test_oracle_reads_literals_the_way_the_reference_does():
    """hand-computed answers on a 4-row table: text and numeric literals against DATE / DATETIME / TIMESTAMP, a TIMESTAMP column against
    a DATETIME one in the fixed UTC+8 zone"""
    from baikaldb_b200.column import make_column
    dt = np.array([dt_image(2024, 1, 31, 8, 0, 0), dt_image(2024, 2, 1), dt_image(1970, 1, 1, 8, 0, 1), 0], dtype=np.uint64)
    ts = np.array([1706659200, 1706745600, 1, 0], dtype=np.uint32)          # 2024-01-31 08:00:00 +08, 2024-02-01 08:00:00 +08, epoch + 1 s
    date = np.array([(2024 * 13 + 1) << 5 | 31, (2024 * 13 + 2) << 5 | 1, (1970 * 13 + 1) << 5 | 1, 0], dtype=np.uint32)
    cols = [make_column(0, 1, T.DATETIME, dt), make_column(0, 2, T.TIMESTAMP, ts), make_column(0, 3, T.DATE, date), make_column(0, 4, T.TIME, np.zeros(4, np.int32)),
            make_column(0, 5, T.INT64, np.zeros(4, np.int64)), make_column(0, 6, T.INT32, np.zeros(4, np.int32))]
    def count(pred):
        res = oracle.execute(_count_where(pred).serialize(), cols)
        return res.columns[0].to_list()[0]
    assert count(P.ge(P.slot_ref(0, 1, T.DATETIME), P.str_lit("2024-01-31 08:00:00"))) == 2
    assert count(P.gt(P.slot_ref(0, 1, T.DATETIME), P.int_lit(20240131))) == 2        # the number reads as text: 2024-01-31 00:00:00
    assert count(P.eq(P.slot_ref(0, 3, T.DATE), P.str_lit("2024-02-01 23:59:59"))) == 1   # cast to DATE drops the time of day
    assert count(P.eq(P.slot_ref(0, 2, T.TIMESTAMP), P.str_lit("2024-01-31 08:00:00"))) == 1
    assert count(P.eq(P.slot_ref(0, 2, T.TIMESTAMP), P.slot_ref(0, 1, T.DATETIME))) == 3  # rows 0, 2 (08:00:01 +08 = 1 s) and the zero pair
    assert count(P.eq(P.slot_ref(0, 3, T.DATE), P.slot_ref(0, 1, T.DATETIME))) == 2       # DATE -> DATETIME is midnight: rows 1 and 3
    assert count(P.in_(P.slot_ref(0, 3, T.DATE), P.str_lit("2024-01-31"), P.str_lit("1970-01-01"))) == 2
    assert count(P.gt(P.add(P.slot_ref(0, 2, T.TIMESTAMP), P.int_lit(86400)), P.int_lit(1706745599))) == 2


@pytest.mark.parametrize("pred, why", [
    (P.eq(P.slot_ref(0, 4, T.TIME), P.slot_ref(0, 1, T.DATETIME)), "current date"),
    (P.eq(P.slot_ref(0, 3, T.DATE), P.time_lit(77)), "current date"),
    (P.gt(P.slot_ref(0, 1, T.DATETIME), P.double_lit(20240131.0)), "DOUBLE literal"),
    (P.gt(P.slot_ref(0, 5, T.INT64), P.str_lit("12")), "STRING"),
    (P.gt(P.add(P.slot_ref(0, 5, T.INT64), P.str_lit("12")), P.int_lit(1)), "STRING"),
])
def test_plans_outside_the_slice_are_refused_when_compiled(pred, why):
    with pytest.raises(_lib.BkgpuError) as e:
        _lib.explain(_count_where(pred).serialize())
    assert e.value.code == _lib.EUNSUPPORTED and why in str(e.value)


def test_same_type_date_comparison_takes_the_direct_kernels():
    text = _lib.explain(_count_where(P.ge(P.slot_ref(0, 3, T.DATE), P.str_lit("2024-01-31"))).serialize())
    assert "direct" in text and f"{(2024 * 13 + 1) << 5 | 31:x}" in text.lower()


@pytest.mark.gpu
@pytest.mark.parametrize("seed", SEEDS)
def test_datetime_fragment_gpu_matches_oracle(seed):
    from tests.util import run_both
    plan, keys = fragment(seed)
    run_both(plan, table(5000 + 41 * seed, seed), keys=keys)


@pytest.mark.gpu
def test_calendar_conversions_on_the_device_cover_every_day():
    """TIMESTAMP -> DATETIME -> DATE on the device for one timestamp per day from 1970 to 2106 (both sides of midnight +08): the counts
    per converted DATE must equal the oracle's"""
    from baikaldb_b200.column import make_column
    from tests.util import run_both
    days = np.arange(0, (1 << 32) // 86400, dtype=np.int64)
    ts = np.concatenate([days * 86400 + 57599, days * 86400 + 57600]).astype(np.uint32)      # 23:59:59 / 00:00:00 at +08
    n = len(ts)
    cols = [make_column(0, 1, T.DATETIME, np.zeros(n, np.uint64)), make_column(0, 2, T.TIMESTAMP, ts), make_column(0, 3, T.DATE, np.zeros(n, np.uint32)),
            make_column(0, 4, T.TIME, np.zeros(n, np.int32)), make_column(0, 5, T.INT64, np.zeros(n, np.int64)), make_column(0, 6, T.INT32, np.zeros(n, np.int32))]
    # GROUP BY (ts compared as DATE with its own DATETIME image) is not expressible; instead: count rows whose DATE image (ts -> DATE, via the
    # comparison's argument type) is >= each of a set of literals, and the MIN / MAX of ts among them
    for text in ("1970-01-02", "1999-12-31", "2000-03-01", "2038-01-19", "2100-03-01", "2106-02-07"):
        aggs = [P.agg_expr("count_star", 1, 1), P.agg_expr("min", 1, 2, None, P.slot_ref(0, 2, T.TIMESTAMP)), P.agg_expr("max", 1, 3, None, P.slot_ref(0, 2, T.TIMESTAMP))]
        pred = P.ge(P.slot_ref(0, 2, T.TIMESTAMP), P.date_lit(int(_lib_parse(text, T.DATE))))     # types {TIMESTAMP, DATE} -> TIMESTAMP: literal converted
        pred2 = P.le(P.slot_ref(0, 1, T.DATETIME), P.slot_ref(0, 2, T.TIMESTAMP))                  # {DATETIME, TIMESTAMP} -> DATETIME: column converted
        pl = P.Plan(P.agg(P.where(P.scan(0), P.and_(pred, pred2)), 1, [], aggs), {0: TUPLE0, 1: [(1, T.INT64), (2, T.TIMESTAMP), (3, T.TIMESTAMP)]})
        run_both(pl, cols, keys=[])
    # every row through ts -> DATETIME, compared with a DATETIME column holding the oracle's conversion: all equal, none different
    import ctypes
    from tests.test_datetime import _O
    want = np.array([_O.bk_oracle_cast_image(int(t), int(T.TIMESTAMP), int(T.DATETIME)) for t in ts], dtype=np.uint64)
    cols[0] = make_column(0, 1, T.DATETIME, want)
    got, _, _ = run_both(_count_where(P.eq(P.slot_ref(0, 1, T.DATETIME), P.slot_ref(0, 2, T.TIMESTAMP))), cols, keys=[])
    assert got[0].to_list()[0] == n
    got, _, _ = run_both(_count_where(P.eq(P.slot_ref(0, 2, T.TIMESTAMP), P.slot_ref(0, 1, T.DATETIME))), cols, keys=[])
    assert got[0].to_list()[0] == n


def _lib_parse(text, prim):
    import ctypes
    out = ctypes.c_uint64()
    raw = text.encode()
    assert _lib.lib().bkgpu_parse_datetime(raw, len(raw), int(prim), ctypes.byref(out)) == 0
    return out.value
