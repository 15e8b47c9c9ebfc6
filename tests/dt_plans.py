"""Deterministic random fragments over DATE / DATETIME / TIMESTAMP / TIME columns (the f4 date/time slice): comparisons with text,
numeric and typed literals, comparisons between columns of different date/time types (calendar conversions on the device), IN lists,
+/- on the images, date/time GROUP BY keys and MIN / MAX arguments."""
import numpy as np

from baikaldb_b200 import plan as P
from baikaldb_b200.column import make_column
from baikaldb_b200.plan import PrimitiveType as T

TUPLE0 = [(1, T.DATETIME), (2, T.TIMESTAMP), (3, T.DATE), (4, T.TIME), (5, T.INT64), (6, T.INT32)]
DT_COLS = [(1, T.DATETIME), (2, T.TIMESTAMP), (3, T.DATE)]


def dt_image(y, mo, d, h=0, mi=0, s=0, us=0):
    return ((y * 13 + mo) << 46) | (d << 41) | (h << 36) | (mi << 30) | (s << 24) | us


def table(n, seed):
    rng = np.random.default_rng(seed)
    y, mo, d = rng.integers(2016, 2026, n), rng.integers(1, 13, n), rng.integers(1, 29, n)
    h, mi, s = rng.integers(0, 24, n), rng.integers(0, 60, n), rng.integers(0, 60, n)
    us = np.where(rng.random(n) < 0.5, 0, rng.integers(0, 1_000_000, n))
    dt = ((y * 13 + mo).astype(np.uint64) << np.uint64(46)) | (d.astype(np.uint64) << np.uint64(41)) | (h.astype(np.uint64) << np.uint64(36)) | \
         (mi.astype(np.uint64) << np.uint64(30)) | (s.astype(np.uint64) << np.uint64(24)) | us.astype(np.uint64)
    dt[rng.random(n) < 0.02] = 0                                        # the zero date
    ts = rng.integers(1_450_000_000, 1_770_000_000, n).astype(np.uint32)    # 2015-12 .. 2026-02
    ts[rng.random(n) < 0.02] = 0
    y2, mo2, d2 = rng.integers(2016, 2026, n), rng.integers(1, 13, n), rng.integers(1, 29, n)
    date = (((y2 * 13 + mo2) << 5) | d2).astype(np.uint32)
    tm = ((rng.integers(0, 30, n) << 12) | (rng.integers(0, 60, n) << 6) | rng.integers(0, 60, n)).astype(np.int32)
    tm = np.where(rng.random(n) < 0.2, -tm, tm).astype(np.int32)
    return [make_column(0, 1, T.DATETIME, dt, rng.random(n) > 0.1), make_column(0, 2, T.TIMESTAMP, ts, rng.random(n) > 0.1),
            make_column(0, 3, T.DATE, date, rng.random(n) > 0.1), make_column(0, 4, T.TIME, tm, rng.random(n) > 0.1),
            make_column(0, 5, T.INT64, rng.integers(-100_000, 100_000, n), rng.random(n) > 0.1), make_column(0, 6, T.INT32, rng.integers(0, 8, n))]


class Gen:
    def __init__(self, seed):
        self.r = np.random.default_rng(seed)

    def pick(self, xs):
        return xs[int(self.r.integers(0, len(xs)))]

    def ymd(self):
        return int(self.r.integers(2015, 2027)), int(self.r.integers(1, 13)), int(self.r.integers(1, 29))

    def text(self):
        y, mo, d = self.ymd()
        h, mi, s = int(self.r.integers(0, 24)), int(self.r.integers(0, 60)), int(self.r.integers(0, 60))
        return self.pick([f"{y:04d}-{mo:02d}-{d:02d}", f"{y:04d}-{mo:02d}-{d:02d} {h:02d}:{mi:02d}:{s:02d}", f"{y:04d}{mo:02d}{d:02d}", f"{y % 100:02d}{mo:02d}{d:02d}",
                          f"{y:04d}/{mo}/{d} {h}:{mi}", f"{y:04d}-{mo:02d}-{d:02d} {h:02d}:{mi:02d}:{s:02d}.5", f"{y:04d}{mo:02d}{d:02d}{h:02d}{mi:02d}{s:02d}", "not a date", f"{y:04d}-13-{d:02d}"])

    def literal_for(self, prim):
        """a literal to hold against a column of date/time type `prim`"""
        k = int(self.r.integers(0, 6))
        y, mo, d = self.ymd()
        if prim == T.TIME:
            return self.pick([P.str_lit(self.pick(["12:30:00", "-03:10:59", "1 02:03:04", "123456", "2024-05-06 07:08:09", "29:59:59"])), P.int_lit(int(self.r.integers(0, 240000))),
                              P.time_lit((int(self.r.integers(0, 30)) << 12) | (int(self.r.integers(0, 60)) << 6) | 7), P.null_lit()])
        if k == 0: return P.str_lit(self.text())
        if k == 1: return P.int_lit(y * 10000 + mo * 100 + d)                 # reads as text: 20240131 -> 2024-01-31
        if k == 2: return P.datetime_lit(dt_image(y, mo, d, int(self.r.integers(0, 24)), int(self.r.integers(0, 60)), 0))
        if k == 3: return P.date_lit(((y * 13 + mo) << 5) | d)
        if k == 4: return P.timestamp_lit(int(self.r.integers(1_450_000_000, 1_770_000_000)))
        return P.str_lit(self.text()) if self.r.random() < 0.8 else P.null_lit()

    def cmp(self):
        return self.pick([P.lt, P.le, P.gt, P.ge, P.eq, P.ne])

    def pred(self, d):
        if d <= 0 or self.r.random() < 0.45:
            k = int(self.r.integers(0, 8))
            s, t = self.pick(DT_COLS)
            c = P.slot_ref(0, s, t)
            if k == 0: return self.cmp()(c, self.literal_for(t))
            if k == 1:                                                          # two columns of different date/time types
                s2, t2 = self.pick(DT_COLS)
                return self.cmp()(c, P.slot_ref(0, s2, t2))
            if k == 2: return P.in_(c, *[self.literal_for(t) for _ in range(3)])
            if k == 3: return self.cmp()(P.slot_ref(0, 4, T.TIME), self.literal_for(T.TIME))
            if k == 4:                                                          # +/- on the image: ts + 3600 > literal seconds, date - date > n
                if self.r.random() < 0.5:
                    return self.cmp()(P.add(P.slot_ref(0, 2, T.TIMESTAMP), P.int_lit(int(self.r.integers(-86400, 86400)))), P.int_lit(int(self.r.integers(1_450_000_000, 1_770_000_000))))
                return self.cmp()(P.minus(c, P.slot_ref(0, s, t)), P.slot_ref(0, 5, T.INT64))
            if k == 5: return self.cmp()(c, P.slot_ref(0, 5, T.INT64))           # an integer column read as the image
            if k == 6: return P.is_null(c)
            return self.cmp()(self.literal_for(t), c)                            # literal first
        k = int(self.r.integers(0, 3))
        if k == 0: return P.and_(self.pred(d - 1), self.pred(d - 1))
        if k == 1: return P.or_(self.pred(d - 1), self.pred(d - 1))
        return P.not_(self.pred(d - 1))


def fragment(seed):
    """[WHERE p] GROUP BY k: COUNT(*), MIN(dt), MAX(dt), COUNT(dt), SUM(ts - n), MIN(time)"""
    g = Gen(1000 + seed)
    s, t = g.pick(DT_COLS)
    keys = [P.slot_ref(0, 6, T.INT32)] if seed % 3 else [P.slot_ref(0, 3, T.DATE)]
    if seed % 7 == 0:
        keys = []
    aggs = [P.agg_expr("count_star", 1, 1), P.agg_expr("min", 1, 2, None, P.slot_ref(0, s, t)), P.agg_expr("max", 1, 3, None, P.slot_ref(0, s, t)),
            P.agg_expr("count", 1, 4, None, P.slot_ref(0, 1, T.DATETIME)), P.agg_expr("sum", 1, 5, None, P.minus(P.slot_ref(0, 2, T.TIMESTAMP), P.int_lit(1_400_000_000))),
            P.agg_expr("min", 1, 6, None, P.slot_ref(0, 4, T.TIME))]
    child = P.where(P.scan(0), g.pred(2)) if seed % 5 else P.scan(0)
    root = P.agg(child, 1, keys, aggs)
    plan = P.Plan(root, {0: TUPLE0, 1: [(1, T.INT64), (2, t), (3, t), (4, T.INT64), (5, T.INT64), (6, T.TIME)]})
    return plan, [f"0_{k.slot_id}" for k in keys]
