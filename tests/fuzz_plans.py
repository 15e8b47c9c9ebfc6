"""Deterministic random fragments for the fuzz tests: expressions over a 5-column nullable table drawn from the operator surface the
GPU path claims (arithmetic, comparisons, three-valued logic, IN, IS NULL, IF / IFNULL / CASE WHEN, ABS / FLOOR / CEIL / ROUND, SQRT / LN / POW /
MOD / SIGN / GREATEST / LEAST, casts),
used as filters, GROUP BY keys and aggregate arguments."""
import numpy as np

from baikaldb_b200 import plan as P
from baikaldb_b200.column import make_column
from baikaldb_b200.plan import PrimitiveType as T

TUPLE0 = [(1, T.INT32), (2, T.INT64), (3, T.DOUBLE), (4, T.UINT32), (5, T.INT32)]


def table(n, seed):
    rng = np.random.default_rng(seed)
    return [make_column(0, 1, T.INT32, rng.integers(0, 12, n), rng.random(n) > 0.1),
            make_column(0, 2, T.INT64, rng.integers(-50, 50, n), rng.random(n) > 0.2),
            make_column(0, 3, T.DOUBLE, np.round(rng.random(n) * 100, 2), rng.random(n) > 0.2),
            make_column(0, 4, T.UINT32, rng.integers(0, 1000, n), rng.random(n) > 0.1),
            make_column(0, 5, T.INT32, rng.integers(-5, 100, n))]


class Gen:
    def __init__(self, seed):
        self.r = np.random.default_rng(seed)

    def pick(self, xs):
        return xs[int(self.r.integers(0, len(xs)))]

    def col(self):
        s, t = self.pick(TUPLE0)
        return P.slot_ref(0, s, t)

    def lit(self):
        k = int(self.r.integers(0, 4))
        if k == 0:
            return P.int_lit(int(self.r.integers(-20, 60)))
        if k == 1:
            return P.double_lit(float(np.round(self.r.random() * 50, 1)))
        if k == 2:
            return P.int_lit(int(self.r.integers(1, 5)))
        return P.null_lit() if self.r.random() < 0.3 else P.int_lit(0)

    def num(self, d):
        """numeric-valued expression"""
        if d <= 0 or self.r.random() < 0.25:
            return self.col() if self.r.random() < 0.7 else self.lit()
        k = int(self.r.integers(0, 17))
        a, b = self.num(d - 1), self.num(d - 1)
        if k == 0: return P.add(a, b)
        if k == 1: return P.minus(a, b)
        if k == 2: return P.multiplies(a, self.col() if self.r.random() < 0.5 else P.int_lit(int(self.r.integers(-3, 4))))
        if k == 3: return P.divides(a, b)
        if k == 4: return P.mod(P.cast_to_signed(a), P.int_lit(int(self.r.integers(1, 7))))
        if k == 5: return P.if_(self.pred(d - 1), a, b)
        if k == 6: return P.ifnull(a, b)
        if k == 7: return P.case_when(self.pred(d - 1), a, self.pred(d - 1), b) if self.r.random() < 0.5 else P.case_when(self.pred(d - 1), a, b)
        if k == 8: return self.pick([P.abs_, P.floor_, P.ceil_, P.round_])(a)
        if k == 9: return P.round_(a, P.int_lit(int(self.r.integers(0, 3))))
        if k == 10: return self.pick([P.cast_to_signed, P.cast_to_double])(a)
        if k == 12: return P.ifnull(self.pick([P.sqrt_, P.ln_])(a), b)
        if k == 13: return self.pick([P.greatest, P.least])(a, b, self.lit())
        if k == 14: return P.sign_(a)
        if k == 15: return P.fmod_(a, self.pick([P.int_lit(3), P.double_lit(2.5), self.col()]))
        if k == 16: return P.pow_(P.divides(a, P.int_lit(50)), P.int_lit(int(self.r.integers(0, 4))))
        return P.uminus(a)

    def pred(self, d):
        if d <= 0 or self.r.random() < 0.3:
            c = self.col()
            k = int(self.r.integers(0, 4))
            if k == 0: return self.pick([P.lt, P.le, P.gt, P.ge, P.eq, P.ne])(c, self.lit())
            if k == 1: return P.is_null(c)
            if k == 2: return P.in_(c, *[P.int_lit(int(v)) for v in self.r.integers(0, 40, 3)])
            return self.pick([P.lt, P.gt, P.ne])(c, self.col())
        k = int(self.r.integers(0, 6))
        if k == 0: return P.and_(self.pred(d - 1), self.pred(d - 1))
        if k == 1: return P.or_(self.pred(d - 1), self.pred(d - 1))
        if k == 2: return P.not_(self.pred(d - 1))
        if k == 3: return self.pick([P.lt, P.ge, P.eq])(self.num(d - 1), self.num(d - 1))
        if k == 4: return P.is_true(self.pred(d - 1))
        return P.is_null(self.num(d - 1))


def fragment(seed):
    """-> (plan, key column names) : [WHERE p] GROUP BY k [, k2]  COUNT(*), SUM(e1), MIN(e2), MAX(e2), AVG(e1), COUNT(e2)"""
    g = Gen(seed)
    e1, e2 = g.num(2), g.num(2)
    keys = [P.slot_ref(0, 1, T.INT32)] if seed % 3 else [P.slot_ref(0, 1, T.INT32), P.slot_ref(0, 5, T.INT32)]
    if seed % 5 == 0:
        keys = [P.cast_to_signed(P.floor_(P.divides(P.slot_ref(0, 3, T.DOUBLE), P.int_lit(25))))]
    aggs = [P.agg_expr("count_star", 1, 1), P.agg_expr("sum", 1, 2, None, e1), P.agg_expr("min", 1, 3, None, e2), P.agg_expr("max", 1, 4, None, e2),
            P.agg_expr("avg", 1, 5, 6, e1), P.agg_expr("count", 1, 7, None, e2)]
    child = P.where(P.scan(0), g.pred(2)) if seed % 4 else P.scan(0)
    root = P.agg(child, 1, keys, aggs)
    # the aggregate tuple is declared with the types the (host-only) lowering infers for each aggregate: the planner would do the same
    import re
    from baikaldb_b200 import _lib
    text = _lib.explain(P.Plan(root, {0: TUPLE0, 1: []}).serialize())
    prims = [int(m) for m in re.findall(r"agg\[\d+\] .*out_prim=(\d+)", text)]
    assert len(prims) == len(aggs)
    slots = {}
    for f, pt in zip(aggs, prims):
        slots[f.final_slot_id] = pt
        if f.intermediate_slot_id != f.final_slot_id:
            slots[f.intermediate_slot_id] = int(T.STRING)
    plan = P.Plan(root, {0: TUPLE0, 1: sorted(slots.items())})
    names = [f"0_{k.slot_id}" if k.node_type == P.ExprNodeType.SLOT_REF else f"-1_{i}" for i, k in enumerate(keys)]
    return plan, names
