"""Helpers shared by the parity tests: run a fragment on the GPU path and on the oracle and compare."""
import math
from typing import Dict, List, Optional, Sequence

import numpy as np

from baikaldb_b200.column import Column, rows_as_set
from baikaldb_b200.exec_node import execute
from oracle import oracle

# names bkgpu_stats.main_kernel_name reports for the lean GROUP BY path (_fx: double sums as fixed-point limbs; wp: opt-in warp-private tables)
LEAN_KERNELS = ("k_agg_group_lean", "k_agg_group_lean_fx", "k_agg_group_wp")
REL_TOL = 1e-6  # north_star: SUM/AVG(double) within 1e-6 relative; everything integer bit-exact


def assert_same_rows(got: Sequence[Column], want: Sequence[Column], keys: Optional[List[str]], rel=REL_TOL, abs_tol=0.0):
    gnames, wnames = [c.name for c in got], [c.name for c in want]
    assert sorted(gnames) == sorted(wnames), (gnames, wnames)
    if keys is None:  # ordered comparison
        assert len(got[0]) == len(want[0]) if got else True
        for nm in gnames:
            a, b = got[gnames.index(nm)].to_list(), want[wnames.index(nm)].to_list()
            _cmp_lists(nm, a, b, rel, abs_tol)
        return
    g, w = rows_as_set(list(got), keys), rows_as_set(list(want), keys)
    assert set(g) == set(w), f"group keys differ: only gpu {sorted(set(g) - set(w))[:5]}, only oracle {sorted(set(w) - set(g))[:5]}"
    for k in g:
        for nm in gnames:
            _cmp_val(f"{k}.{nm}", g[k][gnames.index(nm)], w[k][wnames.index(nm)], rel, abs_tol)


def _cmp_lists(nm, a, b, rel, abs_tol):
    assert len(a) == len(b), (nm, len(a), len(b))
    for i, (x, y) in enumerate(zip(a, b)):
        _cmp_val(f"{nm}[{i}]", x, y, rel, abs_tol)


def _cmp_val(where, x, y, rel, abs_tol):
    if x is None or y is None:
        assert x is None and y is None, (where, x, y)
    elif isinstance(x, bytes):  # AVG intermediate {double sum; int64 count}
        xs, xc = np.frombuffer(x, dtype=np.float64, count=1)[0], np.frombuffer(x, dtype=np.int64, count=1, offset=8)[0]
        ys, yc = np.frombuffer(y, dtype=np.float64, count=1)[0], np.frombuffer(y, dtype=np.int64, count=1, offset=8)[0]
        assert xc == yc, (where, xc, yc)
        assert math.isclose(xs, ys, rel_tol=rel, abs_tol=max(abs_tol, 1e-9)), (where, xs, ys)
    elif isinstance(x, float):
        if math.isnan(x) or math.isnan(y):
            assert math.isnan(x) and math.isnan(y), (where, x, y)
        else:
            assert math.isclose(x, y, rel_tol=rel, abs_tol=abs_tol), (where, x, y)
    else:
        assert x == y, (where, x, y)


def run_both(plan, cols, keys, options: Optional[Dict[str, int]] = None, rel=REL_TOL, abs_tol=0.0, batches=None, check_scanned=True):
    """Run `plan` over `cols` on cuda:0 and on the oracle; assert identical results. Returns (gpu_cols, stats)."""
    want = oracle.execute(plan.serialize(), cols)
    got, stats = execute(plan, batches if batches is not None else cols, device=0, options=options)
    if not want.columns:
        assert not got or len(got[0]) == 0
    else:
        assert_same_rows(got, want.columns, keys, rel, abs_tol)
    if check_scanned:  # (a LIMIT lets the pull-based row engine stop scanning early; a pushed batch is always scanned whole)
        assert stats.rows_scanned == want.rows_scanned, (stats.rows_scanned, want.rows_scanned)
    return got, stats, want
