#!/usr/bin/env python
"""A/B of the lean kernel's FX variant (double sums as fixed-point limbs, csrc/agg_direct.cuh) against the CAS variant, one process:
C2's table at several group counts and selectivities, C3's fused join probe; every FX result is compared with the CAS result of the same
plan and with an independent torch computation.
usage: python scripts/r02_fx_ab.py [--rows 100000000] [--steps 10]   -> one JSON line per case on stdout"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from baikaldb_b200 import _lib, queries
from baikaldb_b200.exec_node import ColumnSource, DeviceColumn, GpuExecNode, RowBatch, RuntimeState
from baikaldb_b200.plan import PrimitiveType as T
import bench_configs as bc

PEAK = bc.PEAK


def fetch(node, st):
    cols, eos, rb = None, False, RowBatch()
    while not eos:
        rc, eos = node.get_next(st, rb)
        assert rc == 0, st.error_msg
        if cols is None:
            cols = {c.name: [c.values] for c in rb.columns}
        else:
            for c in rb.columns:
                cols[c.name].append(c.values)
    return {k: np.concatenate(v) for k, v in cols.items()}


def run(plan, batches, opts, steps, warmup=3):
    """batches: list of lists of DeviceColumn (one push each).  Returns (result columns, ms per step, kernel ms per step, kernel name)."""
    st = RuntimeState(device=0, options=dict(opts))
    node = GpuExecNode(); node.init(plan)
    node.add_child(ColumnSource(batches))
    assert node.open(st) == 0, st.error_msg
    res = fetch(node, st)
    for _ in range(warmup):
        node.reset()
        for b in batches:
            node.push(b)
        node.finish(); res = fetch(node, st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kms = 0.0
    e0.record()
    for _ in range(steps):
        node.reset()
        for b in batches:
            node.push(b)
        node.finish(); res = fetch(node, st)
        kms += node.stats().main_kernel_ms
    e1.record(); torch.cuda.synchronize()
    name = node.stats().main_kernel_name.decode()
    node.close()
    return res, e0.elapsed_time(e1) / steps, kms / steps, name


def by_key(res, key):
    o = np.argsort(res[key], kind="stable")
    return {k: v[o] for k, v in res.items()}


def rel_diff(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    d = np.abs(a - b) / np.maximum(np.abs(b), 1e-300)
    return float(d.max()) if d.size else 0.0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=100_000_000)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--quick", action="store_true", help="only C2 (1000 groups, 50 %) and C3: for A/B of library builds (BKGPU_LIB)")
    a = ap.parse_args()
    torch.cuda.set_device(0)
    n = a.rows
    for groups, k_filter, tag in [(1000, 1 << 19, "C2 1000 groups, 50 %"), (1000, (1 << 20) - (1 << 13), "C2 1000 groups, 99 %"), (1000, 1 << 13, "C2 1000 groups, 1 %"),
                                  (100, 1 << 19, "C2 100 groups, 50 %"), (8, 1 << 19, "C2 8 groups, 50 %")][:1 if a.quick else None]:
        specs = [(0, 1, T.INT32, 0, 1, 0, groups, 1.0), (0, 2, T.INT32, 0, 2, 0, 1 << 20, 1.0), (0, 3, T.DOUBLE, 1, 3, 0, 0, 1.0), (0, 4, T.DOUBLE, 2, 4, 0, 0, 1732.05)]
        ts = [bc.gen(n, s, 2) for s in specs]
        cols = [DeviceColumn(s[0], s[1], int(s[2]), t.data_ptr(), n, 0, t) for s, t in zip(specs, ts)]
        plan = queries.c2_filter_groupby(k_filter)
        out = {}
        for fx in (0, 1, 1):
            res, ms, kms, name = run(plan, [cols], {"lean_fx": fx}, a.steps)
            out.setdefault(fx, []).append((res, ms, kms, name))
        (r0, ms0, k0, n0), (r1, ms1, k1, n1), (r2, _, _, _) = out[0][0], out[1][0], out[1][1]
        keycol = [c for c in r0 if (r0[c].dtype.kind in "iu" and len(np.unique(r0[c])) == len(r0[c]) and r0[c].max() < groups)][0]
        r0, r1, r2 = by_key(r0, keycol), by_key(r1, keycol), by_key(r2, keycol)
        # independent reference: torch over the same device columns
        m = ts[1] < k_filter
        key = ts[0][m].long()
        cnt = torch.bincount(key, minlength=groups)
        s3 = torch.zeros(groups, dtype=torch.float64, device="cuda").index_add_(0, key, ts[2][m])
        s4 = torch.zeros(groups, dtype=torch.float64, device="cuda").index_add_(0, key, ts[3][m])
        ref = {"cnt": cnt.cpu().numpy(), "s3": s3.cpu().numpy(), "s4": s4.cpu().numpy()}
        checks = {"ints_equal_cas": True, "doubles_rel_vs_cas": 0.0, "fx_runs_close": 0.0, "vs_torch": {}}
        for c in r0:
            if r0[c].dtype.kind in "iu" and r0[c].ndim == 1:   # (the AVG blob column is 16 raw bytes per row: its double half is compared through the AVG)
                checks["ints_equal_cas"] &= bool(np.array_equal(r0[c], r1[c]))
            elif r0[c].dtype.kind == "f":
                checks["doubles_rel_vs_cas"] = max(checks["doubles_rel_vs_cas"], rel_diff(r1[c], r0[c]))
                checks["fx_runs_close"] = max(checks["fx_runs_close"], rel_diff(r1[c], r2[c]))   # (the CTAs' partial sums still meet in floating point in the global table)
        fcols = [c for c in r1 if r1[c].dtype.kind == "f"]
        icols = [c for c in r1 if r1[c].dtype.kind in "iu" and c != keycol]
        present = ref["cnt"] > 0
        checks["vs_torch"]["count_exact"] = bool(np.array_equal(r1[icols[0]], ref["cnt"][present])) if icols else None
        if len(fcols) >= 2:   # SUM(0_3), AVG(0_4)
            checks["vs_torch"]["sum_rel"] = rel_diff(r1[fcols[0]], ref["s3"][present])
            checks["vs_torch"]["avg_rel"] = rel_diff(r1[fcols[1]], (ref["s4"] / np.maximum(ref["cnt"], 1))[present])
        print(json.dumps({"lib": os.environ.get("BKGPU_LIB", "default"), "case": tag, "rows": n, "cas": {"kernel": n0, "kernel_ms": k0, "step_ms": ms0, "frac_hbm": 24 * n / (k0 / 1e3) / 1e9 / PEAK},
                          "fx": {"kernel": n1, "kernel_ms": k1, "step_ms": ms1, "frac_hbm": 24 * n / (k1 / 1e3) / 1e9 / PEAK}, "checks": checks}), flush=True)
        del ts, cols
    # ---- C3: fact JOIN dim, GROUP BY the dimension attribute (the probe is fused into the lean kernel) ----
    nf, nd = n, max(n // 10, 1000)
    fs = [(0, 1, T.INT32, 0, 1, 0, nd, 1.0), (0, 2, T.DOUBLE, 1, 2, 0, 0, 1.0)]
    ds = [(1, 1, T.INT32, 4, 11, 0, nd, 1.0), (1, 2, T.INT32, 0, 12, 0, 1000, 1.0)]
    try:
        ft = [bc.gen(nf, s, 3) for s in fs]; dt = [bc.gen(nd, s, 3) for s in ds]
        fc = [DeviceColumn(s[0], s[1], int(s[2]), t.data_ptr(), nf, 0, t) for s, t in zip(fs, ft)]
        dc = [DeviceColumn(s[0], s[1], int(s[2]), t.data_ptr(), nd, 0, t) for s, t in zip(ds, dt)]
        plan = queries.c3_join_groupby()
        o = {}
        for fx in (0, 1):
            res, ms, kms, name = run(plan, [dc, fc], {"lean_fx": fx}, a.steps)
            o[fx] = (res, ms, kms, name)
        r0, r1 = o[0][0], o[1][0]
        keycol = [c for c in r0 if r0[c].dtype.kind in "iu" and len(np.unique(r0[c])) == len(r0[c])][0]
        r0, r1 = by_key(r0, keycol), by_key(r1, keycol)
        ints = all(np.array_equal(r0[c], r1[c]) for c in r0 if r0[c].dtype.kind in "iu")
        dbl = max([rel_diff(r1[c], r0[c]) for c in r0 if r0[c].dtype.kind == "f"] + [0.0])
        algo = 12 * nf + 8 * nd
        print(json.dumps({"lib": os.environ.get("BKGPU_LIB", "default"), "case": "C3 join + GROUP BY", "rows": nf, "cas": {"kernel": o[0][3], "kernel_ms": o[0][2], "step_ms": o[0][1], "frac_hbm_step": algo / (o[0][1] / 1e3) / 1e9 / PEAK},
                          "fx": {"kernel": o[1][3], "kernel_ms": o[1][2], "step_ms": o[1][1], "frac_hbm_step": algo / (o[1][1] / 1e3) / 1e9 / PEAK},
                          "checks": {"ints_equal_cas": bool(ints), "doubles_rel_vs_cas": dbl}}), flush=True)
    except Exception as ex:   # (the C3 generator / plan helper may differ: the A/B above is the point of this script)
        print(json.dumps({"case": "C3 join + GROUP BY", "error": repr(ex)}), flush=True)


if __name__ == "__main__":
    main()
