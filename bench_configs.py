#!/usr/bin/env python
"""bench_configs.py — the OTHER BASELINE.json configs on one B200 (C1 count-where, C3 join+aggregate, C5 top-k):
device-resident synthetic columns, CUDA events around reset+push+finish+get_next, one JSON line per
config with rows/s and achieved GB/s on the config's algorithmic bytes.  The driver's headline line is bench.py's;
this script feeds DESIGN.md / profiles/.   usage: python bench_configs.py [c1] [c3] [c5] [--scale 0.1] [--steps 10]"""
import argparse
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch

from baikaldb_b200 import _lib, datagen, queries
from baikaldb_b200._lib import BkgpuColumn, BkgpuStats
from baikaldb_b200.plan import PrimitiveType as T

L = _lib.lib()
PEAK = 6486.1
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def gen(rows, spec, seed, row0=0):
    """spec: (tuple, slot, prim, dist, column_id, lo, hi, scale) -> (tensor, BkgpuColumn fields)"""
    t, slot, pt, dist, cid, lo, hi, scale = spec
    dt = {T.INT32: torch.int32, T.INT64: torch.int64, T.DOUBLE: torch.float64}[T(pt)]
    x = torch.empty(rows, dtype=dt, device="cuda")
    _lib.check(L.bkgpu_gen_column(0, x.data_ptr(), int(pt), dist, seed, cid, row0, rows, lo, hi, scale))
    return x


def cols_array(specs, tensors):
    arr = (BkgpuColumn * len(specs))()
    for i, (sp, x) in enumerate(zip(specs, tensors)):
        arr[i].tuple_id, arr[i].slot_id, arr[i].prim_type, arr[i].elem_size = sp[0], sp[1], int(sp[2]), 0
        arr[i].values, arr[i].validity, arr[i].length = x.data_ptr(), None, x.numel()
    return arr


def run(name, plan, pushes, rows, algo_bytes, steps, warmup, options=()):
    pb = plan.serialize()
    h = ctypes.c_void_p()
    _lib.check(L.bkgpu_init(ctypes.byref(h), pb, len(pb), 0, None))
    stream = torch.cuda.current_stream()
    _lib.check(L.bkgpu_set_option(h, b"stream", stream.cuda_stream), h)
    for k, v in options:
        _lib.check(L.bkgpu_set_option(h, k, v), h)
    for kv in filter(None, os.environ.get("BKGPU_BENCH_OPTS", "").split(",")):   # A/B of kernel variants
        k, v = kv.split("=")
        _lib.check(L.bkgpu_set_option(h, k.encode(), int(v)), h)
    _lib.check(L.bkgpu_open(h), h)
    out = (BkgpuColumn * 16)()

    def step():
        _lib.check(L.bkgpu_reset(h), h)
        for arr, n, r in pushes:
            _lib.check(L.bkgpu_push(h, arr, n, r, 1), h)
        _lib.check(L.bkgpu_finish(h), h)
        eos, total = ctypes.c_int(0), 0
        while not eos.value:
            ncols, nrows = ctypes.c_int(16), ctypes.c_int64(0)
            _lib.check(L.bkgpu_get_next(h, out, ctypes.byref(ncols), ctypes.byref(nrows), ctypes.byref(eos)), h)
            total += nrows.value
        return total
    for _ in range(warmup):
        n_out = step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kms, kl = 0.0, 0
    e0.record(stream)
    for _ in range(steps):
        n_out = step()
        st = BkgpuStats(); L.bkgpu_get_stats(h, ctypes.byref(st))
        kms += st.main_kernel_ms; kl += st.main_kernel_launches
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    line = {"config": name, "rows": rows, "ms_per_step": ms, "rows_per_s": rows / (ms / 1e3), "algorithmic_bytes": algo_bytes,
            "gbs_whole_step": algo_bytes / (ms / 1e3) / 1e9, "frac_of_measured_hbm_whole_step": algo_bytes / (ms / 1e3) / 1e9 / PEAK,
            "main_kernel": st.main_kernel_name.decode(), "main_kernel_ms_per_step": kms / steps, "result_rows": n_out,
            "gpu_launches_per_step": None}
    print(json.dumps(line), flush=True)
    L.bkgpu_close(h)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("configs", nargs="*", default=["c1", "c3", "c5"])
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    a = ap.parse_args()
    torch.cuda.set_device(0)
    if "c1" in a.configs:
        n = int(100_000_000 * a.scale)   # C1's query at a bandwidth-relevant size (the 1M-row case is a plumbing test)
        specs = [(0, 1, T.INT32, 0, 1, 0, 1 << 20, 1.0)]
        ts = [gen(n, s, 1) for s in specs]
        run("C1 COUNT(*) WHERE int32 < k", queries.c1_count_where(), [(cols_array(specs, ts), 1, n)], n, 4 * n, a.steps, a.warmup)
        del ts
    if "c2n" in a.configs:   # C2 with NULLs: validity bitmaps on both value columns (~30 % NULL), device-resident
        n = int(100_000_000 * a.scale)
        specs = [(0, 1, T.INT32, 0, 1, 0, 1000, 1.0), (0, 2, T.INT32, 0, 2, 0, 1 << 20, 1.0), (0, 3, T.DOUBLE, 1, 3, 0, 0, 1.0), (0, 4, T.DOUBLE, 2, 4, 0, 0, 1732.05)]
        ts = [gen(n, s, 2) for s in specs]
        arr = cols_array(specs, ts)
        bm = [torch.randint(0, 256, ((n + 7) // 8 + 64,), dtype=torch.uint8, device="cuda") | torch.randint(0, 256, ((n + 7) // 8 + 64,), dtype=torch.uint8, device="cuda") for _ in range(2)]
        for i, b in zip((2, 3), bm):
            arr[i].validity = b.data_ptr()
        run("C2 with ~25 % NULLs in both value columns", queries.c2_filter_groupby(), [(arr, 4, n)], n, 24 * n + n // 4, a.steps, a.warmup, options=[(b"group_capacity_log2", 14)])
        del ts, bm
    if "c2m" in a.configs:   # C2's table, MIN / MAX / SUM aggregates: the lean kernel's MM instantiation vs the general kernel
        from baikaldb_b200 import plan as P
        n = int(100_000_000 * a.scale)
        specs = [(0, 1, T.INT32, 0, 1, 0, 1000, 1.0), (0, 2, T.INT32, 0, 2, 0, 1 << 20, 1.0), (0, 3, T.DOUBLE, 1, 3, 0, 0, 1.0), (0, 4, T.DOUBLE, 2, 4, 0, 0, 1732.05)]
        ts = [gen(n, s, 2) for s in specs]
        aggs = [P.agg_expr("count_star", 1, 1), P.agg_expr("min", 1, 2, None, P.slot_ref(0, 3, T.DOUBLE)), P.agg_expr("max", 1, 3, None, P.slot_ref(0, 3, T.DOUBLE)),
                P.agg_expr("sum", 1, 4, None, P.slot_ref(0, 4, T.DOUBLE))]
        root = P.agg(P.where(P.scan(0), P.lt(P.slot_ref(0, 2, T.INT32), P.int_lit(1 << 19))), 1, [P.slot_ref(0, 1, T.INT32)], aggs)
        plan = P.Plan(root, {0: [(1, T.INT32), (2, T.INT32), (3, T.DOUBLE), (4, T.DOUBLE)], 1: P.agg_tuple_slots(aggs, [T.INT64, T.DOUBLE, T.DOUBLE, T.DOUBLE])})
        run("C2 table, COUNT(*), MIN(a), MAX(a), SUM(b)", plan, [(cols_array(specs, ts), 4, n)], n, 24 * n, a.steps, a.warmup, options=[(b"group_capacity_log2", 14)])
        del ts
    if "c5" in a.configs:
        n = int(125_000_000 * a.scale)
        specs = [(0, 1, T.INT64, 3, 1, 0, 0, 1.0), (0, 2, T.INT32, 0, 2, 0, 1 << 30, 1.0)]
        ts = [gen(n, s, 5) for s in specs]
        run("C5 ORDER BY int64 LIMIT 1000 (one region of 125M rows)", queries.c5_topk(1000), [(cols_array(specs, ts), 2, n)], n, 8 * n, a.steps, a.warmup)
        del ts
    if "c5full" in a.configs:   # ORDER BY without LIMIT: the device radix sort (Sorter, src/runtime/sorter.cpp:54-114) — 8-bit passes x 24 B/row
        n = int(125_000_000 * a.scale)
        specs = [(0, 1, T.INT64, 3, 1, 0, 0, 1.0), (0, 2, T.INT32, 0, 2, 0, 1 << 30, 1.0)]
        ts = [gen(n, s, 5) for s in specs]
        from baikaldb_b200 import plan as P
        plan = P.Plan(P.sort(P.scan(0), [P.slot_ref(0, 1, T.INT64)], [True], tuple_id=0), {0: [(1, T.INT64), (2, T.INT32)]})
        run("C5full ORDER BY int64 (no LIMIT), 125M rows: stable LSD radix sort of (key, row id) pairs; main_kernel_ms = the sort on the device, "
            "the step also gathers the payload and copies 1.5 GB of sorted rows to the host", plan, [(cols_array(specs, ts), 2, n)], n,
            (8 + 8 * 24) * n, max(1, a.steps // 5), 1)
        del ts
    if "c3" in a.configs:
        nf, nd = int(100_000_000 * a.scale), int(10_000_000 * a.scale)
        fs = [(0, 1, T.INT32, 0, 1, 0, nd, 1.0), (0, 2, T.DOUBLE, 1, 2, 0, 0, 1.0)]
        ds = [(1, 1, T.INT32, 4, 11, 0, nd, 1.0), (1, 2, T.INT32, 0, 12, 0, 1000, 1.0)]
        ft, dt = [gen(nf, s, 3) for s in fs], [gen(nd, s, 3) for s in ds]
        run("C3 fact JOIN dim ON int32 key, GROUP BY dim attr (100M x 10M)", queries.c3_join_groupby(),
            [(cols_array(ds, dt), 2, nd), (cols_array(fs, ft), 2, nf)], nf, 12 * nf + 8 * nd, max(1, a.steps // 3), 1)


if __name__ == "__main__":
    main()
