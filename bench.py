#!/usr/bin/env python
"""bench.py — rows/s through scan + filter + GROUP BY (BASELINE.json's metric) on N B200s.

A step = one pass of the hot path over one batch of synthetic columns:
    SELECT `0_1`, COUNT(*), SUM(`0_3`), AVG(`0_4`) FROM t WHERE `0_2` < 2^19 GROUP BY `0_1`
N = 1 : config C2 — 100M rows, 4 columns (int32, int32, float64, float64), 1k groups, 24 B/row.
N > 1 : config C4 — one region of 125M rows per GPU (weak scaling; N = 8 is the 1e9-row case), partial
        tables merged by ONE ncclAllGather + merge kernel inside bkgpu_finish.

    value  : device-resident columns, timed with CUDA events on the launching stream, max over ranks
    e2e    : same query through the C ABI with HOST (pinned) columns: H2D inside the timed region
    roofline / cpu_baseline : see DESIGN.md "Measurement"
`--impl reference` times the reference's CPU engine (the Acero plan it builds, all host threads).
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "rows/sec scan+filter+GROUP BY"
K_FILTER = 1 << 19
N_GROUPS = 1000
BYTES_PER_ROW = 24


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--rows", type=int, default=0, help="rows per GPU (default 100M at N=1, 125M per region at N>1)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def env_rank():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def workload_name(n_gpus, rows):
    if n_gpus == 1:
        return f"C2: scan+filter+COUNT/SUM/AVG GROUP BY ({N_GROUPS} groups), {rows // 10**6}M rows x 4 cols (int32,int32,f64,f64), 1xB200"
    return (f"C4: same query, {n_gpus} regions x {rows // 10**6}M rows -> {n_gpus}xB200, partial aggregates merged by one "
            f"ncclAllGather + merge kernel")


# ----------------------------------------------------------------------------------------------
# clocks: sampled DURING the timed region (B200_PROFILING.md)
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region.  NVML in a background thread (one sample every ~2 ms: even a
    10 ms region gets several); `nvidia-smi -lms` as a subprocess when pynvml is not importable."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    REASONS = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap"))

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.p = None
        self.f = None
        self.nvml = None
        self.samples = []
        self.stop_flag = False
        self.thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nvml = None

    def _sample(self):
        n = self.nvml
        try:
            reasons = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        except Exception:
            reasons = 0
        self.samples.append((float(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM)), int(reasons)))

    def _loop(self):
        while not self.stop_flag:
            try:
                self._sample()
            except Exception:
                break
            time.sleep(0.002)

    def start(self):
        if self.nvml:
            import threading
            self.stop_flag = False
            self.thread = threading.Thread(target=self._loop, daemon=True)
            self.thread.start()
            return
        try:
            self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                       "-i", str(self.gpu)], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.nvml:
            try:
                self._sample()   # the region has just ended (the caller synchronised): still the loaded state
            except Exception:
                pass
            self.stop_flag = True
            if self.thread:
                self.thread.join(timeout=2)
            if self.samples:
                bits = 0
                for _, r in self.samples:
                    bits |= r
                out.update(sm_mhz=statistics.median([c for c, _ in self.samples]), sm_max_mhz=self.max_sm,
                           reasons=sorted(name for bit, name in self.REASONS if bits & bit), samples=len(self.samples), source="nvml")
            return out
        if not self.p:
            return out
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1])); mx.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.f.name)
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm), source="nvidia-smi")
        return out


# ----------------------------------------------------------------------------------------------
# reference arm: the CPU engine of the reference's vectorized path (Acero), all host threads
# ----------------------------------------------------------------------------------------------
def host_table_numpy(rows, row0=0):
    from baikaldb_b200 import datagen
    return datagen.c2_table(row0, rows, n_groups=N_GROUPS)


def host_table_via_gpu(rows, row0=0):
    """Same bits as datagen.c2_table, produced by the device generator and copied back (fast path for 1e8 rows)."""
    import numpy as np
    import torch
    from baikaldb_b200 import _lib, datagen
    from baikaldb_b200.column import make_column
    L = _lib.lib()
    dev = torch.cuda.current_device()
    cols = []
    for slot, pt, dist, lo, hi, scale in datagen.C2_COLUMNS:
        if slot == 1:
            hi = N_GROUPS
        dt = torch.int32 if pt == 5 else torch.float64
        t = torch.empty(rows, dtype=dt, device="cuda")
        _lib.check(L.bkgpu_gen_column(dev, t.data_ptr(), int(pt), dist, 2, slot, row0, rows, lo, hi, scale))
        cols.append(make_column(0, slot, pt, t.cpu().numpy()))
        del t
    return cols


def run_reference(args):
    rank, local_rank, world = env_rank()
    if rank != 0:
        return 0
    import pyarrow as pa
    from oracle import acero_oracle as A
    n_gpus = args.gpus
    rows_per_gpu = args.rows or (100_000_000 if n_gpus == 1 else 125_000_000)
    total = rows_per_gpu * n_gpus
    sample = min(total, 100_000_000)
    cores = os.cpu_count() or 1
    pa.set_cpu_count(cores)
    try:
        import torch
        have_gpu = torch.cuda.is_available()
    except Exception:
        have_gpu = False
    cols = host_table_via_gpu(sample) if have_gpu else host_table_numpy(sample)
    table = A.to_table(cols)
    # Acero's table_source hands out 1Mi-row batches; the thread pool works on them in parallel
    for _ in range(max(args.warmup, 1)):
        out = A.c2_filter_groupby(table, K_FILTER, use_threads=True)
    assert out.num_rows == N_GROUPS
    t0 = time.perf_counter()
    for _ in range(args.steps):
        A.c2_filter_groupby(table, K_FILTER, use_threads=True)
    dt = time.perf_counter() - t0
    value = sample * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "rows/s", "n_gpus": n_gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int64/f64", "data": "synthetic",
        "config": {"workload": workload_name(n_gpus, rows_per_gpu), "rows_per_step": sample,
                   "engine": f"Apache Arrow Acero {A.arrow_version()} (reference pins baikalgroup/arrow release-16.1.0): "
                             "table_source -> filter -> aggregate(hash_count_all, hash_sum, hash_mean), use_threads=True"},
        "cpu_baseline": {"value": value, "unit": "rows/s", "cores": cores, "kind": "port",
                         "sample": f"{sample} rows of the workload per step, Acero plan the reference builds, {cores} threads"},
        "e2e": {"value": value, "unit": "rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


# ----------------------------------------------------------------------------------------------
# B200 arm
# ----------------------------------------------------------------------------------------------
_REAL_STDOUT = None


def claim_stdout():
    """The driver reads ONE JSON line from stdout.  Libraries also write there (NCCL prints "NCCL version ..." to stdout when
    NCCL_DEBUG is set on the box): keep the real stdout aside and point fd 1 at stderr for everything else."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    _REAL_STDOUT.write(json.dumps(line) + "\n")
    _REAL_STDOUT.flush()


def main():
    args = parse_args()
    claim_stdout()
    if args.impl == "reference":
        return run_reference(args)
    import numpy as np
    import torch
    import torch.distributed as dist
    from baikaldb_b200 import _lib, datagen, queries
    from baikaldb_b200._lib import BkgpuColumn, BkgpuStats

    rank, local_rank, world = env_rank()
    n_gpus = args.gpus
    if world != n_gpus:
        if world == 1 and n_gpus > 1:
            raise SystemExit("bench.py --gpus N>1 must be launched with torch.distributed.run --nproc-per-node N")
    torch.cuda.set_device(local_rank)
    dev = local_rank
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", dev))
    L = _lib.lib()
    rows = args.rows or (100_000_000 if n_gpus == 1 else 125_000_000)
    row0 = rank * rows
    stream = torch.cuda.current_stream()

    # ---- synthetic region of this rank, generated in HBM ----
    tensors = []
    dcols = (BkgpuColumn * 4)()
    for i, (slot, pt, gdist, lo, hi, scale) in enumerate(datagen.C2_COLUMNS):
        if slot == 1:
            hi = N_GROUPS
        t = torch.empty(rows, dtype=torch.int32 if pt == 5 else torch.float64, device="cuda")
        _lib.check(L.bkgpu_gen_column(dev, t.data_ptr(), int(pt), gdist, 2, slot, row0, rows, lo, hi, scale))
        tensors.append(t)
        dcols[i].tuple_id, dcols[i].slot_id, dcols[i].prim_type, dcols[i].elem_size = 0, slot, int(pt), 0
        dcols[i].values, dcols[i].validity, dcols[i].length = t.data_ptr(), None, rows

    # ---- NCCL communicator of the library (unique id travels over torch.distributed) ----
    comm = ctypes.c_void_p()
    if world > 1:
        idbuf = (ctypes.c_uint8 * 128)()
        if rank == 0:
            _lib.check(L.bkgpu_nccl_unique_id(idbuf))
        idt = torch.tensor(list(idbuf), dtype=torch.uint8, device="cuda")
        dist.broadcast(idt, 0)
        idbuf = (ctypes.c_uint8 * 128)(*idt.cpu().tolist())
        _lib.check(L.bkgpu_nccl_comm_create(ctypes.byref(comm), idbuf, world, rank, dev))

    plan_bytes = queries.c2_filter_groupby(K_FILTER).serialize()
    h = ctypes.c_void_p()
    _lib.check(L.bkgpu_init(ctypes.byref(h), plan_bytes, len(plan_bytes), dev, comm if world > 1 else None))
    _lib.check(L.bkgpu_set_option(h, b"stream", stream.cuda_stream), h)
    _lib.check(L.bkgpu_set_option(h, b"group_capacity_log2", 14), h)
    for kv in filter(None, os.environ.get("BKGPU_BENCH_OPTS", "").split(",")):   # A/B of kernel variants: "l2_lanes=1,no_lean=1"
        k, v = kv.split("=")
        _lib.check(L.bkgpu_set_option(h, k.encode(), int(v)), h)
    _lib.check(L.bkgpu_open(h), h)
    out = (BkgpuColumn * 16)()

    def drain():
        eos = ctypes.c_int(0)
        nrows_total, nbytes = 0, 0
        while not eos.value:
            ncols, nrows = ctypes.c_int(16), ctypes.c_int64(0)
            _lib.check(L.bkgpu_get_next(h, out, ctypes.byref(ncols), ctypes.byref(nrows), ctypes.byref(eos)), h)
            nrows_total += nrows.value
            nbytes += sum(out[i].elem_size for i in range(ncols.value)) * nrows.value
        return nrows_total, nbytes

    def step(cols, on_device):
        _lib.check(L.bkgpu_reset(h), h)
        _lib.check(L.bkgpu_push(h, cols, 4, rows, on_device), h)
        _lib.check(L.bkgpu_finish(h), h)
        return drain()

    def get_stats():
        st = BkgpuStats()
        _lib.check(L.bkgpu_get_stats(h, ctypes.byref(st)), h)
        return st

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(cols, on_device, steps, warmup):
        for _ in range(warmup):
            step(cols, on_device)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sampler = ClockSampler(dev)
        sampler.start()
        launches0 = get_stats().kernel_launches
        e0.record(stream)
        res = None
        kernel_ms, kernel_launches, kernel_bytes, coll_ms = 0.0, 0, 0, 0.0
        for _ in range(steps):
            res = step(cols, on_device)
            st = get_stats()
            kernel_ms += st.main_kernel_ms; kernel_launches += st.main_kernel_launches
            kernel_bytes += st.main_kernel_bytes; coll_ms += st.collective_ms
        e1.record(stream)
        barrier()
        clocks = sampler.stop()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        st = get_stats()
        return {"ms": ms, "res": res, "clocks": clocks, "launches": st.kernel_launches - launches0, "stats": st,
                "kernel_ms": kernel_ms, "kernel_launches": kernel_launches, "kernel_bytes": kernel_bytes, "coll_ms": coll_ms}

    # ---- value: device-resident columns ----
    r = timed(dcols, 1, args.steps, max(args.warmup, 3))
    total_rows = rows * world
    value = total_rows * args.steps / (r["ms"] / 1e3)
    ngroups_out = r["res"][0]

    # ---- one-off verification at full size against torch on the same device columns (not timed) ----
    verified = None
    try:
        key, filt, a, b = tensors
        m = filt < K_FILTER
        cnt = torch.bincount(key[m].to(torch.int64), minlength=N_GROUPS)
        sa = torch.zeros(N_GROUPS, dtype=torch.float64, device="cuda").index_add_(0, key[m].to(torch.int64), a[m])
        sb = torch.zeros(N_GROUPS, dtype=torch.float64, device="cuda").index_add_(0, key[m].to(torch.int64), b[m])
        if world > 1:
            for t in (cnt, sa, sb):
                dist.all_reduce(t)
        _lib.check(L.bkgpu_reset(h), h); _lib.check(L.bkgpu_push(h, dcols, 4, rows, 1), h); _lib.check(L.bkgpu_finish(h), h)
        ncols, nrows, eos = ctypes.c_int(16), ctypes.c_int64(0), ctypes.c_int(0)
        _lib.check(L.bkgpu_get_next(h, out, ctypes.byref(ncols), ctypes.byref(nrows), ctypes.byref(eos)), h)
        n = nrows.value
        by = {}
        for i in range(ncols.value):
            dt = {5: np.int32, 6: np.int64, 12: np.float64}.get(out[i].prim_type)
            if dt is None:
                continue
            by[(out[i].tuple_id, out[i].slot_id)] = np.frombuffer((ctypes.c_uint8 * (n * np.dtype(dt).itemsize)).from_address(out[i].values), dtype=dt).copy()
        order = np.argsort(by[(0, 1)])
        ok = n == N_GROUPS and np.array_equal(by[(0, 1)][order], np.arange(N_GROUPS))
        ok = ok and np.array_equal(by[(1, 1)][order], cnt.cpu().numpy())
        ok = ok and np.allclose(by[(1, 2)][order], sa.cpu().numpy(), rtol=1e-6, atol=0)
        ok = ok and np.allclose(by[(1, 3)][order], (sb / cnt).cpu().numpy(), rtol=1e-6, atol=1e-9)
        verified = bool(ok)
        del m, cnt, sa, sb
    except Exception as e:  # verification must never hide the measurement
        verified = f"error: {e}"

    # ---- e2e: host (pinned) columns through the same calls, H2D inside the timed region ----
    e2e = None
    if not args.no_e2e:
        host = []
        hcols = (BkgpuColumn * 4)()
        for i, t in enumerate(tensors):
            hb = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
            hb.copy_(t)
            host.append(hb)
            hcols[i].tuple_id, hcols[i].slot_id, hcols[i].prim_type, hcols[i].elem_size = 0, dcols[i].slot_id, dcols[i].prim_type, 0
            hcols[i].values, hcols[i].validity, hcols[i].length = hb.data_ptr(), None, rows
        torch.cuda.synchronize()
        e_steps = max(1, min(args.steps, 10))
        re = timed(hcols, 0, e_steps, 1)
        e2e = {"value": total_rows * e_steps / (re["ms"] / 1e3), "unit": "rows/s", "steps": e_steps,
               "ms_per_step": re["ms"] / e_steps, "h2d_bytes_per_step": rows * BYTES_PER_ROW * world,
               "d2h_bytes_per_step": int(re["res"][1]) + 16,
               "h2d_gbs_per_gpu": rows * BYTES_PER_ROW * e_steps / (re["ms"] / 1e3) / 1e9}
        del host

    # ---- roofline of the dominant kernel (algorithmic bytes / CUDA-event duration of its launches) ----
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    achieved = (r["kernel_bytes"] / max(r["kernel_launches"], 1)) / (r["kernel_ms"] / max(r["kernel_launches"], 1) / 1e3) / 1e9 if r["kernel_ms"] > 0 else 0.0
    traffic = None
    try:
        prof = json.load(open(os.path.join(ROOT, "profiles", "latest_traffic.json")))
        entry = prof.get(r["stats"].main_kernel_name.decode())
        if entry:  # DRAM bytes per launch, scaled from the committed ncu capture to this launch's row count
            traffic = entry["dram_bytes_per_row"] * rows
    except Exception:
        pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "kernel": r["stats"].main_kernel_name.decode(), "peak_source": peak_src,
                "kernel_ms_per_launch": r["kernel_ms"] / max(r["kernel_launches"], 1),
                "kernel_share_of_step": r["kernel_ms"] / r["ms"] if r["ms"] else None,
                "collective_ms_per_step": r["coll_ms"] / args.steps}

    # ---- CPU baseline beside it (rank 0, N = 1): the row-engine restatement, one thread ----
    cpu_baseline = None
    if rank == 0 and n_gpus == 1 and not args.no_cpu_baseline:
        from baikaldb_b200.column import make_column
        from oracle import oracle as row_oracle

        def sample_cols(n):
            return [make_column(0, dcols[i].slot_id, dcols[i].prim_type, tensors[i][:n].cpu().numpy()) for i in range(4)]
        probe = sample_cols(1_000_000)
        t0 = time.perf_counter(); row_oracle.execute(plan_bytes, probe); dt = time.perf_counter() - t0
        n_s = int(min(rows, max(2_000_000, 15.0 / (dt / 1e6))))
        cols = sample_cols(n_s)
        t0 = time.perf_counter(); res = row_oracle.execute(plan_bytes, cols); dt = time.perf_counter() - t0
        cpu_baseline = {"value": n_s / dt, "unit": "rows/s", "cores": 1, "kind": "port",
                        "sample": f"first {n_s} rows of the workload, oracle/bk_oracle.c (row-engine restatement, one thread like one bthread per fragment)",
                        "host_cores_available": os.cpu_count(), "seconds": dt, "groups": res.nrows}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "rows/s", "n_gpus": n_gpus, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": r["ms"] / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int64/f64", "data": "synthetic",
            "config": {"workload": workload_name(n_gpus, rows), "rows_per_gpu": rows, "selectivity": 0.5, "groups": N_GROUPS,
                       "algorithmic_bytes_per_row": BYTES_PER_ROW,
                       "l2": f"inputs {rows * BYTES_PER_ROW / 1e9:.1f} GB per GPU >> 126 MB L2: no flush needed",
                       "step": "bkgpu_reset + bkgpu_push(on_device) + bkgpu_finish + bkgpu_get_next"},
            "clocks": {k: r["clocks"][k] for k in ("sm_mhz", "sm_max_mhz", "reasons")},
            "e2e": e2e, "gpu_launches": int(r["launches"]), "roofline": roofline, "cpu_baseline": cpu_baseline,
            "hbm_gbs_whole_step": total_rows * BYTES_PER_ROW * args.steps / (r["ms"] / 1e3) / 1e9 / world,
            "result_groups": ngroups_out, "verified_vs_torch": verified,
        }
        emit(line)
    L.bkgpu_close(h)
    if world > 1:
        L.bkgpu_nccl_comm_destroy(comm)
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
