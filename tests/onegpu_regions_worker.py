"""Worker of the region-merge parity test that needs only ONE GPU: `world` ranks (torch.distributed.run, gloo) share cuda:0.
NCCL refuses two ranks on one device, so each rank runs its region through the C ABI without a communicator, exports the
per-region partial state (bkgpu_partial_export: the compact-row layout the in-library all-gather ships, k_partial_export_rows),
the ranks exchange the buffers over gloo, and every rank folds them with bkgpu_partial_merge (k_partial_merge_rows /
sort_partial_merge: AggFnCall::merge, src/expr/agg_fn_call.cpp:719-822; SelectManagerNode, select_manager_node.cpp:50-51).
Every rank checks the merged result against the oracle run over the WHOLE table."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch
import torch.distributed as dist

from baikaldb_b200 import _lib, datagen, queries
from baikaldb_b200._lib import BkgpuError
from baikaldb_b200.column import make_column
from baikaldb_b200.exec_node import ColumnSource, GpuExecNode, RowBatch, RuntimeState
from tests.util import assert_same_rows
from oracle import oracle

L = _lib.lib()


def fetch(node, st):
    out, rb, eos = [], RowBatch(), False
    while not eos:
        rc, eos = node.get_next(st, rb)
        assert rc == 0, st.error_msg
        out = out or list(rb.columns)
    return out


def run_region(plan, cols, options):
    st = RuntimeState(device=0, options=dict(options))
    node = GpuExecNode()
    node.init(plan)
    node.add_child(ColumnSource([cols]))
    assert node.open(st) == 0, st.error_msg
    return node, st


def exchange_and_merge(node, world):
    h = node.handle()
    nbytes = ctypes.c_size_t(0)
    _lib.check(L.bkgpu_partial_capacity(h, ctypes.byref(nbytes)), h)
    mine = torch.zeros(nbytes.value, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()                                # (the library works on its own stream: torch's fill must have run)
    _lib.check(L.bkgpu_partial_export(h, ctypes.c_void_p(mine.data_ptr()), nbytes.value), h)
    torch.cuda.synchronize()
    parts = [torch.zeros(nbytes.value, dtype=torch.uint8) for _ in range(world)]
    dist.all_gather(parts, mine.cpu())                      # gloo moves the partial states
    allp = torch.cat(parts).cuda()
    torch.cuda.synchronize()
    _lib.check(L.bkgpu_partial_merge(h, ctypes.c_void_p(allp.data_ptr()), nbytes.value, world), h)
    torch.cuda.synchronize()
    return allp


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(0)
    dist.init_process_group("gloo")
    n_region = 120_000
    # ---- C4: GROUP BY over `world` regions (1 200 groups, every rank sees most of them) ----
    plan = queries.c2_filter_groupby()
    for n_groups, opts in ((1200, {}), (9000, {"partial_capacity": 16384})):
        region = datagen.c2_table(rank * n_region, n_region, n_groups=n_groups)
        node, st = run_region(plan, region, opts)
        fetch(node, st)                                     # the region's own result (store side)
        exchange_and_merge(node, world)
        got = fetch(node, st)                               # the merged result (db side)
        whole = datagen.c2_table(0, n_region * world, n_groups=n_groups)
        want = oracle.execute(plan.serialize(), whole)
        assert_same_rows(got, want.columns, ["0_1"])
        node.close(st)
    # ---- a rank that holds more groups than partial_capacity must fail loudly, not ship a truncated state ----
    region = datagen.c2_table(rank * n_region, n_region, n_groups=5000)
    node, st = run_region(plan, region, {"partial_capacity": 1000})
    fetch(node, st)
    try:
        exchange_and_merge(node, world)
        raise AssertionError("export of 5000 groups with partial_capacity=1000 did not fail")
    except BkgpuError as e:
        assert e.code == _lib.ETOOBIG, e
    node.close(st)
    # ---- scalar aggregate (C1 shape): one group per region ----
    c1 = datagen.c1_table(rank * n_region, n_region)
    node, st = run_region(queries.c1_count_where(), c1, {})
    fetch(node, st)
    exchange_and_merge(node, world)
    got = fetch(node, st)
    want = oracle.execute(queries.c1_count_where().serialize(), datagen.c1_table(0, n_region * world))
    assert_same_rows(got, want.columns, [])
    node.close(st)
    # ---- C5: ORDER BY ... LIMIT over regions; duplicates across regions break ties by (region, row) ----
    rng = np.random.default_rng(77)
    keys = rng.integers(0, 5000, n_region * world)
    pay = np.arange(n_region * world, dtype=np.int32)
    whole5 = [make_column(0, 1, 6, keys), make_column(0, 2, 5, pay)]
    mine = [make_column(0, 1, 6, keys[rank * n_region:(rank + 1) * n_region]), make_column(0, 2, 5, pay[rank * n_region:(rank + 1) * n_region])]
    node, st = run_region(queries.c5_topk(1000), mine, {"region_base": rank * n_region})
    fetch(node, st)
    exchange_and_merge(node, world)
    got5 = fetch(node, st)
    want5 = oracle.execute(queries.c5_topk(1000).serialize(), whole5)
    assert len(got5[0]) == len(want5.columns[0]) == 1000, (rank, len(got5), [len(c) for c in got5], [len(c) for c in want5.columns])
    assert_same_rows(got5, want5.columns, None)
    node.close(st)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("ONEGPU_REGIONS_OK", world)


if __name__ == "__main__":
    main()
