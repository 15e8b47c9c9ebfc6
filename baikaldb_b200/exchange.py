"""Hash exchange of column batches between ranks — the host-side plumbing of SURVEY.md §8 f3 beyond aggregates:
repartitioned (distributed) joins and COUNT(DISTINCT) across regions.

The reference moves rows between fragments with ExchangeSenderNode / ExchangeReceiverNode: every row goes to the receiver
``hash(partition exprs) % n`` (src/exec/exchange_sender_node.cpp:867-957, Arrow ``Hashing32`` over the key columns), so that all rows
with equal keys meet in one fragment instance; a join partitions both inputs on the join key, an aggregate with DISTINCT functions
partitions the store-side ``GROUP BY (k, x)`` rows on ``k`` before the db-side MERGE_AGG (select_planner.cpp:612-700).

Here one rank = one GPU (one set of regions).  A batch is a list of ``(tuple_id, slot_id, prim_type, values tensor, valid tensor | None)``;
``exchange`` routes its rows with ONE ``all_to_all_single`` per buffer over ``torch.distributed`` — NCCL over NVLink for CUDA tensors,
gloo for host tensors (how ``tests/test_exchange_gloo.py`` runs it on the CPU).  Partitioning (hash, stable bucket sort) is a handful of
torch ops on the batch's own device; the operators themselves stay in ``libbkgpu.so``: after the exchange every rank runs the ordinary
fragment (``AGG -> JOIN``, ``MERGE_AGG``) over what it received, through ``DeviceColumn``s, and the partial results merge as before.
Only consistency inside this system matters for the hash, so a 64-bit mixer replaces ``Hashing32`` (a rank never sees the reference's
partitions); NULL keys hash like a constant: they all land on one rank (they are one GROUP BY group, and they join nothing).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist

from .column import Column, make_column
from .plan import storage_dtype

Batch = List[Tuple[int, int, int, torch.Tensor, Optional[torch.Tensor]]]

_M1, _M2 = -7046029254386353131, -4265267296055464877   # 0x9E3779B97F4A7C15, 0xC4CEB9FE1A85EC53 as signed 64-bit
_NULL_HASH = 0x5BD1E995


def _lsr(x: torch.Tensor, s: int) -> torch.Tensor:
    """logical shift right of int64 (torch only has the arithmetic one)"""
    return (x >> s) & ((1 << (64 - s)) - 1)


def _mix64(x: torch.Tensor) -> torch.Tensor:
    """murmur3's 64-bit finalizer on int64 tensors (two's-complement wraparound is what the multiplications need)"""
    x = x ^ _lsr(x, 33)
    x = x * _M1
    x = x ^ _lsr(x, 29)
    x = x * _M2
    return x ^ _lsr(x, 32)


def _as_int64(v: torch.Tensor) -> torch.Tensor:
    if v.dtype == torch.float64:
        v = torch.where(v == 0, torch.zeros_like(v), v)      # +0.0 and -0.0 are one key
        return v.view(torch.int64)
    if v.dtype == torch.float32:
        return torch.where(v == 0, torch.zeros_like(v), v).to(torch.float64).view(torch.int64)
    if v.dtype in (torch.uint8, torch.bool):
        return v.to(torch.int64)
    return v.to(torch.int64)     # signed and (numpy-viewed) unsigned integers: equal values give equal images


def destination(batch: Batch, key_slots: Sequence[Tuple[int, int]], world: int) -> torch.Tensor:
    """rank of every row: hash of the key columns named by (tuple_id, slot_id) — the partition exprs of the exchange"""
    by_name = {(t, s): (v, ok) for t, s, _, v, ok in batch}
    h = None
    for ts in key_slots:
        v, ok = by_name[tuple(ts)]
        k = _mix64(_as_int64(v))
        if ok is not None:
            k = torch.where(ok, k, torch.full_like(k, _NULL_HASH))
        h = k if h is None else _mix64(h ^ (k + _M1))
    return (_lsr(h, 1) % world).to(torch.int64)


def exchange(batch: Batch, dest: torch.Tensor, group=None) -> Batch:
    """every row goes to rank dest[row]; returns the rows this rank received (sender order: rank 0's rows first, each sender's rows in
    their original order — the exchange is deterministic)"""
    world = dist.get_world_size(group)
    order = torch.argsort(dest, stable=True)
    send = torch.bincount(dest, minlength=world)
    recv = torch.empty_like(send)
    dist.all_to_all_single(recv, send, group=group)
    send_l, recv_l = send.tolist(), recv.tolist()
    n_recv = int(sum(recv_l))
    out: Batch = []
    for t, s, prim, v, ok in batch:
        moved = v.index_select(0, order).contiguous()
        got = torch.empty((n_recv,) + tuple(v.shape[1:]), dtype=v.dtype, device=v.device)
        dist.all_to_all_single(got, moved, recv_l, send_l, group=group)
        # a column is nullable after the exchange if ANY rank sent NULLs in it: every rank must take part in the same collectives
        flag = torch.tensor([0 if ok is None else 1], dtype=torch.int64, device=dest.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=group)
        got_ok = None
        if int(flag.item()):
            mine = (torch.ones(v.shape[0], dtype=torch.uint8, device=v.device) if ok is None else ok.to(torch.uint8)).index_select(0, order).contiguous()
            rok = torch.empty(n_recv, dtype=torch.uint8, device=v.device)
            dist.all_to_all_single(rok, mine, recv_l, send_l, group=group)
            got_ok = rok.to(torch.bool)
        out.append((t, s, prim, got, got_ok))
    return out


# ---- host-side conversions (tests, the CPU container): numpy Columns <-> torch batches ----
_VIEW = {np.dtype(np.uint32): np.int32, np.dtype(np.uint64): np.int64, np.dtype(np.uint16): np.int16}


def batch_from_columns(cols: Sequence[Column], device: str = "cpu") -> Batch:
    out: Batch = []
    for c in cols:
        a = np.ascontiguousarray(c.values)
        if a.dtype in _VIEW:
            a = a.view(_VIEW[a.dtype])    # torch has no unsigned 32 / 64-bit arithmetic: same bits, signed view
        out.append((c.tuple_id, c.slot_id, c.prim_type, torch.from_numpy(a.copy()).to(device), None if c.valid is None else torch.from_numpy(np.ascontiguousarray(c.valid)).to(device)))
    return out


def columns_from_batch(batch: Batch) -> List[Column]:
    cols = []
    for t, s, prim, v, ok in batch:
        a = v.cpu().numpy()
        want = np.dtype(storage_dtype(prim)) if a.ndim == 1 else a.dtype
        if a.ndim == 1 and a.dtype != want and a.dtype.itemsize == want.itemsize:
            a = a.view(want)
        cols.append(make_column(t, s, prim, a, None if ok is None else ok.cpu().numpy()))
    return cols


def device_columns(batch: Batch):
    """the received batch as DeviceColumns for GpuExecNode.push (CUDA tensors: no copy; the tensors are kept alive by the columns)"""
    from .exec_node import DeviceColumn
    out = []
    for t, s, prim, v, ok in batch:
        if ok is not None:
            raise ValueError("device_columns: pack the validity into an Arrow bitmap first (Column.validity_bitmap) — NULL-free batches only here")
        out.append(DeviceColumn(t, s, int(prim), v.data_ptr(), int(v.shape[0]), 0, v))
    return out
