"""Parity at BASELINE.json's full sizes, through properties that need no CPU oracle run: the tables are generated on the device
(bkgpu_gen_column, bit-identical to datagen.py), the fragment runs through the C ABI on device-resident columns, and the result
is checked against an independent torch computation on the same columns — counts, integer results, keys and row order exact,
double sums within the north star's 1e-6 relative.  C1 100M rows, C2 100M, C3 100M x 10M, C5 125M (one GPU's region)."""
import numpy as np
import pytest

from tests.util import LEAN_KERNELS
import torch

from baikaldb_b200 import _lib, datagen, queries
from baikaldb_b200.exec_node import DeviceColumn, execute
from baikaldb_b200.plan import PrimitiveType as T

pytestmark = pytest.mark.gpu
DT = {T.INT32: torch.int32, T.INT64: torch.int64, T.DOUBLE: torch.float64}


def _gen(tuple_id, slot, prim, dist, seed, column_id, n, lo=0, hi=0, scale=1.0):
    x = torch.empty(n, dtype=DT[prim], device="cuda")
    _lib.check(_lib.lib().bkgpu_gen_column(0, x.data_ptr(), int(prim), dist, seed, column_id, 0, n, lo, hi, scale))
    return x, DeviceColumn(tuple_id, slot, int(prim), x.data_ptr(), n, keepalive=x)


def _by_name(cols):
    return {c.name: c for c in cols}


def test_generator_on_device_equals_host_statement():
    x, _ = _gen(0, 4, T.DOUBLE, datagen.DIST_NORMAL_IH4, 2, 4, 100_000, scale=datagen.NORMAL_SCALE_1E3)
    assert np.array_equal(x.cpu().numpy(), datagen.c2_table(0, 100_000)[3].values)
    p, _ = _gen(1, 1, T.INT32, datagen.DIST_PERMUTATION, 3, 11, 50_000, hi=50_000)
    assert np.array_equal(p.cpu().numpy(), datagen.c3_dim(0, 50_000, 50_000)[0].values)


def test_c1_full_size_count():
    n = 100_000_000
    x, c = _gen(0, 1, T.INT32, 0, 1, 1, n, 0, 1 << 20)
    got, stats = execute(queries.c1_count_where(), [c])
    assert got[0].to_list() == [int((x < (1 << 19)).sum().item())]
    assert stats.rows_scanned == n and stats.rows_filtered == n - got[0].to_list()[0]


def test_c2_full_size_groupby():
    n, g = 100_000_000, 1000
    key, ck = _gen(0, 1, T.INT32, 0, 2, 1, n, 0, g)
    flt, cf = _gen(0, 2, T.INT32, 0, 2, 2, n, 0, 1 << 20)
    a, ca = _gen(0, 3, T.DOUBLE, 1, 2, 3, n)
    b, cb = _gen(0, 4, T.DOUBLE, 2, 2, 4, n, scale=datagen.NORMAL_SCALE_1E3)
    got, stats = execute(queries.c2_filter_groupby(), [ck, cf, ca, cb], options={"group_capacity_log2": 14})
    assert stats.main_kernel_name.decode() in LEAN_KERNELS
    m = flt < (1 << 19)
    k64 = key[m].to(torch.int64)
    cnt = torch.bincount(k64, minlength=g)
    sa = torch.zeros(g, dtype=torch.float64, device="cuda").index_add_(0, k64, a[m])
    sb = torch.zeros(g, dtype=torch.float64, device="cuda").index_add_(0, k64, b[m])
    by = _by_name(got)
    order = np.argsort(by["0_1"].values)
    assert np.array_equal(by["0_1"].values[order], np.arange(g))
    assert np.array_equal(by["1_1"].values[order], cnt.cpu().numpy())                                  # COUNT: bit-exact
    assert np.allclose(by["1_2"].values[order], sa.cpu().numpy(), rtol=1e-6, atol=0)                  # SUM(double): 1e-6 relative
    assert np.allclose(by["1_3"].values[order], (sb / cnt).cpu().numpy(), rtol=1e-6, atol=1e-9)        # AVG
    blob = by["1_4"].values[order].copy().view(np.int64).reshape(g, 2)[:, 1]
    assert np.array_equal(blob, cnt.cpu().numpy())                                                    # AVG intermediate count
    assert stats.rows_filtered == n - int(m.sum().item())


def test_c3_full_size_join_groupby():
    nf, nd, g = 100_000_000, 10_000_000, 1000
    fk, cfk = _gen(0, 1, T.INT32, 0, 3, 1, nf, 0, nd)
    v, cv = _gen(0, 2, T.DOUBLE, 1, 3, 2, nf)
    pk, cpk = _gen(1, 1, T.INT32, datagen.DIST_PERMUTATION, 3, 11, nd, hi=nd)
    attr, cattr = _gen(1, 2, T.INT32, 0, 3, 12, nd, 0, g)
    got, stats = execute(queries.c3_join_groupby(), [[cpk, cattr], [cfk, cv]], options={"group_capacity_log2": 14})
    assert stats.main_kernel_name.decode() in LEAN_KERNELS                                      # probe fused into the aggregate
    attr_of_key = torch.empty(nd, dtype=torch.int64, device="cuda")
    attr_of_key[pk.to(torch.int64)] = attr.to(torch.int64)                                            # pk is a permutation of [0, nd)
    grp = attr_of_key[fk.to(torch.int64)]
    cnt = torch.bincount(grp, minlength=g)
    sv = torch.zeros(g, dtype=torch.float64, device="cuda").index_add_(0, grp, v)
    by = _by_name(got)
    order = np.argsort(by["1_2"].values)
    assert np.array_equal(by["1_2"].values[order], np.arange(g))
    assert np.array_equal(by["2_1"].values[order], cnt.cpu().numpy()) and int(cnt.sum().item()) == nf
    assert np.allclose(by["2_2"].values[order], sv.cpu().numpy(), rtol=1e-6, atol=0)


def test_c5_full_size_topk_is_sorted_and_exact():
    n, k = 125_000_000, 1000
    key, ck = _gen(0, 1, T.INT64, 3, 5, 1, n)
    pay, cp = _gen(0, 2, T.INT32, 0, 5, 2, n, 0, 1 << 30)
    got, _ = execute(queries.c5_topk(k), [ck, cp])
    by = _by_name(got)
    ks = by["0_1"].values
    assert len(ks) == k and np.all(ks[:-1] <= ks[1:])                                                  # sortedness
    top = torch.topk(key, k, largest=False, sorted=True)
    assert np.array_equal(ks, top.values.cpu().numpy())                                                # the k smallest keys, exactly
    assert np.array_equal(by["0_2"].values, pay[top.indices].cpu().numpy())                            # with their own payload
    # idempotence: the top-k of the top-k is itself
    again, _ = execute(queries.c5_topk(k), got)
    assert np.array_equal(_by_name(again)["0_1"].values, ks)
