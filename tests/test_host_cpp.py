"""The C++ host side (host/bkgpu_host.hpp: ExecNode / RuntimeState / pb::Plan mirrors over the C ABI).
CPU: its plan writer emits the same bytes as plan.py and the lowering accepts them.
GPU: the ExecNode tree of host_main.cpp (ColumnScanNode children -> GpuExecNode) returns the oracle's rows on the
same synthetic tables (the C++ generator is a restatement of datagen.py, so equality also pins the generator)."""
import math
import os
import re
import subprocess

import pytest

from baikaldb_b200 import datagen, queries
from oracle import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "baikaldb_b200", "bkgpu_host")
PLANS = {"c1": queries.c1_count_where, "c2": queries.c2_filter_groupby, "c3": queries.c3_join_groupby, "c5": queries.c5_topk}


@pytest.fixture(scope="module")
def host_bin():
    if not os.path.exists(BIN):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "host")])
    return BIN


@pytest.mark.parametrize("cfg", sorted(PLANS))
def test_cpp_plan_writer_matches_python(host_bin, cfg):
    out = subprocess.run([host_bin, "plan", cfg], capture_output=True, text=True, check=True).stdout.strip()
    assert out == PLANS[cfg]().serialize().hex()


@pytest.mark.parametrize("cfg", sorted(PLANS))
def test_cpp_explain(host_bin, cfg):
    out = subprocess.run([host_bin, "explain", cfg], capture_output=True, text=True, check=True).stdout
    assert out.startswith("kind=")


@pytest.mark.parametrize("rows,capacity", [(1, 1), (1000, 64), (5003, 1024), (77, 1000)])
def test_cpp_chunk_row_column_round_trip(host_bin, rows, capacity):
    """f1 adapter on the CPU: MemRow-style values (NULLs, narrow ints, float, bool) -> Chunk column batches of <= capacity rows -> rows"""
    r = subprocess.run([host_bin, "chunk", "-", str(rows), str(capacity)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert f"rows={rows} batches={(rows + capacity - 1) // capacity} mismatches=0" in r.stdout


def _parse(text):
    rows = []
    for line in text.splitlines():
        row = {}
        for tok in line.split():
            name, val = tok.split("=", 1)
            m = re.fullmatch(r"blob\((.*),(-?\d+)\)", val)
            row[name] = None if val == "NULL" else (float(m.group(1)), int(m.group(2))) if m else float(val) if re.search(r"[.eEn]", val) else int(val)
        rows.append(row)
    return rows


def _oracle_rows(plan, cols):
    res = oracle.execute(plan.serialize(), cols)
    names = [c.name for c in res.columns]
    lists = [c.to_list() for c in res.columns]
    rows = []
    for i in range(len(lists[0]) if lists else 0):
        row = {}
        for nm, l in zip(names, lists):
            v = l[i]
            if isinstance(v, bytes):
                import struct
                v = struct.unpack("<dq", v)
            row[nm] = v
        rows.append(row)
    return rows


def _same(a, b):
    if isinstance(a, tuple):
        return a[1] == b[1] and math.isclose(a[0], b[0], rel_tol=1e-6, abs_tol=1e-9)
    if isinstance(a, float) or isinstance(b, float):
        return math.isclose(a, b, rel_tol=1e-6, abs_tol=1e-9)
    return a == b


@pytest.mark.gpu
@pytest.mark.parametrize("cfg,rows,batch,key", [("c1", 300_000, 100_000, None), ("c2", 400_003, 131_072, "0_1"),
                                                ("c3", 200_000, 64_000, "1_2"), ("c5", 300_001, 70_000, None)])
def test_cpp_exec_node_tree_matches_oracle(host_bin, cfg, rows, batch, key):
    r = subprocess.run([host_bin, "run", cfg, str(rows), str(batch)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    got = _parse(r.stdout)
    if cfg == "c1":
        cols = datagen.c1_table(0, rows)
    elif cfg == "c2":
        cols = datagen.c2_table(0, rows)
    elif cfg == "c5":
        cols = datagen.c5_table(0, rows)
    else:
        cols = datagen.c3_dim(0, rows // 10, rows // 10) + datagen.c3_fact(0, rows, rows // 10)
    want = _oracle_rows(PLANS[cfg](), cols)
    assert len(got) == len(want)
    if key is not None:
        got, want = sorted(got, key=lambda x: x[key]), sorted(want, key=lambda x: x[key])
    for g, w in zip(got, want):
        assert set(g) == set(w)
        for nm in g:
            assert _same(g[nm], w[nm]), (nm, g, w)
    if cfg != "c3":
        assert "scan_rows=%d" % rows in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("rows,capacity", [(50_001, 1024), (3_000, 7)])
def test_cpp_row_engine_child_through_chunk_adapter(host_bin, rows, capacity):
    """f1: MemRow-style rows (NULL keys included) -> Chunk column batches of `capacity` rows -> GPU -> rows again."""
    r = subprocess.run([host_bin, "rows", "c2", str(rows), str(capacity)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    got = _parse(r.stdout)
    cols = datagen.c2_table(0, rows)
    from baikaldb_b200.column import make_column
    k = cols[0]
    cols[0] = make_column(k.tuple_id, k.slot_id, k.prim_type, k.values, k.values % 17 != 0)
    want = _oracle_rows(PLANS["c2"](), cols)
    assert len(got) == len(want)
    keyf = lambda x: (x["0_1"] is None, x["0_1"] or 0)
    for g, w in zip(sorted(got, key=keyf), sorted(want, key=keyf)):
        assert set(g) == set(w)
        for nm in g:
            assert (g[nm] is None and w[nm] is None) or _same(g[nm], w[nm]), (nm, g, w)
    assert "scan_rows=%d" % rows in r.stderr


def test_stream_copy_is_a_byte_exact_copy(tmp_path):
    """the non-temporal bounce copy of the pageable-input path (csrc/hostcopy.cpp): every size / alignment combination equals memcpy and
    writes nothing outside its range"""
    exe = str(tmp_path / "stream_copy_check")
    subprocess.check_call(["g++", "-O2", "-std=c++17", os.path.join(ROOT, "tests", "cpp", "stream_copy_check.cpp"), os.path.join(ROOT, "csrc", "hostcopy.cpp"), "-o", exe])
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and "bad=0" in r.stdout, r.stdout + r.stderr
