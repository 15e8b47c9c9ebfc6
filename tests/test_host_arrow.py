"""C++ Arrow bridge (host/bkgpu_arrow.hpp): wire bytes (SerializeSchema + SerializeRecordBatch, src/store/region.cpp:2905-2918) ->
bkgpu_column views -> back.  CPU: views alias the right values / NULLs (hash per column), round trip is the identity, bad inputs are
refused.  GPU: one fragment executed by the C++ binary with IPC in and out equals the oracle."""
import os
import subprocess

import numpy as np
import pyarrow as pa
import pytest

from baikaldb_b200 import arrow_io, datagen, queries
from baikaldb_b200.column import make_column
from baikaldb_b200.plan import PrimitiveType as T
from oracle import oracle
from tests.util import assert_same_rows

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "baikaldb_b200", "bkgpu_arrow_host")


@pytest.fixture(scope="module")
def arrow_bin():
    if not os.path.exists(BIN):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "host"), "arrow", "PYARROW_DIR=" + os.path.dirname(pa.__file__)])
    return BIN


def _cols(n=5000, seed=4):
    rng = np.random.default_rng(seed)
    return [make_column(0, 1, T.INT32, rng.integers(-9, 9, n), rng.random(n) > 0.2), make_column(0, 2, T.INT64, rng.integers(-1 << 50, 1 << 50, n)),
            make_column(0, 3, T.DOUBLE, rng.normal(size=n), rng.random(n) > 0.5), make_column(0, 4, T.UINT32, rng.integers(0, 1 << 32, n, dtype=np.uint64)),
            make_column(0, 5, T.UINT64, rng.integers(0, 1 << 63, n, dtype=np.uint64)), make_column(0, 6, T.FLOAT, rng.normal(size=n).astype(np.float32))]


def _fnv(c):
    h = 1469598103934665603
    raw = np.ascontiguousarray(c.values).view(np.uint8).reshape(len(c), -1)
    valid = np.ones(len(c), bool) if c.valid is None else c.valid
    for r in np.nonzero(valid)[0]:
        for b in raw[r]:
            h = ((h ^ int(b)) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h


def test_describe_views_alias_values_and_nulls(arrow_bin, tmp_path):
    cols = _cols(700)
    s, d = arrow_io.encode(cols)
    (tmp_path / "s").write_bytes(s); (tmp_path / "d").write_bytes(d)
    out = subprocess.run([arrow_bin, "describe", str(tmp_path / "s"), str(tmp_path / "d")], capture_output=True, text=True, check=True).stdout.splitlines()
    assert len(out) == len(cols)
    for line, c in zip(out, cols):
        f = dict(kv.split("=") for kv in line.split()[1:])
        assert line.split()[0] == c.name and int(f["prim"]) == c.prim_type and int(f["rows"]) == len(c)
        assert int(f["nulls"]) == (0 if c.valid is None else int((~c.valid).sum()))
        assert int(f["hash"], 16) == _fnv(c)


def test_round_trip_is_the_identity(arrow_bin, tmp_path):
    cols = _cols()
    s, d = arrow_io.encode(cols)
    (tmp_path / "s").write_bytes(s); (tmp_path / "d").write_bytes(d)
    subprocess.run([arrow_bin, "roundtrip", str(tmp_path / "s"), str(tmp_path / "d"), str(tmp_path / "so"), str(tmp_path / "do")], check=True)
    back = arrow_io.decode((tmp_path / "so").read_bytes(), (tmp_path / "do").read_bytes())
    assert_same_rows(back, cols, None, rel=0.0)
    assert pa.ipc.read_schema(pa.py_buffer((tmp_path / "so").read_bytes())).equals(pa.ipc.read_schema(pa.py_buffer(s)))


def test_refuses_fields_outside_the_path(arrow_bin, tmp_path):
    rb = pa.RecordBatch.from_arrays([pa.array(["a", "b"])], names=["0_1"])
    (tmp_path / "s").write_bytes(rb.schema.serialize().to_pybytes()); (tmp_path / "d").write_bytes(rb.serialize().to_pybytes())
    r = subprocess.run([arrow_bin, "describe", str(tmp_path / "s"), str(tmp_path / "d")], capture_output=True, text=True)
    assert r.returncode != 0 and "outside the zero-copy path" in r.stderr
    rb = pa.RecordBatch.from_arrays([pa.array([1, 2])], names=["price"])
    (tmp_path / "s").write_bytes(rb.schema.serialize().to_pybytes()); (tmp_path / "d").write_bytes(rb.serialize().to_pybytes())
    r = subprocess.run([arrow_bin, "describe", str(tmp_path / "s"), str(tmp_path / "d")], capture_output=True, text=True)
    assert r.returncode != 0 and "<tuple>_<slot>" in r.stderr


@pytest.mark.gpu
def test_cpp_fragment_with_ipc_in_and_out(arrow_bin, tmp_path):
    cols = datagen.c2_table(0, 200_000, n_groups=77)
    plan = queries.c2_filter_groupby()
    s, d = arrow_io.encode(cols)
    (tmp_path / "p").write_bytes(plan.serialize()); (tmp_path / "s").write_bytes(s); (tmp_path / "d").write_bytes(d)
    r = subprocess.run([arrow_bin, "exec", str(tmp_path / "p"), str(tmp_path / "s"), str(tmp_path / "d"), str(tmp_path / "so"), str(tmp_path / "do")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    got = arrow_io.decode((tmp_path / "so").read_bytes(), (tmp_path / "do").read_bytes(), plan.tuples)
    want = oracle.execute(plan.serialize(), cols)
    assert_same_rows(got, want.columns, ["0_1"])
