"""Arrow IPC on the store <-> db wire (src/store/region.cpp:2905-2918, src/exec/fetcher_store.cpp:1130-1160).
CPU: encode/decode round trips keep names, pb types, NULLs and AVG blobs, and alias the Arrow buffers.
GPU: regions answer a fragment over IPC, the db side merges the IPC batches with MERGE_AGG_NODE — all on the GPU path."""
import numpy as np
import pyarrow as pa
import pytest

from baikaldb_b200 import arrow_io, datagen, queries
from baikaldb_b200.column import make_column
from baikaldb_b200.plan import PrimitiveType as T
from oracle import oracle
from tests.util import assert_same_rows


def _mixed(n=1000, seed=3):
    rng = np.random.default_rng(seed)
    blob = np.zeros((n, 16), np.uint8)
    blob.view(np.float64).reshape(n, 2)[:, 0] = rng.normal(size=n)
    blob.view(np.int64).reshape(n, 2)[:, 1] = rng.integers(0, 99, n)
    return [make_column(0, 1, T.INT32, rng.integers(-5, 5, n), rng.random(n) > 0.2), make_column(0, 2, T.INT64, rng.integers(-1 << 40, 1 << 40, n)),
            make_column(0, 3, T.UINT32, rng.integers(0, 1 << 32, n, dtype=np.uint64)), make_column(0, 4, T.UINT64, rng.integers(0, 1 << 63, n, dtype=np.uint64)),
            make_column(0, 5, T.DOUBLE, rng.normal(size=n), rng.random(n) > 0.5), make_column(0, 6, T.FLOAT, rng.normal(size=n).astype(np.float32)),
            make_column(0, 7, T.INT8, rng.integers(-128, 127, n)), make_column(0, 8, T.DATETIME, rng.integers(0, 1 << 50, n, dtype=np.uint64)),
            make_column(1, 4, T.STRING, blob, rng.random(n) > 0.1), make_column(0, 9, T.BOOL, rng.integers(0, 2, n), rng.random(n) > 0.3)]


def test_round_trip_keeps_names_types_nulls_and_blobs():
    cols = _mixed()
    tuples = {0: [(c.slot_id, c.prim_type) for c in cols if c.tuple_id == 0], 1: [(4, T.STRING)]}
    schema, rows = arrow_io.encode(cols)
    sch = pa.ipc.read_schema(pa.py_buffer(schema))
    assert [f.name for f in sch] == [c.name for c in cols]
    assert [str(f.type) for f in sch] == ["int32", "int64", "uint32", "uint64", "double", "float", "int32", "uint64", "large_binary", "bool"]
    back = arrow_io.decode(schema, rows, tuples)
    assert [c.prim_type for c in back] == [c.prim_type for c in cols]
    assert_same_rows(back, cols, None, rel=0.0)


def test_decode_aliases_the_arrow_buffers_and_checks_declared_types():
    cols = _mixed(64)[:2]
    rb = arrow_io.record_batch_from_columns(cols)
    back = arrow_io.columns_from_record_batch(rb)
    assert back[1].values.ctypes.data == rb.column(1).buffers()[1].address          # zero copy
    sliced = arrow_io.columns_from_record_batch(rb.slice(8, 16))
    assert sliced[1].values.tolist() == cols[1].values[8:24].tolist() and len(sliced[0]) == 16
    with pytest.raises(ValueError):
        arrow_io.columns_from_record_batch(rb, {0: [(2, T.DOUBLE)]})                # plan says DOUBLE, wire says int64
    with pytest.raises(ValueError):
        arrow_io.columns_from_record_batch(pa.RecordBatch.from_arrays([pa.array(["a"])], names=["0_1"]))


def test_oracle_runs_on_decoded_batches():
    cols = datagen.c2_table(0, 20_000, n_groups=30)
    plan = queries.c2_filter_groupby()
    want = oracle.execute(plan.serialize(), cols)
    got = oracle.execute(plan.serialize(), arrow_io.decode(*arrow_io.encode(cols), plan.tuples))
    assert_same_rows(got.columns, want.columns, ["0_1"], rel=0.0)


@pytest.mark.gpu
def test_regions_answer_over_ipc_and_the_db_side_merges_them():
    n, regions, groups = 240_000, 3, 200
    frag, merge = queries.c2_filter_groupby(), queries.c2_filter_groupby(merge=True)
    step = n // regions
    answers = []
    for r in range(regions):
        s, b = arrow_io.encode(datagen.c2_table(r * step, step, n_groups=groups))           # the scan's batch on the wire
        rs, rb, stats = arrow_io.execute_ipc(frag, s, b)                                     # store side: fragment on the GPU
        assert stats.rows_scanned == step
        answers.append(arrow_io.decode(rs, rb, merge.tuples))
    merged_in = [make_column(c[0].tuple_id, c[0].slot_id, c[0].prim_type, np.concatenate([x.values for x in c]),
                             None if all(x.valid is None for x in c) else np.concatenate([x.valid if x.valid is not None else np.ones(len(x), bool) for x in c]))
                 for c in zip(*answers)]
    ms, mb, _ = arrow_io.execute_ipc(merge, *arrow_io.encode(merged_in))                    # db side: MERGE_AGG_NODE on the GPU
    got = arrow_io.decode(ms, mb, merge.tuples)
    want = oracle.execute(frag.serialize(), datagen.c2_table(0, n, n_groups=groups))
    assert_same_rows(got, want.columns, ["0_1"], rel=1e-9, abs_tol=1e-9)
