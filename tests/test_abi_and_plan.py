"""CPU-side checks of the product: the C-ABI library loads and exports every symbol include/bkgpu.h declares,
the plan word stream round-trips through the host-side lowering (type inference mirrors the reference), and
the library refuses to run without a GPU instead of falling back."""
import ctypes
import os
import re

import pytest

from baikaldb_b200 import _lib, plan as P, queries
from baikaldb_b200.plan import PrimitiveType as T

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "bkgpu.h")).read()
    declared = set(re.findall(r"\b(bkgpu_[a-z0-9_]+)\s*\(", header))
    declared -= {"bkgpu_plan", "bkgpu_column", "bkgpu_stats"}
    L = ctypes.CDLL(_lib.LIB_PATH)
    missing = [name for name in sorted(declared) if not hasattr(L, name)]
    assert not missing, f"libbkgpu.so lacks {missing}"
    bound = {name for name, _, _ in _lib.SYMBOLS}
    assert declared == bound, (sorted(declared - bound), sorted(bound - declared))


def test_version_and_no_device_is_an_error_not_a_fallback():
    L = _lib.lib()
    assert L.bkgpu_version().startswith(b"bkgpu")

    n = L.bkgpu_device_count()
    if n > 0:
        pytest.skip("a GPU is visible here")
    h = ctypes.c_void_p()
    pb = queries.c1_count_where().serialize()
    rc = L.bkgpu_init(ctypes.byref(h), pb, len(pb), 0, None)
    assert rc == _lib.ENODEV
    assert b"no CPU fallback" in L.bkgpu_last_error(None)


def test_library_was_built_from_the_sources_in_this_tree():
    """provenance: libbkgpu.so is a git-ignored binary that travels to the GPU box; it reports a digest of csrc/ + include/ taken when it was
    compiled (csrc/Makefile: build_id.h), which must equal the digest of the sources that are here now"""
    import glob
    import hashlib
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cs = os.path.join(root, "csrc")
    mk = open(os.path.join(cs, "Makefile")).read()
    names = []
    for var in ("CU", "CPP"):
        line = [l for l in mk.splitlines() if l.startswith(var + " :=")][0]
        names += line.split(":=")[1].split()
    names += [os.path.basename(p) for pat in ("*.h", "*.cuh", "*.inc") for p in glob.glob(os.path.join(cs, pat))]
    names += ["../include/bkgpu.h", "../include/bkgpu_plan.h"]
    h = hashlib.sha256()
    for n in sorted(set(names)):
        h.update(open(os.path.join(cs, n), "rb").read())
    version = _lib.lib().bkgpu_version().decode()
    assert version.endswith("src=" + h.hexdigest()[:16]), (version, h.hexdigest()[:16])


def test_explain_c2_lowering():
    text = _lib.explain(queries.c2_filter_groupby().serialize())
    assert "n_group=1 n_keyw=1 n_agg=3" in text and "direct=1" in text
    # lt_int_int on an INT32 column: compared in INT64 with the literal 2^19 (fn_manager.cpp:316-333)
    assert "direct term[0]: col=0 cmp=19 class=0 const=0x80000" in text
    # SUM(double) and AVG(double) accumulate in double lanes (class 2), COUNT(*) rides on lane 0
    assert "agg[1] kind=2 class=2" in text and "agg[2] kind=3 class=2" in text


def test_type_inference_matches_reference_rules():
    s_i32, s_u32, s_dbl = P.slot_ref(0, 1, T.INT32), P.slot_ref(0, 2, T.UINT32), P.slot_ref(0, 3, T.DOUBLE)
    aggs = [P.agg_expr("count_star", 1, 1)]
    tuples = {0: [(1, T.INT32), (2, T.UINT32), (3, T.DOUBLE)], 1: P.agg_tuple_slots(aggs, [T.INT64])}

    def lower(conj):
        return _lib.explain(P.Plan(P.agg(P.where(P.scan(0), conj), 1, [], aggs), tuples).serialize())
    # int vs unsigned -> UINT64 compare (class 1): any unsigned operand promotes (fn_manager.cpp:325-329)
    assert "CMP       a=19 b=1" in lower(P.lt(s_i32, s_u32))
    # int column vs double column -> DOUBLE compare (class 2) with a cast of the int side
    t = lower(P.lt(s_i32, s_dbl))
    assert "CMP       a=19 b=2" in t and "CAST      a=5 b=12" in t
    # `int32_col < 0.5`: the literal takes the column's type first (scalar_fn_call.cpp:57-67) -> INT64 compare with 0
    t = lower(P.lt(s_i32, P.double_lit(0.5)))
    assert "direct term[0]: col=0 cmp=19 class=0 const=0x0" in t
    # division is always DOUBLE (fn_manager.cpp:357-359)
    t = lower(P.gt(P.divides(s_i32, s_u32), P.int_lit(1)))
    assert "DIV_F64" in t


def test_shared_lanes_for_sum_and_avg_of_same_column():
    s = P.slot_ref(0, 1, T.DOUBLE)
    aggs = [P.agg_expr("sum", 1, 1, None, s), P.agg_expr("avg", 1, 2, 3, s), P.agg_expr("count", 1, 4, None, s)]
    pl = P.Plan(P.agg(P.scan(0), 1, [], aggs), {0: [(1, T.DOUBLE)], 1: P.agg_tuple_slots(aggs, [T.DOUBLE, T.DOUBLE, T.DOUBLE])})
    text = _lib.explain(pl.serialize())
    assert "n_lanes=3" in text  # row count + one shared non-NULL counter + one shared double sum


@pytest.mark.parametrize("mutate,code", [
    (lambda b: b[:40], _lib.EINVAL),                           # truncated
    (lambda b: b"\0\0\0\0" + b[4:], _lib.EINVAL),              # bad magic
    (lambda b: b + b"\1\0\0\0", _lib.EINVAL),                  # trailing words
])
def test_malformed_plans_are_rejected(mutate, code):
    pb = mutate(queries.c2_filter_groupby().serialize())
    buf = ctypes.create_string_buffer(512)
    assert _lib.lib().bkgpu_plan_explain(pb, len(pb), buf, 512) == code


def test_out_of_scope_shapes_report_unsupported():
    aggs = [P.agg_expr("group_concat", 1, 1, None, P.slot_ref(0, 1, T.INT32))]
    pl = P.Plan(P.agg(P.scan(0), 1, [], aggs), {0: [(1, T.INT32)], 1: [(1, int(T.STRING))]})
    with pytest.raises(_lib.BkgpuError) as e:
        _lib.explain(pl.serialize())
    assert e.value.code == _lib.EUNSUPPORTED
    aggs = [P.agg_expr("count_star", 1, 1)]
    pl = P.Plan(P.agg(P.scan(0), 1, [P.slot_ref(0, 1, T.STRING)], aggs), {0: [(1, int(T.STRING))], 1: [(1, int(T.INT64))]})
    with pytest.raises(_lib.BkgpuError) as e:
        _lib.explain(pl.serialize())
    assert e.value.code == _lib.EUNSUPPORTED


def test_datagen_permutation_and_reproducibility():
    import numpy as np
    from baikaldb_b200 import datagen
    a = datagen.permutation(3, 11, 0, 10_000, 10_000)
    assert sorted(a.tolist()) == list(range(10_000))
    b = np.concatenate([datagen.permutation(3, 11, 0, 4000, 10_000), datagen.permutation(3, 11, 4000, 6000, 10_000)])
    assert np.array_equal(a, b)                                 # any region is reproducible on its own
    x = datagen.c2_table(1000, 500)
    y = datagen.c2_table(0, 1500)
    for cx, cy in zip(x, y):
        assert np.array_equal(cx.values, cy.values[1000:])


def test_product_code_never_touches_the_oracle():
    """oracle/ is test infrastructure: nothing under baikaldb_b200/, csrc/, host/ or include/ may import, link or execute it
    (bench.py may, for its cpu_baseline / --impl reference legs only)."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    bad = []
    for sub in ("baikaldb_b200", "csrc", "host", "include"):
        for dirpath, _, files in os.walk(os.path.join(root, sub)):
            for f in files:
                if not f.endswith((".py", ".cpp", ".cu", ".cuh", ".h", ".hpp", "Makefile")):
                    continue
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                if re.search(r"^\s*(from|import)\s+oracle\b|libbk_oracle|bko_execute|oracle/bk_oracle|acero_oracle", text, re.M):
                    bad.append(os.path.join(sub, f))
    assert not bad, bad
