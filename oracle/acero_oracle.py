"""The VECTORIZED half of the reference's CPU path, restated: the Acero plan BaikalDB builds for a
fragment (exec type ``EXEC_ARROW_ACERO``), run on the Arrow C++ engine that ships with pyarrow.

TEST INFRASTRUCTURE ONLY (tests/, smoke(), and the cpu_baseline / ``--impl reference`` legs of
bench.py).  The arithmetic lives in a third-party module absent from /root/reference: Apache Arrow
C++, pinned by the reference to github.com/baikalgroup/arrow tag release-16.1.0
(cmake/arrow.cmake:32-33); here: upstream pyarrow's libarrow_acero (version printed by
``arrow_version()``; skew noted in DESIGN.md).  Declarations follow the reference's call sites:

  FilterNode::build_arrow_declaration   src/exec/filter_node.cpp:581-603   -> "filter"
  AggNode::build_arrow_declaration      src/exec/agg_node.cpp:231-343      -> "project" + "aggregate"
      AggFnCall::transfer_to_arrow_agg_function  src/expr/agg_fn_call.cpp:1365-1392,1459-1528
      (hash_count_all / hash_count / hash_sum / hash_min / hash_max / hash_mean; group-by-nothing is
       rewritten as GROUP BY literal 1)
  SortNode::build_arrow_declaration     src/exec/sort_node.cpp:187-258     -> "order_by" (+ slice = topk)
  JoinNode::build_arrow_declaration     src/exec/join_node.cpp:760-880     -> "hashjoin"
  executor: DeclarationToTable(use_threads = FLAGS vectorlized_parallel_execution, default false)
                                        src/runtime/arrow_io_excutor.cpp:265-292
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np
import pyarrow as pa
import pyarrow.acero as ac
import pyarrow.compute as pc

from baikaldb_b200.column import Column
from baikaldb_b200.plan import PrimitiveType as T

_ARROW_TYPE = {  # src/runtime/chunk.cpp:33-92
    T.BOOL: pa.bool_(), T.INT8: pa.int32(), T.INT16: pa.int32(), T.INT32: pa.int32(), T.TIME: pa.int32(),
    T.INT64: pa.int64(), T.UINT8: pa.uint32(), T.UINT16: pa.uint32(), T.UINT32: pa.uint32(), T.TIMESTAMP: pa.uint32(),
    T.DATE: pa.uint32(), T.UINT64: pa.uint64(), T.DATETIME: pa.uint64(), T.FLOAT: pa.float32(), T.DOUBLE: pa.float64(),
}


def arrow_version() -> str:
    return pa.__version__


def to_table(cols: Sequence[Column]) -> pa.Table:
    """Columns -> Arrow table with the reference's field names "<tuple>_<slot>" (zero-copy for the values)."""
    arrays, names = [], []
    for c in cols:
        mask = None if c.valid is None else ~np.asarray(c.valid, bool)
        vals = c.values.astype(bool) if c.prim_type == T.BOOL else c.values
        arrays.append(pa.array(vals, type=_ARROW_TYPE[T(c.prim_type)], mask=mask))
        names.append(c.name)
    return pa.table(arrays, names=names)


def filter_groupby(table: pa.Table, filter_expr: Optional[pc.Expression], keys: List[str], aggs: List[tuple],
                   use_threads: bool = False) -> pa.Table:
    """aggs: (arrow_function, source_column or None, output_name) — e.g. ("hash_count_all", None, "1_1")."""
    decls = [ac.Declaration("table_source", ac.TableSourceNodeOptions(table))]
    if filter_expr is not None:
        decls.append(ac.Declaration("filter", ac.FilterNodeOptions(filter_expr)))
    group_keys = list(keys)
    if not group_keys:  # group-by-nothing is rewritten as GROUP BY literal 1 (agg_node.cpp:270-296)
        names = table.column_names
        decls.append(ac.Declaration("project", ac.ProjectNodeOptions([pc.field(n) for n in names] + [pc.scalar(1)], names + ["__one"])))
        group_keys = ["__one"]
    agg_specs = []
    for fn, src, out in aggs:
        opts = pc.CountOptions(mode="only_valid") if fn == "hash_count" else None
        agg_specs.append(([] if src is None else src, fn, opts, out))
    decls.append(ac.Declaration("aggregate", ac.AggregateNodeOptions(agg_specs, keys=group_keys)))
    out = ac.Declaration.from_sequence(decls).to_table(use_threads=use_threads)
    if not keys:
        out = out.drop_columns(["__one"])
        if out.num_rows == 0:  # make_default_agg_row_when_no_input (arrow_exec_node.cpp:271-298)
            cols = {name: pa.array([0 if fn in ("hash_count_all", "hash_count") else None], type=out.schema.field(name).type)
                    for fn, _, name in aggs}
            out = pa.table(cols)
    return out


def c1_count_where(table: pa.Table, k: int, use_threads: bool = False) -> pa.Table:
    return filter_groupby(table, pc.field("0_1") < pc.scalar(pa.scalar(k, pa.int32())), [], [("hash_count_all", None, "1_1")], use_threads)


def c2_filter_groupby(table: pa.Table, k: int, use_threads: bool = False) -> pa.Table:
    """SELECT 0_1, COUNT(*), SUM(0_3), AVG(0_4) WHERE 0_2 < k GROUP BY 0_1 as the store-side Acero plan."""
    return filter_groupby(table, pc.field("0_2") < pc.scalar(pa.scalar(k, pa.int32())), ["0_1"],
                          [("hash_count_all", None, "1_1"), ("hash_sum", "0_3", "1_2"), ("hash_mean", "0_4", "1_3")], use_threads)


def c3_join_groupby(fact: pa.Table, dim: pa.Table, use_threads: bool = False) -> pa.Table:
    """hashjoin(probe = fact on 0_1, build = dim on 1_1) -> aggregate GROUP BY 1_2 (join_node.cpp:760-880)."""
    j = ac.Declaration("hashjoin", ac.HashJoinNodeOptions("inner", ["0_1"], ["1_1"]),
                       inputs=[ac.Declaration("table_source", ac.TableSourceNodeOptions(fact)),
                               ac.Declaration("table_source", ac.TableSourceNodeOptions(dim))])
    agg = ac.Declaration("aggregate", ac.AggregateNodeOptions([([], "hash_count_all", None, "2_1"), ("0_2", "hash_sum", None, "2_2")],
                                                              keys=["1_2"]), inputs=[j])
    return agg.to_table(use_threads=use_threads)


def c5_topk(table: pa.Table, k: int, ascending: bool = True, use_threads: bool = False) -> pa.Table:
    """ORDER BY 0_1 LIMIT k: the reference's "topk" node = full SortIndices then slice k (arrow_exec_node.cpp:338-348)."""
    order = "ascending" if ascending else "descending"
    d = ac.Declaration.from_sequence([
        ac.Declaration("table_source", ac.TableSourceNodeOptions(table)),
        ac.Declaration("order_by", ac.OrderByNodeOptions([("0_1", order)], null_placement="at_start" if ascending else "at_end")),
    ])
    return d.to_table(use_threads=use_threads).slice(0, k)


def table_rows(t: pa.Table, key_names: List[str]) -> Dict[tuple, tuple]:
    cols = {n: t.column(n).to_pylist() for n in t.column_names}
    out = {}
    for i in range(t.num_rows):
        out[tuple(cols[k][i] for k in key_names)] = tuple(cols[n][i] for n in t.column_names)
    return out
