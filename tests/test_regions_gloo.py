"""CPU, world_size 2 over gloo: the host-side decomposition the multi-GPU path relies on — regions -> per-rank
partial aggregates -> all_gather -> MERGE_AGG (AggFnCall::merge, src/expr/agg_fn_call.cpp:719-822) equals the
single-region result; per-rank top-k -> all_gather -> final top-k with (region, row) tie order equals the global
top-k (SelectManagerNode merge of sorted runs, select_manager_node.cpp:50-51); hash repartition of the partial rows (each
rank merges the groups it owns, exchange_sender_node.cpp:867-957) yields disjoint partitions whose union is the same
result.  Uses the oracle on each rank."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
sys.path.insert(0, os.environ["BK_ROOT"])
import numpy as np
import torch.distributed as dist
from baikaldb_b200 import datagen, queries
from baikaldb_b200.column import make_column, rows_as_set
from oracle import oracle

dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
n = 40_000
region = datagen.c2_table(rank * n, n, n_groups=200)
part = oracle.execute(queries.c2_filter_groupby().serialize(), region)          # store-side AGG_NODE of this region
payload = [(c.tuple_id, c.slot_id, c.prim_type, c.values, c.valid) for c in part.columns]
gathered = [None] * world
dist.all_gather_object(gathered, payload)
cols = []
for i in range(len(payload)):                                                    # rows of all regions, as the db receives them
    t, s, pt = payload[i][:3]
    vals = np.concatenate([g[i][3] for g in gathered])
    valid = None if all(g[i][4] is None for g in gathered) else np.concatenate([g[i][4] if g[i][4] is not None else np.ones(len(g[i][3]), bool) for g in gathered])
    cols.append(make_column(t, s, pt, vals, valid))
merged = oracle.execute(queries.c2_filter_groupby(merge=True).serialize(), cols)  # db-side MERGE_AGG_NODE
whole = oracle.execute(queries.c2_filter_groupby().serialize(), datagen.c2_table(0, n * world, n_groups=200))
m, w = rows_as_set(merged.columns, ["0_1"]), rows_as_set(whole.columns, ["0_1"])
mn, wn = [c.name for c in merged.columns], [c.name for c in whole.columns]
assert set(m) == set(w)
for k in m:
    assert m[k][mn.index("1_1")] == w[k][wn.index("1_1")]
    assert abs(m[k][mn.index("1_2")] - w[k][wn.index("1_2")]) <= 1e-9 * abs(w[k][wn.index("1_2")])
    assert abs(m[k][mn.index("1_3")] - w[k][wn.index("1_3")]) <= 1e-9 * abs(w[k][wn.index("1_3")]) + 1e-12
# f3 hash repartition: rank r merges only the partial rows of the groups it owns; the union over ranks is the answer
keycol = [i for i in range(len(payload)) if (payload[i][0], payload[i][1]) == (0, 1)][0]
own = (cols[keycol].values.astype(np.int64) % world) == rank
mine_rows = [make_column(c.tuple_id, c.slot_id, c.prim_type, c.values[own], None if c.valid is None else c.valid[own]) for c in cols]
part = oracle.execute(queries.c2_filter_groupby(merge=True).serialize(), mine_rows)
parts = [None] * world
dist.all_gather_object(parts, [(c.name, c.to_list()) for c in part.columns])
union = {}
for pr in parts:
    d = dict(pr)
    for i, k in enumerate(d["0_1"]):
        assert k not in union, "a group came back from two ranks"
        union[k] = (d["1_1"][i], d["1_2"][i], d["1_3"][i])
assert set(union) == {k[0] for k in w}
for k in w:
    assert union[k[0]][0] == w[k][wn.index("1_1")] and abs(union[k[0]][1] - w[k][wn.index("1_2")]) <= 1e-9 * abs(w[k][wn.index("1_2")])
# top-k
rng = np.random.default_rng(3)
keys = rng.integers(0, 300, n * world)
mine = [make_column(0, 1, 6, keys[rank * n:(rank + 1) * n]), make_column(0, 2, 5, np.arange(rank * n, (rank + 1) * n, dtype=np.int32))]
top = oracle.execute(queries.c5_topk(500).serialize(), mine)
runs = [None] * world
dist.all_gather_object(runs, [c.values for c in top.columns])
allk = np.concatenate([r[0] for r in runs]); allp = np.concatenate([r[1] for r in runs])
order = np.lexsort((allp, allk))[:500]                                           # (key, (region,row)) order
wholek = oracle.execute(queries.c5_topk(500).serialize(), [make_column(0, 1, 6, keys), make_column(0, 2, 5, np.arange(n * world, dtype=np.int32))])
assert allk[order].tolist() == wholek.columns[0].to_list() and allp[order].tolist() == wholek.columns[1].to_list()
if rank == 0:
    print("GLOO_OK")
'''


def test_region_decomposition_world2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, BK_ROOT=ROOT)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29611", str(script)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)
    assert r.returncode == 0 and "GLOO_OK" in r.stdout, r.stdout[-1500:] + r.stderr[-3000:]
