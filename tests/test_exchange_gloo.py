"""CPU, world_size 2 over gloo: the hash exchange of baikaldb_b200/exchange.py (ExchangeSender / Receiver, exchange_sender_node.cpp:867-957)
and the two decompositions of SURVEY.md §8 f3 that need it — a REPARTITIONED JOIN (both inputs partitioned on the join key, every rank runs
the ordinary AGG -> JOIN fragment over what it received, the partial aggregates merge as region results do) and COUNT / SUM / AVG (DISTINCT x)
across regions (the store-side GROUP BY (k, x) rows partitioned on k, duplicates of (k, x) merged, then the MERGE_AGG with its *_distinct
functions) — each equal to the single-node result.  The oracle stands in for the per-rank fragment (no GPU here); the exchange code is the
same torch code that moves CUDA tensors over NCCL."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
sys.path.insert(0, os.environ["BK_ROOT"])
import numpy as np
import torch
import torch.distributed as dist
from baikaldb_b200 import exchange as ex, plan as P, queries
from baikaldb_b200.column import make_column, rows_as_set
from baikaldb_b200.plan import PrimitiveType as T
from oracle import oracle
from tests.test_gpu_merge import _distinct_case

dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()

def gather_columns(cols):
    payload = [(c.tuple_id, c.slot_id, c.prim_type, c.values, c.valid) for c in cols]
    got = [None] * world
    dist.all_gather_object(got, payload)
    out = []
    for i in range(len(payload)):
        t, s, pt = payload[i][:3]
        vals = np.concatenate([g[i][3] for g in got])
        valid = None if all(g[i][4] is None for g in got) else np.concatenate([g[i][4] if g[i][4] is not None else np.ones(len(g[i][3]), bool) for g in got])
        out.append(make_column(t, s, pt, vals, valid))
    return out

def shard(cols):   # this rank's region: every world-th row
    return [make_column(c.tuple_id, c.slot_id, c.prim_type, c.values[rank::world], None if c.valid is None else c.valid[rank::world]) for c in cols]

def same(a, b, keys, tol=1e-9):
    ra, rb = rows_as_set(a, keys), rows_as_set(b, keys)
    assert set(ra) == set(rb), (sorted(set(ra) ^ set(rb))[:5])
    for k in ra:
        for x, y in zip(ra[k], rb[k]):
            if isinstance(x, float): assert abs(x - y) <= tol * max(1.0, abs(y)), (k, x, y)
            elif isinstance(x, bytes): assert x[8:] == y[8:]      # AVG blob: the int64 count half
            else: assert x == y, (k, x, y)

# ---- 1. the exchange itself: rows keep their columns together, equal keys meet on one rank, NULL keys meet on one rank ----
rng = np.random.default_rng(11)
n = 30_000
k_all = rng.integers(-1000, 1000, n); v_all = rng.normal(size=n); ok_all = rng.random(n) > 0.1
u_all = rng.integers(0, 1 << 63, n, dtype=np.uint64) * 2 + 1
mine = shard([make_column(0, 1, T.INT64, k_all, ok_all), make_column(0, 2, T.DOUBLE, v_all), make_column(0, 3, T.UINT64, u_all)])
b = ex.batch_from_columns(mine)
got = ex.columns_from_batch(ex.exchange(b, ex.destination(b, [(0, 1)], world)))
everything = gather_columns(got)
order = lambda cols: np.lexsort((cols[2].values, cols[1].values))
eo, ao = order(everything), order([make_column(0, 1, T.INT64, k_all, ok_all), make_column(0, 2, T.DOUBLE, v_all), make_column(0, 3, T.UINT64, u_all)])
assert np.array_equal(everything[1].values[eo], v_all[ao]) and np.array_equal(everything[2].values[eo], u_all[ao])          # nothing lost, nothing doubled
assert np.array_equal(np.where(everything[0].valid[eo], everything[0].values[eo], 0), np.where(ok_all[ao], k_all[ao], 0))   # rows stayed whole
keys_here = set(got[0].values[got[0].valid if got[0].valid is not None else slice(None)].tolist())
sets = [None] * world
dist.all_gather_object(sets, (keys_here, 0 if got[0].valid is None else int((~got[0].valid).sum())))
for i in range(world):
    for j in range(i + 1, world):
        assert not (sets[i][0] & sets[j][0]), "a key value landed on two ranks"
assert sum(1 for s in sets if s[1] > 0) == 1, "NULL keys must meet on one rank"

# ---- 2. repartitioned join: SELECT d.attr, COUNT(*), SUM(f.v) FROM fact f JOIN dim d ON f.k = d.k GROUP BY d.attr ----
rng = np.random.default_rng(7)
nd, nf = 5_000, 60_000
dim_all = [make_column(1, 1, T.INT32, rng.permutation(nd)), make_column(1, 2, T.INT32, rng.integers(0, 50, nd))]
fact_all = [make_column(0, 1, T.INT32, rng.integers(0, nd + 500, nf)), make_column(0, 2, T.DOUBLE, rng.normal(size=nf) * 10)]
plan = queries.c3_join_groupby()
db, fb = ex.batch_from_columns(shard(dim_all)), ex.batch_from_columns(shard(fact_all))
dim_mine = ex.columns_from_batch(ex.exchange(db, ex.destination(db, [(1, 1)], world)))
fact_mine = ex.columns_from_batch(ex.exchange(fb, ex.destination(fb, [(0, 1)], world)))
part = oracle.execute(plan.serialize(), fact_mine + dim_mine)                       # per rank: the ordinary AGG -> JOIN fragment
aggs = [P.agg_expr("count_star", 2, 1), P.agg_expr("sum", 2, 2, None, P.slot_ref(0, 2, T.DOUBLE))]
merge = P.Plan(P.agg(P.scan(1), 2, [P.slot_ref(1, 2, T.INT32)], aggs, merge=True),
               {0: [(1, T.INT32), (2, T.DOUBLE)], 1: [(1, T.INT32), (2, T.INT32)], 2: P.agg_tuple_slots(aggs, [T.INT64, T.DOUBLE])})
merged = oracle.execute(merge.serialize(), gather_columns(part.columns))           # db side: MERGE_AGG over the ranks' partial rows
whole = oracle.execute(plan.serialize(), fact_all + dim_all)
same(merged.columns, whole.columns, ["1_2"])
assert sum(len(c) for c in fact_mine[:1]) > 0

# ---- 3. COUNT / SUM / AVG (DISTINCT x), SUM(v), COUNT(*) GROUP BY k across regions ----
(k, x, xv, v), cols, low, top = _distinct_case(seed=5, n=40_000)
part = oracle.execute(low.serialize(), shard(cols))                                 # store side: GROUP BY (k, x) of this region
pb = ex.batch_from_columns(part.columns)
recv = ex.columns_from_batch(ex.exchange(pb, ex.destination(pb, [(0, 1)], world)))  # partition exprs = the upper aggregate's GROUP BY: k
low_aggs = [P.agg_expr("sum", 1, 1, None, P.slot_ref(0, 3, T.DOUBLE)), P.agg_expr("count_star", 1, 2)]
dedup = P.Plan(P.agg(P.scan(0), 1, [P.slot_ref(0, 1, T.INT32), P.slot_ref(0, 2, T.INT32)], low_aggs, merge=True),
               {0: [(1, T.INT32), (2, T.INT32), (3, T.DOUBLE)], 1: P.agg_tuple_slots(low_aggs, [T.DOUBLE, T.INT64])})
kx = oracle.execute(dedup.serialize(), recv)                                        # the same (k, x) from two regions becomes one row
mine_top = oracle.execute(top.serialize(), kx.columns)                              # MERGE_AGG with count / sum / avg _distinct
union = gather_columns(mine_top.columns)
assert len(set(union[[c.name for c in union].index("0_1")].values.tolist())) == len(union[0]), "a group came back from two ranks"
single = oracle.execute(top.serialize(), oracle.execute(low.serialize(), cols).columns)
same(union, single.columns, ["0_1"])
if rank == 0:
    print("EXCHANGE_OK")
'''


def test_exchange_repartitioned_join_and_distinct_world2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, BK_ROOT=ROOT)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29631", str(script)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0 and "EXCHANGE_OK" in r.stdout, r.stdout[-1500:] + r.stderr[-3000:]
