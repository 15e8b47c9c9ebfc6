"""Randomised fragments (tests/fuzz_plans.py): every generated plan must lower (host-only check) and run on the oracle; on the GPU the
device bytecode / kernels must return the oracle's rows — integers, keys, NULLs exact, doubles within 1e-6."""
import pytest

from baikaldb_b200 import _lib
from oracle import oracle
from tests.fuzz_plans import fragment, table

SEEDS = list(range(48))


@pytest.mark.parametrize("seed", SEEDS)
def test_fuzz_plan_lowers_and_oracle_runs(seed):
    plan, _ = fragment(seed)
    assert _lib.explain(plan.serialize()).startswith("kind=")
    res = oracle.execute(plan.serialize(), table(300, seed))
    assert res.columns is not None       # (a predicate that is never true leaves no group: legitimate)
    for c in res.columns:                # COUNT(*) / COUNT(x) are never NULL
        if (c.tuple_id, c.slot_id) in ((1, 1), (1, 7)):
            assert c.valid is None


@pytest.mark.gpu
@pytest.mark.parametrize("seed", SEEDS)
def test_fuzz_gpu_matches_oracle(seed):
    from tests.util import run_both
    plan, keys = fragment(seed)
    cols = table(3000 + 37 * seed, seed)
    run_both(plan, cols, keys=keys, rel=1e-6, abs_tol=1e-6)
