"""Multi-region parity (configs C4 / C5): one process per GPU over NCCL when the box has several GPUs, and — on any box — `world`
ranks sharing cuda:0 that move the per-region partial state over gloo (the same export / merge kernels, no NCCL)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpu_count():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("world", [2])
def test_regions_across_gpus(world):
    if _gpu_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29617", os.path.join(ROOT, "tests", "mgpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0 and f"MGPU_OK {world}" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]


@pytest.mark.parametrize("world", [2, 3])
def test_regions_on_one_gpu_over_gloo(world):
    """runs on the driver's single-GPU box: k_partial_export_rows / k_partial_merge_rows / the top-k partial merge against the oracle"""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29630 + world), os.path.join(ROOT, "tests", "onegpu_regions_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0 and f"ONEGPU_REGIONS_OK {world}" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]
