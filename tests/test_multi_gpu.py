"""GPU, >= 2 devices: the product's own multi-GPU helpers (kadal_b200.parallel / reliability / sensitivity) over NCCL
against the same calls on one rank -- bit-equal likelihoods and predictions, identical arg-min, identical AK-MCS
scalars, identical Saltelli sums (tools/check_parallel_nccl.py under torchrun).  Skips itself on a one-GPU box; the
sharding logic itself is covered on CPU by tests/test_parallel_gloo.py (gloo, world size 2)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_product_api_over_nccl_matches_single_gpu():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs (have %d)" % n)
    world = min(n, 4)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(29600 + os.getpid() % 300),
           os.path.join(ROOT, "tools", "check_parallel_nccl.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]      # every rank asserts for itself; torchrun fails if one does
    assert " ok: argmin" in r.stdout
