"""-m gpu, needs two GPUs (skipped otherwise): the ring inside the library over NCCL (sphb_comm_init / sphb_ring_step),
two ranks under torchrun, against the single handle (tools/ring_nccl_check.py)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_rank_ring_over_nccl_matches_the_single_handle():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    env = dict(os.environ, RING_CHECK_NX="128", RING_CHECK_STEPS="9")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", "29531", os.path.join(ROOT, "tools", "ring_nccl_check.py")],
                       capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "OK" in r.stdout
