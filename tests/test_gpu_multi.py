"""Multi-GPU test (NCCL, one process per GPU): skipped on boxes with a single GPU.  The CPU-side logic of the same path is
covered by the world-size-2 gloo tests in tests/test_abi_and_host.py."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_set_trainer_nccl_matches_large_batch_step():
    import __graft_entry__ as g
    g.build()
    env = dict(os.environ, NCCL_DEBUG="WARN")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29611", os.path.join(ROOT, "tests", "_ddp_worker.py")], capture_output=True, text=True, timeout=600,
                       env=env, cwd=ROOT)
    assert r.returncode == 0 and "DDP_OK 2" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
