"""GPU suite, multi-rank: launches tests/multigpu_check.py under torchrun with one rank per
GPU when the box has at least two GPUs (the sharded path must be BIT-IDENTICAL to the
single-GPU path; see that script)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("nproc", [2, 4, 8])
def test_sharded_path_bit_identical_to_single_gpu(nproc):
    if _ngpu() < nproc:
        pytest.skip("needs %d GPUs (have %d)" % (nproc, _ngpu()))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc),
           "--master-addr", "127.0.0.1", "--master-port", str(29530 + nproc),
           os.path.join(ROOT, "tests", "multigpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "ALL BIT-IDENTICAL" in r.stdout
