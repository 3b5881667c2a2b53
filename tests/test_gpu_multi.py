"""Multi-GPU parity of record: needs >= 2 GPUs on the box, skipped otherwise.  One process per
GPU (torchrun, NCCL only for the barrier): the sharded run must take the same decisions as the single-GPU run of the
global batch and reproduce its states and gradients; lrnde_allreduce_sum is checked against NCCL."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.skipif(_ngpu() < 2, reason="needs at least 2 GPUs")
@pytest.mark.parametrize("n", [2, 8])
def test_sharded_run_matches_the_single_gpu_run(n):
    if _ngpu() < n:
        pytest.skip(f"needs {n} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", str(29530 + n), os.path.join(ROOT, "tests", "multi_gpu_worker.py")]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    lines = [l for l in r.stdout.splitlines() if "[multi-gpu" in l]
    print("\n".join(lines))
    assert r.returncode == 0, (r.stdout[-3000:], r.stderr[-3000:])
    # (the ranks print at the same time: two reports can share a line)
    assert r.stdout.count("[multi-gpu rank") == 4 * n and r.stdout.count("-> ok") == 4 * n and "MISMATCH" not in r.stdout
