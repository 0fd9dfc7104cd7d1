"""Multi-GPU (NCCL) test of the episode-parallel step: N-rank averaged gradients == mean of the N
single-process gradients, identical on every rank and bitwise stable from run to run (needs >= 2 GPUs;
the CPU counterpart with gloo is tests/test_parallel_gloo.py)."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs at least two GPUs")
def test_nccl_averaged_gradients_equal_the_mean_of_single_rank_gradients():
    world = min(torch.cuda.device_count(), 8)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tests", "nccl_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    rec = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
    assert rec["world"] == world
    # the mean of the single-process gradients is formed in another order than NCCL's: rounding only
    assert rec["worst_rel_err"]["fp32"] < 1e-5 and rec["worst_rel_err"]["tf32"] < 1e-5, rec
