"""On-hardware check of the view-parallel path: the gradient buffer after the NCCL all-reduce of an N-rank
step equals the gradient of one process rendering all 24 views.  Needs >= 2 GPUs (skipped on a 1-GPU box; the CPU/gloo
twin of the host logic is tests/test_parallel_gloo.py); bench.py prints the same figure as `verify` at every N > 1."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_allreduced_gradient_equals_single_gpu():
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29541", os.path.join(ROOT, "bench.py"), "--gpus", "2", "--steps", "2", "--warmup", "3", "--no-e2e",
           "--workloads", "config2"]
    out = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads([ln for ln in out.stdout.splitlines() if ln.startswith("{")][-1])
    assert line["n_gpus"] == 2
    assert line["verify"]["grad_relerr_vs_single_gpu"] <= line["verify"]["tolerance"] == 5e-4, line["verify"]
