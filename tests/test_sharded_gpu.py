"""Frame-sharded multi-GPU forward == single-GPU forward (needs >= 2 GPUs; spawns torchrun)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_frame_sharded_equals_single_gpu():
    n = 8 if torch.cuda.device_count() >= 8 else (4 if torch.cuda.device_count() >= 4 else 2)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", "29541", os.path.join(ROOT, "tests", "dist_sharded_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=400, cwd=ROOT)
    print(r.stdout[-3000:], r.stderr[-3000:])
    assert r.returncode == 0
