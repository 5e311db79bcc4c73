"""Real multi-GPU checks (skipped unless the box has >= 2 GPUs; run with `gpurun --gpus 2`)."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _n_gpus():
    try:
        out = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True, timeout=20).stdout
        return sum(1 for l in out.splitlines() if l.startswith("GPU "))
    except Exception:
        return 0


@pytest.mark.parametrize("exchange", ["nccl", "peer", "hook"])
@pytest.mark.parametrize("mode", ["p2plane", "o3d"])
def test_map_sharded_equals_unsharded(mode, exchange):
    n = _n_gpus()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={min(n, 4)}",
           "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tests", "multigpu", "run_sharded.py"),
           "--mode", mode, "--exchange", exchange]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "-> OK" in r.stdout
