"""Document-sharded path on real GPUs: launches tests/multigpu_check.py under torchrun on 2 GPUs (and on 4 when the box
has them: that adds the hybrid shards x replicas groups).  Skips below 2 GPUs.  The CPU side of the sharded logic is
covered by tests/test_sharded_gloo.py (gloo, world 2)."""
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _n_gpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 4])
def test_sharded_step_on_gpus(world):
    if _n_gpus() < world:
        pytest.skip("needs %d GPUs" % world)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "multigpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + "\n" + r.stderr[-6000:]
    assert "multigpu_check ok" in r.stdout
