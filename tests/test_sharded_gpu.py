"""Driver-run multi-GPU parity: spawns tests/check_sharded.py under torchrun on min(8, device_count) GPUs (one process per
GPU, NCCL).  Every rank's cap, digest slice and leaf range — p2p (copy-engine) and NCCL exchange, device and host inputs — must
equal the oracle's, including 2^16 x 135 (BASELINE.json configs[1]) against the committed golden.  Skipped on a 1-GPU box."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _device_count():
    try:
        out = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True, timeout=60).stdout
        return sum(1 for l in out.splitlines() if l.startswith("GPU "))
    except Exception:
        return 0


def _run(world, extra_env=None):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    env = dict(os.environ, **(extra_env or {}))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "check_sharded.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-4000:]
    assert "MISMATCH" not in r.stdout
    return r.stdout


def test_sharded_commit_all_gpus():
    n = _device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 1 << (min(n, 8).bit_length() - 1)
    out = _run(world)
    assert f"shape (16, 135, 3, 4) world {world} p2p: ok" in out and f"shape (16, 135, 3, 4) world {world} coset: ok" in out
    assert f"shape (16, 135, 3, 4) world {world} stream: ok" in out


def test_sharded_commit_two_gpus():
    n = _device_count()
    if n < 4:                       # with exactly 2 GPUs the test above already ran at world 2
        pytest.skip("needs >= 4 GPUs (world 2 is covered by the all-GPU test on a 2-GPU box)")
    _run(2)
