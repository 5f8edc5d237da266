"""Sharded hot path on >= 2 real GPUs (NCCL): launches tests/dist_worker.py under torchrun.
Skipped on a single-GPU box; the exchange logic itself is covered on CPU by test_parallel_cpu.py."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_sharded_path_on_all_visible_gpus():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "dist_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    first = r.stderr.find("Traceback")          # the first rank to fail, not the launcher's summary of all of them
    assert r.returncode == 0, r.stdout[-1500:] + (r.stderr[first:first + 4000] if first >= 0 else r.stderr[-3000:])
    assert "dist_worker ok" in r.stdout
