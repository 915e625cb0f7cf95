"""GPU, multi-device: the parity suite's core check on EVERY rank of a torchrun launch (one process per GPU,
NCCL): each rank's solver + projection result, gathered to rank 0 over NVLink, equals the CPU oracle bit for
bit. Needs >= 2 GPUs; skipped on a single-GPU box."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu


def test_every_rank_matches_the_oracle():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs (torchrun, one rank per GPU)")
    n = min(n, 8)
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    here = os.path.dirname(os.path.abspath(__file__))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n),
                          "--master-addr", "127.0.0.1", "--master-port", str(port),
                          os.path.join(here, "mgpu_worker.py")], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "MGPU_OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
    for r in range(n):
        assert "rank %d: x ok, dense map ok" % r in out.stdout
