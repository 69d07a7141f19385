"""GPU, >= 2 devices: the real multi-process path (one rank per GPU, NCCL halo exchange over
NVLink) must equal the single-domain run bit for bit.  Skipped on 1-GPU boxes, where
tests/test_slab_gpu.py covers the same kernels with virtual slabs."""
import os
import socket
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]


def test_nccl_slabs_equal_single_domain():
    import torch

    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 8 if n >= 8 else (4 if n >= 4 else 2)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), str(ROOT / "tools" / "mgpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT, env=dict(os.environ))
    assert r.returncode == 0 and "[mgpu] ALL OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
