"""Multi-GPU parity of the sharded extractor: needs >= 2 GPUs on the box (`gpurun --gpus 2`), skipped on a
one-GPU box.  The N-rank logic itself is also covered on CPU (gloo, world size 2) in tests/test_host_logic.py."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


@pytest.mark.parametrize("world", [2])
def test_sharded_gather_equals_single_gpu(world, built_lib):
    import torch

    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs, box has {torch.cuda.device_count()}")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tests", "dist_gather_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0 and f"DIST_GATHER_OK {world}" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
