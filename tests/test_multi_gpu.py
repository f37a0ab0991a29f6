"""GPU, >= 2 devices: the sharded registration step and the sharded matcher (NCCL inside libdunk_b200.so) give results
IDENTICAL to the unsharded path, and contexts on two devices of one process work (tests/multi_gpu_worker.py).
Skipped on a single-GPU box; the host-side shard logic is covered on CPU by tests/test_sharded_merge_gloo.py."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def n_gpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_equals_unsharded(world):
    if n_gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    port = 29500 + world + (os.getpid() % 200)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr",
                          "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "multi_gpu_worker.py")],
                         capture_output=True, text=True, timeout=900, cwd=ROOT)
    print(out.stdout[-3000:], out.stderr[-3000:])
    assert out.returncode == 0 and "MULTI_GPU_OK" in out.stdout
