"""Multi-GPU parity under pytest: launches tests/mg_worker.py with torchrun on 2 GPUs (skipped on a one-GPU box).
The segment + halo partitioning (reference initial `state`, src/filter/fir_node.rs:193-200) and every ordered-gather
path of the C ABI must reproduce the one-GPU stream; see the worker's docstring for the cases."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpu_count():
    try:
        import torch

        return torch.cuda.device_count() if torch.cuda.is_available() else 0
    except Exception:  # noqa: BLE001
        return 0


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2])
def test_segmented_stream_gather_matches_one_gpu(world):
    if _gpu_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29631", os.path.join(ROOT, "tests", "mg_worker.py")]
    p = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    sys.stdout.write(p.stdout[-4000:])
    sys.stderr.write(p.stderr[-4000:])
    assert p.returncode == 0, p.stderr[-2000:]
    assert "mg_worker ok" in p.stdout
