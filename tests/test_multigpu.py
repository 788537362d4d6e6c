"""world_size > 1: host-side partition / gather logic on CPU with gloo (runs everywhere), and
the same through CUDA + NCCL when the box has at least two GPUs."""
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


def _launch(mode, nproc):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tests", "_dist_worker.py"), mode]
    p = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    for r in range(nproc):
        assert f"rank {r} ok" in p.stdout


@pytest.mark.parametrize("world", [2, 3])
def test_strip_gather_and_pose_blocks_gloo(world):
    _launch("gloo", world)


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 3])
def test_strip_group_over_cuda_ipc_on_one_gpu(world):
    """The sort-first exchange step — rank 0's framebuffers shared over CUDA IPC, every other rank pushing its busy tiles
    into them, device-side hand-off flags, two buffers in flight, rank 0 mirroring each frame into host memory — with all
    ranks on GPU 0, so that a one-GPU box runs it too (frames compared with the oracle, TPFs summed over ranks)."""
    _launch("ipc1", world)


@pytest.mark.gpu
def test_strips_and_frame_parallel_nccl():
    import torch

    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least two GPUs")
    _launch("nccl", min(n, 4))
