"""Multi-rank tests: the CPU (gloo, world 2 and 3) check of the strip decomposition logic, and the GPU
checks of the strip solver (world 1 on any GPU box; world 2 when two GPUs are visible)."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = os.path.join(ROOT, "tests", "dist_worker.py")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _launch(mode, world, timeout=600):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), WORKER, mode]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env, cwd=ROOT)


@pytest.mark.parametrize("world", [2, 3])
def test_strip_decomposition_gloo(world):
    r = _launch("cpu", world)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "cpu strips ok" in r.stdout


@pytest.mark.gpu
def test_single_rank_strip_equals_single_gpu(gsb):
    r = _launch("gpu", 1)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("gpu strips ok") == 3


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 3, 4, 8])
def test_n_strips_equal_single_gpu(gsb, world):
    """world row strips (one process per GPU, fused peer-memory halo and NCCL halo) == the 1-GPU solve, bit for bit,
    and every rank stops on the same sweep (tests/dist_worker.py, mode gpu)."""
    if gsb._lib.device_count() < world:
        pytest.skip("needs %d GPUs (gpurun --gpus %d)" % (world, world))
    r = _launch("gpu", world, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("gpu strips ok") == 3 and "gpu strips stop rule ok" in r.stdout


@pytest.mark.gpu
def test_strip_solver_rejects_bad_input(gsb):
    from coursecomputationalphotography_b200 import strips
    s = strips.StripSolver(strips.make_unique_id(), 0, 1, 0)
    with pytest.raises(gsb.GsbError):
        s.poisson_strip(16, 16, 4, 4)  # empty strip
    with pytest.raises(gsb.GsbError) as e:
        s.gauss_seidel_dev(1, 1)  # no matrix yet
    assert e.value.status in (1, 10)
    # a matrix whose parity colouring is improper (wrap-around coupling) is refused
    n = 8
    ro = np.arange(0, 2 * n + 1, 2, dtype=np.int32)
    ci = np.stack([np.arange(n), (np.arange(n) + 2) % n], 1).astype(np.int32)
    ci.sort(axis=1)
    with pytest.raises(gsb.GsbError) as e:
        s.matrix_rows(np.ones(2 * n), ro, ci.ravel(), 0, n, 4)
    assert e.value.status == 8
    s.close()
