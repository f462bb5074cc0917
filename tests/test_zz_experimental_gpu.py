"""Opt-in code paths (off by default): GSB_FUSED_END=1, the end of sweep fused into the last colour phase."""
import ast
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _solve(env, size, channels, sweeps, kernel):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "x_hash.py"), str(size), str(channels), str(sweeps),
                        str(kernel)], capture_output=True, text=True, timeout=600, env=dict(os.environ, **env), cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    line = r.stdout.strip().splitlines()[-1]
    digest, kernel_used, done, _, eps = line.split(" ", 4)
    return digest, int(kernel_used), int(done), ast.literal_eval(eps)


@pytest.mark.parametrize("size,channels,kernel", [(1500, 3, 3), (1500, 1, 4), (200, 3, 3)])
def test_fused_end_of_sweep_equals_separate_kernel(gsb, size, channels, kernel):
    """Same solution bits, same sweep count; the stop norm agrees to rounding (256-thread instead of 1024-thread
    fold).  Kernel 3 passed on B200 in round 1; the window kernel (4) did not, so gsb_plan_can_fuse_end excludes it
    and GSB_FUSED_END=1 must leave it on the separate end-of-sweep kernel (same bits by construction)."""
    base = _solve({"GSB_FUSED_END": "0"}, size, channels, 9, kernel)
    fused = _solve({"GSB_FUSED_END": "1"}, size, channels, 9, kernel)
    assert base[1] == fused[1] == kernel and base[2] == fused[2] == 9
    assert base[0] == fused[0], "solution differs with the fused end of sweep"
    for a, b in zip(base[3], fused[3]):
        assert abs(a - b) <= 1e-12 * abs(a)
