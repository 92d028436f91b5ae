"""Multi-GPU parity (needs >= 2 GPUs on the box; skipped otherwise): tools/dist_check.py under
torchrun — distributed boxblur through the ABI, row-sharded heat2d with NCCL halo exchange,
scatter+all-gather broadcast, row-sharded sgemm, and the peer-memory halo path (heat2d / fp32 blur
reading the neighbours' rows in place over NVLink, 12 even/odd Jacobi sweeps of heat2d_dist with the
neighbour handshake inside the sweep kernels) — all bit-exact against the single-process oracle."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_sharded_paths_match_oracle_on_two_or_more_gpus():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2 if n < 4 else 4
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", "29733", os.path.join(ROOT, "tools", "dist_check.py")],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0 and "DIST_CHECK_OK" in r.stdout, (r.stdout + r.stderr)[-4000:]
