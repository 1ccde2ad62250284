"""Multi-rank test of the row-block distributed solve (csrc/dist.cu + csrc/lines.cu): needs >= 2 GPUs on the box
(skipped otherwise).  Launches scripts/dist_solve.py under torchrun with 2 ranks — peer-memory halo / scalar /
coarse-residual exchange, line preconditioner on the partition — and lets it compare the gathered u with the CPU
oracle's direct solve (||u - u_ref|| <= 1e-10 ||u_ref||)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    from fem_calculator_b200 import _lib
    return int(_lib.load().femb_device_count())


def _run(world, lattice, precond, port, partition="boxes"):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "scripts", "dist_solve.py"), *map(str, lattice), "--check",
           f"--precond={precond}"] + (["--slabs"] if partition == "slabs" else [])
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1]
    assert "check: ||u - u_oracle||" in r.stdout
    return json.loads(line)


@pytest.mark.parametrize("precond,partition", [("lines", "boxes"), ("lines", "slabs"), ("jacobi", "boxes")])
def test_two_rank_solve_matches_oracle(precond, partition):
    if _n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    out = _run(2, (14, 10, 9), precond, {"lines": 29611, "jacobi": 29612}[precond] + (10 if partition == "slabs" else 0), partition)
    assert out["world"] == 2 and out["exchange"] == "p2p" and out["partition"] == partition
    if precond == "lines":
        assert out["precond_used"] == 5 and out["coarse_dim"] > 0, out
        assert out["iterations"] < 300, out


def test_four_rank_box_partition_matches_oracle():
    """2 x 2 boxes: every rank has two neighbours that are not adjacent in rank order, ghosts grouped by owner."""
    if _n_gpus() < 4:
        pytest.skip("needs 4 GPUs")
    out = _run(4, (14, 12, 9), "lines", 29631)
    assert out["world"] == 4 and out["precond_used"] == 5, out


def test_two_rank_modal_on_the_line_preconditioner_matches_oracle():
    """femb_modal on a 2-rank box partition with the line-preconditioned persistent PCG behind K^-1: eigenvalues
    against the oracle's generalized pencil (1e-8), M-orthonormal shapes (scripts/dist_modal.py --check)."""
    if _n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29641", os.path.join(ROOT, "scripts", "dist_modal.py"), "12", "9", "8", "6", "--check"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1]
    out = json.loads(line)
    assert "check: max rel eigenvalue error" in r.stdout
    assert out["modes"] == 6 and out["precond_used"] == 5, out
