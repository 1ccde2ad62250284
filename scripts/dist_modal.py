"""Row-block distributed modal solve under torchrun: `torchrun --nproc-per-node N scripts/dist_modal.py nx ny nz [k] [--check]`."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from fem_calculator_b200 import _lib as L, meshgen, compat
from fem_calculator_b200.api import DistFrameModel
from fem_calculator_b200.sections import calculate_section_properties as csp

args = [a for a in sys.argv[1:] if not a.startswith("--")]
nx, ny, nz = [int(v) for v in args[:3]]
k = int(args[3]) if len(args) > 3 else 20
check = "--check" in sys.argv
precond = L.PRECOND_JACOBI if "--jacobi" in sys.argv else L.PRECOND_AUTO
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
mesh, sec, bc = meshgen.lattice_frame_case(nx, ny, nz, jitter=0.05)
es, props, _ = compat.frame_section_table(mesh, sec, csp)
fixed, f = compat.frame_bc_vectors(mesh, bc, len(mesh.points))
E, nu = meshgen.E_STEEL, meshgen.NU_STEEL
uid = None
if world > 1:
    t = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        t = torch.tensor(list(DistFrameModel.unique_id()), dtype=torch.uint8, device="cuda")
    dist.broadcast(t, 0)
    uid = bytes(t.cpu().numpy().tolist())
def _gather(obj):
    out = [None] * world
    dist.all_gather_object(out, obj)
    return out


use_p2p = world > 1 and not os.environ.get("FEMB_DIST_NO_P2P")
m = DistFrameModel(local)
part = m.setup(mesh.points, mesh.cells_dict["line"], es, props, E, E / (2 * (1 + nu)), fixed, f, rank, world, uid, all_gather=_gather if use_p2p else None)
t0 = time.time()
lam, phi, st = m.modal_dist(k=k, precond=precond) if world > 1 else m.modal(k=k)
wall = time.time() - t0
if rank == 0:
    print(json.dumps({"lattice": [nx, ny, nz], "ndof": len(f), "world": world, "k": k, "modes": len(lam), "device_ms": st["device_ms"], "wall_s": wall,
                      "pcg_iterations": st["iterations"], "rel_residual": st["rel_residual"], "precond_used": st.get("precond_used"),
                      "omega": [float(x) for x in np.sqrt(lam[:4])]}), flush=True)
if check:
    from oracle import ref_sparse as S
    if world > 1:
        parts = [None] * world
        dist.all_gather_object(parts, (part.owned_nodes, phi))
        phig = np.zeros((len(f), phi.shape[1]))
        for nodes, ph in parts:                        # owned rows of every rank, in its ascending global node order
            phig.reshape(-1, 6, phi.shape[1])[nodes] = ph.reshape(-1, 6, phi.shape[1])
    else:
        phig = phi
    if rank == 0:
        K, M = S.frame_assemble(mesh.points, mesh.cells_dict["line"], es, props, E, nu)
        free = np.setdiff1d(np.arange(len(f)), fixed)
        lam_ref, phi_ref = S.frame_modal(K, M, free, k=k + 4)
        rel = np.abs(lam - lam_ref[:k]) / lam_ref[:k]
        g = phig.T @ (M @ phig)
        print(f"check: max rel eigenvalue error {rel.max():.2e}; max |Phi^T M Phi - I| {np.abs(g - np.eye(len(lam))).max():.2e}", flush=True)
        assert rel.max() <= 1e-8 and np.abs(g - np.eye(len(lam))).max() <= 1e-8
m.close()
if world > 1:
    dist.destroy_process_group()
