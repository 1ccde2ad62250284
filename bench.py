#!/usr/bin/env python
"""bench.py — headline benchmark of the FEM hot path on B200 (contract in the task prompt).

Metric (BASELINE.json): DOF/s of the static solve at ~1M DOF.  One "step" = one pass of the
hot path over the synthetic BASELINE config-3 frame (56x56x54 lattice, 169,344 nodes,
1,016,064 DOF, 498,848 elements, box/C/L sections, 5 % seeded node jitter):
    fused element+assembly  ->  BC mask / RHS  ->  Jacobi-PCG to ||r||/||b|| <= 1e-12 with the
    matrix-free (element-by-element) frame operator  ->  reaction recovery K u - f (assembled K).
`value`  : free DOFs / step, device-timed (CUDA events on the library's stream), mesh and
           loads resident in HBM.  Every step re-assembles the 359 MB matrix (> 126 MB L2), which
           evicts the solver's working set between steps (no extra flush needed).
`e2e`    : the same metric through the reference-shaped entry point
           (compat.BeamAnalysisB200.run_simulation: host numpy arrays in, u / reactions /
           stresses out, symbolic analysis + all H2D/D2H inside the timed region).
`roofline`: the dominant kernel (the operator kernel of the PCG), CUDA-event-timed inside the
           timed steps (every 8th launch, events from a pre-created pool), against
           MEASURED_PEAKS.json; `roofline_bsr_spmv`: the assembled-matrix SpMV the same solve would
           use without the matrix-free operator (one extra solve after the timed region);
           `modal` (N=1): device ms of the 20-mode modal solve of the same frame; `assembly`: the
           fused element+assembly kernel (elements/s).
N > 1 (torchrun): independent load cases of the same frame, one per GPU (weak scaling, no
data-path collective; north_star: "independent load cases ... dealt out one batch per GPU").
`--impl reference`: the CPU oracle port of the reference path on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

LATTICE = (56, 56, 54)
JITTER = 0.05
RTOL = 1e-12
SAMPLE_LATTICE = (40, 40, 38)
SAMPLE_ITERS = 400
FALLBACK_HBM_GBS = 6650.0


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clock / throttle sampling during the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            c = [x.strip() for x in ln.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1]))
                smax.append(float(c[2]))
            except ValueError:
                continue
            for nm, v in zip(names, c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def build_case(load_scale=1.0, lattice=LATTICE):
    from fem_calculator_b200 import compat, meshgen
    from fem_calculator_b200.sections import calculate_section_properties as csp
    mesh, sec, bc = meshgen.lattice_frame_case(*lattice, jitter=JITTER,
                                               load=(100.0 * load_scale, 0.0, -1000.0 * load_scale))
    es, props, _ = compat.frame_section_table(mesh, sec, csp)
    fixed, f = compat.frame_bc_vectors(mesh, bc, len(mesh.points))
    return mesh, sec, bc, es, props, fixed, f


def cpu_baseline_sample(full_n_elem, full_n_free, full_iterations):
    """Oracle port on a bounded sample, scaled to the full workload (oracle/cpu_baseline.py)."""
    from oracle import cpu_baseline as CB
    mesh, sec, bc, es, props, fixed, f = build_case(lattice=SAMPLE_LATTICE)
    from fem_calculator_b200 import meshgen
    s = CB.lattice_static_sample(mesh, es, props, bc, meshgen.E_STEEL, meshgen.NU_STEEL, cg_iters=SAMPLE_ITERS)
    dofs, t_full = CB.scaled_static_dof_per_s(s, full_n_elem, full_n_free, full_iterations)
    return dofs, t_full, s


def sample_text(s, iters):
    return (f"oracle/ref_sparse port: {SAMPLE_LATTICE[0]}x{SAMPLE_LATTICE[1]}x{SAMPLE_LATTICE[2]} lattice slice "
            f"({s['n_elem']} elements, {s['n_free']} free DOF): numpy element formation + scipy COO->CSR "
            f"({s['t_assemble']:.2f} s) + {SAMPLE_ITERS} Jacobi-PCG iterations ({s['t_per_iter']*1e3:.2f} ms each); "
            f"scaled linearly in elements to 498,848 elements and to the {iters} iterations the same "
            f"Jacobi-PCG needs at full size for rtol 1e-12; scipy's sparse mat-vec and the numpy vector updates that "
            f"make up >97 % of this time are single-threaded (cores = 1)")


JACOBI_ITERS_FULL = 6931  # Jacobi-PCG iterations at full size, rtol 1e-12 (profiles/r01_*, measured on B200)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_nodes = LATTICE[0] * LATTICE[1] * LATTICE[2]
    n_free = 6 * (n_nodes - LATTICE[0] * LATTICE[1])
    n_elem = (LATTICE[0] - 1) * LATTICE[1] * LATTICE[2] + LATTICE[0] * (LATTICE[1] - 1) * LATTICE[2] + \
        LATTICE[0] * LATTICE[1] * (LATTICE[2] - 1)
    vals, s = [], None
    for i in range(args.warmup + args.steps):
        dofs, t_full, s = cpu_baseline_sample(n_elem, n_free, JACOBI_ITERS_FULL)
        if i >= args.warmup:
            vals.append((dofs, t_full))
    v = statistics.mean(x[0] for x in vals)
    t = statistics.mean(x[1] for x in vals)
    out = {"impl": "reference", "metric": "static_solve_dof_per_s", "value": v, "unit": "DOF/s",
           "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": config_dict(),
           "cpu_baseline": {"value": v, "unit": "DOF/s", "cores": 1, "kind": "port", "sample": sample_text(s, JACOBI_ITERS_FULL)},
           "e2e": {"value": v, "unit": "DOF/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


def config_dict(parallel="1 GPU"):
    return {"workload": "BASELINE configs[2]: synthetic gmsh-like 3D space frame, 56x56x54 lattice, 169,344 nodes / "
                        "1,016,064 DOF (997,248 free), 498,848 Timoshenko elements, box/C/L sections, base fixed, "
                        "loads on all top nodes; static solve K u = F (PCG rtol 1e-12) with reaction recovery",
            "step": "fused element+assembly -> BC -> two-level PCG (matrix-free operator; Jacobi + rigid-body coarse space, "
                    "Galerkin matrix and its inverse rebuilt every step) -> reactions",
            "l2": "every step re-assembles the 359 MB K (> 126 MB L2), evicting the CG working set (~80 MB) between steps",
            "parallelism": parallel}


def run_gpu(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from fem_calculator_b200 import _lib as L
    from fem_calculator_b200 import compat, meshgen
    from fem_calculator_b200.api import FrameModel
    from fem_calculator_b200.sections import calculate_section_properties as csp

    E, nu = meshgen.E_STEEL, meshgen.NU_STEEL
    mesh, sec, bc, es, props, fixed, f = build_case(load_scale=1.0 + 0.25 * rank)  # one load case per GPU
    n_free = len(f) - len(fixed)
    n_elem = len(es)
    m = FrameModel(local)
    m.set_mesh(mesh.points, mesh.cells_dict["line"], es, props, E, E / (2 * (1 + nu)))
    m.assemble()
    m.set_bc(fixed, f)
    # FEMB_PRECOND_AUTO: at this size the two-level preconditioner (Jacobi + rigid-body coarse space, csrc/twolevel.cu);
    # its numeric setup (Galerkin matrix + explicit inverse) is rebuilt inside every timed step, after the assembly
    solve_kw = dict(method=L.SOLVER_PCG, precond=L.PRECOND_AUTO, rtol=RTOL, want_u=False, want_reactions=False)

    def step(profile=0):
        m.assemble()
        _, _, st = m.solve_static(profile=profile, **solve_kw)
        return st

    for _ in range(max(args.warmup, 3)):
        step(profile=8)                       # same code path as the timed steps (fills the event pool)

    def barrier():
        if dist is not None:
            import torch
            dist.barrier()
            torch.cuda.synchronize()

    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    m.timer_start()
    stats = [step(profile=8) for _ in range(args.steps)]
    total_ms = m.timer_stop()
    clocks = sampler.stop()
    barrier()
    if dist is not None:
        import torch
        t = torch.tensor([total_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = world * n_free / (ms_per_step / 1e3)
    launches = sum(s["kernel_launches"] for s in stats) + args.steps  # + one assembly launch per step
    iters = stats[-1]["iterations"]
    n_timed = max(1, sum(s["spmv_timed"] for s in stats))
    spmv_ms = sum(s["spmv_ms"] for s in stats) / n_timed
    update_ms = sum(s["update_ms"] for s in stats) / n_timed
    spmv_share = spmv_ms * sum(s["spmv_launches"] for s in stats) / total_ms if world == 1 else None

    coarse_dim = int(stats[-1].get("coarse_dim", 0))
    if coarse_dim:
        n_pad = (coarse_dim + 63) // 64 * 64
        upd_bytes = 13 * 8 * len(f) + 8 * n_pad * n_pad + 2 * (24 + 4) * len(mesh.points)
    else:
        upd_bytes = 12 * 8 * len(f)
    # per-kernel roofline numbers (algorithmic bytes: DESIGN.md §kernels)
    peak, peak_src = peaks()
    ebe = stats[-1].get("op_used") == L.OP_EBE
    _, spmv_bytes = m.time_kernel(0, 1, 1)
    asm_ms, asm_bytes = m.time_kernel(1, 3, 20)
    spmv_b2b_ms, _ = m.time_kernel(0, 3, 50)
    op_bytes = spmv_bytes
    op_b2b_ms = spmv_b2b_ms
    if ebe:
        op_b2b_ms, op_bytes = m.time_kernel(3, 3, 50)

    def _traffic(name):
        tp = os.path.join(ROOT, "profiles", name)
        if os.path.exists(tp):
            try:
                return json.load(open(tp)).get("dram_bytes_per_launch")
            except Exception:
                return None
        return None

    achieved = op_bytes / (spmv_ms * 1e-3) / 1e9
    if ebe:
        roofline = {"kernel": "frame_ebe_node_kernel<1,1,2,masked,dot,128,4> — matrix-free frame operator y = K_ff x "
                              "(inside PCG, every 8th launch timed)",
                    "bound": "hbm", "achieved": achieved, "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                    "frac": achieved / peak, "bytes_per_launch": op_bytes, "ms_per_launch": spmv_ms,
                    "share_of_step": spmv_share, "traffic": _traffic("ncu_ebe_traffic.json"),
                    "note": "algorithmic bytes are the kernel's own compulsory traffic (16 B/pair + node records + "
                            "coordinates + x + y + mask = 40 MB), 9x fewer than the 359 MB the assembled SpMV streams for "
                            "the same product; the kernel is FP64-issue / gather-latency bound, not HBM bound (ncu: FP64 "
                            "pipe 34-41 %, DRAM 13 % of peak) — see roofline_bsr_spmv for the HBM-bound form of the product",
                    "equivalent_bsr_gbs": spmv_bytes / (spmv_ms * 1e-3) / 1e9}
    else:
        roofline = {"kernel": "bsr_spmv_kernel<6,masked,dot,192,2> (inside PCG, every 8th launch timed)", "bound": "hbm",
                    "achieved": achieved, "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": achieved / peak,
                    "bytes_per_launch": spmv_bytes, "ms_per_launch": spmv_ms, "share_of_step": spmv_share,
                    "traffic": _traffic("ncu_spmv_traffic.json")}
    extra = {
        "pcg": {"iterations": iters, "ms_per_iteration": stats[-1]["device_ms"] / max(1, iters),
                "rel_residual": stats[-1]["rel_residual"],
                "precond": "two-level (Jacobi + 6 rigid-body modes per aggregate)" if coarse_dim else "jacobi",
                "coarse_dim": coarse_dim,
                "form": "Chronopoulos-Gear, 3 kernels/iteration (operator, update + restriction, coarse solve + prolongation)"
                        if coarse_dim else "Chronopoulos-Gear, 2 kernels/iteration",
                "update_kernel_ms": update_ms, "update_kernel_gbs": upd_bytes / (update_ms * 1e-3) / 1e9 if update_ms else None},
        "assembly": {"kernel": "frame_assemble_pairs_persistent_kernel (fused element+assembly)", "ms": asm_ms,
                     "elements_per_s": n_elem / (asm_ms * 1e-3), "achieved_gbs": asm_bytes / (asm_ms * 1e-3) / 1e9,
                     "frac": asm_bytes / (asm_ms * 1e-3) / 1e9 / peak, "bytes_per_launch": asm_bytes},
        "spmv_back_to_back": {"ms": spmv_b2b_ms, "achieved_gbs": spmv_bytes / (spmv_b2b_ms * 1e-3) / 1e9,
                              "frac": spmv_bytes / (spmv_b2b_ms * 1e-3) / 1e9 / peak},
        "operator_back_to_back_ms": op_b2b_ms,
    }
    if coarse_dim:
        extra["pcg"]["update_kernel_note"] = ("tl_update_kernel + tl_coarse_z_kernel together (two launches): 10 + 3 vector "
                                              "passes, the n_pad^2 coarse inverse, coordinates and node lists")
    extra["pcg"]["operator"] = "matrix-free (EBE)" if ebe else "assembled BSR"
    extra["pcg"]["update_kernel_frac"] = (extra["pcg"]["update_kernel_gbs"] / peak) if extra["pcg"]["update_kernel_gbs"] else None
    if coarse_dim and world == 1:
        # the same step with the plain Jacobi preconditioner, outside the timed region (what the coarse space buys)
        jkw = dict(solve_kw, precond=L.PRECOND_JACOBI)
        m.solve_static(**jkw)
        _, _, jst = m.solve_static(profile=8, **jkw)
        extra["jacobi_pcg"] = {"iterations": jst["iterations"], "pcg_ms": jst["device_ms"],
                               "ms_per_iteration": jst["device_ms"] / max(1, jst["iterations"]),
                               "operator_ms": jst["spmv_ms"] / max(1, jst["spmv_timed"]),
                               "update_ms": jst["update_ms"] / max(1, jst["spmv_timed"]),
                               "dof_per_s_pcg_only": n_free / (jst["device_ms"] * 1e-3)}
        _, _, tst = m.solve_static(**solve_kw)      # coarse inverse kept: the solve alone
        extra["pcg"]["solve_only_ms"] = tst["device_ms"]
        extra["pcg"]["coarse_setup_ms"] = stats[-1]["device_ms"] - tst["device_ms"]
    if ebe and world == 1:
        # the HBM-bound form of the same product: one solve with the assembled BSR operator, outside the timed region
        _, _, bst = m.solve_static(profile=8, op=L.OP_BSR, **dict(solve_kw, precond=L.PRECOND_JACOBI))
        nb_t = max(1, bst["spmv_timed"])
        b_ms = bst["spmv_ms"] / nb_t
        b_ach = spmv_bytes / (b_ms * 1e-3) / 1e9
        extra["roofline_bsr_spmv"] = {"kernel": "bsr_spmv_kernel<6,masked,dot,192,2> (inside a PCG with op = BSR, every 8th launch timed)",
                                      "bound": "hbm", "achieved": b_ach, "peak": peak, "unit": "GB/s", "frac": b_ach / peak,
                                      "bytes_per_launch": spmv_bytes, "ms_per_launch": b_ms,
                                      "traffic": _traffic("ncu_spmv_traffic.json"),
                                      "pcg_ms": bst["device_ms"], "pcg_iterations": bst["iterations"],
                                      "dof_per_s_pcg_only": n_free / (bst["device_ms"] * 1e-3)}
    if world == 1 and not args.no_modal:
        # second headline metric: ms per 20-mode modal solve at 1M DOF (device time of femb_modal)
        two_level_modal = os.environ.get("FEMB_BENCH_MODAL_PRECOND", "two_level") != "jacobi"
        lam, _, mst = m.modal(k=20, rtol=1e-8, precond=L.PRECOND_TWO_LEVEL if two_level_modal else L.PRECOND_JACOBI)
        extra["modal"] = {"metric": "ms_per_20_mode_modal", "ms": mst["device_ms"], "modes": int(len(lam)),
                          "pcg_iterations": mst["iterations"], "matrix_passes": mst["spmv_launches"],
                          "rel_residual": mst["rel_residual"], "omega_min_rad_s": float(np.sqrt(lam[0])),
                          "omega_max_rad_s": float(np.sqrt(lam[-1])),
                          "operator": "matrix-free (EBE)" if mst.get("op_used") == L.OP_EBE else "assembled BSR SpMM",
                          "coarse_dim": int(mst.get("coarse_dim", 0)),
                          "method": ("block shift-invert Krylov (block 2), two-level PCG as K^-1, one right-hand side at a time"
                                     if mst.get("coarse_dim") else
                                     "block shift-invert Krylov (block 2), 2-RHS lockstep Jacobi-PCG as K^-1")}
    m.close()

    if world > 1 and not args.no_rowblock:
        # strong-scaling companion (not the headline): the SAME frame split by node slabs across the N
        # GPUs, distributed PCG with peer-memory (NVLink) halo + scalar exchange fused into its kernels
        import torch
        from fem_calculator_b200.api import DistFrameModel
        mesh0, sec0, bc0, es0, props0, fixed0, f0 = build_case(load_scale=1.0)
        t = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            t = torch.tensor(list(DistFrameModel.unique_id()), dtype=torch.uint8, device="cuda")
        dist.broadcast(t, 0)

        def _gather(obj):
            out = [None] * world
            dist.all_gather_object(out, obj)
            return out

        d = DistFrameModel(local)
        d.setup(mesh0.points, mesh0.cells_dict["line"], es0, props0, E, E / (2 * (1 + nu)), fixed0, f0, rank, world,
                bytes(t.cpu().numpy().tolist()), all_gather=_gather)
        d.solve_static_dist(precond=L.PRECOND_JACOBI, rtol=RTOL, want_u=False, want_reactions=False)   # warm-up
        barrier()
        _, _, dst = d.solve_static_dist(precond=L.PRECOND_JACOBI, rtol=RTOL, want_u=False, want_reactions=False)
        tt = torch.tensor([dst["device_ms"]], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        extra["rowblock"] = {"what": "one 1,016,064-DOF frame, node-slab partition over the N GPUs (strong scaling), PCG solve only",
                             "iterations": dst["iterations"], "ms": float(tt.item()),
                             "us_per_iteration": float(tt.item()) / max(1, dst["iterations"]) * 1e3,
                             "dof_per_s": (len(f0) - len(fixed0)) / (float(tt.item()) * 1e-3),
                             "exchange": "peer memory (CUDA IPC over NVLink), fused into SpMV/update" if d.p2p else "NCCL"}
        d.close()

    # end-to-end through the reference-shaped entry point, host buffers in / out
    e2e_steps = max(1, min(args.steps, 3))
    w = compat.BeamAnalysisB200(mesh, sec, bc, E, nu, device=local)
    tc = time.perf_counter()
    w.run_simulation(k_modes=0, solver=L.SOLVER_PCG, rtol=RTOL)  # warm-up (context, allocator, symbolic analysis)
    e2e_first_call_s = time.perf_counter() - tc
    L.io_bytes(reset=True)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        w.run_simulation(k_modes=0, solver=L.SOLVER_PCG, rtol=RTOL)
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    barrier()
    h2d, d2h = L.io_bytes()
    if dist is not None:
        import torch
        t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e = {"value": world * n_free / e2e_s, "unit": "DOF/s", "h2d_bytes_per_step": h2d // e2e_steps,
           "d2h_bytes_per_step": d2h // e2e_steps, "ms_per_step": e2e_s * 1e3, "steps": e2e_steps,
           "first_call_ms": e2e_first_call_s * 1e3,
           "api": "fem_calculator_b200.compat.BeamAnalysisB200.run_simulation (BeamSolver.py:345 signature), "
                  "timed with the host clock around the call; every call takes coordinates, connectivity, sections, "
                  "BCs and loads as host numpy arrays, copies coordinates / section data / loads / BCs to the device and "
                  "returns u / reactions / stresses to the host; the connectivity is compared with the previous call's and "
                  "the symbolic analysis (block pattern, pair records, 62 MB of device tables) is rebuilt and re-uploaded only "
                  "when it changed — first_call_ms is a call that builds it"}

    out = {"metric": "static_solve_dof_per_s", "value": value, "unit": "DOF/s", "n_gpus": world, "steps": args.steps,
           "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": config_dict("1 GPU" if world == 1 else f"{world} independent load cases, one per GPU (no collective)"),
           "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline}
    out.update(extra)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        # Jacobi-PCG iteration count at full size comes from this very run (same algorithm, same rtol)
        # the CPU port runs the Jacobi-PCG (oracle/cpu_baseline.py): scale with ITS iteration count at full size
        cpu_iters = extra.get("jacobi_pcg", {}).get("iterations", iters if not coarse_dim else JACOBI_ITERS_FULL)
        dofs, t_full, s = cpu_baseline_sample(n_elem, n_free, cpu_iters)
        out["cpu_baseline"] = {"value": dofs, "unit": "DOF/s", "cores": 1, "kind": "port",
                               "sample": sample_text(s, cpu_iters), "est_seconds_full": t_full}
    if rank == 0:
        print(json.dumps(out), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-rowblock", action="store_true", help="N>1: skip the row-block (strong-scaling) companion solve")
    ap.add_argument("--no-modal", action="store_true", help="skip the 20-mode modal measurement (N=1 only, ~45 s)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
