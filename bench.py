#!/usr/bin/env python
"""bench.py -- cells/s through one Newton step (assemble J+F, Dirichlet rows, sparse linear solve,
update, residual) of the P-Q system on the 16 Mi-cell structured (r2,r1) mesh.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--nx NX]

Workload (BASELINE.json configs[4]): RectangleMesh 4096x2048 on (0,1)^2
(16 777 216 cells, 16 789 506 dofs, 234 954 756 nnz), manufactured P-Q problem with the SUPG
q-equation (…MixedPrimal_convergence_SUPG.py:118-129), state = nodal interpolant of the exact
solution (every step does the same work: one full Newton iteration from that state).
N > 1 (default): the SAME mesh is strip-partitioned by vertex rows over the N GPUs (RCB along r1; halo
rows, coarse-grid all-gather and Krylov all-reduces by peer-to-peer kernels over NVLink): strong
scaling.  `--instances`: every rank runs an independent problem instance instead (the concentration
sweep of advection_diffusion_mixed.py:273-334, the reference's own data-parallel axis): weak scaling,
no data-path communication.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


# ---------------------------------------------------------------------------------------------
def cpu_newton_step_sample(nx, ny, repeats=1):
    """The CPU restatement of the same step on a bounded sample mesh: plain-C serial cell assembly
    (oracle/cport, what DOLFIN's serial Assembler does) + SuperLU factor/solve (the only sparse
    direct solver in the image, standing in for MUMPS/UMFPACK) + update + residual assembly."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    from oracle import cport, fem, pq
    from oracle.newton import apply_bc_matrix_vec, apply_bc_residual, merge_bcs

    mesh = fem.rectangle_mesh(0.0, 0.0, 1.0, 1.0, nx, ny)
    prm = pq.PQParams(D1=1e-3, D2=1.5e-3, gkind=pq.G_POWER, gcoef=1.5e-3 * 4 * np.pi, q_react=1.0,
                      q_x2p=0.0, q_x2q=1.0, supg_stab=1.0)
    S = pq.PQSystem(mesh, prm)
    X = mesh.coords
    fc = mesh.facets()
    bv = np.unique(fc.verts[fc.exterior])
    qv = np.nonzero(fem.near(X[:, 0], 0.0))[0]
    dofs, gv = merge_bcs([(2 * bv, pq.p_exact(X[bv, 0], X[bv, 1])),
                          (2 * qv + 1, pq.q_exact(X[qv, 0], X[qv, 1]))])
    forcing = pq.manufactured_forcing(prm, x2_on_q=True)
    pts, _ = fem.triangle_rule(4)
    phi = np.stack([1 - pts[:, 0] - pts[:, 1], pts[:, 0], pts[:, 1]], 1)
    P = np.einsum("qi,cid->cqd", phi, X[mesh.cells])
    f2, f3 = forcing(P[:, :, 0], P[:, :, 1])
    fq = np.ascontiguousarray(np.stack([f2, f3], -1))
    u0 = np.empty(S.N)
    u0[0::2] = pq.p_exact(X[:, 0], X[:, 1])
    u0[1::2] = pq.q_exact(X[:, 0], X[:, 1])
    slots = S.cell_slots
    best = None
    for _ in range(repeats):
        u = u0.copy()
        t0 = time.perf_counter()
        vals, b = cport.assemble_cells(prm, mesh.cells, mesh.coords, u, slots, S.col.size, None, fq, True)
        apply_bc_matrix_vec(S.rowptr, S.col, vals, dofs)
        apply_bc_residual(b, u, dofs, gv)
        t1 = time.perf_counter()
        A = sp.csr_matrix((vals, S.col, S.rowptr), shape=(S.N, S.N))
        dx = spla.splu(A.tocsc()).solve(b)
        t2 = time.perf_counter()
        u -= dx
        _, b = cport.assemble_cells(prm, mesh.cells, mesh.coords, u, slots, S.col.size, None, fq, False)
        apply_bc_residual(b, u, dofs, gv)
        float(np.linalg.norm(b))
        t3 = time.perf_counter()
        rec = dict(total=t3 - t0, assemble=t1 - t0, lu=t2 - t1, residual=t3 - t2)
        if best is None or rec["total"] < best["total"]:
            best = rec
    best["cells"] = mesh.num_cells
    return best


REF_NX, REF_NY = 512, 256          # bounded sample of the reference arm / cpu_baseline leg


def reference_arm(args):
    """--impl reference: the reference's CPU path (FEniCS-equivalent restatement: serial C cell loop + SuperLU;
    FEniCS itself is not installable here) timed on this box's host cores; rank 0 only.  Each step is a BOUNDED
    SAMPLE of the workload: the same Newton step on RectangleMesh 512x256 (a direct sparse LU of the 16 Mi-cell
    Jacobian neither fits nor finishes on the host).  The line says so: `config` names the mesh that ran,
    `same_config` is false, and the measured growth exponent of the CPU step time brackets what the per-cell
    ratio means at the full size."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    nx, ny = REF_NX, REF_NY
    for _ in range(args.warmup):
        cpu_newton_step_sample(nx, ny)
    t = []
    rec = None
    for _ in range(args.steps):
        rec = cpu_newton_step_sample(nx, ny)
        t.append(rec["total"])
    ms = 1e3 * float(np.mean(t))
    val = rec["cells"] / (ms * 1e-3)
    small = cpu_newton_step_sample(nx // 2, ny // 2)
    growth = float(np.log(float(np.mean(t)) / small["total"]) / np.log(4.0))
    sample = (f"one Newton step (C serial cell assembly + SuperLU + residual) on RectangleMesh {nx}x{ny} "
              f"= {rec['cells']} cells; assemble {rec['assemble']:.2f}s lu {rec['lu']:.2f}s; step time grows like "
              f"cells^{growth:.2f} ({nx // 2}x{ny // 2} -> {nx}x{ny}, this run; 1.64 from 512x256 to 1024x512, "
              f"profiles/README.md), i.e. cells/s FALLS with the mesh size")
    cfg = reference_config(args, nx, ny, growth)
    out = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": cfg, "same_config": False,
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out))


def reference_config(args, nx, ny, growth=None):
    full_nx = args.nx
    cfg = {"workload": f"BOUNDED SAMPLE of the P-Q CG1xCG1 manufactured SUPG Newton step: RectangleMesh {nx}x{ny} "
                       f"({2 * nx * ny} cells) instead of {full_nx}x{full_nx // 2} ({full_nx * full_nx} cells); "
                       "cells/s of the sample, NOT of the full mesh",
           "nx": nx, "ny": ny, "cells": 2 * nx * ny, "dofs": 2 * (nx + 1) * (ny + 1),
           "full_workload": {"nx": full_nx, "ny": full_nx // 2, "cells": full_nx * full_nx},
           "parallelism": "1 host core (DOLFIN assembles serially; SuperLU is sequential)",
           "linear_solver": "sparse LU (scipy SuperLU, standing in for MUMPS/UMFPACK), one factorisation per step",
           "linear_rtol": None, "same_config_as_gpu_arm": False}
    if growth is not None:
        cfg["cpu_step_time_growth_exponent"] = growth
    return cfg


METRIC = "cells/s through one Newton step (assemble J+F, BCs, linear solve, update, residual)"
UNIT = "cells/s"


def workload_config(args, nx, ny):
    gpus = max(1, args.gpus)
    cells = 2 * nx * ny
    variant = getattr(args, "variant", "supg").upper()
    solver = ("pivoting banded LU (cm_blu_*) + 1 refinement step" if getattr(args, "linear_solver", "gmres") == "lu"
              else f"flexible GMRES({RESTART}) on the fp64 operator + field split + V(1,2) multigrid / zebra line relaxation"
                   + ("" if os.environ.get("CM_F32_OP", "1") == "0" else
                      "; preconditioner detail: the long-line smoother reads its off-line diagonals and line factors "
                      "from an fp32 copy (x, b, residuals, transfers, Galerkin products, Krylov vectors and every "
                      "product / sum are fp64; the solve converges to linear_rtol in the fp64 residual)"))
    if gpus > 1 and not getattr(args, "instances", False):
        kind = "weak scaling: the mesh grows with the GPUs" if getattr(args, "weak", False) else "strong scaling"
        return {"workload": f"P-Q CG1xCG1 manufactured {variant} Newton step on ONE RectangleMesh {nx}x{ny} "
                            f"({cells} cells) strip-partitioned over {gpus} GPUs ({cells // gpus} cells each; {kind})",
                "nx": nx, "ny": ny, "cells": cells, "cells_per_gpu": cells // gpus,
                "dofs": 2 * (nx + 1) * (ny + 1), "dofs_per_gpu": 2 * (nx + 1) * (ny + 1) // gpus,
                "parallelism": f"one mesh strip-partitioned by vertex rows over {gpus} GPUs (boundary rows pushed over "
                               "NVLink by the producing kernels, P2P all-gather / all-reduce kernels)",
                "l2_policy": "inputs larger than L2", "linear_solver": solver, "linear_rtol": LIN_RTOL}
    return {"workload": f"P-Q CG1xCG1 manufactured {variant} Newton step on RectangleMesh {nx}x{ny} "
                        f"({cells} cells)" + (f", one independent instance per GPU x{gpus}" if gpus > 1 else ""),
            "nx": nx, "ny": ny, "cells": cells * gpus, "cells_per_gpu": cells, "dofs_per_gpu": 2 * (nx + 1) * (ny + 1),
            "parallelism": (f"independent problem instance per GPU x{gpus}" if gpus > 1 else "single GPU"),
            "l2_policy": "inputs larger than L2 (Jacobian 1.9 GB, Krylov basis > 1 GB)",
            "linear_solver": solver, "linear_rtol": LIN_RTOL}


RESTART = 40
LIN_RTOL = 1e-10


# ---------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [s.strip() for s in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------
def gpu_arm(args):
    import torch
    import torch.distributed as dist
    import closestmolecule_b200 as cm
    from closestmolecule_b200 import _capi
    from closestmolecule_b200.manufactured import p_exact, q_exact

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); "
                         "use --impl reference for the CPU restatement")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    nx = args.nx
    ny = nx // 2
    if args.weak and world > 1:      # weak scaling: the strip-partitioned mesh grows with the GPUs (nx x nx/2 cells x 2 per GPU)
        ny = ny * world
    lib = _capi.lib()

    # weak scaling keeps the cell shape: the domain grows in r1 with the number of GPUs
    top = float(world) if (args.weak and world > 1) else 1.0
    mesh = cm.RectangleMesh(cm.Point(0, 0), cm.Point(1, top), nx, ny, device=dev)
    P1 = cm.FiniteElement("CG", mesh.ufl_cell(), 1)
    V = cm.FunctionSpace(mesh, cm.MixedElement([P1, P1]))
    bd = cm.MeshFunction("size_t", mesh, 1, 0)
    cm.CompiledSubDomain("near(x[0],0) && on_boundary").mark(bd, 31)
    if args.variant == "supg":
        form = cm.PQPrimalForm.manufactured(V, "supg")
        bcs = [cm.DirichletBC(V.sub(0), p_exact, "on_boundary"), cm.DirichletBC(V.sub(1), q_exact, bd, 31)]
    elif args.variant == "galerkin":   # BASELINE config 2 (…MixedPrimal_convergence.py:108-131): needs the LU path
        bd.set_all(0)
        cm.CompiledSubDomain("near(x[0],1) && on_boundary").mark(bd, 31)
        form = cm.PQPrimalForm.manufactured(V, "galerkin")
        bcs = [cm.DirichletBC(V.sub(0), p_exact, "on_boundary"), cm.DirichletBC(V.sub(1), q_exact, bd, 31)]
    else:   # C0IP variants of BASELINE config 3 (…_C0IP.py:141-147 / test_for_paper_convergence.py:140-153)
        cm.CompiledSubDomain("near(x[0],0) && on_boundary").mark(bd, 30)
        cm.CompiledSubDomain("near(x[0],1) && on_boundary").mark(bd, 31)
        cm.CompiledSubDomain("(near(x[1],0) || near(x[1],1)) && on_boundary").mark(bd, 32)
        if args.variant == "c0ip_minus":
            form = cm.PQPrimalForm.manufactured(V, "c0ip", D1=0.01, D2=0.015, stab=0.05, sign=-1.0,
                                                ext_terms=[(bd, 30, cm.EF_ADV)])
            bcs = [cm.DirichletBC(V.sub(0), p_exact, "on_boundary"), cm.DirichletBC(V.sub(1), q_exact, bd, 31)]
        else:
            form = cm.PQPrimalForm.manufactured(
                V, "c0ip", D1=1.0, D2=1.5, stab=0.05, sign=+1.0, c0ip_h="facet", degree=6, pen_stab2=1.0,
                ext_terms=[(bd, 30, cm.EF_ADV), (bd, 32, cm.EF_ADV), (bd, 31, cm.EF_DATA),
                           (bd, "on_boundary", cm.EF_PENALTY)])
            bcs = [cm.DirichletBC(V.sub(0), p_exact, "on_boundary")]
    strong = world > 1 and not args.instances
    X = mesh.coords
    if strong:
        from closestmolecule_b200.multigpu import DistributedPQNewton
        solver = DistributedPQNewton(form, bcs, restart=RESTART, rtol_lin=LIN_RTOL)
        uvec = solver.u
        own = solver.own
    else:
        u = cm.Function(V)
        solver = cm.NonlinearVariationalSolver(cm.NonlinearVariationalProblem(form, u, bcs))
        solver.parameters["newton_solver"]["linear_solver"] = "mumps" if args.linear_solver == "lu" else "gmres"
        solver.parameters["newton_solver"]["krylov_solver"]["relative_tolerance"] = LIN_RTOL
        if args.variant != "supg":
            solver.parameters["newton_solver"]["krylov_solver"]["restart"] = 100
        uvec = u.vector()
        own = slice(0, uvec.numel())
    u0 = torch.empty_like(uvec)
    u0[0::2] = p_exact(X[:, 0], X[:, 1])
    u0[1::2] = q_exact(X[:, 0], X[:, 1])
    cells = mesh.num_cells()
    Vn = mesh.num_vertices()
    nnz = solver.asm.topo.nnz

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ev = lambda: torch.cuda.Event(enable_timing=True)

    def step():
        uvec[own] = u0[own]
        return solver.newton_step(compute_residual=True)

    sampler = ClockSampler(local)
    if rank == 0:       # started before the warm-up so that nvidia-smi is already streaming samples
        sampler.start()  # when the timed region begins; the samples cover warm-up + timed steps
    for _ in range(args.warmup):
        step()
    barrier()
    l0 = int(lib.cm_launch_count())
    by0 = float(lib.cm_algorithmic_bytes())
    lib.cm_profile_reset()
    # CUDA events around the roofline kernels, on the launching stream.  Multi-GPU: off, so that the
    # preconditioner applications can be replayed as CUDA graphs (the roofline is reported at N=1)
    lib.cm_profile_enable(0 if (world > 1 or os.environ.get('CM_BENCH_NOPROF')) else 1)
    solver.linear_iterations = []
    solver.phase_events = []      # CUDA events on the launching stream around each phase
    t0, t1 = ev(), ev()
    barrier()
    t0.record()
    rnorm = None
    for k_ in range(args.steps):
        if k_ == 1:
            # the first timed step carries the CUDA-event kernel timing (roofline object); the others
            # replay the preconditioner as CUDA graphs, which cannot hold timing events
            lib.cm_profile_enable(0)
        rnorm = step()
    t1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    launches = int(lib.cm_launch_count()) - l0
    step_bytes = (float(lib.cm_algorithmic_bytes()) - by0) / args.steps
    lib.cm_profile_enable(0)
    import ctypes as _C
    prof = {}
    prof_steps = 1 if args.steps > 1 else args.steps
    for cls, name in ((0, ROOF_KERNEL), (1, "pq_stencil_spmv_kernel"), (2, "pq_sweep_kernel (J+F)")):
        n_, ms_, by_ = _C.c_int64(0), _C.c_double(0.0), _C.c_double(0.0)
        lib.cm_profile_report(cls, _C.byref(n_), _C.byref(ms_), _C.byref(by_))
        if n_.value:
            prof[name] = {"launches_per_step": n_.value / prof_steps, "avg_us": 1e3 * ms_.value / n_.value,
                          "ms_per_step": ms_.value / prof_steps,
                          "algorithmic_bytes_per_launch": by_.value / n_.value,
                          "gbs": by_.value / (ms_.value * 1e-3) / 1e9}
    ms_step = t0.elapsed_time(t1) / args.steps
    ph = np.array([[m[i].elapsed_time(m[i + 1]) for i in range(4)] for m in solver.phase_events])
    solver.phase_events = None
    asm_ms, setup_ms, solve_ms, resid_ms = [float(v) for v in ph.mean(0)]
    lin_its = float(np.mean(solver.linear_iterations))

    # ---- e2e: host buffers in, host buffers out, through the public Newton-step call ------------
    h_in = torch.empty(u0[own].numel(), dtype=torch.float64, pin_memory=True)
    h_in.copy_(u0[own])
    h_out = torch.empty(h_in.numel(), dtype=torch.float64, pin_memory=True)   # empty_like would not be pinned
    e2e_steps = max(1, min(args.steps, 3))
    dst = uvec[own]
    for _ in range(1):              # untimed warm-up of the copy path
        dst.copy_(h_in, non_blocking=True)
        h_out.copy_(dst, non_blocking=True)
    barrier()
    marks = []
    e0, e1 = ev(), ev()
    side = torch.cuda.Stream()
    upd = torch.cuda.Event()

    def push_result():
        # the new iterate is final once the update is enqueued: its device-to-host copy runs on a side stream
        # while the step evaluates the residual norm of the new iterate (the step's scalar result)
        upd.record()
        side.wait_event(upd)
        with torch.cuda.stream(side):
            h_out.copy_(dst, non_blocking=True)

    e0.record()
    for _ in range(e2e_steps):
        a, b_, c_, d_ = ev(), ev(), ev(), ev()
        a.record()
        dst.copy_(h_in, non_blocking=True)
        b_.record()
        r = solver.newton_step(compute_residual=True, on_updated=push_result)
        c_.record()
        torch.cuda.current_stream().wait_stream(side)     # the result is on the host
        d_.record()
        torch.cuda.current_stream().synchronize()
        marks.append((a, b_, c_, d_))
    e1.record()
    barrier()
    e2e_ms = e0.elapsed_time(e1) / e2e_steps
    e2e_parts = [float(np.mean([m[i].elapsed_time(m[i + 1]) for m in marks])) for i in range(3)]

    # ---- parity (untimed): the answer of the timed path against an independent solve -----------------------
    def lin_info():
        try:
            return solver.info() if strong else solver._linear.info()
        except Exception as e:      # LU path etc.
            return {"note": type(e).__name__}

    def one_step_solution(kind, mesh_, form_, bcs_, u_init):
        uu = cm.Function(form_.space)
        sv = cm.NonlinearVariationalSolver(cm.NonlinearVariationalProblem(form_, uu, bcs_))
        sv.parameters["newton_solver"]["linear_solver"] = kind
        sv.parameters["newton_solver"]["krylov_solver"]["relative_tolerance"] = LIN_RTOL
        uu.vector().copy_(u_init)
        sv.newton_step(compute_residual=False)
        return uu.vector().clone(), (sv.linear_iterations[-1] if sv.linear_iterations else None)

    # ---- full-size property (untimed, one GPU): the Newton update of the timed path satisfies the ASSEMBLED system
    # J dx = F -- true residual through the generic CSR kernel (cm_spmv on the DOLFIN-layout values, which shares
    # nothing with the stencil blocks / multigrid / Krylov kernels of the solve)
    full_check = None
    if not strong and world == 1 and not args.no_parity and args.linear_solver == "gmres":
        uvec.copy_(u0)
        solver.newton_step(compute_residual=False)     # leaves J (with Dirichlet rows) in asm.vals, F in asm.b, dx in _dx
        asm_, t_ = solver.asm, solver.asm.topo
        r_ = torch.empty_like(asm_.b)
        _capi.check(lib.cm_spmv(asm_.bs, t_.nv, _capi.ptr(t_.nrowptr), _capi.ptr(t_.ncol), _capi.ptr(asm_.vals),
                                _capi.ptr(solver._dx), _capi.ptr(r_), _capi.stream_ptr()))
        r_.sub_(asm_.b)
        # the solve runs on the row-equilibrated system: measure the residual in the same scaling (rows / max |entry|),
        # and unscaled
        bn = float(torch.linalg.vector_norm(asm_.b))
        full_check = {"what": f"|| J dx - F ||_2 / || F ||_2 at the full size ({nx}x{ny}), J dx by cm_spmv on the CSR values",
                      "rel_residual_unscaled": float(torch.linalg.vector_norm(r_)) / bn if bn > 0 else None,
                      "gmres_its": int(solver.linear_iterations[-1])}
        del r_

    parity = None
    if args.variant == "supg" and args.linear_solver == "gmres" and not args.no_parity:
        if strong:
            # the strip-partitioned Newton step against the single-GPU Newton step from the same state (rank 0)
            step()
            u_dist = solver.gather_solution().clone()
            if rank == 0:
                u_one, its_one = one_step_solution("gmres", mesh, form, bcs, u0)
                parity = {"what": f"state after one Newton step, {world} GPUs (strip-partitioned) vs 1 GPU, same mesh",
                          "rel_l2_vs_n1": float(torch.linalg.vector_norm(u_dist - u_one) / torch.linalg.vector_norm(u_one)),
                          "gmres_its": int(solver.linear_iterations[-1]), "gmres_its_n1": its_one}
                del u_one
            barrier()
        elif world == 1:
            # the multigrid-Krylov Newton step against the pivoting direct solve (cm_blu_*) on the largest mesh whose
            # factors fit comfortably: same form, same state, 512x256
            pm = cm.RectangleMesh(cm.Point(0, 0), cm.Point(1, 1), REF_NX, REF_NY, device=dev)
            pV = cm.FunctionSpace(pm, cm.MixedElement([P1, P1]))
            pbd = cm.MeshFunction("size_t", pm, 1, 0)
            cm.CompiledSubDomain("near(x[0],0) && on_boundary").mark(pbd, 31)
            pform = cm.PQPrimalForm.manufactured(pV, "supg")
            pbcs = [cm.DirichletBC(pV.sub(0), p_exact, "on_boundary"), cm.DirichletBC(pV.sub(1), q_exact, pbd, 31)]
            pu0 = torch.empty(2 * pm.num_vertices(), dtype=torch.float64, device=dev)
            pu0[0::2] = p_exact(pm.coords[:, 0], pm.coords[:, 1])
            pu0[1::2] = q_exact(pm.coords[:, 0], pm.coords[:, 1])
            uk, its_k = one_step_solution("gmres", pm, pform, pbcs, pu0)
            ud, _ = one_step_solution("mumps", pm, pform, pbcs, pu0)
            du = torch.linalg.vector_norm(ud - pu0)
            parity = {"what": f"Newton update of the GMRES+multigrid path vs the pivoting direct LU (cm_blu_*), "
                              f"RectangleMesh {REF_NX}x{REF_NY}, same state",
                      "rel_l2_state": float(torch.linalg.vector_norm(uk - ud) / torch.linalg.vector_norm(ud)),
                      "rel_l2_update": float(torch.linalg.vector_norm(uk - ud) / du) if float(du) > 0 else None,
                      "gmres_its": its_k}
            del uk, ud

    zk0 = prof.get(ROOF_KERNEL)
    # the long-line smoother reads its seven operator streams from an fp32 pack unless CM_F32_OP=0:
    # 52 / 28 algorithmic bytes per vertex of the colour (with couplings / from a zero guess) instead of 80 / 40
    f32_op = os.environ.get("CM_F32_OP", "1") != "0" and nx + 1 >= 1024
    b_cpl, b_zero = (52.0, 28.0) if f32_op else (80.0, 40.0)
    prof_mix = 0.2
    if zk0:
        unit = 0.5 * (nx + 1) * (ny + 1)
        prof_mix = float(min(1.0, max(0.0, (b_cpl - zk0["algorithmic_bytes_per_launch"] / unit) / (b_cpl - b_zero))))

    tt = torch.tensor([ms_step, e2e_ms, asm_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms_step, e2e_ms, asm_ms = [float(v) for v in tt.tolist()]

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
        bytes_asm = 12 * cells + 16 * Vn + 8 * (2 * Vn) + 8 * (4 * nnz) + 8 * (2 * Vn)
        bytes_asm_rank = bytes_asm / (world if strong else 1)      # every rank assembles its own rows
        zk = prof.get(ROOF_KERNEL)
        traffic = ncu_traffic(prof_mix, f32_op) if (zk and world == 1) else None
        solver_info = lin_info()
        out = {
            "metric": METRIC, "value": (1 if strong else world) * cells / (ms_step * 1e-3), "unit": UNIT,
            "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "strong" if (strong and not args.weak) else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, nx, ny),
            "newton_step_ms": ms_step, "linear_iterations_per_step": lin_its,
            "assembly_ms": asm_ms, "assembly_cells_per_s": cells / (asm_ms * 1e-3),
            "phases_ms": {"assemble_J_F": asm_ms, "bc_and_solver_setup": setup_ms,
                          "krylov_solve": solve_ms, "update_and_residual": resid_ms},
            "residual_norm_after_step": rnorm,
            "e2e": {"value": (1 if strong else world) * cells / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": int(h_in.numel() * 8), "d2h_bytes_per_step": int(h_out.numel() * 8 + 8),
                    "h2d_ms": e2e_parts[0], "step_ms": e2e_parts[1], "d2h_ms": e2e_parts[2],
                    "note": ("bytes per rank" if world > 1 else "whole state vector") +
                            "; d2h_ms is the part of the result copy that is still exposed after the step: it starts on a "
                            "side stream as soon as the update is enqueued and overlaps the residual evaluation"},
            "gpu_launches": launches,
            "clocks": clocks,
            "solver": solver_info,
            # dominant kernel of the step by time share (profiles/): the fused zebra half sweep (couplings +
            # x-line solves) on the finest grid, p- and q-block launches; durations from CUDA events on the launching
            # stream inside the timed region (first timed step); reported on one GPU only
            "roofline": {"kernel": "zebra_ring_kernel (fused zebra half sweep, finest grid)", "bound": "hbm",
                         "achieved": zk["gbs"] if zk else None, "peak": peak, "unit": "GB/s",
                         "frac": (zk["gbs"] / peak) if zk else None, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": zk["algorithmic_bytes_per_launch"] if zk else None,
                         "algorithmic_bytes_per_vertex": (
                             "52 with couplings (b 8, 4 off-line diagonals + 3 line factors from the fp32 pack 28, the "
                             "neighbour lines of x once 8, x out 8), 28 from a zero guess; x, b and all arithmetic fp64; "
                             "with CM_F32_OP=0 the same kernel streams the fp64 operator: 80 / 40 B, 0.67 of the peak "
                             "but 17 us slower per launch (profiles/README.md)" if f32_op else
                             "80 with couplings (b, 4 off-line diagonals, 3 factors, the neighbour lines of x once, "
                             "x out), 40 from a zero guess"),
                         "operator_storage": "fp32 pack" if f32_op else "fp64",
                         "avg_launch_us": zk["avg_us"] if zk else None,
                         "share_of_step": (zk["ms_per_step"] / ms_step) if zk else None,
                         "traffic": traffic["bytes_per_launch"] if traffic else None,
                         "traffic_source": traffic["source"] if traffic else None},
            # whole Newton step: algorithmic bytes of every kernel launched by the library in the step
            # (each array once per pass; per rank) over the step time and the measured HBM peak
            "step_roofline": {"algorithmic_bytes_per_step": step_bytes, "gbs": step_bytes / (ms_step * 1e-3) / 1e9,
                              "frac": step_bytes / (ms_step * 1e-3) / 1e9 / peak, "peak": peak,
                              "gmres_iterations": lin_its, "note": "per rank" if world > 1 else "whole job"},
            "kernels": {**prof, "pq_sweep_kernel roofline frac (per rank)": (bytes_asm_rank / (asm_ms * 1e-3) / 1e9) / peak},
            "parity": parity,
            "full_size_check": full_check,
        }
        if world == 1 and not args.no_cpu:
            nxs, nys = REF_NX, REF_NY
            rec = cpu_newton_step_sample(nxs, nys)
            out["cpu_baseline"] = {
                "value": rec["cells"] / rec["total"], "unit": UNIT, "cores": 1, "kind": "port",
                "sample": f"one Newton step (C serial cell assembly {rec['assemble']:.2f}s + SuperLU "
                          f"{rec['lu']:.2f}s + residual {rec['residual']:.2f}s) on RectangleMesh "
                          f"{nxs}x{nys} = {rec['cells']} cells"}
        print(json.dumps(out))
    if strong:
        solver.close()
    if world > 1:
        dist.destroy_process_group()


ROOF_KERNEL = "zebra_ring_kernel (fine grid)"
NCU_CSV = os.path.join(ROOT, "profiles", "ncu_metrics_r2.csv")


def ncu_traffic(mix, f32_op=True):
    """dram read + write bytes per launch of the roofline kernel from the committed `ncu --set full` export
    (profiles/ncu_metrics_r2.csv, written by tools/ncu_export.py from the .ncu-rep of the same kernels), weighted
    by the launch mix of this run (mix = fraction of zero-guess launches).  None when no capture is committed."""
    import csv
    if not os.path.exists(NCU_CSV):
        return None
    rows = {}
    sfx = "_f32op" if f32_op else ""
    for r in csv.DictReader(open(NCU_CSV)):
        if r.get("class") in ("zebra_fine_coupled" + sfx, "zebra_fine_zero" + sfx):
            rows[r["class"]] = float(r["dram_read_bytes"]) + float(r["dram_write_bytes"])
    if len(rows) != 2:
        return None
    return {"bytes_per_launch": mix * rows["zebra_fine_zero" + sfx] + (1.0 - mix) * rows["zebra_fine_coupled" + sfx],
            "source": f"profiles/ncu_metrics_r2.csv (ncu --set full), {mix:.2f} zero-guess / {1 - mix:.2f} coupled launches"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--nx", type=int, default=4096, help="mesh is nx x nx/2 (default: the 16 Mi-cell mesh)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-parity", action="store_true", help="skip the (untimed) parity check after the timed region")
    ap.add_argument("--linear-solver", default="gmres", choices=["gmres", "lu"],
                    help="gmres: multigrid-preconditioned Krylov path (default, the 16 Mi-cell hot path); "
                         "lu: pivoting banded LU on the GPU (needed by --variant galerkin)")
    ap.add_argument("--variant", default="supg", choices=["supg", "galerkin", "c0ip_minus", "c0ip_paper"],
                    help="q-equation variant (default: the SUPG headline workload)")
    ap.add_argument("--weak", action="store_true",
                    help="N > 1: strip-partition a mesh of N x (nx x nx/2 x 2) cells (fixed work per GPU) "
                         "instead of the same 16 Mi-cell mesh")
    ap.add_argument("--instances", action="store_true",
                    help="N>1: independent problem instance per GPU (weak scaling) instead of strip partitioning")
    args = ap.parse_args()
    if args.impl == "reference":
        reference_arm(args)
    else:
        gpu_arm(args)


if __name__ == "__main__":
    main()
