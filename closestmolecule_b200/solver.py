"""Newton solver front end mirroring NonlinearVariationalProblem / NonlinearVariationalSolver as the
reference scripts configure them (advection_diffusion_convergence.py:207-215,
advection_diffusion_mixed.py:219-226, …MixedPrimal_convergence.py:132-140), with DOLFIN's
NewtonSolver semantics: start from the current u, residual convergence test
||F|| < atol or ||F||/||F0|| < rtol, full steps x <- x - relaxation*dx, RuntimeError on failure.

linear_solver 'mumps' / 'umfpack' / 'lu' / 'default' select the direct path (banded LU with partial
pivoting on the GPU, cm_blu_*) whenever its factors fit the memory budget
parameters['newton_solver']['lu_solver']['max_bytes'] (default: 80 % of the free device memory);
beyond that size (the 16 Mi-cell meshes) they fall back to the Krylov path, like 'gmres' / 'bicgstab'.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _capi
from .assembler import make_assembler
from .forms import PQPrimalForm


class LinearSolveError(RuntimeError):
    pass


class StructuredPQLinearSolver:
    """Field-split GMRES + multigrid on structured meshes through a solver context (cm_pqs_create /
    cm_pqs_ctx_setup / cm_pqs_ctx_solve): the context caches the CUDA graphs of one GMRES iteration across
    the Newton steps and reads the tuning knobs (environment) once, here."""

    def __init__(self, asm, restart=40):
        mesh = asm.mesh
        if not mesh.structured:
            raise LinearSolveError("the P-Q linear solver needs a structured (RectangleMesh-numbered) mesh")
        self.asm, self.restart = asm, int(restart)
        self.nx, self.ny = mesh.nx, mesh.ny
        self.lib = _capi.lib()
        nbytes = int(self.lib.cm_pqs_workspace_bytes(self.nx, self.ny, self.restart))
        if nbytes <= 0:
            raise LinearSolveError("cm_pqs_workspace_bytes failed")
        self.ws = torch.empty(nbytes, dtype=torch.uint8, device=mesh.device)
        self.nbytes = nbytes
        self.ctx = self.lib.cm_pqs_create(self.nx, self.ny, self.restart, _capi.ptr(self.ws), nbytes, None)
        if not self.ctx:
            raise LinearSolveError("cm_pqs_create failed: " + self.lib.cm_last_error().decode())
        self.last_iterations = 0
        self.last_residual = 0.0

    def __del__(self):
        ctx, self.ctx = getattr(self, "ctx", None), None
        if ctx:
            self.lib.cm_pqs_destroy(ctx)

    def info(self):
        """{levels, fused coarse cycle from level, first replicated level, kernel generation, ...}"""
        out = (C.c_int32 * 8)()
        _capi.check(self.lib.cm_pqs_ctx_info(self.ctx, out))
        keys = ("levels", "coarse_cycle_from", "first_replicated_level", "generation", "kernels_per_iteration_graph",
                "coarse_cluster", "graphs", "wide_q_block")
        return dict(zip(keys, list(out)))

    def blocks(self):
        """Where the assembly sweep may write the stencil form of the Jacobian (None: extract it in setup).
        CM_ASSEMBLE_BLOCKS=0 keeps the extraction pass."""
        import os
        if os.environ.get("CM_ASSEMBLE_BLOCKS", "1") == "0" or not getattr(self.asm, "writes_blocks", lambda: False)():
            return None
        if getattr(self, "_blocks", None) is None:
            self._blocks = _capi.PQBlocksC()
            _capi.check(self.lib.cm_pqs_ctx_blocks(self.ctx, C.byref(self._blocks)))
        return self._blocks

    def setup(self, blocks_ready=False):
        if blocks_ready:
            _capi.check(self.lib.cm_pqs_ctx_setup_from_blocks(self.ctx, _capi.stream_ptr()))
            return
        t = self.asm.topo
        _capi.check(self.lib.cm_pqs_ctx_setup(self.ctx, _capi.ptr(t.nrowptr), _capi.ptr(t.ncol),
                                              _capi.ptr(self.asm.vals), _capi.stream_ptr()))

    def solve(self, b, x, rtol, atol, maxit):
        it, rr = C.c_int32(0), C.c_double(0.0)
        rc = self.lib.cm_pqs_ctx_solve(self.ctx, _capi.ptr(b), _capi.ptr(x), rtol, atol, maxit, C.byref(it),
                                       C.byref(rr), _capi.stream_ptr())
        self.last_iterations, self.last_residual = it.value, rr.value
        if rc < 0:
            _capi.check(rc)
        return rc == 0


class BandedLULinearSolver:
    """Sparse LU stand-in for PETScLUSolver('mumps'|'umfpack') (advection_diffusion_convergence.py:212,
    ...MixedPrimal_convergence.py:136): banded LU with partial pivoting, block-tridiagonal organisation
    (cm_blu_factor / cm_blu_solve) + iterative refinement on the assembled CSR.  Works for any
    nonsingular Jacobian (zero q-q diagonal of the Galerkin form included) on any mesh; memory is
    32*(bs*nbn)^2 bytes per block of nbn nodes, nbn = node bandwidth of the renumbered pattern."""

    def __init__(self, asm, max_bytes=None, refinement_steps=1):
        self.asm = asm
        self.lib = _capi.lib()
        self.bs = asm.bs
        t = asm.topo
        self.perm, self.iperm, self.nbn = self.ordering(asm.mesh, t)
        self.nbytes = int(self.lib.cm_blu_workspace_bytes(self.bs, t.nv, self.nbn))
        if self.nbytes <= 0:
            raise LinearSolveError("cm_blu_workspace_bytes failed")
        if max_bytes is None:
            max_bytes = int(0.8 * torch.cuda.mem_get_info(asm.mesh.device)[0])
        if self.nbytes > max_bytes:
            raise LinearSolveError(
                f"the LU factors need {self.nbytes / 2**30:.1f} GiB (node bandwidth {self.nbn}), more than "
                f"the budget of {max_bytes / 2**30:.1f} GiB: use linear_solver='gmres'")
        self.ws = torch.empty(self.nbytes, dtype=torch.uint8, device=asm.mesh.device)
        self.ctx = self.lib.cm_blu_create()
        if not self.ctx:
            raise LinearSolveError("cm_blu_create failed: " + self.lib.cm_last_error().decode())
        self.refinement_steps = int(refinement_steps)
        self._r = torch.empty_like(asm.b)
        self._d = torch.empty_like(asm.b)
        self.last_iterations = 0
        self.last_residual = 0.0

    def __del__(self):
        ctx, self.ctx = getattr(self, "ctx", None), None
        if ctx:
            self.lib.cm_blu_destroy(ctx)

    @staticmethod
    def ordering(mesh, topo):
        """(perm, iperm, nbn): node renumbering with a small bandwidth.  RectangleMesh numbering: the
        grid lines along the shorter direction; anything else: reverse Cuthill-McKee of the node pattern
        (host-side setup, once per mesh)."""
        dev, V = topo.nrowptr.device, topo.nv
        if mesh.structured and mesh.nx > mesh.ny:
            v = torch.arange(V, device=dev)
            perm = (v % (mesh.nx + 1)) * (mesh.ny + 1) + v // (mesh.nx + 1)
        elif mesh.structured:
            perm = torch.arange(V, device=dev)
        else:
            import numpy as np
            import scipy.sparse as sp
            from scipy.sparse.csgraph import reverse_cuthill_mckee
            rp, col = topo.nrowptr.cpu().numpy(), topo.ncol.cpu().numpy()
            G = sp.csr_matrix((np.ones(col.size, dtype=np.int8), col, rp), shape=(V, V))
            order = reverse_cuthill_mckee(G, symmetric_mode=True).astype(np.int64)   # order[pos] = node
            perm = torch.empty(V, dtype=torch.int64, device=dev)
            perm[torch.as_tensor(order, device=dev)] = torch.arange(V, device=dev)
        iperm = torch.empty_like(perm)
        iperm[perm] = torch.arange(V, device=dev)
        deg = (topo.nrowptr[1:] - topo.nrowptr[:-1]).long()
        rows = torch.repeat_interleave(torch.arange(V, device=dev), deg)
        bw = int((perm[rows] - perm[topo.ncol.long()]).abs().max())
        nbn = max(8, -(-bw // 8) * 8)
        return perm.to(torch.int32).contiguous(), iperm.to(torch.int32).contiguous(), nbn

    def setup(self):
        t = self.asm.topo
        info = C.c_int32(0)
        _capi.check(self.lib.cm_blu_factor(
            self.ctx, self.bs, t.nv, self.nbn, _capi.ptr(self.perm), _capi.ptr(self.iperm),
            _capi.ptr(t.nrowptr), _capi.ptr(t.ncol), _capi.ptr(self.asm.vals), _capi.ptr(self.ws),
            self.nbytes, C.byref(info), _capi.stream_ptr()))

    def _solve(self, b, x):
        _capi.check(self.lib.cm_blu_solve(self.ctx, self.bs, self.asm.topo.nv, self.nbn,
                                          _capi.ptr(self.perm), _capi.ptr(b), _capi.ptr(x),
                                          _capi.ptr(self.ws), self.nbytes, _capi.stream_ptr()))

    def _residual(self, b, x):
        t = self.asm.topo
        _capi.check(self.lib.cm_spmv(self.bs, t.nv, _capi.ptr(t.nrowptr), _capi.ptr(t.ncol),
                                     _capi.ptr(self.asm.vals), _capi.ptr(x), _capi.ptr(self._r),
                                     _capi.stream_ptr()))
        torch.sub(b, self._r, out=self._r)
        return self._r

    def solve(self, b, x, rtol=0.0, atol=0.0, maxit=0):
        self._solve(b, x)
        for _ in range(self.refinement_steps):
            self._solve(self._residual(b, x), self._d)
            x.add_(self._d)
        bn = float(torch.linalg.vector_norm(b))
        self.last_residual = float(torch.linalg.vector_norm(self._residual(b, x))) / (bn if bn > 0 else 1.0)
        self.last_iterations = 0
        if not (self.last_residual == self.last_residual):
            raise LinearSolveError("LU solve produced NaN (numerically singular Jacobian)")
        return True


class JacobiKrylovLinearSolver:
    """GMRES(restart) or BiCGStab + (block-)Jacobi on the node CSR layouts, any mesh (cm_krylov_solve /
    cm_bicgstab_solve)."""

    def __init__(self, asm, restart=60, method="gmres"):
        if method not in ("gmres", "bicgstab"):
            raise ValueError(f"unknown Krylov method '{method}'")
        self.asm, self.restart, self.method = asm, int(restart), method
        self.lib = _capi.lib()
        t = asm.topo
        nbytes = int(self.lib.cm_krylov_workspace_bytes(asm.bs, t.nv, self.restart) if method == "gmres"
                     else self.lib.cm_bicgstab_workspace_bytes(asm.bs, t.nv))
        self.ws = torch.empty(nbytes, dtype=torch.uint8, device=asm.mesh.device)
        self.nbytes = nbytes
        self.last_iterations = 0
        self.last_residual = 0.0

    def setup(self):
        pass

    def solve(self, b, x, rtol, atol, maxit):
        t = self.asm.topo
        it, rr = C.c_int32(0), C.c_double(0.0)
        if self.method == "gmres":
            rc = self.lib.cm_krylov_solve(self.asm.bs, t.nv, _capi.ptr(t.nrowptr), _capi.ptr(t.ncol),
                                          _capi.ptr(t.diag_slot), _capi.ptr(self.asm.vals), _capi.ptr(b),
                                          _capi.ptr(x), rtol, atol, maxit, self.restart, _capi.ptr(self.ws),
                                          self.nbytes, C.byref(it), C.byref(rr), _capi.stream_ptr())
        else:
            rc = self.lib.cm_bicgstab_solve(self.asm.bs, t.nv, _capi.ptr(t.nrowptr), _capi.ptr(t.ncol),
                                            _capi.ptr(t.diag_slot), _capi.ptr(self.asm.vals), _capi.ptr(b),
                                            _capi.ptr(x), rtol, atol, maxit, _capi.ptr(self.ws), self.nbytes,
                                            C.byref(it), C.byref(rr), _capi.stream_ptr())
        self.last_iterations, self.last_residual = it.value, rr.value
        if rc < 0:
            _capi.check(rc)
        return rc == 0


class NonlinearVariationalProblem:
    def __init__(self, F, u, bcs=None, J=None):
        self.form, self.u = F, u
        if bcs is None:
            bcs = []
        self.bcs = list(bcs) if isinstance(bcs, (list, tuple)) else [bcs]


_LU_ALIASES = ("mumps", "umfpack", "lu", "default", "superlu", "petsc")


class NonlinearVariationalSolver:
    def __init__(self, problem):
        self.problem = problem
        self.parameters = {
            "nonlinear_solver": "newton",
            "newton_solver": self._defaults(),
            "snes_solver": self._defaults(),
        }
        self.asm = make_assembler(problem.form)
        self._linear = None
        self._linear_key = None
        self.history = []
        self.linear_iterations = []
        self.phase_events = None   # set to [] to collect CUDA events around the phases of each step

    @staticmethod
    def _defaults():
        # DOLFIN NewtonSolver defaults + the Krylov controls of this implementation
        return {"linear_solver": "default", "absolute_tolerance": 1e-10, "relative_tolerance": 1e-9,
                "maximum_iterations": 50, "relaxation_parameter": 1.0, "error_on_nonconvergence": True,
                "report": False,
                "lu_solver": {"max_bytes": None, "refinement_steps": 1, "allow_krylov_fallback": False},
                "krylov_solver": {"relative_tolerance": 1e-11, "absolute_tolerance": 0.0,
                                  "maximum_iterations": 400, "restart": 40}}

    def _linear_solver(self, prm):
        key = (prm["linear_solver"], prm["krylov_solver"]["restart"], prm["lu_solver"]["max_bytes"],
               prm["lu_solver"]["refinement_steps"], prm["lu_solver"].get("allow_krylov_fallback", False))
        if self._linear is not None and key != self._linear_key:
            self._linear = None                           # the parameters changed since the solver was built
        if self._linear is None:
            self._linear_key = key
            name = prm["linear_solver"]
            if name in _LU_ALIASES:
                try:
                    self._linear = BandedLULinearSolver(self.asm, prm["lu_solver"]["max_bytes"],
                                                        prm["lu_solver"]["refinement_steps"])
                    return self._linear
                except LinearSolveError as e:
                    # too large for a direct solve.  The caller asked for an exact solve: switching to the
                    # Krylov path (rtol-limited) needs an explicit opt-in, and is announced
                    if not prm["lu_solver"].get("allow_krylov_fallback", False):
                        raise LinearSolveError(
                            f"linear_solver='{name}': {e}; set parameters['newton_solver']['lu_solver']"
                            "['allow_krylov_fallback'] = True to solve with GMRES to "
                            "krylov_solver['relative_tolerance'] instead") from e
                    import warnings
                    warnings.warn(f"linear_solver='{name}' falls back to the Krylov path: {e}")
                    self.lu_fallback_reason = str(e)
            elif name not in ("gmres", "bicgstab"):
                raise ValueError(f"unknown linear_solver '{name}'")
            restart = prm["krylov_solver"]["restart"]
            if isinstance(self.problem.form, PQPrimalForm):
                # the field-split / multigrid preconditioner is applied inexactly (fixed sweeps), which GMRES tolerates
                # and BiCGStab's short recurrences do not reward: 'bicgstab' selects the same GMRES path here
                self._linear = StructuredPQLinearSolver(self.asm, restart)
            else:
                self._linear = JacobiKrylovLinearSolver(self.asm, restart,
                                                        "bicgstab" if name == "bicgstab" else "gmres")
        return self._linear

    def _prepare(self):
        kind = self.parameters["nonlinear_solver"]
        prm = self.parameters[kind + "_solver"]
        if getattr(self, "_bc", None) is None:
            self._bc = self.asm._bc_tables(self.problem.bcs)
            self._dx = torch.empty_like(self.problem.u.vector())
        return prm, self._linear_solver(prm)

    def refresh_bcs(self):
        """Re-evaluate the Dirichlet tables (time-dependent boundary data) before the next assembly."""
        self._bc = None

    def residual_norm(self):
        """||F(u)||_2 with the Dirichlet rows set to u - g (NewtonSolver's convergence measure)."""
        prm, _ = self._prepare()
        u = self.problem.u.vector()
        self.asm.assemble(u, want_jac=False)
        self.asm.apply_bcs(self._bc, u, with_matrix=False)
        return self._norm(self.asm.b)

    def _vec_scratch(self):
        if getattr(self, "_scr", None) is None:
            lib, dev = _capi.lib(), self.problem.u.vector().device
            self._scr = torch.empty(int(lib.cm_vec_scratch_doubles()) + 8, dtype=torch.float64, device=dev)
            self._sc = torch.zeros(4, dtype=torch.float64, device=dev)
        return self._scr, self._sc

    def _norm(self, v):
        """||v||_2 by the library's fused two-stage dot kernel (deterministic) + one 8-byte read."""
        scr, sc = self._vec_scratch()
        _capi.check(_capi.lib().cm_dot_batch(v.numel(), _capi.ptr(v), v.numel(), 1, _capi.ptr(v), _capi.ptr(sc),
                                             _capi.ptr(scr), _capi.stream_ptr()))
        return float(sc[0].sqrt())

    def _update(self, u, dx, relax):
        """u <- u - relax*dx (the library's fused axpy kernel)."""
        scr, sc = self._vec_scratch()
        sc[1] = relax
        _capi.check(_capi.lib().cm_axpy_batch(u.numel(), _capi.ptr(dx), dx.numel(), 1, _capi.ptr(sc[1:2]), -1.0,
                                              _capi.ptr(u), None, _capi.ptr(scr), _capi.stream_ptr()))

    def newton_step(self, compute_residual=True, on_updated=None):
        """One Newton iteration at the current u: fused J+F assembly, Dirichlet rows, linear-solver
        setup, Krylov solve, u <- u - relaxation*dx, then (optionally) the new residual norm.
        on_updated: called right after the update has been enqueued (the new iterate is final from there on in
        stream order) -- e.g. to start its device-to-host copy on a side stream while the residual is evaluated."""
        prm, lin = self._prepare()
        kp = prm["krylov_solver"]
        asm, u = self.asm, self.problem.u.vector()
        marks = [] if self.phase_events is not None else None

        def mark():
            if marks is not None:
                e = torch.cuda.Event(enable_timing=True)
                e.record()
                marks.append(e)
        mark()
        # structured P-Q path: the sweep writes the solver's stencil blocks next to the CSR values (no extraction pass)
        blk = lin.blocks() if hasattr(lin, "blocks") else None
        if blk is not None:
            asm.assemble(u, want_jac=True, blocks=blk)
            mark()
            asm.apply_bcs(self._bc, u, with_matrix=True, blocks=blk)
            lin.setup(blocks_ready=True)
        else:
            asm.assemble(u, want_jac=True)
            mark()
            asm.apply_bcs(self._bc, u, with_matrix=True)
            lin.setup()
        mark()
        ok = lin.solve(asm.b, self._dx, kp["relative_tolerance"], kp["absolute_tolerance"],
                       kp["maximum_iterations"])
        mark()
        self.linear_iterations.append(lin.last_iterations)
        if not ok:
            raise LinearSolveError(
                f"Krylov solver did not converge in {lin.last_iterations} iterations "
                f"(relative residual {lin.last_residual:.3e})")
        self._update(u, self._dx, prm["relaxation_parameter"])
        if on_updated is not None:
            on_updated()
        r = self.residual_norm() if compute_residual else None
        mark()
        if marks is not None:   # (assemble J+F, BC + solver setup, Krylov solve, update + residual)
            self.phase_events.append(marks)
        return r

    def solve(self):
        prm, lin = self._prepare()
        atol, rtol = prm["absolute_tolerance"], prm["relative_tolerance"]
        maxit = prm["maximum_iterations"]
        r0 = self.residual_norm()
        r = r0
        self.history = [r0]
        self.linear_iterations = []
        it = 0
        conv = r < atol or (r0 > 0 and r / r0 < rtol)
        while not conv and it < maxit:
            r = self.newton_step()
            it += 1
            self.history.append(r)
            if prm["report"]:
                print(f"Newton iteration {it}: r (abs) = {r:.3e} (tol = {atol:.3e}) "
                      f"r (rel) = {r / r0:.3e} (tol = {rtol:.3e})  [linear its {lin.last_iterations}]")
            if r != r:
                raise RuntimeError("Newton solver diverged (NaN residual)")
            conv = r < atol or r / r0 < rtol
        if not conv and prm["error_on_nonconvergence"]:
            raise RuntimeError("Newton solver did not converge because maximum number of iterations "
                               "reached")  # DOLFIN's message text
        return it, conv


class LinearVariationalSolver:
    """solve(a == L, u, bcs) for the linear scalar form (SphericalLaplace_convergence.py:135): one
    Newton step from u = 0 of the affine residual."""

    def __init__(self, form, u, bcs):
        self.inner = NonlinearVariationalSolver(NonlinearVariationalProblem(form, u, bcs))
        self.inner.parameters["newton_solver"]["maximum_iterations"] = 1
        self.inner.parameters["newton_solver"]["error_on_nonconvergence"] = False

    def solve(self):
        self.inner.problem.u.vector().zero_()
        return self.inner.solve()


def solve(form, u, bcs=None):
    """solve(FF == 0, u, bcs) (SphericalAdvectionDiffusionReaction_convergence.py:112 comment) with
    DOLFIN's default Newton parameters."""
    return NonlinearVariationalSolver(NonlinearVariationalProblem(form, u, bcs)).solve()
