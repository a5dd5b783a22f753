"""Strip-partitioned multi-GPU Newton solve of the P-Q system.

One process per GPU (torchrun).  The structured mesh is cut into strips of whole vertex rows (RCB
along r1: the r2-lines of the line smoother stay intact); every rank holds GLOBALLY indexed arrays
and computes its own rows.  torch.distributed is used only to bootstrap (exchange of CUDA-IPC
handles, barriers); all data-path communication -- halo rows, the coarse-level all-gather and the
Krylov all-reduces -- is done by the library's peer-to-peer kernels over NVLink.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.distributed as dist

from . import _capi
from .assembler import PQAssembler
from .solver import LinearSolveError


class _RawCuda:
    """Minimal __cuda_array_interface__ wrapper so torch can view library-owned device memory."""

    def __init__(self, ptr, numel, owner):
        self.__cuda_array_interface__ = {"shape": (numel,), "typestr": "<f8", "data": (ptr, False),
                                         "version": 2}
        self._owner = owner


def strip_partition(ny, world):
    """Strip partition of the ny+1 vertex rows of a structured mesh over `world` ranks (RCB along r1 keeps the
    r2-lines of the line solver intact): row_begin[r] .. row_begin[r+1] are the rows rank r owns.  Strip
    boundaries are multiples of 8 rows, so that they stay EVEN rows on the three finest multigrid levels (the
    zebra colouring of the inline halo protocol, csrc/cm_halo.cuh, needs even strip starts on every distributed
    level) and every strip keeps >= 4 rows down to the first replicated level; the last rank takes the rest."""
    nrows = ny + 1
    per = (nrows // world) // 8 * 8
    if per < 8:
        raise ValueError("too few vertex rows per rank (need >= 8)")
    return [r * per for r in range(world)] + [nrows]


class DistContext:
    def __init__(self, mesh, restart=40):
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed must be initialised (torchrun)")
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        if self.world > 8:
            raise ValueError("at most 8 GPUs of one node")
        self.mesh, self.restart = mesh, int(restart)
        self.nx, self.ny = mesh.nx, mesh.ny
        self.lib = _capi.lib()
        self.row_begin = strip_partition(self.ny, self.world)
        self.rb, self.re = self.row_begin[self.rank], self.row_begin[self.rank + 1]
        self.nbytes = int(self.lib.cm_pqs_workspace_bytes(self.nx, self.ny, self.restart))
        self.base = self.lib.cm_ipc_alloc(self.nbytes)
        if not self.base:
            raise _capi.CMError(self.lib.cm_last_error().decode())
        handle = (C.c_ubyte * 64)()
        _capi.check(self.lib.cm_ipc_get_handle(self.base, handle))
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(handle))
        self.D = _capi.DistC()
        self.D.rank, self.D.world = self.rank, self.world
        self._peers = []
        for r, h in enumerate(handles):
            if r == self.rank:
                self.D.peer_base[r] = self.base
            else:
                buf = (C.c_ubyte * 64).from_buffer_copy(h)
                p = self.lib.cm_ipc_open(buf)
                if not p:
                    raise _capi.CMError(self.lib.cm_last_error().decode())
                self.D.peer_base[r] = p
                self._peers.append(p)
        for r in range(self.world + 1):
            self.D.row_begin[r] = self.row_begin[r]
        self.state_off = int(self.lib.cm_pqs_state_offset(self.nx, self.ny, self.restart))
        V = mesh.num_vertices()
        self.u = torch.as_tensor(_RawCuda(self.base + self.state_off, 2 * V, self), device=mesh.device)
        dist.barrier()

    def exchange_state(self):
        n1 = self.nx + 1
        _capi.check(self.lib.cm_dist_exchange_rows(C.byref(self.D), self.state_off, 2 * n1, self.rb, self.re,
                                                   _capi.stream_ptr()))

    def allgather_state(self):
        n1 = self.nx + 1
        _capi.check(self.lib.cm_dist_allgather_rows(C.byref(self.D), self.state_off, 2 * n1, self.rb, self.re,
                                                    _capi.stream_ptr()))

    def allreduce(self, t):
        _capi.check(self.lib.cm_dist_allreduce_sum(C.byref(self.D), _capi.ptr(t), t.numel(), _capi.stream_ptr()))
        return t

    def close(self):
        torch.cuda.synchronize()
        dist.barrier()
        for p in self._peers:
            self.lib.cm_ipc_close(p)
        self._peers = []
        dist.barrier()
        if self.base:
            self.lib.cm_ipc_free(self.base)
            self.base = None


class DistributedPQNewton:
    """NewtonSolver semantics on a strip-partitioned structured mesh."""

    def __init__(self, form, bcs, restart=40, rtol_lin=1e-11, maxit_lin=400):
        mesh = form.space.mesh
        self.form, self.mesh = form, mesh
        self.ctx = DistContext(mesh, restart)
        self.asm = PQAssembler(form)
        if self.asm.Kq is not None or self.asm.topo.with_interior_facets:
            raise LinearSolveError("the strip-partitioned path handles the 7-point (cell) pattern only")
        self.lib = _capi.lib()
        self.rtol_lin, self.maxit_lin = rtol_lin, maxit_lin
        ctx = self.ctx
        n1 = mesh.nx + 1
        dofs, vals, diag = self.asm._bc_tables(bcs)
        row = (dofs.long() // 2) // n1
        own = (row >= ctx.rb) & (row < ctx.re)
        self.bc = (dofs[own].contiguous(), vals[own].contiguous(), diag[own].contiguous())
        self.dx = torch.zeros_like(self.asm.b)
        self.u = ctx.u
        self.own = slice(2 * ctx.rb * n1, 2 * ctx.re * n1)
        self.scal = torch.zeros(1, dtype=torch.float64, device=mesh.device)
        self.pqs = self.lib.cm_pqs_create(ctx.nx, ctx.ny, ctx.restart, ctx.base, ctx.nbytes, C.byref(ctx.D))
        if not self.pqs:
            raise LinearSolveError("cm_pqs_create failed: " + self.lib.cm_last_error().decode())
        self.linear_iterations = []
        self.history = []
        self.phase_events = None

    def _blocks(self):
        """the solver's stencil blocks, written by the assembly sweep itself (CM_ASSEMBLE_BLOCKS=0: extracted in setup)"""
        import os
        if os.environ.get("CM_ASSEMBLE_BLOCKS", "1") == "0":
            return None
        if getattr(self, "_blk", None) is None:
            self._blk = _capi.PQBlocksC()
            _capi.check(self.lib.cm_pqs_ctx_blocks(self.pqs, C.byref(self._blk)))
        return self._blk

    def _assemble(self, want_jac):
        ctx, asm, form, mesh = self.ctx, self.asm, self.form, self.mesh
        ctx.exchange_state()
        p_old = form.p_old.vector() if form.p_old is not None else None
        blk = self._blocks() if want_jac else None
        if blk is not None:
            _capi.check(self.lib.cm_assemble_pq_cells_structured_blocks(
                C.byref(form.params), mesh.nx, mesh.ny, _capi.ptr(mesh.coords), _capi.ptr(asm.topo.nrowptr),
                _capi.ptr(self.u), _capi.ptr(p_old), _capi.ptr(asm.load), _capi.ptr(asm.vals), _capi.ptr(asm.b),
                C.byref(blk), ctx.rb, ctx.re, _capi.stream_ptr()))
        else:
            _capi.check(self.lib.cm_assemble_pq_cells_structured_rows(
                C.byref(form.params), mesh.nx, mesh.ny, _capi.ptr(mesh.coords), _capi.ptr(asm.topo.nrowptr),
                _capi.ptr(self.u), _capi.ptr(p_old), _capi.ptr(asm.load), _capi.ptr(asm.vals) if want_jac else None,
                _capi.ptr(asm.b), ctx.rb, ctx.re, _capi.stream_ptr()))
        asm.apply_bcs(self.bc, self.u, with_matrix=want_jac, blocks=blk)
        return blk is not None

    def residual_norm(self):
        self._assemble(False)
        self.scal[0] = torch.sum(self.asm.b[self.own] ** 2)
        self.ctx.allreduce(self.scal)
        return float(self.scal.sqrt())

    def newton_step(self, compute_residual=True, on_updated=None):
        """on_updated: called right after the update of the owned rows has been enqueued (see solver.py)."""
        ctx, asm = self.ctx, self.asm
        marks = [] if self.phase_events is not None else None

        def mark():
            if marks is not None:
                e = torch.cuda.Event(enable_timing=True)
                e.record()
                marks.append(e)
        mark()
        blocks_ready = self._assemble(True)
        mark()
        t = asm.topo
        if blocks_ready:
            _capi.check(self.lib.cm_pqs_ctx_setup_from_blocks(self.pqs, _capi.stream_ptr()))
        else:
            _capi.check(self.lib.cm_pqs_ctx_setup(self.pqs, _capi.ptr(t.nrowptr), _capi.ptr(t.ncol), _capi.ptr(asm.vals),
                                                  _capi.stream_ptr()))
        mark()
        it, rr = C.c_int32(0), C.c_double(0.0)
        rc = self.lib.cm_pqs_ctx_solve(self.pqs, _capi.ptr(asm.b), _capi.ptr(self.dx), self.rtol_lin, 0.0,
                                       self.maxit_lin, C.byref(it), C.byref(rr), _capi.stream_ptr())
        self.linear_iterations.append(it.value)
        if rc < 0:
            _capi.check(rc)
        if rc != 0:
            raise LinearSolveError(f"GMRES did not converge in {it.value} iterations (rel. residual {rr.value:.3e})")
        mark()
        self.u[self.own] -= self.dx[self.own]
        if on_updated is not None:
            on_updated()
        r = self.residual_norm() if compute_residual else None
        mark()
        if marks is not None:
            self.phase_events.append(marks)
        return r

    def solve(self, atol=1e-10, rtol=1e-9, maxit=50):
        r0 = r = self.residual_norm()
        self.history = [r0]
        it = 0
        conv = r < atol or (r0 > 0 and r / r0 < rtol)
        while not conv and it < maxit:
            r = self.newton_step()
            it += 1
            self.history.append(r)
            conv = r < atol or r / r0 < rtol
        if not conv:
            raise RuntimeError("Newton solver did not converge because maximum number of iterations reached")
        return it, conv

    def gather_solution(self):
        """Full solution vector on every rank."""
        self.ctx.allgather_state()
        torch.cuda.current_stream().synchronize()
        return self.u

    def info(self):
        out = (C.c_int32 * 8)()
        _capi.check(self.lib.cm_pqs_ctx_info(self.pqs, out))
        keys = ("levels", "coarse_cycle_from", "first_replicated_level", "generation", "kernels_per_iteration_graph",
                "coarse_cluster", "graphs", "wide_q_block")
        return dict(zip(keys, list(out)))

    def close(self):
        torch.cuda.synchronize()
        if self.pqs:
            self.lib.cm_pqs_destroy(self.pqs)
            self.pqs = None
        self.ctx.close()
