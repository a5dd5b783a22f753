"""Ad-hoc timing of the assembly and SpMV kernels on the 16 Mi-cell mesh (development aid)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import closestmolecule_b200 as cm
from closestmolecule_b200 import _capi
from closestmolecule_b200.assembler import PQAssembler

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
ny = nx // 2
t0 = time.time()
mesh = cm.RectangleMesh(cm.Point(0, 0), cm.Point(1, 1), nx, ny)
V = cm.FunctionSpace(mesh, cm.MixedElement([cm.FiniteElement("CG", "triangle", 1)] * 2))
form = cm.PQPrimalForm.manufactured(V, "supg")
asm = PQAssembler(form)
torch.cuda.synchronize()
print("setup s", time.time() - t0, "cells", mesh.num_cells(), "nnz", asm.topo.nnz, "mem GB",
      torch.cuda.max_memory_allocated() / 1e9)
X = mesh.coords
u = torch.empty(2 * mesh.num_vertices(), dtype=torch.float64, device=mesh.device)
u[0::2] = torch.sin(3.14159 * X[:, 0]) * torch.sin(3.14159 * X[:, 1])
u[1::2] = 1.5 + torch.cos(3.14159 * X[:, 0] * X[:, 1])

def timeit(f, n=10):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

C = mesh.num_cells(); Vn = mesh.num_vertices(); nnz = asm.topo.nnz
t = timeit(lambda: asm.assemble(u, True))
bytes_asm = 12 * C + 16 * Vn + 16 * Vn + 8 * 4 * nnz + 16 * Vn
print(f"assemble J+F: {t:.3f} ms  {C / t / 1e6:.2f} Gcells/s  algGB/s {bytes_asm / t / 1e6:.1f}")
t = timeit(lambda: asm.assemble(u, False))
print(f"assemble F only: {t:.3f} ms")
x = torch.randn_like(u); y = torch.empty_like(u)
tp = asm.topo
f = lambda: _capi.check(_capi.lib().cm_spmv(2, tp.nv, _capi.ptr(tp.nrowptr), _capi.ptr(tp.ncol), _capi.ptr(asm.vals), _capi.ptr(x), _capi.ptr(y), _capi.stream_ptr()))
t = timeit(f)
bytes_spmv = 8 * 4 * nnz + 4 * nnz + 4 * Vn + 32 * Vn
print(f"spmv: {t:.3f} ms  GB/s {bytes_spmv / t / 1e6:.1f}")
a = torch.empty(1 << 29, dtype=torch.float64, device='cuda'); bq = torch.empty_like(a)
t = timeit(lambda: bq.copy_(a))
print(f"copy 4GiB: {t:.3f} ms GB/s {2 * a.numel() * 8 / t / 1e6:.1f}")
