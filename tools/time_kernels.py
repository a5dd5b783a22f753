"""Micro-benchmark of the structured-solver building blocks on a (nx, nx/2) grid (development aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from closestmolecule_b200 import _capi

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
ny = nx // 2
n1 = nx + 1
n = n1 * (ny + 1)
dev = torch.device("cuda")
lib = _capi.lib()
torch.manual_seed(0)
A = torch.zeros(7, ny + 1, n1, dtype=torch.float64, device=dev)
off = [(-1, -1), (0, -1), (-1, 0), (0, 0), (1, 0), (0, 1), (1, 1)]
for k, (dx, dy) in enumerate(off):
    v = -0.1 - 0.05 * torch.rand(ny + 1, n1, dtype=torch.float64, device=dev) if k != 3 else 1.0 + torch.rand(ny + 1, n1, dtype=torch.float64, device=dev)
    if dx == -1: v[:, 0] = 0
    if dx == 1: v[:, -1] = 0
    if dy == -1: v[0, :] = 0
    if dy == 1: v[-1, :] = 0
    A[k] = v
A = A.reshape(7, n).contiguous()
lf, iu, cu = (torch.empty(n, dtype=torch.float64, device=dev) for _ in range(3))
err = torch.zeros(1, dtype=torch.int32, device=dev)
b = torch.randn(n, dtype=torch.float64, device=dev)
x = torch.zeros(n, dtype=torch.float64, device=dev)
r = torch.empty_like(x)
tmp = torch.empty_like(x)
P = _capi.ptr
st = _capi.stream_ptr()

def timeit(f, reps=20):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    big = torch.empty(64 << 20, dtype=torch.float64, device=dev)
    t = 0.0
    for _ in range(reps):
        big.fill_(1.0)          # flush L2 (512 MB > 126 MB)
        e0.record(); f(); e1.record(); torch.cuda.synchronize()
        t += e0.elapsed_time(e1)
    return t / reps * 1e3

t = timeit(lambda: _capi.check(lib.cm_stencil_line_factor(nx, ny, P(A), P(lf), P(iu), P(cu), P(err), st)))
print(f"line_factor: {t:.1f} us  ({(3 * 8 + 3 * 8) * n / t / 1e6:.0f} GB/s)")
assert int(err.item()) == 0
for colour in (0, 1):
    for xz in (1, 0):
        t = timeit(lambda: _capi.check(lib.cm_stencil_zebra(nx, ny, colour, P(A), P(lf), P(iu), P(cu), P(b), P(x), xz, P(tmp), st)))
        byts = (n / 2) * (8 + 24 + 8 + (0 if xz else 32 + 16))
        print(f"zebra colour {colour} x_zero {xz}: {t:.1f} us  ({byts / t / 1e6:.0f} GB/s algorithmic)")
t = timeit(lambda: _capi.check(lib.cm_stencil_residual(nx, ny, P(A), P(b), P(x), P(r), st)))
print(f"residual: {t:.1f} us ({80 * n / t / 1e6:.0f} GB/s)")
# correctness of one zebra sweep against the residual: after solving lines of colour c, the residual on those lines is ~0
x.zero_()
_capi.check(lib.cm_stencil_zebra(nx, ny, 0, P(A), P(lf), P(iu), P(cu), P(b), P(x), 1, P(tmp), st))
_capi.check(lib.cm_stencil_zebra(nx, ny, 1, P(A), P(lf), P(iu), P(cu), P(b), P(x), 0, P(tmp), st))
_capi.check(lib.cm_stencil_residual(nx, ny, P(A), P(b), P(x), P(r), st))
R = r.reshape(ny + 1, n1)
print("max |r| on odd lines after the sweep:", float(R[1::2].abs().max()), " even lines:", float(R[0::2].abs().max()))
if os.environ.get("CM_ZEBRA_DEBUG") and int(os.environ["CM_ZEBRA_DEBUG"]) & 32:
    import ctypes, numpy as np
    lib.cm_dev_zebra_times.argtypes = [ctypes.c_void_p, ctypes.c_int32]
    for xz in (1, 0):
        _capi.check(lib.cm_stencil_zebra(nx, ny, 0, P(A), P(lf), P(iu), P(cu), P(b), P(x), xz, P(tmp), st))
        torch.cuda.synchronize()
        buf = np.zeros(1025 * 8, dtype=np.uint64)
        lib.cm_dev_zebra_times(buf.ctypes.data, buf.size)
        tt = buf.reshape(-1, 8)[:, :6].astype(np.int64)
        t0 = tt[:, 0].min()
        d = np.diff(tt, axis=1)
        print("x_zero", xz, "phase durations ns (mean) [hybrid: wait, LDS+issue, fwd, bwd, store]:", d.mean(0).round(0), " CTA total", (tt[:, 5] - tt[:, 0]).mean().round(0),
              " kernel span", tt[:, 5].max() - t0, " CTA start spread (percentiles 50/90/100)", np.percentile(tt[:, 0] - t0, [50, 90, 100]).round(0))
