"""GPU microbenchmark of the fused zebra half sweep (cm_mg_zebra) on long lines: CUDA-event time per launch
for the zero-guess and the coupled variant, fraction of the measured HBM peak (40 / 80 bytes per vertex), and
the in-kernel phase stamps.   python tools/time_ring.py [nx ny]   (env: CM_RING_S, CM_RING_DBG)"""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from closestmolecule_b200 import _capi  # noqa: E402

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
ny = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
reps = 20
dev = torch.device("cuda")
n1, n = nx + 1, (nx + 1) * (ny + 1)
g = torch.Generator(device=dev).manual_seed(1)
A = -(0.1 + 0.05 * torch.rand(7, n, generator=g, device=dev, dtype=torch.float64))
A[3] = 1.0 + torch.rand(n, generator=g, device=dev, dtype=torch.float64)
Av = A.view(7, ny + 1, n1)
Av[0, :, 0] = 0; Av[2, :, 0] = 0; Av[4, :, -1] = 0; Av[6, :, -1] = 0
Av[0, 0, :] = 0; Av[1, 0, :] = 0; Av[5, -1, :] = 0; Av[6, -1, :] = 0
lf, iu, cu = (torch.empty(n, dtype=torch.float64, device=dev) for _ in range(3))
err = torch.zeros(1, dtype=torch.int32, device=dev)
b = torch.randn(n, generator=g, device=dev, dtype=torch.float64)
x = torch.zeros(n, dtype=torch.float64, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
lib, P, st = _capi.lib(), _capi.ptr, _capi.stream_ptr()
_capi.check(lib.cm_stencil_line_factor(nx, ny, P(A), P(lf), P(iu), P(cu), P(err), st))
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
    os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6537.6
out = {"nx": nx, "ny": ny, "S": os.environ.get("CM_RING_S")}
for name, colour, xz, bpv in (("zero", 0, 1, 40.0), ("coupled", 1, 0, 80.0), ("coupled0", 0, 0, 80.0)):
    ts = []
    for r in range(reps + 3):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _capi.check(lib.cm_mg_zebra(nx, ny, colour, P(A), P(lf), P(iu), P(cu), P(b), P(x), xz, st))
        e1.record()
        torch.cuda.synchronize()
        if r >= 3:
            ts.append(e0.elapsed_time(e1) * 1e3)
    nl = (ny + 2 - colour) // 2
    by = nl * n1 * bpv
    us = float(np.median(ts))
    out[name] = {"us": round(us, 1), "min_us": round(min(ts), 1), "gbs": round(by / us / 1e3, 1),
                 "frac": round(by / us / 1e3 / peak, 3)}
if os.environ.get("CM_RING_DBG"):
    buf = (C.c_uint64 * (64 * 16 * 8))()
    _capi.check(lib.cm_mg_ring_debug_times(buf, 64 * 16 * 8))
    T = np.array(buf, dtype=np.uint64).reshape(64, 16, 8).astype(np.int64)
    d = np.diff(T[:8, :6, :], axis=2) / 1e3      # us between consecutive stamps
    out["phases_us (x-lower, rhs streams, factors, fwd, bwd, barrier, store) mean over 8 CTAs x lines 1..5"] = [
        round(float(v), 2) for v in d[:, 1:6, :].mean(axis=(0, 1))]
    out["line_period_us"] = round(float(np.diff(T[:8, :6, 0], axis=1).mean() / 1e3), 2)
print(json.dumps(out))
