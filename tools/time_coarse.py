"""Phase stamps of the fused coarse V-cycle kernel inside a real solve (GPU).  env CM_COARSE_NCTA / CM_COARSE_N1."""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from closestmolecule_b200 import _capi  # noqa: E402
from test_gpu_mg import _supg_solver  # noqa: E402

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
os.environ["CM_COARSE_DBG"] = "1"
os.environ["CM_GRAPH"] = "0"
asm, make = _supg_solver(nx, nx // 2)
lin = make()
lin.setup()
dx = torch.zeros_like(asm.b)
lin.solve(asm.b, dx, 1e-10, 0.0, 200)
torch.cuda.synchronize()
buf = (C.c_uint64 * (64 * 16 * 8))()
_capi.check(_capi.lib().cm_mg_ring_debug_times(buf, 64 * 16 * 8))
T = np.array(buf, dtype=np.uint64).astype(np.int64)
info = lin.info()
nlev = info["levels"] - info["coarse_cycle_from"]
nst = 1 + (nlev - 1) * 3 + 1 + (nlev - 1) * 5
T = T[:nst]
d = np.diff(T) / 1e3
print("info", info, "its", lin.last_iterations)
print("total us", round(float((T[-1] - T[0]) / 1e3), 1))
k = 0
for lev in range(nlev - 1):
    print(f"down level {info['coarse_cycle_from'] + lev}: c0 {d[k]:.1f} c1 {d[k+1]:.1f} resid+restrict {d[k+2]:.1f}")
    k += 3
print(f"coarsest: {d[k]:.1f}")
k += 1
for lev in range(nlev - 2, -1, -1):
    print(f"up level {info['coarse_cycle_from'] + lev}: prolong {d[k]:.1f} sweeps {d[k+1]:.1f} {d[k+2]:.1f} {d[k+3]:.1f} {d[k+4]:.1f}")
    k += 5
