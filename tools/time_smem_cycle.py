"""Phase stamps (%globaltimer) of the shared-memory resident coarse cycle inside a real solve (GPU)."""
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
buf = (C.c_uint64 * 64)()
n = _capi.lib().cm_mg_smem_debug_times(buf, 64)
T = np.array(buf, dtype=np.uint64).astype(np.int64)[:n]
print("info", lin.info(), "its", lin.last_iterations, "stamps", n)
print("total us", round(float((T[-1] - T[0]) / 1e3), 2))
print("phases us (start | copies issued | level 0 resident | per level: pre, restrict ... | coarsest | per level: post ... | stored):")
print(" ".join(f"{v:.2f}" for v in np.diff(T) / 1e3))
