"""Experiment: is cuSOLVER's panel LU (the bottleneck of cm_blu_factor) launch-bound on the host, and does a
CUDA-graph replay of the same call help?  Times torch.linalg.lu_factor_ex (cusolverDnDgetrf) eagerly and as a
replayed graph for the panel shapes of the direct solver."""
import sys
import time

import torch

torch.backends.cuda.preferred_linalg_library("cusolver")
dev = torch.device("cuda")
for nb in (528, 1040):
    A0 = torch.randn(nb, 2 * nb, dtype=torch.float64, device=dev)      # column-major 2nb x nb panel = A0.T
    A = A0.clone()
    for _ in range(3):
        LU, piv, info = torch.linalg.lu_factor_ex(A.T.clone().T if False else A0.T)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(10):
        LU, piv, info = torch.linalg.lu_factor_ex(A0.T)
    e1.record()
    t_host = (time.perf_counter() - t0) / 10
    torch.cuda.synchronize()
    eager = e0.elapsed_time(e1) / 10
    msg = f"panel {2 * nb} x {nb}: eager {eager:.3f} ms per getrf (host enqueue {1e3 * t_host:.3f} ms)"
    try:
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(2):
                LU, piv, info = torch.linalg.lu_factor_ex(A0.T)
        torch.cuda.current_stream().wait_stream(s)
        with torch.cuda.graph(g):
            LU, piv, info = torch.linalg.lu_factor_ex(A0.T)
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(10):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        msg += f", graph replay {e0.elapsed_time(e1) / 10:.3f} ms"
    except Exception as ex:  # noqa: BLE001
        msg += f", graph capture failed: {str(ex)[:120]}"
    print(msg, flush=True)
