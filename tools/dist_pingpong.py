"""torchrun --nproc-per-node N tools/dist_pingpong.py: latency of the P2P halo exchange / all-reduce."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import closestmolecule_b200 as cm
from closestmolecule_b200 import _capi
from closestmolecule_b200.multigpu import DistContext

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
mesh = cm.RectangleMesh(cm.Point(0, 0), cm.Point(1, 1), 4096, 256, device=torch.device("cuda", local))
ctx = DistContext(mesh, restart=4)
lib = _capi.lib()
def timeit(f, n=200):
    for _ in range(20): f()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
t = timeit(ctx.exchange_state)
s = torch.ones(8, dtype=torch.float64, device=mesh.device)
t2 = timeit(lambda: ctx.allreduce(s))
t3 = timeit(lambda: s.add_(1.0))
if dist.get_rank() == 0:
    print(f"world {dist.get_world_size()}: halo exchange (2 rows x 8194 doubles) {t:.1f} us, allreduce(8) {t2:.1f} us, "
          f"empty torch kernel {t3:.1f} us", flush=True)
ctx.close()
dist.destroy_process_group()
