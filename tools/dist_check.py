"""torchrun --nproc-per-node N tools/dist_check.py [nx]: the strip-partitioned Newton solve against the
single-GPU solve of the same problem (run on every rank redundantly)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import closestmolecule_b200 as cm
from closestmolecule_b200.manufactured import p_exact, q_exact
from closestmolecule_b200.multigpu import DistributedPQNewton

rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
nx = int(sys.argv[1]) if len(sys.argv) > 1 else 256
ny = nx // 2
dev = torch.device("cuda", local)
mesh = cm.RectangleMesh(cm.Point(0, 0), cm.Point(1, 1), nx, ny, device=dev)
P1 = cm.FiniteElement("CG", mesh.ufl_cell(), 1)
V = cm.FunctionSpace(mesh, cm.MixedElement([P1, P1]))
bd = cm.MeshFunction("size_t", mesh, 1, 0)
cm.CompiledSubDomain("near(x[0],0) && on_boundary").mark(bd, 31)
form = cm.PQPrimalForm.manufactured(V, "supg")
bcs = [cm.DirichletBC(V.sub(0), p_exact, "on_boundary"), cm.DirichletBC(V.sub(1), q_exact, bd, 31)]
# single-GPU reference
u = cm.Function(V)
s = cm.NonlinearVariationalSolver(cm.NonlinearVariationalProblem(form, u, bcs))
s.parameters["newton_solver"]["linear_solver"] = "gmres"   # the same Krylov path as the distributed solver
s.parameters["newton_solver"]["absolute_tolerance"] = 1e-8
s.parameters["newton_solver"]["relative_tolerance"] = 1e-8
it1, _ = s.solve()
ref = u.vector().clone()
# distributed
d = DistributedPQNewton(form, bcs)
it2, _ = d.solve(atol=1e-8, rtol=1e-8)
sol = d.gather_solution()
err = float(torch.linalg.vector_norm(sol - ref) / torch.linalg.vector_norm(ref))
print(f"rank {rank}: single-GPU Newton {it1} its (linear {s.linear_iterations}), distributed {it2} its "
      f"(linear {d.linear_iterations}), rel L2 diff {err:.3e}, residual history {['%.3e' % h for h in d.history]}", flush=True)
assert it1 == it2 and err < 1e-8
d.close()
dist.destroy_process_group()
