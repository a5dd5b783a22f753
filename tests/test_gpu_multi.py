"""Strip-partitioned multi-GPU Newton solve (2 ranks on 2 GPUs) against the single-GPU solve: same
Newton and GMRES iteration counts, solutions equal to rounding.  Skipped on boxes with one GPU
(the N>1 host logic is covered on the CPU by tests/test_dist_gloo.py)."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, nx, out, env):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank),
                      MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    os.environ.update(env)
    import torch.distributed as dist
    import closestmolecule_b200 as cm
    from closestmolecule_b200.manufactured import p_exact, q_exact
    from closestmolecule_b200.multigpu import DistributedPQNewton
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
    dev = torch.device("cuda", rank)
    mesh = cm.RectangleMesh(cm.Point(0, 0), cm.Point(1, 1), nx, nx // 2, device=dev)
    P1 = cm.FiniteElement("CG", mesh.ufl_cell(), 1)
    V = cm.FunctionSpace(mesh, cm.MixedElement([P1, P1]))
    bd = cm.MeshFunction("size_t", mesh, 1, 0)
    cm.CompiledSubDomain("near(x[0],0) && on_boundary").mark(bd, 31)
    form = cm.PQPrimalForm.manufactured(V, "supg")
    bcs = [cm.DirichletBC(V.sub(0), p_exact, "on_boundary"), cm.DirichletBC(V.sub(1), q_exact, bd, 31)]
    u = cm.Function(V)
    s = cm.NonlinearVariationalSolver(cm.NonlinearVariationalProblem(form, u, bcs))
    s.parameters["newton_solver"]["linear_solver"] = "gmres"   # the same Krylov path as the distributed solver
    s.parameters["newton_solver"]["absolute_tolerance"] = 1e-8
    s.parameters["newton_solver"]["relative_tolerance"] = 1e-8
    it1, _ = s.solve()
    d = DistributedPQNewton(form, bcs)
    it2, _ = d.solve(atol=1e-8, rtol=1e-8)
    sol = d.gather_solution()
    err = float(torch.linalg.vector_norm(sol - u.vector()) / torch.linalg.vector_norm(u.vector()))
    if rank == 0:
        torch.save(dict(it1=it1, it2=it2, lin1=s.linear_iterations, lin2=d.linear_iterations, err=err,
                        info=d.info()), out)
    d.close()
    dist.destroy_process_group()


# nx = 128 with the fused coarse cycle pushed down to 17-point lines: two distributed levels on the short-line
# kernel; nx = 128 default: everything below the q-sweeps replicated (first replicated level 0); nx = 512: the
# 513-unknown lines of the short-line kernel; nx = 1024: the long-line (ring) kernel pushes / waits inline;
# CM_SOLVER_V1: the first-generation path with explicit exchange launches (same counters and flags)
@pytest.mark.parametrize("nx,env", [(128, {"CM_COARSE_N1": "17"}), (128, {}), (512, {}), (1024, {}),
                                    (1024, {"CM_DIST_LD": "1"}), (512, {"CM_SOLVER_V1": "1"})])
def test_two_gpu_strip_partition_matches_single_gpu(nx, env, tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    out = str(tmp_path / "res.pt")
    mp.spawn(_worker, args=(2, _free_port(), nx, out, env), nprocs=2, join=True)
    r = torch.load(out)
    assert r["it1"] == r["it2"] and r["lin1"] == r["lin2"], r
    assert r["err"] < 1e-10, r
    assert r["info"]["generation"] == (1 if "CM_SOLVER_V1" in env else 2), r
