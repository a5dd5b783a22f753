"""In-situ kernel timeline of one Newton step (torch.profiler / CUPTI activity records: production mode, CUDA
graphs and the second stream on), per rank.  Prints, for the kernels of ONE mid-solve GMRES iteration, name,
grid, start offset, duration and the gap to the previous kernel on the same stream, plus per-kernel-class
totals over the whole step.
usage (1 GPU):  python tools/trace_step.py [--nx 4096] [--out gpurun_out/trace.json]
      (N GPUs): python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/trace_step.py
"""
import argparse
import collections
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def short(n):
    n = n.replace("(anonymous namespace)::", "")
    n = re.sub(r"\(.*", "", n)
    m = re.match(r"(void )?(cm::)?([A-Za-z_0-9:]+)(<[^>]*>)?", n)
    return (m.group(3) + (m.group(4) or "")) if m else n


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nx", type=int, default=4096)
    ap.add_argument("--out", default="")
    ap.add_argument("--iteration", type=int, default=5, help="GMRES iteration whose kernels are listed")
    args = ap.parse_args()
    # CUPTI records the body of a WHILE graph node once, not per loop trip: trace the one-graph-per-iteration form
    os.environ.setdefault("CM_GRAPH_WHILE", "0")
    import torch
    import torch.distributed as dist
    import closestmolecule_b200 as cm
    from closestmolecule_b200.manufactured import p_exact, q_exact

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    nx, ny = args.nx, args.nx // 2
    mesh = cm.RectangleMesh(cm.Point(0, 0), cm.Point(1, 1), nx, ny, device=dev)
    P1 = cm.FiniteElement("CG", mesh.ufl_cell(), 1)
    V = cm.FunctionSpace(mesh, cm.MixedElement([P1, P1]))
    bd = cm.MeshFunction("size_t", mesh, 1, 0)
    cm.CompiledSubDomain("near(x[0],0) && on_boundary").mark(bd, 31)
    form = cm.PQPrimalForm.manufactured(V, "supg")
    bcs = [cm.DirichletBC(V.sub(0), p_exact, "on_boundary"), cm.DirichletBC(V.sub(1), q_exact, bd, 31)]
    X = mesh.coords
    if world > 1:
        from closestmolecule_b200.multigpu import DistributedPQNewton
        solver = DistributedPQNewton(form, bcs, restart=40, rtol_lin=1e-10)
        uvec, own = solver.u, solver.own
    else:
        u = cm.Function(V)
        solver = cm.NonlinearVariationalSolver(cm.NonlinearVariationalProblem(form, u, bcs))
        solver.parameters["newton_solver"]["linear_solver"] = "gmres"
        solver.parameters["newton_solver"]["krylov_solver"]["relative_tolerance"] = 1e-10
        uvec, own = u.vector(), slice(0, u.vector().numel())
    u0 = torch.empty_like(uvec)
    u0[0::2] = p_exact(X[:, 0], X[:, 1])
    u0[1::2] = q_exact(X[:, 0], X[:, 1])

    def step():
        uvec[own] = u0[own]
        return solver.newton_step(compute_residual=True)

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        step()
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    ks = []
    for e in evs:
        nm = e.name
        if nm.startswith("Memcpy") or nm.startswith("Memset"):
            continue
        ks.append((e.time_range.start, e.time_range.end - e.time_range.start, short(nm)))
    ks.sort()
    if not ks:
        print("no kernel records")
        return
    t_first = ks[0][0]
    span = ks[-1][0] + ks[-1][1] - t_first
    agg = collections.defaultdict(lambda: [0, 0.0])
    for s, d, n in ks:
        agg[n][0] += 1
        agg[n][1] += d
    busy = sum(a[1] for a in agg.values())
    lines = [f"# rank {rank}/{world}  nx={nx}: {len(ks)} kernels, span {span / 1e3:.3f} ms, sum of durations "
             f"{busy / 1e3:.3f} ms (streams overlap)"]
    lines.append("kernel,launches,total_us,avg_us")
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f'"{n}",{c},{t:.1f},{t / c:.2f}')
    # one GMRES iteration = the kernels after one Givens kernel up to and including the next
    ends = [i for i, k in enumerate(ks) if k[2].startswith("gd_givens")]
    if len(ends) > args.iteration + 1:
        a, b = ends[args.iteration] + 1, ends[args.iteration + 1] + 1
        lines.append(f"# GMRES iteration {args.iteration}: {b - a} kernels, "
                     f"{(ks[b - 1][0] + ks[b - 1][1] - ks[a][0]):.1f} us from the first kernel to the end of the Givens kernel, "
                     f"{(ks[min(b, len(ks) - 1)][0] - ks[a][0]):.1f} us to the start of the next iteration")
        lines.append("offset_us,dur_us,gap_us,kernel")
        prev_end = ks[a][0]
        for s, d, n in ks[a:b]:
            lines.append(f"{s - ks[a][0]:.1f},{d:.1f},{s - prev_end:.1f},{n}")
            prev_end = max(prev_end, s + d)
    if ends:
        # everything outside the GMRES cycle: assembly, Dirichlet rows, solver setup, start of the solve / end of the step
        def dump(title, seq):
            lines.append(f"# {title}: {len(seq)} kernels")
            lines.append("offset_us,dur_us,gap_us,kernel")
            prev_end = seq[0][0] if seq else 0.0
            for s_, d_, n_ in seq:
                lines.append(f"{s_ - t_first:.1f},{d_:.1f},{s_ - prev_end:.1f},{n_}")
                prev_end = max(prev_end, s_ + d_)
        dump("step start up to the first Givens kernel", ks[:ends[0] + 1])
        dump("after the last Givens kernel", ks[ends[-1] + 1:])
    txt = "\n".join(lines)
    out = args.out or os.path.join(ROOT, "gpurun_out", f"trace_n{world}_r{rank}.csv")
    if rank in (0, world // 2):
        with open(out if world == 1 or args.out == "" else args.out + f".r{rank}", "w") as f:
            f.write(txt + "\n")
    if rank == 0:
        print(txt[:6000])
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
