"""world_size-2 gloo test (CPU) of the N>1 host logic: instance sharding and ordered gathering."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank),
                      MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from closestmolecule_b200 import dist as cmd
    r, w, _ = cmd.init("gloo")
    assert (r, w) == (rank, world)
    conc = [1, 2, 3, 4, 5, 6, 7, 8, 9]                       # advection_diffusion_mixed.py:296
    mine = cmd.shard_instances(conc, r, w)
    assert [i for i, _ in mine] == list(range(r, 9, w))
    table = cmd.sweep(conc, lambda c: [float(c), 4.0 * c / (1.0 + c)], 2)
    t = cmd.max_over_ranks([float(rank + 1), 10.0 - rank])
    if rank == 0:
        torch.save((table, t), out)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sweep(tmp_path):
    out = str(tmp_path / "res.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    table, t = torch.load(out)
    assert table.shape == (9, 2) and not torch.isnan(table).any()
    assert table[:, 0].tolist() == [float(c) for c in range(1, 10)]
    assert torch.allclose(table[:, 1], 4.0 * table[:, 0] / (1.0 + table[:, 0]))
    assert t == [2.0, 10.0]


# ---------------------------------------------------------------------------------------------------------
# strip partition of the multi-GPU Newton solve (closestmolecule_b200/multigpu.py): ownership, alignment and the
# ghost-row semantics of the halo protocol, emulated with gloo send / recv on the CPU
# ---------------------------------------------------------------------------------------------------------
OFFS = [(-1, -1), (0, -1), (-1, 0), (0, 0), (1, 0), (0, 1), (1, 1)]     # 7-point stencil of the right-diagonal meshes


def _stencil_rows(A, x, rows):
    """y[rows] = (A x)[rows] for the stencil stored as 7 diagonals; x must hold valid rows rows-1 .. rows+1"""
    import numpy as np
    ny1, n1 = x.shape
    y = np.zeros((len(rows), n1))
    xp = np.pad(x, 1)
    for k, (dx, dy) in enumerate(OFFS):
        for a, j in enumerate(rows):
            y[a] += A[k, j] * xp[1 + j + dy, 1 + dx:1 + dx + n1]
    return y


def _strip_worker(rank, world, port, out):
    import numpy as np
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank),
                      MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from closestmolecule_b200.multigpu import strip_partition
    nx, ny = 24, 64
    rb_all = strip_partition(ny, world)
    rb, re = rb_all[rank], rb_all[rank + 1]
    # every rank sees the same partition, strips tile the rows, start on multiples of 8 and keep >= 4 rows on the
    # two distributed multigrid levels (rows >> 1)
    got = [None] * world
    dist.all_gather_object(got, (rb, re))
    assert got[0][0] == 0 and got[-1][1] == ny + 1
    assert all(got[r][1] == got[r + 1][0] for r in range(world - 1))
    assert all(b % 8 == 0 and ((e - b) >> 1) >= 4 for b, e in got)
    rng = np.random.default_rng(5)           # same seed on every rank: globally indexed arrays, like the GPU path
    n1 = nx + 1
    A = rng.standard_normal((7, ny + 1, n1))
    x_true = rng.standard_normal((ny + 1, n1))
    # a rank holds valid data on its own rows only; ghost rows arrive from the neighbours (row rb goes down,
    # row re-1 goes up: the "down" / "up" lanes of csrc/cm_halo.cuh)
    x = np.full((ny + 1, n1), np.nan)
    x[rb:re] = x_true[rb:re]
    reqs = []
    if rank > 0:
        reqs.append(dist.isend(torch.from_numpy(x[rb].copy()), rank - 1))
    if rank + 1 < world:
        reqs.append(dist.isend(torch.from_numpy(x[re - 1].copy()), rank + 1))
    if rank + 1 < world:
        t = torch.empty(n1, dtype=torch.float64)
        dist.recv(t, rank + 1)
        x[re] = t.numpy()
    if rank > 0:
        t = torch.empty(n1, dtype=torch.float64)
        dist.recv(t, rank - 1)
        x[rb - 1] = t.numpy()
    for r_ in reqs:
        r_.wait()
    lo, hi = max(rb - 1, 0), min(re + 1, ny + 1)
    xs = np.where(np.isnan(x), 0.0, x)
    assert not np.isnan(x[lo:hi]).any() and (np.isnan(x[:lo]).all() if lo > 0 else True)
    y_own = _stencil_rows(A, xs, list(range(rb, re)))
    parts = [None] * world
    dist.all_gather_object(parts, (rb, re, y_own))
    if rank == 0:
        y = np.concatenate([p[2] for p in sorted(parts, key=lambda p: p[0])])
        y_ref = _stencil_rows(A, x_true, list(range(ny + 1)))
        torch.save(float(np.max(np.abs(y - y_ref))), out)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_strip_partition_and_ghost_rows(tmp_path):
    out = str(tmp_path / "strip.pt")
    mp.spawn(_strip_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    assert torch.load(out) == 0.0


def test_strip_partition_rules():
    from closestmolecule_b200.multigpu import strip_partition
    import pytest
    assert strip_partition(2048, 8) == [0, 256, 512, 768, 1024, 1280, 1536, 1792, 2049]
    assert strip_partition(2048, 1) == [0, 2049]
    assert strip_partition(100, 3) == [0, 32, 64, 101]
    with pytest.raises(ValueError):
        strip_partition(20, 4)
