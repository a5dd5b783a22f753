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
