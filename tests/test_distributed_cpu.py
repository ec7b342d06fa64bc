"""CPU test of the N>1 host logic (mode R: particle shards, no data-path collective) with the
gloo backend at world_size 2."""

import os
import socket

import numpy as np
import torch.multiprocessing as mp

from parcels_b200 import distributed as D
from parcels_b200.particle import create_particle_data


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = 11
    full = create_particle_data(nparticles=n, ngrids=1, initial=dict(x=np.arange(n), y=np.arange(n) * 2.0, z=np.zeros(n),
                                                                     t=np.zeros(n), particle_id=np.arange(n)))  # fmt: skip
    mine = D.shard_particles(full, rank, world)
    # each rank "advects" its shard independently (stand-in: shift x) and rank 1 deletes one particle
    mine["x"] = mine["x"] + 100 * (rank + 1)
    if rank == 1:
        mine = {k: np.delete(v, 0, axis=0) for k, v in mine.items()}
    merged = D.gather_particles(mine, dist, dst=0)
    tmax = D.allreduce_max(10.0 + rank, dist)
    ssum = D.allreduce_sum(len(mine["x"]), dist)
    if rank == 0:
        q.put((merged["particle_id"].tolist(), merged["x"].tolist(), tmax, ssum))
    dist.destroy_process_group()


def test_shard_bounds_balanced():
    assert D.shard_bounds(10, 4) == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert D.shard_bounds(0, 2) == [(0, 0), (0, 0)]
    b = D.shard_bounds(10**7 + 3, 8)
    assert b[0][0] == 0 and b[-1][1] == 10**7 + 3 and all(b[i][1] == b[i + 1][0] for i in range(7))


def test_mode_r_world2_gloo():
    ctx = mp.get_context("fork")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ids, xs, tmax, ssum = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ids == [0, 1, 2, 3, 4, 5, 7, 8, 9, 10]  # rank 1 owned 6..10 and deleted id 6
    assert xs[:6] == [100 + i for i in range(6)] and xs[6] == 207.0
    assert tmax == 11.0 and ssum == 10.0
