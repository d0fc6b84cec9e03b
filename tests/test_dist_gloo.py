"""world_size-2 gloo test (CPU) of the N > 1 path: chunk sharding with no data-path collective, and the
max-over-ranks timing reduction bench.py uses."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from scenesplat_b200.sharding import assign_chunks, job_throughput


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, sizes, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = assign_chunks(sizes, world, "lpt")[rank]
    units = float(sum(sizes[i] for i in mine))
    ms = 10.0 * (rank + 1)  # pretend device time
    t = torch.tensor([ms])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    u = torch.tensor([units], dtype=torch.float64)
    dist.all_reduce(u)
    got = [None] * world
    dist.all_gather_object(got, mine)
    dist.barrier()
    if rank == 0:
        q.put((float(t.item()), float(u.item()), got))
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_chunk_sharding_gloo():
    sizes = [300, 100, 250, 50, 400, 120, 80]
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, sizes, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(100)
        assert p.exitcode == 0
    tmax, units, parts = q.get()
    assert tmax == 20.0 and units == float(sum(sizes))
    assert sorted(i for p in parts for i in p) == list(range(len(sizes)))  # disjoint cover, no exchange needed
    assert job_throughput([units], [tmax]) == units / 0.02
