"""N>1 host logic on CPU: world_size-2 gloo processes run the object-sharded plan and the final record gather."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from onepose_b200 import sharding


def test_lpt_partition_is_balanced_and_complete():
    costs = [sharding.object_cost(f, n, m) for f, n, m in [(300, 1000, 2000), (100, 500, 800), (250, 900, 2500),
                                                          (50, 300, 900), (400, 1200, 2400), (120, 700, 1500), (90, 400, 1000)]]
    for world in (1, 2, 4, 8):
        plan = sharding.partition_lpt(costs, world)
        assert sorted(i for p in plan for i in p) == list(range(len(costs)))
        loads = [sum(costs[i] for i in p) for p in plan]
        # LPT guarantee: makespan <= (4/3 - 1/(3m)) * OPT, and OPT >= max(mean load, largest item)
        opt_lb = max(sum(costs) / world, max(costs))
        assert max(loads) <= (4 / 3) * opt_lb + 1e-9
    assert sharding.partition_lpt(costs, 2) == sharding.partition_lpt(costs, 2)       # deterministic


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, costs, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        plan = sharding.partition_lpt(costs, world)
        mine = plan[rank]
        # stand-in for "run the matcher on my objects": frames processed and a checksum of the object ids
        rec = torch.tensor([float(len(mine)), float(sum(costs[i] for i in mine)), float(sum(mine))], dtype=torch.float64)
        allrec = sharding.gather_records(rec)
        q.put((rank, allrec.tolist(), mine))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_world2_gloo_sharded_run_and_gather():
    costs = [5.0, 3.0, 8.0, 1.0, 4.0, 7.0]
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, costs, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=90) for _ in range(world)]
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    results.sort()
    # every rank sees the same gathered table, it covers all objects exactly once, equals the local plan
    assert results[0][1] == results[1][1]
    table = results[0][1]
    assert sum(row[0] for row in table) == len(costs)
    assert sum(row[2] for row in table) == sum(range(len(costs)))
    for rank, _, mine in results:
        assert table[rank][0] == len(mine) and table[rank][1] == sum(costs[i] for i in mine)
    assert sorted(i for _, _, mine in results for i in mine) == list(range(len(costs)))


def test_gather_without_process_group_is_identity():
    rec = torch.tensor([1.0, 2.0], dtype=torch.float64)
    assert torch.equal(sharding.gather_records(rec), rec[None])
