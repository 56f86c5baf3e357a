"""world_size-2 gloo test (CPU) of the multi-GPU plumbing: sharding of independent areas + timing/trial reductions +
gather of optimised unknowns.  The data path itself has no collective."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from srba_b200 import parallel as P


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = P.shard_indices(11, rank, world)
    # each rank "solves" its areas: pretend area i costs i+1 ms and i+2 trials
    ms = float(sum(i + 1 for i in mine))
    trials = float(sum(i + 2 for i in mine))
    tmax, ttot = P.reduce_timing(ms, trials)
    (m1, m2), (s1,) = P.reduce_max_sum([ms, -ms], [trials])   # the helper bench.py / tools/bench_configs.py use
    assert (m1, s1) == (tmax, ttot) and m2 <= 0
    ids = torch.tensor(mine, dtype=torch.int64)
    vals = torch.stack([ids.double(), ids.double() * 2], dim=1)
    gi, gv = P.gather_unknowns(ids, vals)
    q.put((rank, mine, tmax, ttot, sorted(gi.tolist()), float(gv[:, 1].sum())))
    dist.barrier()
    dist.destroy_process_group()


def test_sharding_and_reductions_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    out.sort()
    all_items = sorted(out[0][1] + out[1][1])
    assert all_items == list(range(11))                      # every area solved exactly once
    exp_ms = max(sum(i + 1 for i in range(0, 11, 2)), sum(i + 1 for i in range(1, 11, 2)))
    for rank, mine, tmax, ttot, gi, gsum in out:
        assert tmax == exp_ms                                # max over ranks
        assert ttot == sum(i + 2 for i in range(11))         # whole-job trial count
        assert gi == list(range(11)) and gsum == 2.0 * sum(range(11))


def test_single_process_identity():
    assert P.reduce_timing(3.5, 7) == (3.5, 7.0)
    assert P.shard_indices(5, 0, 1) == [0, 1, 2, 3, 4]
