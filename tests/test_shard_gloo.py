"""N > 1 host logic on CPU: world_size-2 gloo run of the frame sharding + counter gather that bench.py uses
(NCCL on the GPU box carries exactly the same calls)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from xfeatslam_b200 import shard


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
    mine = shard.frame_indices(rank, world, 37)
    kp = sum(100 + i for i in mine)               # fake per-frame keypoint counts
    elapsed = 10.0 + 5.0 * rank                    # rank 1 is the slow one
    counters, tmax = shard.gather_counters(len(mine), kp, len(mine) - 1, elapsed)
    q.put((rank, mine, counters.tolist(), tmax, shard.whole_job_fps(counters, tmax)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_and_gather():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, f0, c0, t0, fps0), (r1, f1, c1, t1, fps1) = res
    assert sorted(f0 + f1) == list(range(37)) and not set(f0) & set(f1)      # disjoint cover
    assert all(shard.owner_of(i, world) == 0 for i in f0) and all(shard.owner_of(i, world) == 1 for i in f1)
    assert c0 == c1                                                           # every rank sees the same table
    assert [row[0] for row in c0] == [19, 18]
    assert c0[0][1] == sum(100 + i for i in f0) and c0[1][1] == sum(100 + i for i in f1)
    assert t0 == t1 == 15.0                                                   # max over ranks
    assert abs(fps0 - 37 / 0.015) < 1e-6 and fps0 == fps1


def test_single_process_path():
    counters, tmax = shard.gather_counters(5, 50, 4, 2.0)
    assert counters.shape == (1, 4) and counters[0, 0] == 5 and tmax == 2.0
    assert shard.frame_indices(0, 1, 4) == [0, 1, 2, 3]
