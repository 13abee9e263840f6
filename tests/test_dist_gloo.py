"""world_size-2 gloo run of the multi-process plumbing (image sharding, max-over-ranks timing,
gradient bucket allreduce) on CPU."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cim_b200 import dist as cdist


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank),
                      MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    r, w, _ = cdist.init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    mine = cdist.shard_images(7, r, w)
    cdist.barrier()
    slow = cdist.max_over_ranks(1.0 + rank)            # slowest rank defines the step time
    total = cdist.sum_over_ranks(len(mine))
    grads = [torch.full((3, 4), float(rank + 1)), torch.full((5,), float(10 * (rank + 1)))]
    cdist.allreduce_mean_(grads)
    bucket = torch.full((11,), float(rank + 1))        # the step's head-gradient bucket path
    finish = cdist.allreduce_mean_async_(bucket)
    assert finish is not None
    finish()
    assert torch.equal(bucket, torch.full((11,), 1.5))
    out[rank] = (mine, slow, total, grads[0][0, 0].item(), grads[1][0].item())
    dist.destroy_process_group()


def test_two_rank_gloo():
    world, port = 2, _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
        res = dict(out)
    assert res[0][0] == [0, 2, 4, 6] and res[1][0] == [1, 3, 5]
    for r in (0, 1):
        assert res[r][1] == 2.0 and res[r][2] == 7.0
        assert res[r][3] == 1.5 and res[r][4] == 15.0


def test_single_process_is_a_noop():
    assert cdist.shard_images(3, 0, 1) == [0, 1, 2]
    assert cdist.max_over_ranks(0.25) == 0.25
    assert cdist.allreduce_mean_async_(torch.ones(3)) is None
