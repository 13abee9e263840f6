"""One process per GPU; the CIM head path shards by image with no data-path collective
(every stage is per image, the reference already runs one image per GPU,
lib/roi_data/minibatch.py:28-29).  The only collective is the allreduce of the head gradients,
the counterpart of the reference's comm.reduce_add_coalesced (lib/nn/parallel/_functions.py:39).
"""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Initialise torch.distributed from RANK / WORLD_SIZE / MASTER_* (torchrun).  Returns
    (rank, world_size, local_rank); a no-op single-process setup when WORLD_SIZE is unset."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        # NCCL prints its version banner (NCCL_DEBUG=VERSION/INFO) to stdout; rank 0's stdout carries
        # exactly one JSON line in bench.py, so send NCCL's log to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend=backend, rank=rank, world_size=world,
                                    device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def pin_rank_to_cores(local_rank, local_world):
    """Give every rank of a node its own slice of the host cores (Linux): the per-step host work of a rank -- kernel
    launches, the sampling hop's wait + random draw, the staging of the next inputs -- then does not migrate or share a
    core with another rank's.  Call before allocating pinned buffers.  Returns the core list, or None when the node
    has fewer cores than ranks or affinity is not supported."""
    try:
        avail = sorted(os.sched_getaffinity(0))
    except (AttributeError, OSError):
        return None
    per = len(avail) // max(local_world, 1)
    if local_world <= 1 or per < 1:
        return None
    cores = avail[local_rank * per:(local_rank + 1) * per]
    try:
        os.sched_setaffinity(0, cores)
    except OSError:
        return None
    return cores


def shard_images(n_images, rank, world):
    """Images rank `rank` owns: g, g + world, ... (image-level data parallel)."""
    return list(range(rank, n_images, world))


def shutdown():
    if dist.is_initialized():
        dist.destroy_process_group()


def barrier():
    if dist.is_initialized():
        dist.barrier()


def max_over_ranks(value, device="cpu"):
    """MAX of a python float over all ranks (step time = the slowest rank)."""
    if not dist.is_initialized():
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value, device="cpu"):
    if not dist.is_initialized():
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def allreduce_mean_(tensors):
    """Average gradient tensors over ranks in ONE flattened bucket (the reference averages losses
    over GPUs, lib/utils/training_stats.py:76, and sums gradients on GPU 0)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return tensors
    flat = torch.cat([t.reshape(-1) for t in tensors])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    flat /= dist.get_world_size()
    off = 0
    for t in tensors:
        n = t.numel()
        t.copy_(flat[off:off + n].view_as(t))
        off += n
    return tensors


def allreduce_mean_async_(bucket):
    """Start averaging one flat gradient bucket over the ranks (in place) and return a callable that makes
    the CURRENT stream wait for the result, or None in a single-process run.  NCCL runs the collective on
    its own stream, so kernels launched between the two calls overlap with it."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return None
    work = dist.all_reduce(bucket, op=dist.ReduceOp.AVG if dist.get_backend() == "nccl" else dist.ReduceOp.SUM,
                           async_op=True)
    world = dist.get_world_size()
    nccl = dist.get_backend() == "nccl"

    def finish():
        work.wait()
        if not nccl:
            bucket.div_(world)
    return finish
