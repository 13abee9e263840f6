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


def _parse_cpulist(text):
    """'0-3,8,10-11' -> [0, 1, 2, 3, 8, 10, 11] (the format of /sys/devices/system/node/nodeN/cpulist)."""
    out = []
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        out.extend(range(int(lo), int(hi or lo) + 1))
    return out


def gpu_numa_node(index):
    """NUMA node of the host memory closest to CUDA device `index` (sysfs of its PCI function), or -1 when the
    platform does not say (single-node hosts, containers without sysfs, no CUDA)."""
    try:
        p = torch.cuda.get_device_properties(index)
        path = f"/sys/bus/pci/devices/{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0/numa_node"
        with open(path) as f:
            return int(f.read().strip())
    except Exception:
        return -1


def plan_rank_cores(local_rank, local_world, avail, rank_nodes, node_cpus):
    """Pure planning half of pin_rank_to_cores (unit-tested on the CPU): `avail` = the cores this process may use,
    rank_nodes[r] = NUMA node of rank r's GPU (-1 unknown), node_cpus[n] = cores of node n.  A rank whose GPU's node is
    known and has usable cores gets an equal slice of THAT node's cores among the ranks that share the node -- its
    pinned staging buffers (first touch) and its copy engine's reads then stay on the socket the GPU hangs off.  If any
    rank cannot be placed that way (unknown node, a cpuset that does not reach the other socket), ALL ranks take equal
    slices of the usable cores by local rank, so the slices stay disjoint."""
    avail = sorted(avail)
    usable = set(avail)

    def node_plan(r):
        node = rank_nodes[r] if r < len(rank_nodes) else -1
        if node < 0 or node not in node_cpus:
            return None
        cpus = [c for c in sorted(node_cpus[node]) if c in usable]
        peers = [q for q in range(local_world) if rank_nodes[q] == node]
        per = len(cpus) // max(len(peers), 1)
        if per < 1:
            return None
        i = peers.index(r)
        return cpus[i * per:(i + 1) * per]

    plans = [node_plan(r) for r in range(local_world)]
    if all(p is not None for p in plans):              # every rank can sit on its GPU's node: disjoint by construction
        return plans[local_rank]
    per = len(avail) // max(local_world, 1)
    if per < 1:
        return None
    return avail[local_rank * per:(local_rank + 1) * per]


def pin_rank_to_cores(local_rank, local_world):
    """Give every rank of a node its own slice of the host cores (Linux), on the NUMA node of its GPU when sysfs tells
    (plan_rank_cores): the per-step host work of a rank -- kernel launches, the staging of the next inputs -- then does
    not migrate or share a core with another rank's, and the pinned buffers it allocates afterwards (first touch) are
    local to the GPU that reads them.  Call before allocating pinned buffers.  Returns the core list, or None when the
    node has fewer cores than ranks or affinity is not supported."""
    try:
        avail = sorted(os.sched_getaffinity(0))
    except (AttributeError, OSError):
        return None
    if local_world <= 1:
        return None
    rank_nodes = [gpu_numa_node(r) for r in range(local_world)]
    node_cpus = {}
    for n in set(rank_nodes):
        if n >= 0:
            try:
                with open(f"/sys/devices/system/node/node{n}/cpulist") as f:
                    node_cpus[n] = _parse_cpulist(f.read())
            except (OSError, ValueError):
                pass
    cores = plan_rank_cores(local_rank, local_world, avail, rank_nodes, node_cpus)
    if not cores:
        return None
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
