"""Multi-GPU plumbing for the path: envs are independent units, so a job of
`total_envs` shards by contiguous global env id ranges, one process per GPU, and NOTHING is
exchanged on the step path.  torch.distributed is used only to agree on timing / totals."""
import os

import torch
import torch.distributed as dist


def shard_range(rank, world_size, total_envs):
    """Contiguous block partition [base, base+count) of global env ids for `rank`.
    Blocks differ by at most one env; ids are global so Philox streams (and therefore results)
    do not depend on world_size."""
    if not (0 <= rank < world_size):
        raise ValueError("rank %d outside world of %d" % (rank, world_size))
    q, r = divmod(int(total_envs), int(world_size))
    base = rank * q + min(rank, r)
    return base, q + (1 if rank < r else 0)


def make_shard(env_id, total_envs, rank, world_size, device=None, seed=1337, **kw):
    """make() this rank's shard of a `total_envs`-env job."""
    from .register import make
    base, count = shard_range(rank, world_size, total_envs)
    return make(env_id, num_envs=count, device=device, seed=seed, env_id_base=base, **kw)


def aggregate_throughput(local_units, local_seconds, device="cpu", group=None):
    """Whole-job throughput = (sum over ranks of units) / (max over ranks of time).
    Works on any backend (nccl on GPUs, gloo in the CPU tests); returns (value, total_units, max_seconds)."""
    if not (dist.is_available() and dist.is_initialized()):
        return local_units / local_seconds, local_units, local_seconds
    t = torch.tensor([float(local_seconds)], dtype=torch.float64, device=device)
    u = torch.tensor([float(local_units)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    dist.all_reduce(u, op=dist.ReduceOp.SUM, group=group)
    return float(u.item()) / float(t.item()), float(u.item()), float(t.item())


def bind_to_gpu_numa_node(device_index):
    """Pin the calling process to the CPUs of the NUMA node its GPU hangs off, so that host buffers pinned afterwards
    (step_host) are allocated next to that GPU's PCIe root.  With one process per GPU this keeps the D2H streams of the
    ranks from crossing the socket interconnect.  Best effort: returns the node number, or None if the topology cannot
    be read (then nothing is changed)."""
    try:
        p = torch.cuda.get_device_properties(device_index)
        bdf = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        with open("/sys/bus/pci/devices/%s/numa_node" % bdf) as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
            cpus = set()
            for part in f.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except Exception:
        return None
