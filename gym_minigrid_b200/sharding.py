"""Multi-GPU plumbing for the path: envs are independent units, so a job of
`total_envs` shards by contiguous global env id ranges, one process per GPU, and NOTHING is
exchanged on the step path.  torch.distributed is used only to agree on timing / totals."""
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
