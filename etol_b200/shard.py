"""Sharding of a batch of independent VGP instances over the ranks of one node, and the one exchange
step of the path: an all-gather of per-instance results (BASELINE.json config 5, SURVEY.md
section 8e). Backend-agnostic torch.distributed code: NCCL over NVLink on the GPU box, gloo in the
CPU test-suite."""
import torch
import torch.distributed as dist


def shard_range(n_instances, world, rank):
    """Contiguous instance range [lo, hi) of `rank`; ranges differ by at most one instance."""
    base, rem = divmod(n_instances, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def padded_shard(n_instances, world):
    """Instances per rank when every rank must contribute equally to all_gather_into_tensor."""
    return (n_instances + world - 1) // world


def gather_rows(local, n_instances=None):
    """All-gather per-instance rows (local: [b_local, width], same b_local on every rank) into
    [world*b_local, width] on every rank; trims the padding when n_instances is given."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local if n_instances is None else local[:n_instances]
    world = dist.get_world_size()
    out = torch.empty((world * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local.contiguous())
    return out if n_instances is None else out[:n_instances]
