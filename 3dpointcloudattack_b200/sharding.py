"""Batch sharding for one-process-per-GPU attack runs.

Every primitive of the path is per-sample, so the adversarial batch is cut along B: rank r of G
owns samples [start, stop).  Nothing is exchanged inside the attack loop; one all-gather at the
end collects per-sample losses and the perturbed clouds.  Two couplings of the reference's loops
are neutralised here:
  * losses are averaged over the batch (attack/CW/CW_attack.py:160-165) -> scale local means by
    B_local / B_global so every sample sees the gradient it would see in the unsharded run;
  * random initial perturbations are drawn per sample from a generator seeded by the GLOBAL
    sample id, so results do not depend on G.
"""
import torch
import torch.distributed as dist


def shard_range(global_batch, rank, world):
    """Contiguous near-equal split: the first (global_batch % world) ranks own one more sample."""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world of size {world}")
    base, extra = divmod(global_batch, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard(x, rank, world):
    start, stop = shard_range(x.shape[0], rank, world)
    return x[start:stop]


def mean_scale(local_batch, global_batch):
    """Factor that turns `loss_local.mean()` into this rank's share of the global batch mean."""
    return float(local_batch) / float(global_batch)


def per_sample_noise(shape_per_sample, first_sample, count, sigma, seed=0, device="cpu"):
    """[count, *shape] N(0, sigma^2) noise, sample k drawn from seed + global id (CW_attack.py:94)."""
    out = []
    for k in range(count):
        g = torch.Generator().manual_seed(int(seed) * 1000003 + first_sample + k)
        out.append(torch.randn(shape_per_sample, generator=g) * sigma)
    return torch.stack(out).to(device) if out else torch.empty((0,) + tuple(shape_per_sample), device=device)


def gather_batch(local, global_batch, group=None):
    """All-gather a per-sample tensor [B_local, ...] into [global_batch, ...] on every rank (the single
    collective of the path: NCCL all_gather_into_tensor over NVLink; gloo on CPU for tests)."""
    if not dist.is_available() or not dist.is_initialized():
        return local
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    base, extra = divmod(global_batch, world)
    width = base + (1 if extra else 0)                      # padded shard size (equal on all ranks)
    pad = torch.zeros((width,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    if dist.get_backend(group) == "nccl":
        out = torch.empty((world * width,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, pad.contiguous(), group=group)
        parts = list(out.split(width))
    else:
        parts = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(parts, pad.contiguous(), group=group)
    keep = []
    for r in range(world):
        s, e = shard_range(global_batch, r, world)
        keep.append(parts[r][:e - s])
    del rank
    return torch.cat(keep, 0)
