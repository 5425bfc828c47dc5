"""Env sharding over GPUs: one process per GPU, contiguous env-index ranges, no collective
on the step path; ONE all-gather concatenates rollout tensors when a consumer wants them
(SURVEY section 8(e)).  Philox streams are keyed by the GLOBAL env id, so results do not
depend on the number of ranks."""
import torch
import torch.distributed as dist


def shard_range(n_total, rank, world):
    """contiguous [lo, hi) of rank `rank`; the first n_total % world ranks get one extra env"""
    if not 0 <= rank < world:
        raise ValueError("rank %d outside world of %d" % (rank, world))
    base, rem = divmod(n_total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def make_sharded(env_cls, n_total, rank=None, world=None, device=None, **kwargs):
    """build this rank's shard of an n_total-env batched env"""
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    lo, hi = shard_range(n_total, rank, world)
    return env_cls(num_envs=hi - lo, env_offset=lo, device=device, **kwargs)


def gather_rollout(local, n_total=None, group=None, dim=0):
    """all-gather a per-rank rollout tensor ([n_local, ...] or [T, n_local, ...] with dim=1)
    into the global one.  Shards may differ by one env, so tensors are padded to the largest
    shard for the collective and trimmed afterwards.  NCCL over NVLink for CUDA tensors,
    gloo for CPU tensors."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    n_local = torch.tensor([local.shape[dim]], device=local.device, dtype=torch.int64)
    sizes = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(sizes, n_local, group=group)
    sizes = [int(s.item()) for s in sizes]
    m = max(sizes)
    pad = local
    if local.shape[dim] < m:
        shp = list(local.shape)
        shp[dim] = m - local.shape[dim]
        pad = torch.cat([local, local.new_zeros(shp)], dim=dim)
    pad = pad.contiguous()
    outs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(outs, pad, group=group)
    out = torch.cat([o.narrow(dim, 0, s) for o, s in zip(outs, sizes)], dim=dim)
    if n_total is not None and out.shape[dim] != n_total:
        raise RuntimeError("gathered %d envs, expected %d" % (out.shape[dim], n_total))
    return out
