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


def gather_rollout(local, n_total=None, group=None, dim=0, stream=None):
    """all-gather a per-rank rollout tensor ([n_local, ...] or [T, n_local, ...] with dim=1)
    into the global one: ONE `all_gather_into_tensor` (NCCL over NVLink for CUDA tensors, gloo
    for CPU tensors), no host round trip -- shard sizes follow from `shard_range`, they are
    not exchanged.  Shards may differ by one env; they are padded to the largest one for the
    collective and trimmed afterwards (n_total is needed then; with equal shards it is optional).

    stream: a CUDA side stream to run the collective on, so that it overlaps the next rollout
    on the current stream.  The side stream first waits for the current stream (the producer
    of `local`); the call returns (tensor, event) and the consumer waits for the event
    (`torch.cuda.current_stream().wait_event(event)`) before it reads the tensor."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local if stream is None else (local, None)
    world = dist.get_world_size(group)
    n_local = local.shape[dim]
    if n_total is None:
        n_total = n_local * world                      # equal shards
    sizes = [hi - lo for lo, hi in (shard_range(n_total, r, world) for r in range(world))]
    if sizes[dist.get_rank(group)] != n_local:
        raise RuntimeError("this rank holds %d envs, shard_range(%d, rank, %d) says %d"
                           % (n_local, n_total, world, sizes[dist.get_rank(group)]))
    m = max(sizes)

    def run():
        x = local.movedim(dim, 0)
        if n_local < m:
            x = torch.cat([x, x.new_zeros((m - n_local,) + tuple(x.shape[1:]))], dim=0)
        x = x.contiguous()
        out = x.new_empty((world * m,) + tuple(x.shape[1:]))
        dist.all_gather_into_tensor(out, x, group=group)
        if min(sizes) < m:
            out = torch.cat([out[r * m:r * m + s] for r, s in enumerate(sizes)], dim=0)
        return out.movedim(0, dim)

    if stream is None:
        return run()
    stream.wait_stream(torch.cuda.current_stream(local.device))
    with torch.cuda.stream(stream):
        out = run()
        local.record_stream(stream)
        ev = torch.cuda.Event()
        ev.record(stream)
    return out, ev
