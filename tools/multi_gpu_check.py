#!/usr/bin/env python
"""Multi-GPU check of SURVEY 8(e), run under torchrun on N GPUs of one node:

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/multi_gpu_check.py

Every rank steps its contiguous shard of one n_total-env VSS-v0 world (no collective on the step
path), the rollout observations are concatenated with ONE NCCL all-gather, and rank 0 compares
them bit for bit with the same world stepped unsharded on its own GPU (Philox streams are keyed
by the global env id, so the sharding must not show)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from rsoccer_b200 import envs  # noqa: E402
from rsoccer_b200.sharding import gather_rollout, make_sharded, shard_range  # noqa: E402


def main():
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    # RS_CHECK_PER_RANK=16384: the shards run the scalar-form kernel, the unsharded world the packed one
    n_total, T = int(os.environ.get("RS_CHECK_PER_RANK", "65536")) * world + 37, 24
    env = make_sharded(envs.VSSVecEnv, n_total, rank, world, device=dev, seed=5, max_episode_steps=9)
    lo, hi = shard_range(n_total, rank, world)
    g = torch.Generator().manual_seed(3)
    acts = torch.rand(T, n_total, 2, generator=g) * 2 - 1
    obs0, _ = env.reset()
    traj = [obs0.clone()]
    rew = []
    for t in range(T):
        o, r, d, tr, _ = env.step(acts[t, lo:hi].to(dev))
        traj.append(o.clone()); rew.append(r.clone())
    full = gather_rollout(torch.stack(traj), n_total=n_total, dim=1)        # [T + 1, n_total, 40] on every rank
    full_r = gather_rollout(torch.stack(rew), n_total=n_total, dim=1)
    ok = True
    if rank == 0:
        ref = envs.VSSVecEnv(num_envs=n_total, device=dev, seed=5, max_episode_steps=9)
        o, _ = ref.reset()
        same = [bool(torch.equal(o, full[0]))]
        for t in range(T):
            o, r, d, tr, _ = ref.step(acts[t].to(dev))
            same.append(bool(torch.equal(o, full[t + 1]) and torch.equal(r, full_r[t])))
        ok = all(same)
        print("multi_gpu_check: world=%d n_total=%d steps=%d (episodes of 9 steps -> auto-reset exercised) "
              "sharded + NCCL all-gather == unsharded, bit-exact: %s" % (world, n_total, T, ok), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
