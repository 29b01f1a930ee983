"""Env-sharded vector environment: one process per GPU, `torch.distributed` plumbing.

B200-native replacement for the reference's `DistributedBreedGym`
(breedgym/vector/vec_env.py:150-236), which spawns one subprocess per device
and ships every observation back through host pipes.  Here each rank owns a
contiguous block of the E logical envs on its own GPU and steps it
independently; all constants (germplasm, thresholds, effects) are replicated
and the shared cross key / per-env reset keys are derived locally from the
same seed, so the union of the shards is bit-identical to one `VecBreedGym`
with E envs.  The only exchange is the all-gather of the per-env rewards
(float32[E/G] per rank, NCCL over NVLink when the backend is nccl).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist

from .vec_env import VecBreedGym


def shard_range(total: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous block of `total` envs owned by `rank`: (begin, count); sizes differ by at most 1."""
    if not 0 <= rank < world:
        raise ValueError("rank out of range")
    base, rem = divmod(total, world)
    return rank * base + min(rank, rem), base + (1 if rank < rem else 0)


def shard_counts(total: int, world: int) -> List[int]:
    return [shard_range(total, world, r)[1] for r in range(world)]


def allgather_rewards(local: torch.Tensor, counts: Sequence[int], group=None) -> torch.Tensor:
    """Concatenate the per-rank reward vectors (`counts[r]` entries from rank r) on every rank."""
    world = len(counts)
    if world == 1:
        return local.clone()
    if len(set(counts)) == 1:
        out = torch.empty(sum(counts), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    cap = max(counts)
    padded = torch.zeros(cap, dtype=local.dtype, device=local.device)
    padded[: local.numel()] = local
    parts = [torch.empty(cap, dtype=local.dtype, device=local.device) for _ in range(world)]
    dist.all_gather(parts, padded, group=group)
    return torch.cat([p[:c] for p, c in zip(parts, counts)])


class ShardedVecBreedGym:
    """`VecBreedGym` with `total_envs` logical envs partitioned over the ranks of a process group.

    `step(actions)` takes THIS rank's actions `[count, n, 2]` and returns the local
    observation handle with the rewards of ALL envs.
    """

    def __init__(self, total_envs: int, group=None, device: Optional[int] = None, **kwargs):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.total_envs = total_envs
        self.begin, self.count = shard_range(total_envs, self.world, self.rank)
        self.counts = shard_counts(total_envs, self.world)
        if device is None:
            device = torch.cuda.current_device()
        self.env = VecBreedGym(num_envs=self.count, env_shard=(self.begin, total_envs), device=device, **kwargs)
        self.num_envs = total_envs

    def __getattr__(self, name):
        if name.startswith("_") or name == "env":
            raise AttributeError(name)
        return getattr(self.env, name)

    def local_slice(self) -> slice:
        return slice(self.begin, self.begin + self.count)

    def reset(self, seed: Optional[int] = None, options: Optional[dict] = None):
        return self.env.reset(seed=seed, options=options)

    def step(self, local_actions):
        will_reward = self.env.reward_shaping or self.env.step_idx + 1 == self.env.num_generations
        obs, rews, ter, tru, infos = self.env.step(local_actions)
        on_device = isinstance(rews, torch.Tensor)  # info_device="device": rewards never leave the GPUs
        if will_reward:
            local = rews if on_device else torch.from_numpy(np.ascontiguousarray(rews, dtype=np.float32)).to(self.env.device)
            rews = allgather_rewards(local, self.counts, self.group)
            if not on_device:
                rews = rews.cpu().numpy()
        elif on_device:
            rews = torch.zeros(self.total_envs, dtype=torch.float32, device=self.env.device)
        else:
            rews = np.zeros(self.total_envs)
        ter = np.full(self.total_envs, bool(ter[0]) if len(ter) else False)
        tru = np.full(self.total_envs, bool(tru[0]) if len(tru) else False)
        return obs, rews, ter, tru, infos
