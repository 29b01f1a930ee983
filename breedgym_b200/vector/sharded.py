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

import ctypes
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist

from .. import _lib
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


class RewardGather:
    """The path's one collective through the C ABI (`bg_allgather_f32`: ncclAllGather, no host round trip; the
    communicator is created and warmed once).  `torch.distributed` is used only to hand NCCL's unique id to the other
    ranks.  Equal shard sizes only (the sharded env pads otherwise through torch).

    The collective runs on its OWN stream: it waits (event) for the step that produced the local rewards, copies them
    into a private staging buffer and all-gathers from there, so the step stream never waits for the slowest rank --
    only whoever consumes the gathered rewards does.  `__call__(local, wait=True)` makes the step stream wait for the
    result right away (plain in-order semantics); `wait=False` leaves that to the caller (`self.done` is the event).
    """

    def __init__(self, simulator, world: int, rank: int, count: int, group=None):
        lib = _lib.load()
        ident = np.zeros(128, dtype=np.uint8)
        if rank == 0:
            _lib.check(lib.bg_comm_unique_id(_lib.nptr(ident)))
        box = [ident.tobytes()]
        dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        ident = np.frombuffer(box[0], dtype=np.uint8).copy()
        self._comm = ctypes.c_void_p()
        _lib.check(lib.bg_comm_create(simulator._engine, _lib.nptr(ident), world, rank, ctypes.byref(self._comm)))
        self._fn = lib.bg_allgather_f32
        self._destroy = lib.bg_comm_destroy
        self.world, self.count = world, count
        self.device = simulator.device
        self.stream = torch.cuda.Stream(device=self.device)
        self._stream_ptr = ctypes.c_void_p(self.stream.cuda_stream)
        self.ready = torch.cuda.Event()
        self.done = torch.cuda.Event()
        # two staging / receive buffers, alternating: the rewards of an episode stay valid while the next episode runs
        self._stage = [torch.empty(count, dtype=torch.float32, device=self.device) for _ in range(2)]
        self._out = [torch.empty(world * count, dtype=torch.float32, device=self.device) for _ in range(2)]
        self._pos = 0

    def __call__(self, local: torch.Tensor, wait: bool = True) -> torch.Tensor:
        self._pos ^= 1
        stage, out = self._stage[self._pos], self._out[self._pos]
        main = torch.cuda.current_stream(self.device)
        self.ready.record(main)
        self.stream.wait_event(self.ready)
        with torch.cuda.stream(self.stream):
            stage.copy_(local, non_blocking=True)  # the step's reward buffer is free again as soon as this has run
        rc = self._fn(self._comm, stage.data_ptr(), out.data_ptr(), self.count, self._stream_ptr)
        if rc:
            _lib.check(rc)
        self.done.record(self.stream)
        if wait:
            main.wait_event(self.done)
        return out

    def close(self):
        if self._comm:
            self._destroy(self._comm)
            self._comm = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ShardedVecBreedGym:
    """`VecBreedGym` with `total_envs` logical envs partitioned over the ranks of a process group.

    `step(actions)` takes THIS rank's actions `[count, n, 2]` and returns the local
    observation handle with the rewards of ALL envs.  `collective`: "native" = `bg_allgather_f32`
    (NCCL through the C ABI, on its own stream), "torch" = `torch.distributed` (gloo in the CPU
    tests), "auto" = native when the group's backend is nccl and the shards are equal.
    `async_rewards` (native collective, `info_device="device"`): the step stream does not wait for the
    all-gather -- the following steps overlap it; wait for `env.rewards_done` (a CUDA event) before
    reading the returned rewards on another stream, or call `env.wait_rewards()`.
    """

    def __init__(self, total_envs: int, group=None, device: Optional[int] = None, collective: str = "auto",
                 async_rewards: bool = False, **kwargs):
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
        if collective not in ("auto", "native", "torch"):
            raise ValueError("collective must be 'auto', 'native' or 'torch'")
        equal = len(set(self.counts)) == 1
        if collective == "auto":
            collective = "native" if (self.world > 1 and equal and dist.get_backend(group) == "nccl") else "torch"
        if collective == "native" and not equal:
            raise ValueError("the native reward all-gather needs equal shard sizes")
        self.collective = collective
        self._gather = RewardGather(self.env.simulator, self.world, self.rank, self.count, group) \
            if (collective == "native" and self.world > 1) else None
        self._zeros = None
        self.async_rewards = bool(async_rewards) and self._gather is not None

    @property
    def rewards_done(self):
        return self._gather.done if self._gather is not None else None

    def wait_rewards(self):
        """Make the current stream wait for the last reward all-gather (no-op without the native collective)."""
        if self._gather is not None:
            torch.cuda.current_stream(self.env.device).wait_event(self._gather.done)

    def __getattr__(self, name):
        if name.startswith("_") or name == "env":
            raise AttributeError(name)
        return getattr(self.env, name)

    def local_slice(self) -> slice:
        return slice(self.begin, self.begin + self.count)

    def reset(self, seed: Optional[int] = None, options: Optional[dict] = None):
        return self.env.reset(seed=seed, options=options)

    def gather_rewards(self, local: torch.Tensor) -> torch.Tensor:
        """float32[count] on this rank's GPU -> float32[total_envs] on every rank (enqueued, not synchronised)."""
        if self.world == 1:
            return local
        if self._gather is not None:
            return self._gather(local, wait=not self.async_rewards)
        return allgather_rewards(local, self.counts, self.group)

    def step(self, local_actions):
        env = self.env
        will_reward = env.reward_shaping or env.step_idx + 1 == env.num_generations
        obs, rews, ter, tru, infos = env.step(local_actions)
        on_device = isinstance(rews, torch.Tensor)  # info_device="device": rewards never leave the GPUs
        if will_reward:
            local = rews if on_device else torch.from_numpy(np.ascontiguousarray(rews, dtype=np.float32)).to(env.device)
            rews = self.gather_rewards(local)
            if not on_device:
                rews = rews.cpu().numpy()
        elif on_device:
            if self._zeros is None:
                self._zeros = torch.zeros(self.total_envs, dtype=torch.float32, device=env.device)
            rews = self._zeros
        else:
            rews = np.zeros(self.total_envs)
        ter = np.full(self.total_envs, bool(ter[0]) if len(ter) else False)
        tru = np.full(self.total_envs, bool(tru[0]) if len(tru) else False)
        return obs, rews, ter, tru, infos
