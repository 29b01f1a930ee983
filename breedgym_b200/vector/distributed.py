"""Single-process multi-device vector env (API mirror of `DistributedBreedGym`).

The reference (breedgym/vector/vec_env.py:150-236) spawns one subprocess per
device through Gymnasium's AsyncVectorEnv and moves every observation through
host pipes.  This keeps the constructor / reset / step contract -- shard i is a
`VecBreedGym(envs_per_device, autoreset=False)` on device i, `reset(seed=s)`
seeds shard i with s + i (AsyncVectorEnv's convention), observations come back
as ONE `[num_envs, n, m, 2]` array-like, infos carry AsyncVectorEnv's `_key`
presence masks -- but holds all shards in one process and leaves the
observations on their GPUs (`ShardedObservation`: the shards' packed
populations, concatenated on the host only when somebody asks for the array).
Unlike the reference (whose `lambda` closes over the loop variable, so every
worker lands on the LAST device), shard i really runs on `devices[i]`.
For throughput use `ShardedVecBreedGym` (one process per GPU, NCCL reward all-gather).
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np
import torch

from ..gym_compat import spaces
from .vec_env import _VecBreedGym


class ShardedObservation:
    """`bool[num_envs, n, m, 2]` made of per-device `PackedPopulation` shards (env-major, in device order)."""

    __array_priority__ = 1000

    def __init__(self, shards):
        self.shards = list(shards)

    @property
    def shape(self):
        first = self.shards[0].shape
        return (sum(s.shape[0] for s in self.shards),) + tuple(first[1:])

    @property
    def dtype(self):
        return np.dtype(np.bool_)

    def __len__(self):
        return self.shape[0]

    def __array__(self, dtype=None, copy=None):
        a = np.concatenate([np.asarray(s) for s in self.shards], axis=0)
        return a if dtype is None else a.astype(dtype)

    def __getitem__(self, idx):
        if isinstance(idx, (int, np.integer)):
            i = int(idx) + (len(self) if idx < 0 else 0)
            for s in self.shards:
                if i < s.shape[0]:
                    return s[i]
                i -= s.shape[0]
            raise IndexError(idx)
        return np.asarray(self)[idx]

    def __repr__(self):
        return f"ShardedObservation(shape={self.shape}, shards={len(self.shards)})"


class DistributedBreedGym:
    def __init__(self, envs_per_device: int, initial_population, devices: Optional[List[int]] = None, **kwargs):
        if devices is None:
            devices = list(range(torch.cuda.device_count()))
        self.devices = [d if isinstance(d, (int, np.integer)) else torch.device(d).index for d in devices]
        self.envs_per_device = envs_per_device
        kwargs.pop("autoreset", None)
        kwargs.pop("device", None)
        self.envs = [
            _VecBreedGym(envs_per_device, initial_population=initial_population, device=int(d), autoreset=False, **kwargs)
            for d in self.devices
        ]
        self.num_envs = envs_per_device * len(self.devices)
        first = self.envs[0]
        self.single_observation_space = first.single_observation_space
        self.single_action_space = first.single_action_space
        obs_shape = (self.num_envs, *first.single_observation_space.shape)
        act_shape = (self.num_envs, *first.single_action_space.shape)
        self.observation_space = spaces.Box(low=0, high=1, shape=obs_shape, dtype=np.int8)
        self.action_space = spaces.Box(low=0, high=first.individual_per_gen, shape=act_shape, dtype=np.int32)
        self._pending = None

    def reset(self, seed: Optional[int] = None, options: Optional[dict] = None):
        obs, infos = [], []
        for i, env in enumerate(self.envs):
            o, info = env.reset(seed=None if seed is None else seed + i, options=options)
            obs.append(o)
            infos.append(info)
        return ShardedObservation(obs), self._merge_infos(infos)

    # AsyncVectorEnv's two-phase step (vec_env.py:197-219)
    def step_async(self, actions):
        actions = np.asarray(actions)
        if actions.shape[0] != self.num_envs:
            raise ValueError(f"actions must have {self.num_envs} leading entries, got {actions.shape}")
        self._pending = actions.reshape(len(self.devices), self.envs_per_device, *actions.shape[1:])

    def step_wait(self):
        if self._pending is None:
            raise RuntimeError("step_wait called without step_async")
        actions, self._pending = self._pending, None
        results = [env.step(a) for env, a in zip(self.envs, actions)]
        obs = ShardedObservation([r[0] for r in results])
        rews = np.concatenate([np.asarray(r[1]) for r in results]).flatten()
        ter = [r[2] for r in results]
        tru = [r[3] for r in results]
        assert all(t == ter[0] for t in ter) and all(t == tru[0] for t in tru)
        return (obs, rews, np.full((self.num_envs,), ter[0]), np.full((self.num_envs,), tru[0]),
                self._merge_infos([r[4] for r in results]))

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    def _merge_infos(self, infos):
        """AsyncVectorEnv._add_info (vec_env.py:221-236): per key one `[num_envs, ...]` array plus the `_key` mask."""
        out = {}
        for k in infos[0].keys():
            vals = [i[k].cpu().numpy() if isinstance(i[k], torch.Tensor) else np.asarray(i[k]) for i in infos]
            out[k] = np.concatenate(vals, axis=0)
            out[f"_{k}"] = np.ones(self.num_envs, dtype=bool)
        return out

    def close(self):
        self.envs = []
