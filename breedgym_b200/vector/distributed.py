"""Single-process multi-device vector env (API mirror of `DistributedBreedGym`).

The reference (breedgym/vector/vec_env.py:150-236) spawns one subprocess per
device through Gymnasium's AsyncVectorEnv and moves every observation through
host memory.  This keeps the constructor / reset / step contract -- shard i is a
`VecBreedGym(envs_per_device, autoreset=False)` on device i, `reset(seed=s)`
seeds shard i with s + i (AsyncVectorEnv's convention) -- but holds all shards in
one process and leaves the observations on their GPUs.  For throughput use
`ShardedVecBreedGym` (one process per GPU).
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np
import torch

from .vec_env import _VecBreedGym


class DistributedBreedGym:
    def __init__(self, envs_per_device: int, initial_population, devices: Optional[List[int]] = None, **kwargs):
        if devices is None:
            devices = list(range(torch.cuda.device_count()))
        self.devices = [d if isinstance(d, int) else torch.device(d).index for d in devices]
        self.envs_per_device = envs_per_device
        kwargs.pop("autoreset", None)
        self.envs = [
            _VecBreedGym(envs_per_device, initial_population=initial_population, device=d, autoreset=False, **kwargs)
            for d in self.devices
        ]
        self.num_envs = envs_per_device * len(self.devices)
        first = self.envs[0]
        self.single_observation_space = first.single_observation_space
        self.single_action_space = first.single_action_space
        from ..gym_compat import spaces

        obs_shape = (self.num_envs, *first.single_observation_space.shape)
        act_shape = (self.num_envs, *first.single_action_space.shape)
        self.observation_space = spaces.Box(low=0, high=1, shape=obs_shape, dtype=np.int8)
        self.action_space = spaces.Box(low=0, high=first.individual_per_gen, shape=act_shape, dtype=np.int32)

    def reset(self, seed: Optional[int] = None, options: Optional[dict] = None):
        obs, infos = [], []
        for i, env in enumerate(self.envs):
            o, info = env.reset(seed=None if seed is None else seed + i, options=options)
            obs.append(o)
            infos.append(info)
        return obs, self._merge_infos(infos)

    def step(self, actions):
        actions = np.asarray(actions).reshape(len(self.devices), self.envs_per_device, *np.shape(actions)[1:])
        results = [env.step(a) for env, a in zip(self.envs, actions)]
        obs = [r[0] for r in results]
        rews = np.concatenate([np.asarray(r[1]) for r in results])
        ter = [r[2] for r in results]
        tru = [r[3] for r in results]
        assert all(t == ter[0] for t in ter) and all(t == tru[0] for t in tru)
        return (obs, rews, np.full((self.num_envs,), ter[0]), np.full((self.num_envs,), tru[0]),
                self._merge_infos([r[4] for r in results]))

    @staticmethod
    def _merge_infos(infos):
        out = {}
        for k in infos[0].keys():
            vals = [i[k].cpu().numpy() if isinstance(i[k], torch.Tensor) else np.asarray(i[k]) for i in infos]
            out[k] = np.concatenate(vals, axis=0)
        return out

    def close(self):
        self.envs = []
