"""Action translators on top of VecBreedGym (drop-in for breedgym.vector.vec_wrappers).

Mirrors breedgym/vector/vec_wrappers.py:15-155: `SelectionScores` (scores ->
top-k -> random subset of the diallel -> repeat), `PairScores` (n x n pair scores
-> top-n pairs, softmax-proportional offspring counts) and `RavelIndex`.  The
index math runs on the host (`breedgym_b200.jaxlike`); the resulting
`int32[E, n, 2]` pairs feed the unchanged `VecBreedGym.step` hot path.
"""
from __future__ import annotations

from math import ceil, prod
from typing import Optional

import numpy as np
import torch

from .. import jaxlike
from ..gym_compat import VectorWrapper, spaces
from ..simulator import Simulator
from .vec_env import VecBreedGym


def _to_host(a) -> np.ndarray:
    return a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)


class SelectionScores(VectorWrapper):
    def __init__(self, vec_env: Optional[VecBreedGym] = None, k: Optional[int] = None,
                 n_crosses: Optional[int] = None, **kwargs):
        if vec_env is None:
            vec_env = VecBreedGym(**kwargs)
        super().__init__(vec_env)

        if k is None:
            k = self.individual_per_gen // 10
        elif k > self.individual_per_gen:
            raise ValueError(f"Cannot select {k} best from a population of ",
                             f"{self.individual_per_gen} individuals")
        self.k = k

        max_crosses = self.k * (self.k - 1) // 2
        if n_crosses is None:
            n_crosses = max_crosses
        elif n_crosses > max_crosses:
            raise ValueError("Incompatible value for k and n_crosses. ",
                             f"With k={k}, the maximum number of crosses is {max_crosses}")
        self.n_crosses = n_crosses

        if self.n_crosses > self.individual_per_gen:
            raise ValueError("Invalid combination for k and n_crosses. ",
                             f"Resulting population size will be {n_crosses} ",
                             f"that is grater than {self.individual_per_gen}")

        self.single_action_space = spaces.Box(-1e5, 1e5, shape=(self.individual_per_gen,))
        self.action_space = spaces.Box(-1e5, 1e5, shape=(self.num_envs, self.individual_per_gen))

    def _convert_action(self, action: np.ndarray, random_key: np.ndarray) -> np.ndarray:
        n = self.individual_per_gen
        _, best_pop = jaxlike.top_k(action, self.k)
        diallel = Simulator._diallel_indices(best_pop)
        sel = jaxlike.choice_no_replace(random_key, len(diallel), self.n_crosses, self.simulator.rng_layout)
        cross_indices = diallel[sel]
        return jaxlike.repeat_total(cross_indices, int(ceil(n / self.n_crosses)), n)

    def _convert_actions(self, actions: np.ndarray, random_keys: np.ndarray) -> np.ndarray:
        return np.stack([self._convert_action(a, k) for a, k in zip(actions, random_keys)]).astype(np.int32)

    def step(self, actions):
        begin, total = self.env.env_shard
        random_keys = self.simulator._split(self.random_key, total + 1)
        self.env.random_key = random_keys[0]
        mine = random_keys[1 + begin:1 + begin + self.num_envs]
        low_level_actions = self._convert_actions(_to_host(actions), mine)
        return super().step(low_level_actions)


def _pairs_from_scores(action: np.ndarray, n_crosses: int) -> np.ndarray:
    """Top-n pairs of an n x n score matrix with softmax-proportional offspring counts."""
    best_values, best_crosses = jaxlike.top_k(action.reshape(-1), n_crosses)
    offspring_per_cross = jaxlike.softmax_f32(best_values) * np.float32(n_crosses)
    cross_indices = np.stack((best_crosses // n_crosses, best_crosses % n_crosses), axis=1)
    return jaxlike.repeat_total(cross_indices, np.ceil(offspring_per_cross).astype(np.int32), n_crosses)


class PairScores(VectorWrapper):
    def __init__(self, vec_env: Optional[VecBreedGym] = None, **kwargs):
        if vec_env is None:
            vec_env = VecBreedGym(**kwargs)
        super().__init__(vec_env)

        self.n_crosses = self.individual_per_gen
        action_shape = self.n_crosses, self.n_crosses
        self.single_action_space = spaces.Box(-1e5, 1e5, shape=action_shape)
        self.action_space = spaces.Box(-1e5, 1e5, shape=(self.num_envs, *action_shape))

    def _convert_actions(self, actions: np.ndarray) -> np.ndarray:
        return np.stack([_pairs_from_scores(a, self.n_crosses) for a in actions]).astype(np.int32)

    def step(self, actions):
        low_level_actions = self._convert_actions(_to_host(actions))
        obs, rew, ter, tru, infos = super().step(low_level_actions)
        infos["low_level_actions"] = low_level_actions
        return obs, rew, ter, tru, infos


class RavelIndex(VectorWrapper):
    def __init__(self, vec_env):
        super().__init__(vec_env)
        self.action_shape = tuple(self.env.single_action_space.shape)
        n_elems = prod(self.action_shape)
        n_vec = np.full((self.individual_per_gen,), n_elems)
        self.single_action_space = spaces.MultiDiscrete(n_vec)
        self.action_space = spaces.MultiDiscrete(np.broadcast_to(n_vec[None, ...], (self.num_envs, *n_vec.shape)))

    def _convert_actions(self, actions: np.ndarray) -> np.ndarray:
        return np.stack(np.unravel_index(actions, self.action_shape), axis=-1).astype(np.int32)

    def step(self, actions):
        return super().step(self._convert_actions(_to_host(actions)))
