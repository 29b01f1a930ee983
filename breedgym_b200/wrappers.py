"""Single-env action wrappers (drop-in for `breedgym.wrappers`).

Behavioural mirror of breedgym/wrappers.py:16-118.  `SimplifiedBreedGym`
observes {GEBV, corrcoef} per individual and acts with {n_bests, n_crosses}:
truncation selection of the `n_bests` by `f_index`, a random subset of
`n_crosses` pairs of their diallel (drawn from the env's `np_random`), each
repeated ceil(n / n_crosses) times and cut to n.  `KBestBreedGym` acts with the
single integer `n_bests` and crosses the full diallel.
"""
from __future__ import annotations

from math import ceil, sqrt
from typing import Callable, Optional

import numpy as np
import torch

from .breedgym import BreedGym
from .gym_compat import Wrapper, spaces
from .utils.index_functions import yield_index


def _unit_box(n: int):
    return spaces.Box(-1, 1, shape=(n,))


class SimplifiedBreedGym(Wrapper):

    metadata = BreedGym.metadata

    def __init__(self, env: Optional[BreedGym] = None, individual_per_gen: int = 2250,
                 f_index: Optional[Callable] = None, **kwargs):
        super().__init__(BreedGym(**kwargs) if env is None else env)
        n = self.individual_per_gen = individual_per_gen
        self.f_index = yield_index(self.env.simulator.GEBV_model) if f_index is None else f_index
        self.observation_space = spaces.Dict({"GEBV": _unit_box(n), "corrcoef": _unit_box(n)})
        self.action_space = spaces.Dict({
            "n_bests": spaces.Discrete(n - 1, start=2),
            "n_crosses": spaces.Discrete(n, start=1),
        })

    def reset(self, seed: Optional[int] = None, options: Optional[dict] = None):
        options = dict(options or {}, n_individuals=self.individual_per_gen)
        if "index" in options:
            self.f_index = options["index"]
        _, info = self.env.reset(seed=seed, options=options)
        return self._simplified_obs(), info

    def _plan_crosses(self, n_bests: int, n_crosses: int) -> np.ndarray:
        """(parent, parent) pairs among the already-selected `n_bests` (indices 0..n_bests-1)."""
        n = self.individual_per_gen
        pairs = self.simulator._diallel_indices(np.arange(n_bests))
        chosen = self.np_random.choice(len(pairs), n_crosses, replace=False)
        return np.repeat(pairs[chosen], ceil(n / n_crosses), axis=0)[:n]

    def step(self, action: dict):
        n_bests, n_crosses = action["n_bests"], action["n_crosses"]
        if n_bests < 2:
            raise ValueError("n_bests must be higher or equal to 2")
        if n_crosses > self.individual_per_gen:
            raise ValueError("n_crosses must be lower or equal to individual_per_gen")

        base = self.unwrapped
        base.population, _ = self.simulator.select(population=base.population, k=n_bests, f_index=self.f_index)
        _, rew, terminated, truncated, info = self.env.step(self._plan_crosses(n_bests, n_crosses))
        return self._simplified_obs(), rew, terminated, truncated, info

    @staticmethod
    def _correlation(population) -> np.ndarray:
        """Cosine similarity of every individual's centred dosage (dosage - 1) with the population mean."""
        centred = population.to_bool().sum(dim=-1).to(torch.float32) - 1.0
        mean_ind = centred.mean(dim=0)
        scale = torch.linalg.norm(centred, dim=-1) * torch.linalg.norm(mean_ind)
        return (centred @ mean_ind / scale).cpu().numpy()

    def _simplified_obs(self) -> dict:
        return {
            "GEBV": self.GEBV["Yield"].to_numpy(),
            "corrcoef": SimplifiedBreedGym._correlation(self.population),
        }


class KBestBreedGym(SimplifiedBreedGym):
    def __init__(self, *args, **kwargs) -> None:
        super().__init__(*args, **kwargs)
        # largest x with x (x - 1) / 2 < individual_per_gen
        max_best = int((1 + sqrt(1 + 8 * self.individual_per_gen)) // 2)
        self.action_space = spaces.Discrete(max_best - 1, start=2)

    def step(self, action: int):
        return super().step({"n_bests": action, "n_crosses": action * (action - 1) // 2})
