"""Single-environment Gymnasium surface (drop-in for `breedgym:BreedGym`).

Mirrors breedgym/breedgym.py:23-242 of the reference: same constructor
arguments, spaces, `reset`/`step` contract, caching of `GEBV` / `corrcoef` by
object identity, and reward rule.  The population is a device-resident
`PackedPopulation`; `population[action]` is a lazy view and `simulator.cross`
fuses the parent gather into the meiosis kernel.
"""
from __future__ import annotations

from pathlib import Path
from typing import Optional, Tuple, Union

import numpy as np
import pandas as pd

from .gym_compat import Env, spaces
from .population import PackedPopulation
from .simulator import Simulator
from .utils.paths import DATA_PATH

GENOME_FILE = DATA_PATH.joinpath("small_geno.npy")


class BreedGym(Env):

    metadata = {"render_modes": ["matplotlib"], "render_fps": 1}

    def __init__(
        self,
        initial_population: Union[str, Path, np.ndarray, PackedPopulation] = GENOME_FILE,
        num_generations: int = 10,
        reward_shaping: bool = False,
        render_mode: Optional[str] = None,
        render_kwargs: Optional[dict] = None,
        **kwargs,
    ):
        self.simulator = Simulator(**kwargs)
        if isinstance(initial_population, (str, Path)):
            germplasm = self.simulator.load_population(initial_population)
        else:
            germplasm = self.simulator.as_packed(initial_population)
        self.device = self.simulator.device
        self.germplasm = germplasm

        self.num_generations = num_generations
        self.reward_shaping = reward_shaping

        self._population = None
        self._GEBV, self._GEBV_cache = None, False
        self._corrcoef, self._corrcoef_cache = None, False
        self._set_spaces(self.germplasm.shape)

        self.step_idx = None
        self.episode_idx = -1
        self.render_mode = render_mode
        if self.render_mode is not None:
            self.render_kwargs = dict(render_kwargs or {})
            self.render_kwargs.setdefault("colors", ["b", "g", "r", "c", "m"])
            self.render_kwargs.setdefault("offset", 0)
            self.render_kwargs.setdefault("traits", self.simulator.trait_names)
            self.render_kwargs.setdefault("other_features", [lambda: self.corrcoef])
            self.render_kwargs.setdefault("feature_names", ["corrcoef"])
            self.render_kwargs.setdefault("episode_names", "Episode {:d}")
            self.axs = self._make_axs()

    # ---- spaces ------------------------------------------------------------------
    def _set_spaces(self, shape):
        n = shape[0]
        self.observation_space = spaces.Box(0, 1, shape=shape, dtype=np.bool_)
        self.action_space = spaces.Sequence(spaces.Tuple((spaces.Discrete(n), spaces.Discrete(n))))

    def _update_spaces(self):
        shape = self.population.shape
        if getattr(self, "_spaces_shape", None) != shape:  # the spaces depend on the population size only
            self._set_spaces(shape)
            self._spaces_shape = shape

    # ---- gym API -----------------------------------------------------------------
    def reset(self, seed: Optional[int] = None, options: Optional[dict] = None):
        super().reset(seed=seed)
        if seed is not None:
            self.simulator.set_seed(seed=seed)

        self.step_idx = 0
        self.episode_idx += 1
        if options is not None and "n_individuals" in options.keys():
            selected = self.np_random.choice(len(self.germplasm), options["n_individuals"], replace=False)
            self.population = self.germplasm[selected]
        else:
            self.population = self.germplasm

        self._update_spaces()
        info = self._get_info()
        if self.render_mode is not None:
            self._render_step(info)
        return self.population, info

    def step(self, action):
        """`action`: int array `n x 2`, one (parent, parent) index pair per cross."""
        action = np.asarray(action)
        if action.ndim != 2 or action.shape[1] != 2:
            raise ValueError(f"action must have shape (n, 2), got {action.shape}")
        # `parents = self.population[action]; self.simulator.cross(parents)` and the GEBV of the offspring
        # (breedgym/breedgym.py:142-143, 233) in one library call
        self.population, gebv = self.simulator.cross_and_score(self.simulator.as_packed(self.population), action)
        self._GEBV = self.simulator._gebv_frame(gebv)
        self._GEBV_cache = True
        self.step_idx += 1
        self._update_spaces()

        info = self._get_info()
        if self.render_mode is not None:
            self._render_step(info)

        truncated = self.step_idx == self.num_generations
        if self.reward_shaping or truncated:
            reward = np.mean(self.GEBV.to_numpy())
        else:
            reward = 0
        return self.population, reward, False, truncated, info

    def _get_info(self):
        return {"GEBV": self.GEBV}

    @property
    def population(self):
        return self._population

    @population.setter
    def population(self, new_pop):
        self._population = new_pop
        self._GEBV_cache = False
        self._corrcoef_cache = False

    @property
    def GEBV(self) -> pd.DataFrame:
        """GEBV of every individual for every trait (`n x t` DataFrame), cached per population."""
        if not self._GEBV_cache:
            self._GEBV = self.simulator.GEBV(self.population)
            self._GEBV_cache = True
        return self._GEBV

    @property
    def corrcoef(self):
        if not self._corrcoef_cache:
            self._corrcoef = self.simulator.corrcoef(self.population)
            self._corrcoef_cache = True
        return self._corrcoef

    # ---- rendering: out of scope -- the constructor arguments are accepted for drop-in
    # compatibility, the plotting itself (breedgym/breedgym.py:159-222 of the reference) is not part of the hot path
    def _make_axs(self):
        return self.render_kwargs.get("axs")

    def _render_step(self, info: dict):
        pass

    def render(self, file_name: Optional[Union[str, Path]] = None):
        if self.render_mode is not None:
            raise NotImplementedError("breedgym_b200 does not render: plot env.GEBV / env.corrcoef with the reference's "
                                      "matplotlib code (breedgym/breedgym.py:159-222) if needed")
