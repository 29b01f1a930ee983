"""Wheat breeding-program wrapper (drop-in for breedgym.vector.WheatBreedGym).

Mirrors breedgym/vector/breeding_programs_env.py:12-72: cross the top pairs ->
`plant_per_line` double haploids per line -> keep `k_per_line` per line ->
global selection back to `individual_per_gen`.  Like the reference's vmapped
calls, every env shares the simulator's key for each of the two random stages.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import _lib
from ..gym_compat import VectorWrapper, spaces
from ..population import PackedPopulation
from .vec_wrappers import _pairs_from_scores, _to_host


class WheatBreedGym(VectorWrapper):
    def __init__(self, vec_env, n_lines=200, plant_per_line=100, k_per_line=5):
        super().__init__(vec_env)
        self.n_lines = n_lines
        self.plant_per_line = plant_per_line
        self.k_per_line = k_per_line
        action_shape = self.n_lines, self.n_lines
        self.single_action_space = spaces.Box(-1e5, 1e5, shape=action_shape)
        self.action_space = spaces.Box(-1e5, 1e5, shape=(self.num_envs, *action_shape))

    def _convert_actions(self, actions) -> torch.Tensor:
        return _pairs_from_scores(actions, self.n_lines, self.device, self.simulator)

    def _index(self, pop: PackedPopulation) -> torch.Tensor:
        return self.simulator.GEBV_model(pop).sum(dim=-1)

    def _take(self, pop_words: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
        """pop_words [G, n_src, 2, W], idx [G, k] -> [G, k, 2, W] (per-group gather on the GPU)."""
        sim = self.simulator
        G, n_src = pop_words.shape[:2]
        k = idx.shape[1]
        out = sim._empty_words(G, k)
        idx = idx.to(torch.int32).contiguous()
        _lib.check(_lib.load().bg_gather_individuals(sim._engine, pop_words.contiguous().data_ptr(), idx.data_ptr(),
                                                     out.data_ptr(), G, n_src, k, n_src, sim._stream()))
        return out

    def _topk(self, values: torch.Tensor, k: int) -> torch.Tensor:
        # descending, ties -> lower index (lax.top_k): the library's radix-select kernel
        return self.simulator._top_k(values, k)

    def step(self, actions):
        env, sim = self.env, self.simulator
        E = self.num_envs
        pairs = self._convert_actions(actions)
        pop = self.cross(pairs)  # [E, n_lines]
        assert pop.shape[1] == self.n_lines

        # double haploids: one key for all envs, one launch (vmap(simulator.double_haploid, in_axes=(None, 0)))
        dh = sim.double_haploid(pop, n_offspring=self.plant_per_line).words
        if dh.dim() == 4:  # plant_per_line == 1 comes back squeezed
            dh = dh.unsqueeze(2)
        W = sim.words_per_row
        # best k_per_line of every line
        lines = dh.reshape(E * self.n_lines, self.plant_per_line, 2, W)
        best = self._topk(self._index(PackedPopulation(sim, lines)), self.k_per_line)
        kept = self._take(lines, best).reshape(E, self.n_lines * self.k_per_line, 2, W)
        # global selection
        best = self._topk(self._index(PackedPopulation(sim, kept)), self.individual_per_gen)
        env.populations = PackedPopulation(sim, self._take(kept, best))

        env.step_idx += 1
        infos = self.get_info()
        done = env.step_idx == self.num_generations
        if self.reward_shaping or done:
            rews = np.asarray(np.max(_to_host(infos["GEBV"]), axis=(1, 2)))
        else:
            rews = np.zeros(E)

        if done and self.autoreset:
            self.reset()

        terminated = np.full(E, False)
        truncated = np.full(E, done)
        return self.populations, rews, terminated, truncated, infos
