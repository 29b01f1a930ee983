"""Action translators on top of VecBreedGym (drop-in for breedgym.vector.vec_wrappers).

Mirrors breedgym/vector/vec_wrappers.py:15-155: `SelectionScores` (scores ->
top-k -> random subset of the diallel -> repeat), `PairScores` (n x n pair scores
-> top-n pairs, softmax-proportional offspring counts) and `RavelIndex`.  The
index math is batched over the envs and runs on the GPU: `SelectionScores` with a
stable sort + the library's jax-compatible permutation kernel (a NumPy version of
the same translation is kept for very large k and as the cross-check), `PairScores`
with a stable sort of the E x n^2 scores; the resulting `int32[E, n, 2]` pairs feed
the unchanged `VecBreedGym.step` hot path.
"""
from __future__ import annotations

from math import ceil, prod
from typing import Optional

import numpy as np
import torch

from .. import _lib, jaxlike
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

    def _convert_actions(self, actions: np.ndarray, random_keys: np.ndarray) -> np.ndarray:
        """Scores `[E, n]` + one key per env -> parent pairs `int32[E, n, 2]`, vectorised over the envs:
        top-k (ties -> lower index), the k(k-1)/2 pairs of the best in upper-triangular order, a random subset of
        `n_crosses` of them (`jax.random.choice(replace=False)` = first entries of a permutation) and each pair
        repeated ceil(n / n_crosses) times, cut to n."""
        n = self.individual_per_gen
        _, best = jaxlike.top_k(np.asarray(actions), self.k)                       # [E, k]
        ia, ib = np.triu_indices(self.k, k=1)
        diallel = np.stack([best[:, ia], best[:, ib]], axis=-1)                    # [E, C(k,2), 2]
        if self.n_crosses > diallel.shape[1]:
            raise ValueError("Cannot take a larger sample than population when 'replace=False'")
        perm = jaxlike.permutation_batch(random_keys, diallel.shape[1], self.simulator.rng_layout)
        chosen = np.take_along_axis(diallel, perm[:, : self.n_crosses, None], axis=1)  # [E, n_crosses, 2]
        out = np.repeat(chosen, int(ceil(n / self.n_crosses)), axis=1)[:, :n]
        if out.shape[1] < n:  # jnp.repeat(total_repeat_length=n) pads with the last entry
            out = np.concatenate([out, np.repeat(out[:, -1:], n - out.shape[1], axis=1)], axis=1)
        return np.ascontiguousarray(out, dtype=np.int32)

    _DEVICE_MAX_PAIRS = 16000  # the on-device permutation sorts the C(k, 2) pairs in shared memory

    def _convert_actions_device(self, actions, random_key: np.ndarray) -> torch.Tensor:
        """The same translation on the GPU, batched over the envs: a stable descending sort (= top-k, ties -> lower
        index: `bg_topk`), the library's permutation kernel for the random subset (env g uses keys[1 + g] of
        split(random_key, total + 1): exactly the draw `VecBreedGym.reset` makes, `bg_reset_indices`) and one kernel
        for the pair lookup + repeat (`bg_diallel_pairs`): three launches, no host round trip."""
        sim, dev = self.simulator, self.device
        E, n, k, nc = self.num_envs, self.individual_per_gen, self.k, self.n_crosses
        begin, total = self.env.env_shard
        s = torch.as_tensor(actions, device=dev).to(torch.float32).reshape(E, n)  # jax computes in float32 as well
        best = sim._top_k(s, k).to(torch.int32).contiguous()                          # [E, k] (bg_topk: descending, ties -> lower index)
        n_pairs = k * (k - 1) // 2
        perm = torch.empty((E, nc), dtype=torch.int32, device=dev)
        key = np.ascontiguousarray(random_key, dtype=np.uint32)
        _lib.check(_lib.load().bg_reset_indices(sim._engine, _lib.nptr(key), total, begin, E, n_pairs, nc, sim._layout_id,
                                                perm.data_ptr(), sim._stream()))
        # chosen entries of the upper-triangular pair list, repeated ceil(n / nc) times, cut / padded to n: one kernel
        out = torch.empty((E, n, 2), dtype=torch.int32, device=dev)
        _lib.check(_lib.load().bg_diallel_pairs(sim._engine, best.data_ptr(), perm.data_ptr(), E, k, nc, n, out.data_ptr(), sim._stream()))
        return out

    def step(self, actions):
        begin, total = self.env.env_shard
        key = np.array(self.random_key, dtype=np.uint32, copy=True)
        n_pairs = self.k * (self.k - 1) // 2
        if n_pairs <= self._DEVICE_MAX_PAIRS and self.n_crosses <= n_pairs:
            self.env.random_key = _lib.key_split_at(key, 0, total + 1, self.simulator.rng_layout)
            low_level_actions = self._convert_actions_device(actions, key)
        else:
            random_keys = self.simulator._split(key, total + 1)
            self.env.random_key = random_keys[0]
            mine = random_keys[1 + begin:1 + begin + self.num_envs]
            low_level_actions = self._convert_actions(_to_host(actions), mine)
        return super().step(low_level_actions)


def _pairs_from_scores(scores, n_crosses: int, device, sim: Optional[Simulator] = None) -> torch.Tensor:
    """Pair scores `[E, n, n]` -> parent pairs `int32[E, n_crosses, 2]` on the GPU.

    Per env (breedgym/vector/vec_wrappers.py:100-112): the n_crosses best pairs (descending, ties -> lower flat
    index), offspring per pair = ceil(softmax(best values) * n_crosses), `jnp.repeat(..., total_repeat_length=n_crosses)`
    (truncate, or pad with the last pair).  The top-k over the E x n^2 scores is the library's radix-select kernel
    (`bg_topk`) when a simulator is given; the torch formulation below it is the CPU cross-check."""
    s = torch.as_tensor(scores, dtype=torch.float32, device=device)
    E, n = s.shape[0], s.shape[-1]
    flat = s.reshape(E, -1) + 0.0  # -0.0 -> +0.0: the two compare equal, so they must get the same key
    L = flat.shape[1]  # scores are [E, a, n]: flat index -> (index // n, index % n)
    if sim is not None and flat.is_cuda and n_crosses <= 1024:
        flat = flat.contiguous()
        vals = torch.empty((E, n_crosses), dtype=torch.float32, device=flat.device)
        idx32 = torch.empty((E, n_crosses), dtype=torch.int32, device=flat.device)
        lib = _lib.load()
        _lib.check(lib.bg_topk(sim._engine, flat.data_ptr(), E, L, n_crosses, vals.data_ptr(), idx32.data_ptr(), sim._stream()))
        # softmax -> counts -> prefix sums -> repeat -> (row, column): one kernel (`bg_pairs_from_topk`, csrc/pairs.cu)
        pairs = torch.empty((E, n_crosses, 2), dtype=torch.int32, device=flat.device)
        _lib.check(lib.bg_pairs_from_topk(sim._engine, vals.data_ptr(), idx32.data_ptr(), E, n_crosses, n, pairs.data_ptr(), sim._stream()))
        return pairs
    else:
        # top-n_crosses with jax.lax.top_k's order (descending, ties -> lower flat index) WITHOUT sorting all n^2 scores:
        # 64-bit keys = (order-preserving integer image of the float32 score, inverted flat index) are unique, so a plain
        # (unstable) top-k of the keys is that order exactly
        bits = flat.view(torch.int32)
        ordered = bits ^ ((bits >> 31) & 0x7FFFFFFF)  # monotone in the float value (sign-magnitude -> two's complement)
        low = (L - 1) - torch.arange(L, device=s.device, dtype=torch.int64)
        keys = (ordered.to(torch.int64) << 32) + low
        top = torch.topk(keys, n_crosses, dim=1, largest=True, sorted=True).values
        idx = (L - 1) - (top & 0xFFFFFFFF)
        vals = torch.gather(flat, 1, idx)
    reps = torch.ceil(torch.softmax(vals, dim=1) * n_crosses).to(torch.int64)
    ends = torch.cumsum(reps, dim=1)  # pair b fills output slots [ends[b-1], ends[b])
    slots = torch.arange(n_crosses, device=s.device).expand(E, n_crosses).contiguous()
    which = torch.searchsorted(ends, slots, right=True).clamp_(max=n_crosses - 1)
    flat = torch.gather(idx, 1, which)
    return torch.stack((flat // n, flat % n), dim=-1).to(torch.int32)


class PairScores(VectorWrapper):
    def __init__(self, vec_env: Optional[VecBreedGym] = None, **kwargs):
        if vec_env is None:
            vec_env = VecBreedGym(**kwargs)
        super().__init__(vec_env)

        self.n_crosses = self.individual_per_gen
        action_shape = self.n_crosses, self.n_crosses
        self.single_action_space = spaces.Box(-1e5, 1e5, shape=action_shape)
        self.action_space = spaces.Box(-1e5, 1e5, shape=(self.num_envs, *action_shape))

    def _convert_actions(self, actions) -> torch.Tensor:
        return _pairs_from_scores(actions, self.n_crosses, self.device, self.simulator)

    def step(self, actions):
        low_level_actions = self._convert_actions(actions)
        obs, rew, ter, tru, infos = super().step(low_level_actions)
        infos["low_level_actions"] = low_level_actions.cpu().numpy()
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
