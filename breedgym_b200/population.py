"""Device-resident bit-packed populations.

The reference keeps populations as `bool[..., n, m, 2]` device arrays
(breedgym/breedgym.py:42, breedgym/vector/vec_env.py:52).  Here the state lives
in HBM as two bit planes per individual (`int32[..., n, 2, Wpad]`, see
include/breedgym_b200.h); `PackedPopulation` quacks like the reference's array
(`shape`, `len`, indexing, `np.asarray`) and materialises the byte layout only
when asked to.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np
import torch

from . import _lib


def _is_int_index(idx) -> bool:
    """A single integer ARRAY index (list / ndarray / tensor).  Tuples are multi-axis indices, handled separately."""
    if isinstance(idx, list):
        try:
            idx = np.asarray(idx)
        except ValueError:
            return False
    if isinstance(idx, torch.Tensor):
        return not (idx.dtype.is_floating_point or idx.dtype == torch.bool)
    return isinstance(idx, np.ndarray) and idx.dtype.kind in "iu"


class PackedPopulation:
    """`bool[*lead, n, m, 2]` population stored as bit planes on the GPU."""

    __array_priority__ = 1000

    def __init__(self, sim, words: torch.Tensor):
        assert words.dtype == torch.int32 and words.shape[-1] == sim.words_per_row and words.shape[-2] == 2
        self.sim = sim
        self.words = words
        self._bool = None

    @classmethod
    def _trusted(cls, sim, words: torch.Tensor) -> "PackedPopulation":
        """Wrap words the library has just produced (skips the layout checks: the env's per-step path)."""
        self = cls.__new__(cls)
        self.sim = sim
        self.words = words
        self._bool = None
        return self

    # ---- array-like surface --------------------------------------------------
    @property
    def shape(self) -> Tuple[int, ...]:
        return tuple(self.words.shape[:-2]) + (self.sim.n_markers, 2)

    @property
    def ndim(self) -> int:
        return self.words.dim()

    @property
    def dtype(self):
        return np.dtype(np.bool_)

    @property
    def device(self):
        return self.words.device

    def __len__(self) -> int:
        return self.words.shape[0]

    def __repr__(self):
        return f"PackedPopulation(shape={self.shape}, device={self.words.device})"

    def to_bool(self) -> torch.Tensor:
        """Materialise `bool[*lead, n, m, 2]` on the device (cached)."""
        if self._bool is None:
            rows = int(np.prod(self.words.shape[:-2])) if self.words.dim() > 2 else 1
            out = torch.empty(self.shape, dtype=torch.bool, device=self.words.device)
            w = self.words.contiguous()
            _lib.check(_lib.load().bg_unpack(self.sim._engine, w.data_ptr(), out.data_ptr(), rows, self.sim._stream()))
            self._bool = out
        return self._bool

    def numpy(self) -> np.ndarray:
        return self.to_bool().cpu().numpy()

    def __array__(self, dtype=None, copy=None):
        a = self.numpy()
        return a if dtype is None else a.astype(dtype)

    def __eq__(self, other):
        return self.numpy() == np.asarray(other)

    __hash__ = None

    def __iter__(self):
        for i in range(len(self)):
            yield self[i]

    def reshape(self, *shape):
        """Reshape of the leading (env / individual) dims only."""
        if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
            shape = tuple(shape[0])
        if tuple(shape[-2:]) != (self.sim.n_markers, 2):
            return self.numpy().reshape(*shape)
        return PackedPopulation(self.sim, self.words.reshape(*shape[:-2], 2, self.sim.words_per_row))

    def __getitem__(self, idx):
        if isinstance(idx, (int, np.integer)):
            i = int(idx)
            if self.words.dim() == 3:  # one individual: small, hand back its bool[m, 2] tensor
                i = i + len(self) if i < 0 else i
                return PackedPopulation(self.sim, self.words[i:i + 1]).to_bool()[0]
            return PackedPopulation(self.sim, self.words[i])
        if isinstance(idx, slice):
            return PackedPopulation(self.sim, self.words[idx])
        if isinstance(idx, tuple):
            # multi-axis index.  The reference's vector-env idiom `populations[arange(E)[:, None, None], actions]`
            # (breedgym/vector/vec_env.py:89-90, breeding_programs_env.py) is kept lazy: a per-env parents view that
            # `Simulator.cross` hands to the kernel as index pairs; everything else is plain numpy indexing.
            if len(idx) == 2 and self.words.dim() == 4:
                env_ix, act = idx
                act_a = act if isinstance(act, torch.Tensor) else np.asarray(act)
                env_a = np.asarray(env_ix.cpu() if isinstance(env_ix, torch.Tensor) else env_ix)
                E = self.words.shape[0]
                if (act_a.ndim == 3 and act_a.shape[0] == E and act_a.shape[2] == 2 and _is_int_index(act_a)
                        and env_a.shape == (E, 1, 1) and np.array_equal(env_a.ravel(), np.arange(E))):
                    return ParentsView(self, act_a)
            idx = tuple(i.cpu().numpy() if isinstance(i, torch.Tensor) else i for i in idx)
            return self.numpy()[idx]
        if _is_int_index(idx) and self.words.dim() == 3:
            ia = idx if isinstance(idx, torch.Tensor) else np.asarray(idx)
            if ia.ndim == 1:
                return self.sim._gather(self, ia)
            if ia.ndim == 2 and ia.shape[1] == 2:
                return ParentsView(self, ia)
        if isinstance(idx, torch.Tensor):
            idx = idx.cpu().numpy()
        return self.numpy()[idx]


class ParentsView:
    """Lazy `population[action]` (breedgym/breedgym.py:142): `bool[n, 2, m, 2]`.

    Never materialised on the cross path: `Simulator.cross` hands the index
    pairs to the kernel, which reads the parents' bit planes in place.
    """

    def __init__(self, population: PackedPopulation, pairs):
        self.population = population
        self.pairs = pairs

    @property
    def shape(self):
        return tuple(self.pairs.shape) + self.population.shape[-2:]

    def __len__(self):
        return len(self.pairs)

    def __array__(self, dtype=None, copy=None):
        pairs = self.pairs.cpu().numpy() if isinstance(self.pairs, torch.Tensor) else np.asarray(self.pairs)
        pop = self.population.numpy()
        if pairs.ndim == 3:  # per-env view: populations[arange(E)[:, None, None], actions] -> [E, n, 2, m, 2]
            n = pop.shape[1]
            pairs = np.clip(np.where(pairs < 0, pairs + n, pairs), 0, n - 1)
            a = pop[np.arange(pop.shape[0])[:, None, None], pairs]
        else:
            n = len(self.population)
            pairs = np.clip(np.where(pairs < 0, pairs + n, pairs), 0, n - 1)
            a = pop[pairs]
        return a if dtype is None else a.astype(dtype)
