"""`Simulator` -- drop-in for the `chromax.Simulator` surface BreedGym uses.

The reference's operator API for the hot path is the Python object
`chromax.Simulator`: constructed at
breedgym/breedgym.py:36 and breedgym/vector/vec_env.py:45, then
`.load_population`, `.set_seed`, `.cross`, `.GEBV`, `.GEBV_model`, `.corrcoef`,
`.select`, `._diallel_indices`, `.double_haploid`.  This class keeps the same
names, argument meaning and error behaviour, does the host-only work (map
parsing, the Threefry key chain) and enqueues the sm_100a kernels of
libbreedgym_b200 through the C ABI.  Per element, Python computes nothing.
"""
from __future__ import annotations

import ctypes
import random
from pathlib import Path
from typing import Callable, List, Optional, Sequence, Tuple, Union

import numpy as np
import pandas as pd
import torch

from . import _lib
from .population import PackedPopulation, ParentsView


class TraitModel:
    """chromax.trait_model.TraitModel: `dot(sum(pop, -1), effects) + offset`."""

    def __init__(self, sim: "Simulator", marker_effects: np.ndarray, offset: float = 0.0, own_engine: bool = False):
        self._sim = sim
        self.marker_effects = np.ascontiguousarray(marker_effects, dtype=np.float32)
        self.offset = offset
        self.n_traits = self.marker_effects.shape[1]
        # a model other than the simulator's GEBV model (the GxE model) scores through an engine of its own: the same
        # kernels over its own fixed-point effect tables
        self._engine = None
        if own_engine:
            self._engine = ctypes.c_void_p()
            lib = _lib.load()
            _lib.check(lib.bg_engine_create(sim.device.index, ctypes.byref(self._engine)))
            _lib.check(lib.bg_engine_set_map(self._engine, _lib.nptr(sim.recombination_vec), _lib.nptr(self.marker_effects),
                                             sim.n_markers, self.n_traits, 0.0))

    def __call__(self, population) -> torch.Tensor:
        """`float32[..., n_traits]` on the simulator's device."""
        out = self._sim._gebv(self._sim.as_packed(population), self._engine, self.n_traits)
        return out + self.offset if self.offset else out

    def __del__(self):
        eng = getattr(self, "_engine", None)
        if eng:
            try:
                _lib.load().bg_engine_destroy(eng)
            except Exception:
                pass
            self._engine = None

    @property
    def positive_mask(self) -> np.ndarray:
        return self.marker_effects > 0

    @property
    def max(self) -> np.ndarray:
        return 2 * np.sum(self.marker_effects, axis=0, where=self.positive_mask) + self.offset

    @property
    def min(self) -> np.ndarray:
        return 2 * np.sum(self.marker_effects, axis=0, where=~self.positive_mask) + self.offset

    @property
    def mean(self) -> np.ndarray:
        return np.sum(self.marker_effects, axis=0) + self.offset

    @property
    def var(self) -> np.ndarray:
        return np.sum(self.marker_effects**2, axis=0) / 2


def _resolve_device(device) -> torch.device:
    if device is None:
        idx = torch.cuda.current_device() if torch.cuda.is_available() else 0
        return torch.device("cuda", idx)
    if isinstance(device, (int, np.integer)):
        return torch.device("cuda", int(device))
    dev = torch.device(device)
    if dev.type != "cuda":
        raise ValueError("breedgym_b200 runs on CUDA devices only (no CPU fallback)")
    return torch.device("cuda", dev.index if dev.index is not None else 0)


class Simulator:
    """Breeding simulator bound to one B200."""

    def __init__(
        self,
        genetic_map: Union[str, Path, pd.DataFrame],
        trait_names: Optional[List[str]] = None,
        chr_column: str = "CHR.PHYS",
        position_column: str = "cM",
        recombination_column: str = "RecombRate",
        mutation: float = 0.0,
        h2: Optional[np.ndarray] = None,
        seed: Optional[int] = None,
        device=None,
        backend=None,
        rng_layout: str = "legacy",
        key_schedule: str = "S2",
        engine_options: Optional[dict] = None,
    ):
        if rng_layout not in _lib.LAYOUT_ID:
            raise ValueError(f"rng_layout must be one of {list(_lib.LAYOUT_ID)}")
        if key_schedule not in _lib.SCHEDULE_ID:
            raise ValueError(f"key_schedule must be one of {list(_lib.SCHEDULE_ID)}")
        self.rng_layout = rng_layout
        self.key_schedule = key_schedule
        self._layout_id = _lib.LAYOUT_ID[rng_layout]
        self._schedule_id = _lib.SCHEDULE_ID[key_schedule]
        self._key_state = np.zeros(2, dtype=np.uint32)   # random_key, advanced in place by the C key chain
        self._chain_out = np.zeros(6, dtype=np.uint32)   # [k, next k, the k after that]
        self._key_ptr, self._chain_ptr = _lib.nptr(self._key_state), _lib.nptr(self._chain_out)
        self._chain_out_addr = self._chain_out.ctypes.data
        self._chain_fn = _lib.load().bg_key_chain_next
        self.mutation = float(mutation)
        self.device = _resolve_device(device)
        self._device_index = self.device.index

        if not isinstance(genetic_map, pd.DataFrame):
            genetic_map = pd.read_table(genetic_map, sep="\t")
        if trait_names is None:
            skip = {"MRK.NAME", chr_column, position_column, recombination_column}
            trait_names = [c for c in genetic_map.columns if c not in skip]
        self.trait_names = list(trait_names)
        self.n_markers = len(genetic_map)
        chrom = genetic_map[chr_column].to_numpy()
        first = np.ones(self.n_markers, dtype=bool)
        first[1:] = chrom[1:] != chrom[:-1]
        starts = np.flatnonzero(first)
        self.chr_lens = np.diff(np.append(starts, self.n_markers))

        if recombination_column in genetic_map.columns:
            rec = np.array(genetic_map[recombination_column].to_numpy(), dtype=np.float64, copy=True)
            rec[1:] = rec[:-1].copy()  # "recombine now" instead of "recombine after"
        elif position_column in genetic_map.columns:
            cm = np.asarray(genetic_map[position_column].to_numpy(), dtype=np.float64)
            rec = np.zeros(self.n_markers, dtype=np.float64)
            rec[1:] = 0.5 * (1.0 - np.exp(-2.0 * (cm[1:] - cm[:-1]) / 100.0))  # Haldane
        else:
            raise ValueError(f"genetic map needs a '{recombination_column}' or a '{position_column}' column")
        rec[first] = 0.5
        self.recombination_vec = rec.astype(np.float32)
        if h2 is None:
            h2 = np.full((len(self.trait_names),), 0.5)
        self.h2 = np.asarray(h2)

        effects = genetic_map[self.trait_names].to_numpy(dtype=np.float32)
        self.GEBV_model = TraitModel(self, effects)

        self._engine = ctypes.c_void_p()
        lib = _lib.load()
        _lib.check(lib.bg_engine_create(self.device.index, ctypes.byref(self._engine)))
        for name, value in (engine_options or {}).items():
            self.set_option(name, value)
        eff = self.GEBV_model.marker_effects
        _lib.check(lib.bg_engine_set_map(self._engine, _lib.nptr(self.recombination_vec), _lib.nptr(eff),
                                         self.n_markers, eff.shape[1], self.mutation))
        self.words_per_row = _lib.words_per_row(self.n_markers)

        if seed is None:
            seed = random.randint(0, 2**32)
        self.random_key = _lib.key_data(seed)
        # chromax draws the GxE effects at construction, consuming one split: `random_key, split_key = split(random_key)`.
        # The draw itself (m x T normals) is deferred until somebody asks for a phenotype (`GxE_model`).
        halves = self._split(self.random_key, 2)
        self.random_key = halves[0]
        self._gxe_key = halves[1].copy()
        self._gxe_model = None
        self.h2 = np.full(len(self.trait_names), 0.5, dtype=np.float32) if h2 is None else np.asarray(h2, dtype=np.float32)

    def __del__(self):
        eng = getattr(self, "_engine", None)
        if eng:
            try:
                _lib.load().bg_engine_destroy(eng)
            except Exception:
                pass
            self._engine = None

    # ---- helpers ---------------------------------------------------------------
    def set_option(self, name: str, value: int):
        """Engine tuning / cross-check switch (`bg_engine_set_option`; none changes results)."""
        _lib.check(_lib.load().bg_engine_set_option(self._engine, name.encode(), int(value)))

    def _stream(self):
        # raw handle of torch's current stream on this device (the Python Stream object costs ~9 us per call)
        return ctypes.c_void_p(torch._C._cuda_getCurrentRawStream(self._device_index))

    def _split(self, key, num=2) -> np.ndarray:
        return _lib.key_split(key, num, self.rng_layout)

    @property
    def random_key(self) -> np.ndarray:
        """Raw data (uint32[2]) of the simulator's jax-style PRNG key."""
        return self._key_state

    @random_key.setter
    def random_key(self, value):
        self._key_state[:] = np.asarray(value, dtype=np.uint32)

    def _next_key(self, lookahead: bool = False):
        """`random_key, k = split(random_key)` (chromax Simulator.cross); with `lookahead` also the k
        the NEXT call will get (valid unless `set_seed` intervenes).  One C call, no allocation; the
        returned arrays are views of a scratch buffer that the next call overwrites."""
        _lib.check(self._chain_fn(self._key_ptr, self._layout_id, self._chain_ptr))
        return (self._chain_out[:2], self._chain_out[2:4]) if lookahead else self._chain_out[:2].copy()

    def _layout(self) -> int:
        return _lib.LAYOUT_ID[self.rng_layout]

    def _schedule(self) -> int:
        return _lib.SCHEDULE_ID[self.key_schedule]

    def _empty_words(self, *lead) -> torch.Tensor:
        return torch.empty((*lead, 2, self.words_per_row), dtype=torch.int32, device=self.device)

    def _index_tensor(self, idx) -> torch.Tensor:
        if isinstance(idx, torch.Tensor):
            return idx.to(device=self.device, dtype=torch.int32).contiguous()
        a = np.ascontiguousarray(np.asarray(idx), dtype=np.int32)
        return torch.from_numpy(a).to(self.device)

    def as_packed(self, population) -> PackedPopulation:
        """Accept a PackedPopulation, or a bool array/tensor `[..., n, m, 2]` (packed on the device)."""
        if isinstance(population, PackedPopulation):
            if population.sim.words_per_row != self.words_per_row or population.words.device != self.device:
                raise ValueError("population belongs to a different simulator / device")
            return population
        if isinstance(population, ParentsView):
            population = np.asarray(population)
        if isinstance(population, torch.Tensor):
            b = population.to(device=self.device)
        else:
            b = torch.from_numpy(np.ascontiguousarray(np.asarray(population))).to(self.device)
        if b.dim() < 3 or b.shape[-1] != 2 or b.shape[-2] != self.n_markers:
            raise ValueError(f"population must have shape (..., n, {self.n_markers}, 2), got {tuple(b.shape)}")
        b = (b != 0).contiguous()
        lead = tuple(b.shape[:-2])
        rows = int(np.prod(lead))
        words = self._empty_words(*lead)
        _lib.check(_lib.load().bg_pack(self._engine, b.data_ptr(), words.data_ptr(), rows, self._stream()))
        return PackedPopulation(self, words)

    def _gather(self, population: PackedPopulation, idx) -> PackedPopulation:
        """`population[idx]` for a 1-D integer index (whole individuals)."""
        it = self._index_tensor(idx).reshape(1, -1)
        n = it.shape[1]
        src = population.words.contiguous()
        out = self._empty_words(n)
        _lib.check(_lib.load().bg_gather_individuals(self._engine, src.data_ptr(), it.data_ptr(), out.data_ptr(),
                                                     1, src.shape[0], n, 0, self._stream()))
        return PackedPopulation(self, out)

    def _gebv(self, population: PackedPopulation, engine=None, n_traits: Optional[int] = None) -> torch.Tensor:
        w = population.words.contiguous()
        lead = tuple(w.shape[:-2])
        rows = int(np.prod(lead))
        T = self.GEBV_model.n_traits if n_traits is None else n_traits
        out = torch.empty((*lead, T), dtype=torch.float32, device=self.device)
        _lib.check(_lib.load().bg_gebv(engine or self._engine, w.data_ptr(), rows, out.data_ptr(), self._stream()))
        return out

    def _top_k(self, values: torch.Tensor, k: int) -> torch.Tensor:
        """`jax.lax.top_k(values, k)[1]` along the last axis on the GPU (descending, ties -> lower index): the library's
        radix-select kernel (`bg_topk`); int64 indices `[..., k]`."""
        v = values.to(device=self.device, dtype=torch.float32)
        lead, length = tuple(v.shape[:-1]), v.shape[-1]
        if k > 1024 or k > length:
            if k > length:
                raise ValueError(f"k={k} must not exceed the number of candidates {length}")
            return torch.sort(v, dim=-1, descending=True, stable=True).indices[..., :k]
        flat = v.reshape(-1, length).contiguous()
        rows = flat.shape[0]
        vals = torch.empty((rows, k), dtype=torch.float32, device=self.device)
        idx = torch.empty((rows, k), dtype=torch.int32, device=self.device)
        _lib.check(_lib.load().bg_topk(self._engine, flat.data_ptr(), rows, length, k, vals.data_ptr(), idx.data_ptr(), self._stream()))
        return idx.long().reshape(*lead, k)

    # ---- chromax.Simulator surface -----------------------------------------------
    def set_seed(self, seed: int):
        self.random_key = _lib.key_data(seed)

    def load_population(self, file_name: Union[str, Path]) -> PackedPopulation:
        file_name = Path(file_name)
        if file_name.suffix == ".npy":
            pop = np.load(file_name)
        else:
            pop = np.loadtxt(file_name, dtype="bool")
            pop = pop.reshape(pop.shape[0], self.n_markers, 2)
        return self.as_packed(pop)

    def save_population(self, population, file_name: Union[str, Path]):
        np.save(file_name, np.asarray(population), allow_pickle=False)

    def cross(self, parents) -> PackedPopulation:
        """Offspring of `parents[n, 2, m, 2]` (or of a lazy `population[action]` view)."""
        k = self._next_key()
        if isinstance(parents, ParentsView):
            pop, pairs = parents.population, parents.pairs
        else:
            arr = parents if isinstance(parents, torch.Tensor) else np.asarray(parents)
            if arr.ndim != 4 or arr.shape[1] != 2:
                raise ValueError(f"parents must have shape (n, 2, m, 2), got {tuple(arr.shape)}")
            n = arr.shape[0]
            pop = self.as_packed(arr.reshape(2 * n, *arr.shape[2:]))
            pairs = np.arange(2 * n, dtype=np.int32).reshape(n, 2)
        return self._cross_indexed(pop, pairs, k)

    def _cross_indexed(self, pop: PackedPopulation, pairs, k: np.ndarray) -> PackedPopulation:
        src = pop.words.contiguous()
        it = self._index_tensor(pairs)
        if src.dim() == 3:  # one population
            E, n_src, n = 1, src.shape[0], it.shape[0]
            out = self._empty_words(n)
        else:  # [E, n_src, 2, Wpad] with pairs [E, n, 2]
            E, n_src, n = src.shape[0], src.shape[1], it.shape[1]
            out = self._empty_words(E, n)
        k = np.ascontiguousarray(k, dtype=np.uint32)
        _lib.check(_lib.load().bg_cross(self._engine, src.data_ptr(), it.data_ptr(), out.data_ptr(), E, n_src, n,
                                        _lib.nptr(k), self._layout(), self._schedule(), self._stream()))
        return PackedPopulation(self, out)

    def cross_and_score(self, pop: PackedPopulation, pairs: np.ndarray):
        """One population's step in ONE C call (`bg_vec_step` with one env: unique-key cross + GEBV): host pairs
        `int[n, 2]` in through a pinned staging buffer, offspring on the GPU and their GEBVs `float32[n, T]` (numpy,
        already on the host) out, one stream synchronisation.  Same kernels and the same key as `cross` followed by
        `GEBV` (breedgym/breedgym.py:142-143, 233): bit-identical results, a third fewer host round trips."""
        src = pop.words.contiguous()
        if src.dim() != 3:
            raise ValueError("cross_and_score expects one population (n, m, 2)")
        a = np.asarray(pairs)
        if a.ndim != 2 or a.shape[1] != 2:
            raise ValueError(f"pairs must have shape (n, 2), got {a.shape}")
        n, T = a.shape[0], self.GEBV_model.n_traits
        io = getattr(self, "_one_io", None)
        if io is None or io["n"] != n:
            act_pin = torch.empty((n, 2), dtype=torch.int32, pin_memory=True)
            gebv_pin = torch.empty((n, T), dtype=torch.float32, pin_memory=True)
            io = self._one_io = {"n": n, "keep": (act_pin, gebv_pin), "act_np": act_pin.numpy(), "gebv_np": gebv_pin.numpy(),
                                 "act_pin": act_pin.data_ptr(), "gebv_pin": gebv_pin.data_ptr(),
                                 "act_dev": torch.empty((n, 2), dtype=torch.int32, device=self.device),
                                 "gebv_dev": torch.empty((n, T), dtype=torch.float32, device=self.device)}
        io["act_np"][...] = a
        out = self._empty_words(n)
        # the key chain (`random_key, k = split(random_key)`) advances inside the call
        _lib.check(_lib.load().bg_vec_step(self._engine, src.data_ptr(), out.data_ptr(), io["act_pin"], io["act_dev"].data_ptr(),
                                           1, src.shape[0], n, self._key_ptr, self._layout_id, self._schedule_id,
                                           io["gebv_dev"].data_ptr(), None, io["gebv_pin"], None, self._stream()))
        gebv = io["gebv_np"].copy()
        if self.GEBV_model.offset:
            gebv = gebv + self.GEBV_model.offset
        return PackedPopulation(self, out), gebv

    def _gebv_frame(self, gebv: np.ndarray) -> pd.DataFrame:
        cols = getattr(self, "_trait_index", None)
        if cols is None or list(cols) != list(self.trait_names):  # building the column Index is most of DataFrame()'s cost
            cols = self._trait_index = pd.Index(self.trait_names)
        return pd.DataFrame(gebv, columns=cols, copy=False)

    def cross_envs(self, populations: PackedPopulation, actions) -> PackedPopulation:
        """`vmap(cross)(populations[arange, actions])` with ONE key for all envs
        (breedgym/vector/vec_env.py:75-77, 89-91)."""
        return self._cross_indexed(populations, actions, self._next_key())

    # ---- phenotype / GxE (chromax Simulator.phenotype, create_environments; scripts/time_wheat.py:17-50) ----------
    @property
    def GxE_model(self) -> TraitModel:
        """Genotype-by-environment effects: `normal(split_key, (m, T))` rescaled per trait to the variance
        `(1 - h2) / h2 * GEBV_model.var`, offset 1 (chromax Simulator.__init__, recalled in SURVEY App. B -- parity
        unpinned, like every chromax detail).  Built on first use, scored by the same kernels as the GEBV."""
        if self._gxe_model is None:
            from . import jaxlike

            m, T = self.n_markers, len(self.trait_names)
            env = jaxlike.normal(self._gxe_key, m * T, self.rng_layout).reshape(m, T)
            target = (1 - self.h2) / self.h2 * self.GEBV_model.var
            env = (env * np.sqrt(target / (np.sum(env**2, axis=0) / 2)).astype(np.float32)).astype(np.float32)
            self._gxe_model = TraitModel(self, env, offset=1, own_engine=True)
        return self._gxe_model

    def create_environments(self, num_environments: int) -> np.ndarray:
        """`random_key, k = split(random_key)`; `normal(k, (num_environments,))`: one scalar per environment."""
        from . import jaxlike

        k = self._next_key()
        return jaxlike.normal(k, int(num_environments), self.rng_layout)

    def phenotype(self, population, *, num_environments: Optional[int] = None, environments=None) -> torch.Tensor:
        """Mean over the environments of `GEBV(pop) + env * GxE(pop)` -> `float32[..., n_traits]` on the device
        (chromax `_phenotype`).  One environment is drawn when neither argument is given."""
        if num_environments is not None and environments is not None:
            raise ValueError("You cannot specify both the number of environments and the environments.")
        if environments is None:
            environments = self.create_environments(1 if num_environments is None else num_environments)
        pop = self.as_packed(population)
        g = self.GEBV_model(pop)
        e = self.GxE_model(pop)
        envs = torch.as_tensor(np.asarray(environments, dtype=np.float32).reshape(-1), device=self.device)
        view = (-1,) + (1,) * g.dim()
        return torch.mean(g.unsqueeze(0) + envs.view(view) * e.unsqueeze(0), dim=0)

    def double_haploid(self, population, n_offspring: int = 1) -> PackedPopulation:
        """`(n, m, 2)` -> `(n, n_offspring, m, 2)` (squeezed for one offspring); a batch `(E, n, m, 2)` is the
        reference's `vmap(double_haploid, in_axes=(None, 0))`: ONE key for all envs, one launch."""
        pop = self.as_packed(population)
        if pop.words.dim() not in (3, 4):
            raise ValueError("double_haploid expects one population (n, m, 2) or a batch (E, n, m, 2)")
        k = np.ascontiguousarray(self._next_key(), dtype=np.uint32)
        lead = tuple(pop.words.shape[:-2])
        E, n = (1, lead[0]) if len(lead) == 1 else lead
        out = self._empty_words(*lead, n_offspring)
        _lib.check(_lib.load().bg_double_haploid(self._engine, pop.words.contiguous().data_ptr(), out.data_ptr(), E, n,
                                                 n_offspring, _lib.nptr(k), self._layout(), self._schedule(),
                                                 self._stream()))
        if n_offspring == 1:
            out = out.squeeze(-3)
        return PackedPopulation(self, out)

    def GEBV(self, population) -> pd.DataFrame:
        return self._gebv_frame(self.GEBV_model(population).cpu().numpy())

    @property
    def max_gebv(self):
        return self.GEBV_model.max

    @property
    def min_gebv(self):
        return self.GEBV_model.min

    @property
    def mean_gebv(self):
        return self.GEBV_model.mean

    def corrcoef(self, population) -> np.ndarray:
        """Correlation of every individual's 2m alleles with the population mean."""
        flat = self.as_packed(population).to_bool().reshape(len(population), -1).to(torch.float32)
        stacked = torch.cat([flat.mean(dim=0, keepdim=True), flat], dim=0)
        return torch.corrcoef(stacked)[0, 1:].cpu().numpy()

    def select(self, population, k: int, f_index: Optional[Callable] = None) -> Tuple[PackedPopulation, np.ndarray]:
        """Top-k individuals by `f_index` (default: GEBV summed over traits); ties -> lower index."""
        pop = self.as_packed(population)
        if pop.words.dim() != 3:
            raise ValueError("select expects one population (n, m, 2)")
        if k > len(pop):
            raise ValueError(f"k={k} must not exceed the population size {len(pop)}")
        if f_index is None:
            values = self.GEBV_model(pop).sum(dim=-1)
        else:
            values = f_index(pop)
        if not isinstance(values, torch.Tensor):
            values = torch.from_numpy(np.ascontiguousarray(np.asarray(values)))
        if values.dim() != 1:
            raise ValueError("f_index must return one value per individual")
        best = self._top_k(values, k).cpu().numpy()
        return self._gather(pop, best), best

    @staticmethod
    def _diallel_indices(indices: Sequence[int]) -> np.ndarray:
        indices = np.asarray(indices)
        a, b = np.triu_indices(len(indices), k=1)
        return np.stack([indices[a], indices[b]], axis=1)

    def diallel(self, population, n_offspring: int = 1) -> PackedPopulation:
        pop = self.as_packed(population)
        pairs = np.repeat(self._diallel_indices(np.arange(len(pop))), n_offspring, axis=0)
        return self.cross(pop[pairs])

    def random_crosses(self, population, n_crosses: int, n_offspring: int = 1):
        pop = self.as_packed(population)
        rng = np.random.default_rng(int(self._next_key()[1]))
        pairs = rng.integers(0, len(pop), size=(n_crosses, 2))
        return self.cross(pop[np.repeat(pairs, n_offspring, axis=0)]), pairs
