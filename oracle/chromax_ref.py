"""NumPy restatement of the chromax semantics BreedGym calls.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

`chromax` (PyPI, un-pinned in /root/reference/pyproject.toml:24) is absent
from /root/reference and not installable here; its published algorithm is
restated and anchored on the reference's own call sites:

  Simulator(...)            breedgym/breedgym.py:36, breedgym/vector/vec_env.py:45
  .load_population          breedgym/breedgym.py:38, vec_env.py:50
  .set_seed                 breedgym/breedgym.py:117, vec_env.py:114
  .cross(parents)           breedgym/breedgym.py:143, vec_env.py:75-77
  .GEBV / .GEBV_model       breedgym/breedgym.py:233, vec_env.py:133
  .corrcoef                 breedgym/breedgym.py:240
  .select / ._diallel_indices   breedgym/wrappers.py:75-80, vec_wrappers.py:66
  .double_haploid           breedgym/vector/breeding_programs_env.py:41

Everything operates on UNPACKED bool arrays `(n, m, 2)` exactly like the
reference; no bit tricks are shared with the CUDA path.

Key-schedule variants (the part of App. B flagged uncertain):
  "S1": the per-gamete key is used directly for the recombination draw
  "S2": the per-gamete key is first split into (recombination, mutation) keys
"""
from __future__ import annotations

from pathlib import Path
from typing import Optional, Sequence

import numpy as np

from . import jax_prng as jp

SCHEDULES = ("S1", "S2")


# --------------------------------------------------------------------------
# genetic map -> recombination vector, marker effects
# --------------------------------------------------------------------------
def read_genetic_map(path):
    import pandas as pd

    return pd.read_table(path, sep="\t")


def recombination_vector(genetic_map, chr_column="CHR.PHYS", position_column="cM",
                         recombination_column="RecombRate") -> np.ndarray:
    """float32[m] per-marker recombination probability (chromax Simulator ctor).

    RecombRate column: shifted by one ("recombine now" semantics); else Haldane
    from cM distances; then every first marker of a chromosome gets 0.5.
    """
    chrom = genetic_map[chr_column].to_numpy()
    m = len(chrom)
    if recombination_column in genetic_map.columns:
        r = np.array(genetic_map[recombination_column].to_numpy(), dtype=np.float64, copy=True)
        r[1:] = r[:-1].copy()
    elif position_column in genetic_map.columns:
        cm = np.asarray(genetic_map[position_column].to_numpy(), dtype=np.float64)
        r = np.zeros(m, dtype=np.float64)
        d = (cm[1:] - cm[:-1]) / 100.0
        r[1:] = 0.5 * (1.0 - np.exp(-2.0 * d))
    else:
        raise ValueError("genetic map needs a recombination or a position column")
    first = np.ones(m, dtype=bool)
    first[1:] = chrom[1:] != chrom[:-1]
    r[first] = 0.5
    return r.astype(np.float32)


def trait_columns(genetic_map, trait_names=None, chr_column="CHR.PHYS",
                  position_column="cM", recombination_column="RecombRate"):
    if trait_names is not None:
        return list(trait_names)
    skip = {"MRK.NAME", chr_column, position_column, recombination_column}
    return [c for c in genetic_map.columns if c not in skip]


def marker_effects(genetic_map, trait_names) -> np.ndarray:
    return genetic_map[list(trait_names)].to_numpy(dtype=np.float32)


def chr_lens(genetic_map, chr_column="CHR.PHYS") -> np.ndarray:
    chrom = genetic_map[chr_column].to_numpy()
    starts = np.flatnonzero(np.concatenate([[True], chrom[1:] != chrom[:-1]]))
    return np.diff(np.concatenate([starts, [len(chrom)]]))


# --------------------------------------------------------------------------
# meiosis / cross / double haploid
# --------------------------------------------------------------------------
def meiosis(individual: np.ndarray, r: np.ndarray, k, mutation: float = 0.0,
            schedule: str = "S2", layout: str = "legacy") -> np.ndarray:
    """One gamete of `individual[m,2]` (chromax.functional._meiosis).

    u = uniform(key_rec,(m,)); s = u < r; mask = inclusive cumulative XOR over
    the WHOLE genome; hap[j] = individual[j, mask[j]]; optional mutation XOR.
    """
    m = individual.shape[0]
    if schedule == "S2":
        ks = jp.split(k, 2, layout)
        k_rec, k_mut = ks[0], ks[1]
    elif schedule == "S1":
        k_rec, k_mut = k, None
    else:
        raise ValueError(schedule)
    u = jp.uniform(k_rec, m, layout)
    sites = u < np.asarray(r, dtype=np.float32)
    mask = np.bitwise_xor.accumulate(sites.astype(np.uint8)).astype(np.intp)
    hap = np.take_along_axis(np.asarray(individual, dtype=bool), mask[:, None], axis=1)[:, 0]
    if mutation > 0.0:
        if k_mut is None:  # S1 predates the mutation feature
            raise ValueError("schedule S1 has no mutation key")
        um = jp.uniform(k_mut, m, layout)
        hap = hap ^ (um < np.float32(mutation))
    return hap


def cross(parents: np.ndarray, r: np.ndarray, k, mutation: float = 0.0,
          schedule: str = "S2", layout: str = "legacy") -> np.ndarray:
    """chromax.functional.cross: parents bool[n,2,m,2] -> offspring bool[n,m,2].

    keys = split(k, 2n) reshaped (n,2): key index 2i+p drives the gamete of
    parents[i,p], which becomes out[i,:,p].
    """
    parents = np.asarray(parents, dtype=bool)
    n, _, m, _ = parents.shape
    keys = jp.split(k, 2 * n, layout).reshape(n, 2, 2)
    out = np.empty((n, m, 2), dtype=bool)
    for i in range(n):
        for p in range(2):
            out[i, :, p] = meiosis(parents[i, p], r, keys[i, p], mutation, schedule, layout)
    return out


def cross_rows(parents: np.ndarray, r: np.ndarray, k, rows: Sequence[int], n_total: int,
               mutation: float = 0.0, schedule: str = "S2", layout: str = "legacy") -> np.ndarray:
    """Gametes for a SAMPLE of (offspring, parent) rows q = 2i+p of an n_total cross.

    `parents[len(rows), m, 2]` holds the parent individual of each sampled row.
    Every row is an independent function of (k, q): lets full-size configs be
    checked on a few rows without materialising the whole tensor.
    """
    keys = jp.split(k, 2 * n_total, layout)
    out = np.empty((len(rows), parents.shape[1]), dtype=bool)
    for t, q in enumerate(rows):
        out[t] = meiosis(parents[t], r, keys[q], mutation, schedule, layout)
    return out


def double_haploid(pop: np.ndarray, r: np.ndarray, k, n_offspring: int = 1,
                   mutation: float = 0.0, schedule: str = "S2", layout: str = "legacy") -> np.ndarray:
    """chromax.functional.double_haploid: bool[n,m,2] -> bool[n,n_offspring,m,2]."""
    pop = np.asarray(pop, dtype=bool)
    n, m, _ = pop.shape
    keys = jp.split(k, n * n_offspring, layout).reshape(n, n_offspring, 2)
    out = np.empty((n, n_offspring, m, 2), dtype=bool)
    for i in range(n):
        for o in range(n_offspring):
            hap = meiosis(pop[i], r, keys[i, o], mutation, schedule, layout)
            out[i, o, :, 0] = hap
            out[i, o, :, 1] = hap
    return out


# --------------------------------------------------------------------------
# trait model
# --------------------------------------------------------------------------
def gebv(pop: np.ndarray, effects: np.ndarray, offset=0.0, dtype=np.float64) -> np.ndarray:
    """TraitModel.__call__: dot(sum(pop,-1), effects[m,T]) + offset -> [...,T].

    Computed in float64 from the float32 effects: the "true" value both the
    reference's float32 dot and the CUDA kernel are compared against.
    """
    dosage = np.asarray(pop, dtype=bool).sum(axis=-1, dtype=np.int8)
    eff = np.asarray(effects, dtype=np.float32).astype(dtype)
    return dosage.astype(dtype) @ eff + offset


def gebv_f32(pop: np.ndarray, effects: np.ndarray) -> np.ndarray:
    """The same dot in float32 (what the reference's XLA dot produces, up to order)."""
    return gebv(pop, effects, dtype=np.float32).astype(np.float32)


def gxe_effects(effects: np.ndarray, split_key, h2=None, layout="legacy") -> np.ndarray:
    """chromax Simulator.__init__ (recalled): GxE marker effects = normal(split_key, (m, T)) rescaled so
    that var(GxE) = (1 - h2) / h2 * var(GEBV) per trait (TraitModel.var = sum(effects^2) / 2); h2 defaults to 0.5."""
    m, T = effects.shape
    h2 = np.full(T, 0.5, dtype=np.float32) if h2 is None else np.asarray(h2, dtype=np.float32)
    env = jp.normal(split_key, m * T, layout).reshape(m, T)
    var_g = (np.sum(effects.astype(np.float32) ** 2, axis=0) / 2).astype(np.float32)
    var_e = (np.sum(env ** 2, axis=0) / 2).astype(np.float32)
    return (env * np.sqrt(((1 - h2) / h2 * var_g) / var_e).astype(np.float32)).astype(np.float32)


def phenotype(pop: np.ndarray, effects: np.ndarray, gxe: np.ndarray, environments: np.ndarray, dtype=np.float64) -> np.ndarray:
    """chromax `_phenotype`: mean over environments of GEBV(pop) + env * GxE(pop), GxE = TraitModel(gxe, offset=1)."""
    g = gebv(pop, effects, dtype=dtype)
    e = gebv(pop, gxe, offset=1.0, dtype=dtype)
    envs = np.asarray(environments, dtype=dtype)
    return np.mean(g[None] + envs.reshape(-1, *([1] * g.ndim)) * e[None], axis=0)


def corrcoef(pop: np.ndarray) -> np.ndarray:
    """Simulator.corrcoef: correlation of each flattened individual with the mean."""
    pop = np.asarray(pop, dtype=bool)
    flat = pop.reshape(pop.shape[0], -1).astype(np.float64)
    mean = flat.mean(axis=0, keepdims=True)
    cc = np.corrcoef(np.concatenate([mean, flat], axis=0))
    return cc[0, 1:]


def simplified_correlation(pop: np.ndarray) -> np.ndarray:
    """SimplifiedBreedGym._correlation (breedgym/wrappers.py:92-98)."""
    mono = np.asarray(pop, dtype=bool).sum(axis=-1).astype(np.float64) - 1.0
    mean_ind = mono.mean(axis=0)
    norms = np.linalg.norm(mono, axis=-1) * np.linalg.norm(mean_ind)
    return mono @ mean_ind / norms


def diallel_indices(idx: np.ndarray) -> np.ndarray:
    """Simulator._diallel_indices: all unordered pairs, upper-triangular order."""
    idx = np.asarray(idx)
    a, b = np.triu_indices(len(idx), k=1)
    return np.stack([idx[a], idx[b]], axis=1)


def select(pop: np.ndarray, k: int, index_values: np.ndarray):
    """Simulator.select given the index values: top-k, ties -> lower index."""
    _, best = jp.top_k(index_values, k)
    return pop[best], best


# --------------------------------------------------------------------------
# stateful facade reproducing the key chain
# --------------------------------------------------------------------------
class OracleSimulator:
    """Key-chain-faithful stand-in for chromax.Simulator (cross/GEBV only)."""

    def __init__(self, r: np.ndarray, effects: np.ndarray, seed: int = 0, mutation: float = 0.0,
                 schedule: str = "S2", layout: str = "legacy"):
        self.r = np.asarray(r, dtype=np.float32)
        self.effects = np.asarray(effects, dtype=np.float32)
        if self.effects.ndim == 1:
            self.effects = self.effects[:, None]
        self.mutation = float(mutation)
        self.schedule = schedule
        self.layout = layout
        self.set_seed(seed)

    def set_seed(self, seed: int):
        self.random_key = jp.key(seed)

    def next_cross_key(self):
        ks = jp.split(self.random_key, 2, self.layout)
        self.random_key = ks[0]
        return ks[1]

    def cross(self, parents: np.ndarray) -> np.ndarray:
        k = self.next_cross_key()
        return cross(parents, self.r, k, self.mutation, self.schedule, self.layout)

    def double_haploid(self, pop: np.ndarray, n_offspring: int = 1) -> np.ndarray:
        k = self.next_cross_key()
        out = double_haploid(pop, self.r, k, n_offspring, self.mutation, self.schedule, self.layout)
        return out[:, 0] if n_offspring == 1 else out

    def GEBV_model(self, pop: np.ndarray) -> np.ndarray:
        return gebv(pop, self.effects)


# --------------------------------------------------------------------------
# environment-level restatements (breedgym/vector/vec_env.py)
# --------------------------------------------------------------------------
def normalize_index(idx: np.ndarray, n: int) -> np.ndarray:
    """jnp `x[idx]` semantics: negatives wrap once, then clamp into range."""
    idx = np.asarray(idx).astype(np.int64)
    idx = np.where(idx < 0, idx + n, idx)
    return np.clip(idx, 0, n - 1)


def vec_reset(germplasm: np.ndarray, n: int, num_envs: int, random_key, layout="legacy"):
    """VecBreedGym.reset (vec_env.py:109-130): returns (new_key, populations, indices)."""
    keys = jp.split(random_key, num_envs + 1, layout)
    idx = np.stack([jp.choice_no_replace(keys[1 + e], len(germplasm), n, layout) for e in range(num_envs)])
    return keys[0], np.asarray(germplasm)[idx], idx


def vec_step(sim: OracleSimulator, populations: np.ndarray, actions: np.ndarray):
    """VecBreedGym.step core (vec_env.py:88-94): ONE key for all envs (a4 in SURVEY §8)."""
    E, n = populations.shape[:2]
    k = sim.next_cross_key()
    act = normalize_index(actions, n)
    out = np.empty((E, act.shape[1]) + populations.shape[2:], dtype=bool)
    for e in range(E):
        parents = populations[e][act[e]]
        out[e] = cross(parents, sim.r, k, sim.mutation, sim.schedule, sim.layout)
    return out
