#!/usr/bin/env python
"""Small run of every kernel for compute-sanitizer (memcheck): tiny shapes, results checked against the oracle."""
import sys
from pathlib import Path

import numpy as np
import pandas as pd
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from breedgym_b200.simulator import Simulator  # noqa: E402
from breedgym_b200.vector import VecBreedGym  # noqa: E402
from oracle import chromax_ref as cr  # noqa: E402
from oracle import jax_prng as jp  # noqa: E402

rng = np.random.default_rng(0)
for m, T in ((77, 1), (1000, 3), (4100, 16)):
    df = pd.DataFrame({"CHR.PHYS": np.arange(m) // (m // 3 + 1), "RecombRate": rng.random(m) * 0.05})
    for t in range(T):
        df[f"t{t}"] = rng.standard_normal(m).astype(np.float32)
    sim = Simulator(genetic_map=df, device=0, seed=1)
    pop = rng.random((9, m, 2)) < 0.5
    pairs = rng.integers(0, 9, (7, 2))
    key = jp.key(5)
    packed = sim.as_packed(pop)
    off = sim._cross_indexed(packed, pairs, key)
    assert np.array_equal(np.asarray(off), cr.cross(pop[pairs], sim.recombination_vec, key))
    pops = rng.random((3, 9, m, 2)) < 0.5
    acts = rng.integers(0, 9, (3, 7, 2))
    voff = sim._cross_indexed(sim.as_packed(pops), acts, key)
    osim = cr.OracleSimulator(sim.recombination_vec, sim.GEBV_model.marker_effects, seed=0)
    g = sim.GEBV_model(voff).cpu().numpy()
    assert np.allclose(g, cr.gebv(np.asarray(voff), osim.effects), rtol=1e-5, atol=0)
    dh = sim.double_haploid(pop, 2)
    assert dh.shape == (9, 2, m, 2)
    sel, idx = sim.select(packed, 3)
env = VecBreedGym(num_envs=3, initial_population=rng.random((20, 1000, 2)) < 0.5,
                  genetic_map=ROOT / "breedgym_b200/data/sample_with_r_genetic_map.txt", individual_per_gen=12,
                  num_generations=3, device=0)
env.reset(seed=3)
for _ in range(7):
    env.step(rng.integers(0, 12, (3, 12, 2)))
torch.cuda.synchronize()
print("sanitize smoke ok")
