#!/usr/bin/env python
"""BASELINE config C1: single-env BreedGym (370 x 10 000, 10 generations, random crosses): steps/s through the Gym API."""
import cProfile
import gc
import pstats
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from breedgym_b200.breedgym import BreedGym  # noqa: E402

germ = np.random.default_rng(0).random((370, 10000, 2)) < 0.5
env = BreedGym(initial_population=germ, genetic_map=ROOT / "breedgym_b200/data/small_genetic_map.txt", trait_names=["Yield"], device=0)
rng = np.random.default_rng(1)
acts = [rng.integers(0, 370, (370, 2)) for _ in range(8)]


def episode():
    env.reset(seed=7)
    for g in range(10):
        obs, rew, ter, tru, info = env.step(acts[g % 8])
    return rew


for _ in range(20):
    episode()
gc.collect()
gc.freeze()
torch.cuda.synchronize()
t0 = time.perf_counter()
n = 100
for _ in range(n):
    episode()
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print(f"C1 single env: {1e6 * dt / (10 * n):.1f} us per step ({10 * n / dt:.0f} env-steps/s, {370 * 10000 * 10 * n / dt / 1e9:.2f} G offspring-markers/s)")
pr = cProfile.Profile()
pr.enable()
for _ in range(30):
    episode()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
