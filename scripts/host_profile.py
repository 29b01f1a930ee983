#!/usr/bin/env python
"""cProfile of the host side of VecBreedGym.step (device-resident actions, no D2H)."""
import cProfile
import pstats
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from breedgym_b200.vector import VecBreedGym  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "device"
germ = np.random.default_rng(0).random((370, 10000, 2)) < 0.5
env = VecBreedGym(num_envs=64, initial_population=germ, genetic_map=ROOT / "breedgym_b200/data/small_genetic_map.txt",
                  trait_names=["Yield"], individual_per_gen=370, device=0, info_device=mode)
env.reset(seed=7)
acts_np = np.random.default_rng(1).integers(0, 370, (64, 370, 2), dtype=np.int32)
acts = torch.from_numpy(acts_np).cuda() if mode == "device" else acts_np
for _ in range(3000):  # ~0.3 s: lets the SM clock ramp up
    env.step(acts)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(500):
    env.step(acts)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"mode={mode}: enqueue {1e6 * (t1 - t0) / 500:.1f} us/step, with final sync {1e6 * (t2 - t0) / 500:.1f} us/step")
pr = cProfile.Profile()
pr.enable()
for _ in range(500):
    env.step(acts)
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
