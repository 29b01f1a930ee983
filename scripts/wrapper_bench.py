#!/usr/bin/env python
"""Steps/s of the action wrappers at the C2 shape (64 envs x 370 x 10k): SelectionScores, PairScores."""
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from breedgym_b200.vector import PairScores, SelectionScores, VecBreedGym  # noqa: E402

germ = np.random.default_rng(0).random((370, 10000, 2)) < 0.5
kw = dict(num_envs=64, initial_population=germ, genetic_map=ROOT / "breedgym_b200/data/small_genetic_map.txt",
          trait_names=["Yield"], individual_per_gen=370, device=0)
for name, make in (("SelectionScores", lambda: SelectionScores(VecBreedGym(**kw), k=20)),
                   ("PairScores", lambda: PairScores(VecBreedGym(**kw)))):
    env = make()
    _, infos = env.reset(seed=7)
    for phase in ("warm", "timed"):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        steps = 30 if phase == "warm" else 100
        for _ in range(steps):
            g = infos["GEBV"].squeeze(-1)
            act = g if name == "SelectionScores" else torch.as_tensor(g, device="cuda")[:, :, None] + torch.as_tensor(g, device="cuda")[:, None, :]
            _, _, _, _, infos = env.step(act)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    print(f"{name}: {1e6 * dt / steps:.0f} us/step = {64 * steps / dt:.0f} env-steps/s")
