#!/usr/bin/env python
"""Look for periodic stalls: per-block wall time of VecBreedGym.step over many blocks (host and device mode)."""
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from breedgym_b200.vector import VecBreedGym  # noqa: E402

germ = np.random.default_rng(0).random((370, 10000, 2)) < 0.5
rng = np.random.default_rng(1)
acts_np = [rng.integers(0, 370, (64, 370, 2), dtype=np.int32) for _ in range(8)]
nrep = int(sys.argv[2]) if len(sys.argv) > 2 else 4
mode = sys.argv[1] if len(sys.argv) > 1 else "host"
envs = [VecBreedGym(num_envs=64, initial_population=germ, genetic_map=ROOT / "breedgym_b200/data/small_genetic_map.txt",
                    trait_names=["Yield"], individual_per_gen=370, device=0, info_device=mode) for _ in range(nrep)]
for e in envs:
    e.reset(seed=7)
acts = [torch.from_numpy(a).cuda() for a in acts_np] if mode == "device" else acts_np
t_start = time.perf_counter()
out = []
for blk in range(60):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(200):
        envs[i % nrep].step(acts[i % 8])
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    out.append((round(t0 - t_start, 2), round(1e6 * (t1 - t0) / 200, 1)))
print(mode, nrep, out)
