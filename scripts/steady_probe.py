#!/usr/bin/env python
"""Steady-state step loop of ONE env set (64 envs x 370 x 10 000, device mode): microseconds per step over a long
run, with the library's own timing of the mask batches (option timing=1 prints each batch's device time).

    BG_OPT_TIMING=1 python scripts/steady_probe.py [steps] [num_generations]
"""
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from breedgym_b200.vector import VecBreedGym  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 800
gens = int(sys.argv[2]) if len(sys.argv) > 2 else 10
replicas = int(sys.argv[3]) if len(sys.argv) > 3 else 1
torch.cuda.set_stream(torch.cuda.Stream(priority=-1))
germ = np.random.default_rng(0).random((370, 10000, 2)) < 0.5
rng = np.random.default_rng(1)
acts = [torch.from_numpy(rng.integers(0, 370, (64, 370, 2), dtype=np.int32)).cuda() for _ in range(8)]
envs = []
for r in range(replicas):
    env = VecBreedGym(num_envs=64, initial_population=germ, genetic_map=ROOT / "breedgym_b200/data/small_genetic_map.txt",
                      trait_names=["Yield"], individual_per_gen=370, num_generations=gens, device=0, info_device="device")
    env.reset(seed=7)
    envs.append(env)
for i in range(200):
    envs[i % replicas].step(acts[i % 8])
torch.cuda.synchronize()
for blk in range(3):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    a.record()
    for i in range(steps):
        envs[i % replicas].step(acts[i % 8])
    b.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    print(f"block {blk}: {1e3 * a.elapsed_time(b) / steps:.2f} us/step on the device, host {1e6 * (t1 - t0) / steps:.1f} us/step "
          f"({replicas} env set(s), {gens} generations per episode)", flush=True)
