#!/usr/bin/env python
"""Where does a VecBreedGym.step go?  CPU time vs GPU time per block of steps, device and host mode."""
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
for mode in ("device", "host"):
    env = VecBreedGym(num_envs=64, initial_population=germ, genetic_map=ROOT / "breedgym_b200/data/small_genetic_map.txt",
                      trait_names=["Yield"], individual_per_gen=370, device=0, info_device=mode)
    env.reset(seed=7)
    acts = [torch.from_numpy(a).cuda() for a in acts_np] if mode == "device" else acts_np
    for i in range(3000):
        env.step(acts[i % 8])
    torch.cuda.synchronize()
    for blk in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        a.record()
        for i in range(200):
            env.step(acts[i % 8])
        b.record()
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        print(f"{mode} block {blk}: cpu-enqueue {1e6 * (t1 - t0) / 200:.1f} us/step, gpu-events {1e3 * a.elapsed_time(b) / 200:.1f} us/step, "
              f"wall {1e6 * (t2 - t0) / 200:.1f} us/step", flush=True)
