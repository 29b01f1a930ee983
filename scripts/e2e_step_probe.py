"""Diagnostics: per-step wall time of the host-facing vector-env step (numpy actions in, numpy GEBV out) in blocks of 20 steps,
as bench.py's e2e leg runs it at the driver's K = 20."""
import gc, sys, time, json
from pathlib import Path
import numpy as np
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench
from breedgym_b200.vector import VecBreedGym

germ, gmap = bench.workload_inputs()
torch.cuda.set_stream(torch.cuda.Stream(priority=-1))
envs = []
for _ in range(4):
    env = VecBreedGym(num_envs=64, initial_population=germ, genetic_map=gmap, trait_names=["Yield"], individual_per_gen=370,
                      num_generations=10, device=0)
    env.reset(seed=7)
    envs.append(env)
rng = np.random.default_rng(1)
acts = [rng.integers(0, 370, (64, 370, 2), dtype=np.int32) for _ in range(16)]
for i in range(44):
    envs[i % 4].step(acts[i % 16])
gc.collect(); gc.freeze()
out = []
for blk in range(6):
    ts = []
    torch.cuda.synchronize()
    for i in range(20):
        t0 = time.perf_counter()
        envs[i % 4].step(acts[i % 16])
        ts.append(1e6 * (time.perf_counter() - t0))
    out.append({"block": blk, "mean_us": round(float(np.mean(ts)), 1), "median_us": round(float(np.median(ts)), 1), "max_us": round(float(np.max(ts)), 1),
                "steps_us": [round(t) for t in ts]})
    time.sleep(0.05 * blk)
print(json.dumps(out))
