"""Diagnostics: is the device-resident step loop host-bound?  Times, for K steps of the C2 workload in device mode,
the host's enqueue time (perf_counter around the loop, no sync) next to the device time (CUDA events)."""
import sys, time, json
from pathlib import Path
import numpy as np
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench
from breedgym_b200.vector import VecBreedGym

def main():
    K = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
    germ, gmap = bench.workload_inputs()
    dev = torch.device("cuda", 0)
    envs = []
    for _ in range(4):
        env = VecBreedGym(num_envs=64, initial_population=germ, genetic_map=gmap, trait_names=["Yield"], individual_per_gen=370,
                          num_generations=10, device=0, info_device="device")
        env.reset(seed=7)
        envs.append(env)
    rng = np.random.default_rng(1)
    acts = [torch.from_numpy(rng.integers(0, 370, (64, 370, 2), dtype=np.int32)).to(dev) for _ in range(16)]
    out = []
    for rep in range(reps):
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        t0 = time.perf_counter()
        for i in range(K):
            envs[i % 4].step(acts[i % 16])
        t1 = time.perf_counter()
        e.record()
        torch.cuda.synchronize()
        out.append({"rep": rep, "host_us_per_step": 1e6 * (t1 - t0) / K, "dev_us_per_step": 1e3 * s.elapsed_time(e) / K})
    print(json.dumps(out))

main()
