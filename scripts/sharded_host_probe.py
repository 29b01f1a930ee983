#!/usr/bin/env python
"""Where does the sharded step's host time go?  Run under torchrun (one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
        scripts/sharded_host_probe.py

Same workload as bench.py's value leg (4 replicas of 64 envs per GPU, device-resident actions, high-priority step
stream); per variant: host microseconds per step split into reward steps / other steps, the slowest calls, and
the device time per step between two CUDA events.
"""
import gc
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from breedgym_b200.vector import ShardedVecBreedGym  # noqa: E402
from breedgym_b200.vector import sharded as sharded_mod  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    torch.cuda.set_stream(torch.cuda.Stream(device=dev, priority=-1))
    germ = np.random.default_rng(0).random((370, 10000, 2)) < 0.5
    kw = dict(initial_population=germ, genetic_map=ROOT / "breedgym_b200/data/small_genetic_map.txt", trait_names=["Yield"],
              individual_per_gen=370, num_generations=10, device=local, info_device="device")
    rng = np.random.default_rng(1 + rank)
    acts = [torch.from_numpy(rng.integers(0, 370, (64, 370, 2), dtype=np.int32)).to(dev) for _ in range(16)]
    variants = sys.argv[1:] or ["async", "sync", "torch", "nogather"]
    for variant in variants:
        env_kw = dict(kw)
        if variant == "torch":
            mk = lambda: ShardedVecBreedGym(total_envs=64 * world, collective="torch", **env_kw)
        elif variant == "sync":
            mk = lambda: ShardedVecBreedGym(total_envs=64 * world, async_rewards=False, **env_kw)
        else:
            mk = lambda: ShardedVecBreedGym(total_envs=64 * world, async_rewards=True, **env_kw)
        envs = [mk() for _ in range(4)]
        for e in envs:
            e.reset(seed=7)
            if variant == "nogather":
                e.gather_rewards = lambda local_rewards: local_rewards
        for i in range(88):
            envs[i % 4].step(acts[i % 16])
        torch.cuda.synchronize()
        dist.barrier()
        gc.collect()
        gc.freeze()
        for K in (20, 400):
            torch.cuda.synchronize()
            dist.barrier()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            stream = torch.cuda.current_stream(dev)
            times, kinds = [], []
            a.record(stream)
            t00 = time.perf_counter()
            for i in range(K):
                env = envs[i % 4]
                kinds.append(env.env.step_idx + 1 == env.env.num_generations)
                t0 = time.perf_counter()
                env.step(acts[i % 16])
                times.append(time.perf_counter() - t0)
            host = time.perf_counter() - t00
            b.record(stream)
            torch.cuda.synchronize()
            dist.barrier()
            t = np.array(times) * 1e6
            k = np.array(kinds)
            print(f"[rank {rank}] {variant:9s} K={K:4d}: device {1e3 * a.elapsed_time(b) / K:6.1f} us/step, host {1e6 * host / K:6.1f} us/step; "
                  f"reward steps {t[k].mean() if k.any() else 0:7.1f} us (n={int(k.sum())}, max {t[k].max() if k.any() else 0:7.1f}), "
                  f"other steps {t[~k].mean():6.1f} us (max {t[~k].max():7.1f}, p90 {np.percentile(t[~k], 90):6.1f})", flush=True)
        gc.unfreeze()
        del envs
        torch.cuda.synchronize()
        dist.barrier()
    # split of one reward step's host time (async variant): profile the pieces of RewardGather.__call__
    env = ShardedVecBreedGym(total_envs=64 * world, async_rewards=True, **kw)
    env.reset(seed=7)
    g = env._gather
    local_rewards = torch.zeros(64, dtype=torch.float32, device=dev)
    for rep in range(3):
        torch.cuda.synchronize()
        dist.barrier()
        marks = [time.perf_counter()]
        main_stream = torch.cuda.current_stream(dev)
        g.ready.record(main_stream)
        marks.append(time.perf_counter())
        g.stream.wait_event(g.ready)
        marks.append(time.perf_counter())
        with torch.cuda.stream(g.stream):
            g._stage[0].copy_(local_rewards, non_blocking=True)
        marks.append(time.perf_counter())
        rc = g._fn(g._comm, g._stage[0].data_ptr(), g._out[0].data_ptr(), g.count, g._stream_ptr)
        marks.append(time.perf_counter())
        g.done.record(g.stream)
        marks.append(time.perf_counter())
        d = np.diff(np.array(marks)) * 1e6
        print(f"[rank {rank}] gather pieces rep {rep}: record {d[0]:.1f}, wait_event {d[1]:.1f}, staged copy {d[2]:.1f}, "
              f"bg_allgather_f32 {d[3]:.1f}, record {d[4]:.1f} us (rc {rc})", flush=True)
    torch.cuda.synchronize()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
