#!/usr/bin/env python
"""Multi-GPU parity check (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 \
        scripts/multi_gpu_check.py

Every rank steps its shard of a ShardedVecBreedGym (NCCL reward all-gather) and, on the same GPU, the full
unsharded VecBreedGym; the shard must equal its slice of the full env bit for bit and the gathered rewards
must equal the full env's rewards.
"""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from breedgym_b200.vector import ShardedVecBreedGym, VecBreedGym  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    data = ROOT / "breedgym_b200" / "data"
    total, n = 2 * world + 1, 40  # ragged shards on purpose
    kw = dict(initial_population=data / "sample_geno.npy", genetic_map=data / "sample_with_r_genetic_map.txt",
              individual_per_gen=n, num_generations=4)
    shard = ShardedVecBreedGym(total_envs=total, device=local, **kw)
    full = VecBreedGym(num_envs=total, device=local, **kw)
    ps, _ = shard.reset(seed=3)
    pf, _ = full.reset(seed=3)
    sl = shard.local_slice()
    assert np.array_equal(np.asarray(ps), np.asarray(pf)[sl]), "reset differs"
    rng = np.random.default_rng(0)  # same stream on every rank
    for step in range(9):  # crosses two autoresets
        act = rng.integers(0, n, (total, n, 2))
        ps, rs, _, ts, infos_s = shard.step(act[sl])
        pf, rf, _, tf, infos_f = full.step(act)
        assert np.array_equal(np.asarray(ps), np.asarray(pf)[sl]), f"step {step}: populations differ"
        assert np.array_equal(infos_s["GEBV"], infos_f["GEBV"][sl]), f"step {step}: GEBV differs"
        assert rs.shape == (total,) and np.array_equal(rs, rf), f"step {step}: gathered rewards differ"
        assert np.array_equal(ts, tf)
    # equal shards: the reward all-gather goes through the C ABI (bg_allgather_f32 = ncclAllGather on the step's stream),
    # device-resident infos, observation ring; compared with the unsharded env on the same GPU
    total2 = 3 * world
    shard2 = ShardedVecBreedGym(total_envs=total2, device=local, info_device="device", async_rewards=True, **kw)
    assert shard2.collective == "native", shard2.collective
    full2 = VecBreedGym(num_envs=total2, device=local, info_device="device", **kw)
    shard2.reset(seed=5)
    full2.reset(seed=5)
    sl2 = shard2.local_slice()
    dev = torch.device("cuda", local)
    for step in range(9):
        act = rng.integers(0, n, (total2, n, 2)).astype(np.int32)
        ps, rs, _, ts, infos_s = shard2.step(torch.from_numpy(act[sl2]).to(dev))
        pf, rf, _, tf, infos_f = full2.step(torch.from_numpy(act).to(dev))
        shard2.wait_rewards()  # async_rewards: the all-gather ran on its own stream
        assert np.array_equal(np.asarray(ps), np.asarray(pf)[sl2]), f"native step {step}: populations differ"
        assert torch.equal(infos_s["GEBV"], infos_f["GEBV"][sl2]), f"native step {step}: GEBV differs"
        assert rs.shape == (total2,) and torch.equal(rs, rf), f"native step {step}: gathered rewards differ"
        assert np.array_equal(ts, tf)
    dist.barrier()
    if rank == 0:
        print(f"native all-gather ok: {world} ranks, {total2} envs, bg_allgather_f32 == unsharded rewards on every rank")
    dist.barrier()
    if rank == 0:
        print(f"multi-gpu check ok: {world} ranks, {total} envs, sharded == unsharded bit for bit")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
