#!/usr/bin/env python
"""Multi-GPU parity check (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 \
        scripts/multi_gpu_check.py

Every rank steps its shard of a ShardedVecBreedGym and, on the same GPU, the full unsharded VecBreedGym; the shard
must equal its slice of the full env bit for bit and the exchanged rewards must equal the full env's rewards --
for each reward exchange: peer memory (ragged shards, host mode; equal shards, device mode, asynchronous), the
NCCL all-gather through the C ABI, torch.distributed.
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
    rng = np.random.default_rng(0)  # same stream on every rank
    dev = torch.device("cuda", local)

    def host_mode(collective, expect):
        shard = ShardedVecBreedGym(total_envs=total, device=local, collective=collective, **kw)
        assert shard.collective == expect, (shard.collective, expect)
        full = VecBreedGym(num_envs=total, device=local, **kw)
        ps, _ = shard.reset(seed=3)
        pf, _ = full.reset(seed=3)
        sl = shard.local_slice()
        assert np.array_equal(np.asarray(ps), np.asarray(pf)[sl]), "reset differs"
        for step in range(9):  # crosses two autoresets
            act = rng.integers(0, n, (total, n, 2))
            ps, rs, _, ts, infos_s = shard.step(act[sl])
            pf, rf, _, tf, infos_f = full.step(act)
            assert np.array_equal(np.asarray(ps), np.asarray(pf)[sl]), f"{expect} step {step}: populations differ"
            assert np.array_equal(infos_s["GEBV"], infos_f["GEBV"][sl]), f"{expect} step {step}: GEBV differs"
            assert rs.shape == (total,) and np.array_equal(rs, rf), f"{expect} step {step}: exchanged rewards differ"
            assert np.array_equal(ts, tf)
        if shard._peer is not None:
            assert shard._peer.timeouts() == 0
        dist.barrier()
        shard.close()
        dist.barrier()
        if rank == 0:
            print(f"{expect} exchange ok (host mode): {world} ranks, {total} envs in ragged shards, sharded == unsharded bit for bit", flush=True)

    def device_mode(collective, expect):
        total2 = 3 * world
        shard2 = ShardedVecBreedGym(total_envs=total2, device=local, info_device="device", async_rewards=True, collective=collective, **kw)
        assert shard2.collective == expect, (shard2.collective, expect)
        full2 = VecBreedGym(num_envs=total2, device=local, info_device="device", **kw)
        shard2.reset(seed=5)
        full2.reset(seed=5)
        sl2 = shard2.local_slice()
        for step in range(13):  # three episode ends: the windows' two halves are both re-used
            act = rng.integers(0, n, (total2, n, 2)).astype(np.int32)
            ps, rs, _, ts, infos_s = shard2.step(torch.from_numpy(act[sl2]).to(dev))
            pf, rf, _, tf, infos_f = full2.step(torch.from_numpy(act).to(dev))
            shard2.wait_rewards()  # async_rewards: the step stream did not wait for the other ranks
            assert np.array_equal(np.asarray(ps), np.asarray(pf)[sl2]), f"{expect} step {step}: populations differ"
            assert torch.equal(infos_s["GEBV"], infos_f["GEBV"][sl2]), f"{expect} step {step}: GEBV differs"
            assert rs.shape == (total2,) and torch.equal(rs, rf), f"{expect} step {step}: exchanged rewards differ"
            assert np.array_equal(ts, tf)
        if shard2._peer is not None:
            assert shard2._peer.timeouts() == 0
        dist.barrier()
        shard2.close()
        dist.barrier()
        if rank == 0:
            print(f"{expect} exchange ok (device mode, asynchronous): {world} ranks, {total2} envs, == unsharded rewards on every rank", flush=True)

    host_mode("auto", "peer")      # ragged shards through the peer windows
    device_mode("auto", "peer")
    device_mode("native", "native")  # bg_allgather_f32 = ncclAllGather through the C ABI, on its own stream
    host_mode("torch", "torch")
    if rank == 0:
        print(f"multi-gpu check ok: {world} ranks", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
