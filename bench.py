#!/usr/bin/env python
"""Benchmark of the breedgym hot path on B200 ("Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one vector-env step of BASELINE.json's config[1] per GPU:
64 envs x 370 individuals x 10 000 markers (small_genetic_map.txt, trait Yield),
i.e. mask generation + blend of all envs (cross), GEBV of every offspring, the
max-GEBV reward and the on-device autoreset on every 10th step.  N > 1 shards
64 x N envs, 64 per GPU (weak scaling), one process per GPU; the only collective
is the reward all-gather on episode ends.

  value      env-steps/s, inputs resident in HBM, CUDA-event timed over the K steps; the working set
             (4 replicas of the workload, round-robin) is larger than L2
  e2e        the same through VecBreedGym.step with HOST actions in / GEBV+rewards out
  roofline   dominant kernel: algorithmic bytes / CUDA-event time vs measured HBM peak
  cpu_baseline / --impl reference: the C oracle (OpenMP) on the host cores -- jax/chromax are
             not installable on this image, so the oracle port stands in for the JAX-CPU path
"""
from __future__ import annotations

import argparse
import gc
import json
import os
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

ENVS_PER_GPU = 64
N_IND = 370
N_MARKERS = 10_000
NUM_GENERATIONS = 10
REPLICAS = 4         # independent copies of the workload stepped round-robin (working set > L2)
B_ALG_CROSS = 0.75   # bytes per offspring-marker: read 2 parents x 2 bits, write 2 bits
B_ALG_GEBV = 0.25    # bytes per individual-marker: read 2 bits
METRIC = "env_steps_per_sec"
UNIT = "env-steps/s"
WORKLOAD = "C2 vector env: 64 envs/GPU x 370 individuals x 10000 markers, 10 generations, Yield"


def workload_inputs():
    data = ROOT / "breedgym_b200" / "data"
    germ = np.random.default_rng(0).random((N_IND, N_MARKERS, 2)) < 0.5  # small_geno.npy is not in the snapshot
    return germ, data / "small_genetic_map.txt"


def config_dict(n_gpus, replicas=REPLICAS):
    REPLICAS = replicas  # noqa: N806 (shadow for the f-strings below)
    return {
        "workload": WORKLOAD,
        "envs_total": ENVS_PER_GPU * n_gpus,
        "envs_per_gpu": ENVS_PER_GPU,
        "individuals": N_IND,
        "markers": N_MARKERS,
        "traits": 1,
        "num_generations": NUM_GENERATIONS,
        "parallelism": f"env-sharded x{n_gpus}" if n_gpus > 1 else "single GPU",
        "l2": f"inputs larger than L2: {REPLICAS} independent replicas stepped round-robin "
              f"({REPLICAS} x {2 * ENVS_PER_GPU * N_IND * 2560 / 1e6:.0f} MB of populations between reuse, L2 = 126 MB); "
              f"per-kernel breakdown: 256 MiB flush",
        "observation": "packed bit planes resident in HBM (bool observation materialised on request only)",
        "rng": "threefry2x32 legacy layout, key schedule S2, seed 7",
    }


# --------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle's C restatement on the host cores
# --------------------------------------------------------------------------------------
def cpu_steps(n_steps, warmup, envs=ENVS_PER_GPU, min_seconds=0.0):
    """Times full vector-env steps (cross of all envs + GEBV + max reward) with the C oracle."""
    from oracle import c_oracle as co
    from oracle import chromax_ref as cr
    from oracle import jax_prng as jp

    co.set_threads(os.cpu_count() or 1)  # all host cores, whatever OMP_NUM_THREADS the launcher exported
    germ, gmap = workload_inputs()
    g = cr.read_genetic_map(gmap)
    r = cr.recombination_vector(g)
    eff = cr.marker_effects(g, ["Yield"])
    rng = np.random.default_rng(1)
    _, pops, _ = cr.vec_reset(germ, N_IND, envs, jp.key(7))
    key = jp.key(7)
    times = []
    i = 0
    while True:
        act = rng.integers(0, N_IND, (envs, N_IND, 2), dtype=np.int32)
        ks = jp.split(key, 2)
        key, k = ks[0], ks[1]
        t0 = time.perf_counter()
        pops = co.cross_envs(pops, act, r, k)
        gebv = co.gebv(pops, eff)
        _ = gebv.max(axis=(1, 2))
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
        i += 1
        if len(times) >= n_steps and sum(times) >= min_seconds:
            break
    total = float(sum(times))
    return {"env_steps_per_sec": envs * len(times) / total, "seconds": total, "steps": len(times),
            "cores": co.num_threads(), "envs": envs}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    # bounded sample: every step processes `envs` of the 64 envs of a GPU's share, chosen so that the whole
    # --steps K --warmup W run stays around a minute whatever K is (throughput per env-step does not depend on it)
    calib = cpu_steps(1, 0, envs=8)
    per_env_step = 1.0 / calib["env_steps_per_sec"]
    budget_s = 60.0
    envs = int(max(1, min(ENVS_PER_GPU, budget_s / (per_env_step * max(1, args.steps + args.warmup)))))
    res = cpu_steps(args.steps, args.warmup, envs=envs)
    sample = (f"{res['steps']} vector-env steps of {res['envs']} envs x {N_IND} x {N_MARKERS} each "
              f"(cross + GEBV + reward), C oracle, OpenMP")
    line = {
        "impl": "reference",
        "metric": METRIC, "value": res["env_steps_per_sec"], "unit": UNIT,
        "n_gpus": args.gpus, "steps": res["steps"], "warmup": args.warmup,
        "ms_per_step": 1e3 * res["seconds"] / res["steps"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u32 bit planes + int64 fixed point (CPU port: u8 + f64)", "data": "synthetic",
        "config": config_dict(args.gpus),
        "offspring_markers_per_sec": res["env_steps_per_sec"] * N_IND * N_MARKERS,
        "cpu_baseline": {"value": res["env_steps_per_sec"], "unit": UNIT, "cores": res["cores"], "kind": "port",
                         "sample": sample,
                         "note": "jax/chromax are not installable on this image; the oracle's C restatement "
                                 "of the reference algorithm is timed instead"},
        "e2e": {"value": res["env_steps_per_sec"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# --------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index, period=0.02):
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._thread = None
        self.period = period
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def _sample_once(self):
        try:
            self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
        except Exception:
            pass

    def start(self):
        if self.nv is not None:
            self._sample_once()  # a short timed region may end before the thread's first tick
            self._thread = threading.Thread(target=self._loop, daemon=True)
            self._thread.start()

    def stop(self):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()
        if self.nv is not None:
            self._sample_once()
        return {
            "sm_mhz": float(np.median(self.samples)) if self.samples else None,
            "sm_max_mhz": self.max_mhz,
            "reasons": sorted(self.reasons),
            "samples": len(self.samples),
        }


# --------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------
def measured_peak_gbs():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def run_ours(args):
    import torch
    import torch.distributed as dist

    from breedgym_b200 import _lib
    from breedgym_b200.vector import VecBreedGym
    from breedgym_b200.vector.sharded import allgather_rewards, shard_counts, shard_range

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    total_envs = ENVS_PER_GPU * world
    begin, count = shard_range(total_envs, world, rank)
    counts = shard_counts(total_envs, world)
    germ, gmap = workload_inputs()
    lib = _lib.load()
    K, W = args.steps, max(args.warmup, 3)
    REPLICAS = max(1, args.replicas)

    def make_env(info_device):
        env = VecBreedGym(num_envs=count, initial_population=germ, genetic_map=gmap, trait_names=["Yield"],
                          individual_per_gen=N_IND, num_generations=NUM_GENERATIONS, device=local_rank,
                          info_device=info_device, env_shard=(begin, total_envs))
        env.reset(seed=7)
        return env

    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def flush_l2():
        flush_buf.fill_(1)

    def spin_up(seconds=0.4):
        """Keep the GPU busy long enough for the SM clock to leave its idle state before anything is timed."""
        t0 = time.perf_counter()
        while time.perf_counter() - t0 < seconds:
            for _ in range(20):
                flush_buf.fill_(0)
            torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    rng = np.random.default_rng(1 + rank)
    n_act = 16  # distinct action batches, cycled
    acts_host = [rng.integers(0, N_IND, (count, N_IND, 2), dtype=np.int32) for _ in range(n_act)]
    acts_dev = [torch.from_numpy(a).to(dev) for a in acts_host]

    # ---------------- value: device-resident inputs, pipelined, inputs larger than L2 ----------------
    # REPLICAS independent copies of the workload are stepped round-robin: between two steps of the
    # same replica ~REPLICAS x 120 MB of other populations stream through the 126 MB L2, so every
    # step reads its population from HBM without a flush kernel polluting the pipeline.
    envs = [make_env("device") for _ in range(REPLICAS)]
    env = envs[0]

    def device_step(i):
        _, rews, _, tru, _ = envs[i % REPLICAS].step(acts_dev[i % n_act])
        if world > 1 and bool(tru[0]):
            allgather_rewards(rews, counts)  # the one collective of the path (NCCL)

    def timed(fn, steps):
        """K steps bracketed by barrier + synchronize; device time between two CUDA events."""
        # a full (generation 2) Python garbage collection walks every object the imported packages hold: 40-500 ms,
        # once every few thousand steps, wherever it happens to fall (seen as one block 5-10x slower than the others).
        # Collect now and park the survivors in the permanent generation so that none falls inside the timed region.
        gc.collect()
        gc.freeze()
        stream = torch.cuda.current_stream(dev)
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        start.record(stream)
        for i in range(steps):
            fn(i)
        end.record(stream)
        barrier()
        return start.elapsed_time(end) * 1e-3  # seconds

    spin_up()
    for i in range(max(W, REPLICAS)):
        device_step(i)
    # untimed dress rehearsal of the timed region: on a freshly provisioned box the first seconds of a process run with
    # cold page / instruction caches on the host (enqueue 3x slower: the device loop turns host-bound)
    timed(device_step, K)
    sampler = ClockSampler(local_rank)
    launches0 = lib.bg_kernel_launches()
    sampler.start()
    t_value = max_over_ranks(timed(device_step, K))
    launches = lib.bg_kernel_launches() - launches0
    clocks = sampler.stop()

    # ---------------- per-kernel breakdown (same inputs, same flush policy) ----------------
    sim = env.simulator
    pop_words = env.populations.words
    E = count
    Wpad = sim.words_per_row
    mask = torch.empty((2 * N_IND, Wpad), dtype=torch.int32, device=dev)
    out_words = torch.empty_like(pop_words)
    gebv_out = torch.empty((E, N_IND, 1), dtype=torch.float32, device=dev)
    key = np.array([0, 12345], dtype=np.uint32)
    stream = torch.cuda.current_stream(dev)
    sptr = sim._stream()
    # the step = meiosis_masks (side stream, one step ahead) + cross_gebv_fused; blend_envs / gebv are the two
    # kernels of the unfused path (BG_NO_FUSE=1), timed for comparison
    bk = {"meiosis_masks": 0.0, "cross_gebv_fused": 0.0, "blend_envs": 0.0, "gebv": 0.0}
    reps = min(K, 50)
    spin_up(0.2)
    # masks of `key` land in one of the engine's slots here: the timed bg_cross_gebv calls launch the fused kernel only
    _lib.check(lib.bg_cross_gebv(sim._engine, pop_words.data_ptr(), acts_dev[0].data_ptr(), out_words.data_ptr(), E, N_IND, N_IND,
                                 _lib.nptr(key), 0, 2, gebv_out.data_ptr(), sptr))
    for it in range(3 + reps):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(8)]
        flush_l2()
        ev[0].record(stream)
        _lib.check(lib.bg_meiosis_masks(sim._engine, mask.data_ptr(), 2 * N_IND, _lib.nptr(key), 0, 2, sptr))
        ev[1].record(stream)
        flush_l2()
        ev[2].record(stream)
        _lib.check(lib.bg_cross_gebv(sim._engine, pop_words.data_ptr(), acts_dev[it % n_act].data_ptr(), out_words.data_ptr(), E,
                                     N_IND, N_IND, _lib.nptr(key), 0, 2, gebv_out.data_ptr(), sptr))
        ev[3].record(stream)
        flush_l2()
        ev[4].record(stream)
        _lib.check(lib.bg_blend_envs(sim._engine, pop_words.data_ptr(), acts_dev[it % n_act].data_ptr(), mask.data_ptr(),
                                     None, out_words.data_ptr(), E, N_IND, N_IND, sptr))
        ev[5].record(stream)
        flush_l2()
        ev[6].record(stream)
        _lib.check(lib.bg_gebv(sim._engine, out_words.data_ptr(), E * N_IND, gebv_out.data_ptr(), sptr))
        ev[7].record(stream)
        torch.cuda.synchronize()
        if it >= 3:
            bk["meiosis_masks"] += ev[0].elapsed_time(ev[1])
            bk["cross_gebv_fused"] += ev[2].elapsed_time(ev[3])
            bk["blend_envs"] += ev[4].elapsed_time(ev[5])
            bk["gebv"] += ev[6].elapsed_time(ev[7])
    bk = {k: v / reps for k, v in bk.items()}  # ms per launch
    peak, peak_src = measured_peak_gbs()
    om = E * N_IND * N_MARKERS
    # algorithmic bytes: cross 0.75 B per offspring-marker; the fused kernel scores the offspring it
    # has just built, so its GEBV adds no bytes; the standalone GEBV reads 0.25 B per individual-marker
    alg = {"meiosis_masks": None, "cross_gebv_fused": B_ALG_CROSS * om, "blend_envs": B_ALG_CROSS * om, "gebv": B_ALG_GEBV * om}
    kernels = {}
    for k, ms in bk.items():
        kernels[k] = {"ms": ms, "algorithmic_bytes": alg[k],
                      "achieved_gbs": (alg[k] / (ms * 1e-3) / 1e9) if alg[k] else None}
        if alg[k]:
            kernels[k]["frac_of_hbm_peak"] = kernels[k]["achieved_gbs"] / peak
    dom = "cross_gebv_fused"  # the one kernel on the step's critical path
    roofline = {"kernel": dom, "bound": "hbm", "achieved": kernels[dom]["achieved_gbs"], "peak": peak, "unit": "GB/s",
                "frac": kernels[dom]["achieved_gbs"] / peak, "traffic": None, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg[dom], "ms_per_launch": bk[dom]}
    prof = ROOT / "profiles" / "traffic.json"  # dram bytes per launch from the committed ncu capture
    if prof.exists():
        try:
            roofline["traffic"] = json.loads(prof.read_text()).get(dom)
        except Exception:
            pass

    # ---------------- e2e: public API, host actions in, GEBV + rewards out ----------------
    del envs
    if args.skip_e2e:
        if rank == 0:
            print(json.dumps({"value": total_envs * K / t_value, "ms_per_step": 1e3 * t_value / K, "replicas": REPLICAS,
                              "kernels_ms": bk}), flush=True)
        return 0
    envs_h = [make_env("host") for _ in range(REPLICAS)]

    def host_step(i):
        obs, rews, ter, tru, infos = envs_h[i % REPLICAS].step(acts_host[i % n_act])
        if world > 1 and bool(tru[0]):
            allgather_rewards(torch.from_numpy(np.ascontiguousarray(rews, dtype=np.float32)).to(dev), counts)

    spin_up()
    for i in range(max(W, REPLICAS)):
        host_step(i)
    timed(host_step, min(K, 500))  # untimed rehearsal, as above
    # timed in blocks (same total K): a block far slower than the others points at the box (clock state, a descheduled
    # host thread), not at the path; reported beside the total
    nblk = 4 if K >= 400 else 1
    blocks = []
    sampler_e2e = ClockSampler(local_rank)
    sampler_e2e.start()
    base = 0
    for b in range(nblk):
        kb = K // nblk + (1 if b < K % nblk else 0)
        blocks.append(timed(lambda i, base=base: host_step(base + i), kb) / kb)
        base += kb
    clocks_e2e = sampler_e2e.stop()
    t_e2e = max_over_ranks(sum(bt * (K // nblk + (1 if b < K % nblk else 0)) for b, bt in enumerate(blocks)))
    h2d = count * N_IND * 2 * 4
    d2h = count * N_IND * 4 + (count * 4) / NUM_GENERATIONS

    # ---------------- cpu baseline (rank 0, N=1 only) ----------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        res = cpu_steps(2, 1, envs=ENVS_PER_GPU, min_seconds=10.0)
        cpu = {"value": res["env_steps_per_sec"], "unit": UNIT, "cores": res["cores"], "kind": "port",
               "sample": f"{res['steps']} full steps of {res['envs']} envs x {N_IND} x {N_MARKERS} "
                         f"(cross + GEBV + reward) in {res['seconds']:.1f} s, C oracle with OpenMP"}

    if rank == 0:
        value = total_envs * K / t_value
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": 1e3 * t_value / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32 bit planes (cross) + int64 fixed point (GEBV)", "data": "synthetic",
            "config": config_dict(world, REPLICAS),
            "offspring_markers_per_sec": value * N_IND * N_MARKERS,
            "clocks": clocks,
            "e2e": {"value": total_envs * K / t_e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": "VecBreedGym.step(numpy actions) -> numpy GEBV / rewards, one sync per step",
                    "us_per_step_by_block": [round(1e6 * bt, 1) for bt in blocks], "clocks": clocks_e2e},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "kernels": kernels,
            "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


_REAL_STDOUT = None


def _quiet_stdout():
    """The contract is ONE JSON line on stdout: libraries that write there on their own (NCCL prints its version
    banner to stdout when NCCL_DEBUG is set) are sent to stderr; the JSON line goes to the real stdout."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)
        sys.stdout = _REAL_STDOUT


def main():
    _quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--replicas", type=int, default=REPLICAS, help="independent workload copies stepped round-robin")
    ap.add_argument("--skip-e2e", action="store_true", help="diagnostics: only the device-resident value")
    ap.add_argument("--envs-per-gpu", type=int, default=ENVS_PER_GPU,
                    help="envs per GPU (default 64 = BASELINE config C2; 512 = one GPU's share of C5, 4096 envs on 8 GPUs)")
    args = ap.parse_args()
    if args.envs_per_gpu != ENVS_PER_GPU:
        globals()["ENVS_PER_GPU"] = args.envs_per_gpu
        globals()["WORKLOAD"] = WORKLOAD.replace("C2 vector env: 64 envs/GPU", f"vector env: {args.envs_per_gpu} envs/GPU")
        if args.replicas == REPLICAS and args.envs_per_gpu >= 256:
            args.replicas = 1  # one population pair already exceeds L2
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
