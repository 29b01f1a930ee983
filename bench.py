#!/usr/bin/env python
"""Benchmark of the breedgym hot path on B200 ("Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one vector-env step of BASELINE.json's config[1] per GPU:
64 envs x 370 individuals x 10 000 markers (small_genetic_map.txt, trait Yield),
i.e. the fused cross + GEBV kernel over all envs (the crossover masks of the following
steps are generated on a side stream), the max-GEBV reward and the on-device autoreset
on every 10th step.  N > 1 shards 64 x N envs, 64 per GPU (weak scaling), one process
per GPU; the only collective is the reward all-gather on episode ends (ncclAllGather
through the C ABI, on the step's stream).

  value      env-steps/s, inputs resident in HBM, CUDA-event timed over the K steps; the working set
             (4 replicas of the workload, round-robin) is larger than L2
  e2e        the same through VecBreedGym.step with HOST actions in / GEBV+rewards out
  roofline   dominant kernel: algorithmic bytes / CUDA-event time vs measured HBM peak
  cpu_baseline / --impl reference: the C oracle (OpenMP) on the host cores, drawing the crossover masks once per
             step for all envs as the reference's vmap does -- jax/chromax are not installable on this image, so
             the oracle port stands in for the JAX-CPU path
  c5         BASELINE config[4]: 512 envs per GPU (weak) and 4096 envs in total (strong) records
  configs    (N = 1) the other named shapes: C1 single env through the Gym API, C3 wheat-scale cross + GEBV,
             C4 10 000 x 1 M markers cross + 16-trait GEBV, each with its own roofline and CPU baseline
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


def config_dict(n_gpus, replicas=REPLICAS, envs_per_gpu=None, envs_total=None):
    epg = ENVS_PER_GPU if envs_per_gpu is None else envs_per_gpu
    return {
        "workload": WORKLOAD,
        "envs_total": epg * n_gpus if envs_total is None else envs_total,
        "envs_per_gpu": epg,
        "individuals": N_IND,
        "markers": N_MARKERS,
        "traits": 1,
        "num_generations": NUM_GENERATIONS,
        "parallelism": f"env-sharded x{n_gpus}" if n_gpus > 1 else "single GPU",
        "l2": f"inputs larger than L2: {replicas} independent replicas stepped round-robin "
              f"({replicas} x {2 * epg * N_IND * 2560 / 1e6:.0f} MB of populations between reuse, L2 = 126 MB); "
              f"step kernel: back to back over the replicas' populations, other kernels behind a 256 MiB flush",
        "window": "the replicas are staggered (episode ends and mask batches spread over the round-robin); the closing event of "
                  "the timed region waits for everything the steps started on the library's side stream (bg_engine_join: masks "
                  "of following steps, prefetched resets) and for the reward exchange",
        "observation": "packed bit planes resident in HBM (bool observation materialised on request only)",
        "rng": "threefry2x32 legacy layout, key schedule S2, seed 7",
        "stream": "steps on a CUDA stream of priority -1 (torch.cuda.Stream(priority=-1)); the library's mask lookahead runs on its own side stream at priority 0",
    }


# --------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle's C restatement on the host cores
# --------------------------------------------------------------------------------------
def cpu_steps(n_steps, warmup, envs=ENVS_PER_GPU, min_seconds=0.0, shared_masks=True):
    """Times full vector-env steps (cross of all envs + GEBV + max reward) with the C oracle.  shared_masks: the 2n
    crossover masks are drawn once per step and reused by every env (how the reference's vmap executes,
    breedgym/vector/vec_env.py:75-77); False re-draws them per env (E x the Threefry work)."""
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
        pops = co.cross_envs(pops, act, r, k, shared_masks=shared_masks)
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


CPU_NOTE = ("jax/chromax are not installable on this image; the oracle's C restatement of the reference algorithm is timed "
            "instead (masks drawn once per step for all envs, as the reference's vmap does)")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    # bounded sample: every step processes `envs` of the 64 envs of ONE GPU's share (rank 0 only, whatever --gpus is),
    # chosen so that the whole --steps K --warmup W run stays around a minute (throughput per env-step barely depends
    # on it: the per-step mask draw is amortised over fewer envs, which makes a smaller sample slightly pessimistic)
    calib = cpu_steps(1, 0, envs=8)
    per_env_step = 1.0 / calib["env_steps_per_sec"]
    budget_s = 60.0
    envs = int(max(1, min(ENVS_PER_GPU, budget_s / (per_env_step * max(1, args.steps + args.warmup)))))
    res = cpu_steps(args.steps, args.warmup, envs=envs)
    redraw = cpu_steps(1, 0, envs=min(envs, 8), shared_masks=False)
    sample = (f"{res['steps']} vector-env steps of {res['envs']} envs x {N_IND} x {N_MARKERS} each (cross + GEBV + reward) "
              f"on rank 0's host cores, C oracle with OpenMP, shared-mask port (2n masks drawn once per step)")
    cfg = config_dict(1, envs_per_gpu=res["envs"])
    cfg["parallelism"] = f"{res['cores']} host threads (CPU arm: runs on rank 0 only, --gpus {args.gpus} is ignored)"
    cfg["l2"] = "n/a (CPU)"
    line = {
        "impl": "reference",
        "metric": METRIC, "value": res["env_steps_per_sec"], "unit": UNIT,
        "n_gpus": args.gpus, "steps": res["steps"], "warmup": args.warmup,
        "ms_per_step": 1e3 * res["seconds"] / res["steps"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8 alleles + f64 GEBV (CPU port)", "data": "synthetic",
        "config": cfg,
        "offspring_markers_per_sec": res["env_steps_per_sec"] * N_IND * N_MARKERS,
        "cpu_baseline": {"value": res["env_steps_per_sec"], "unit": UNIT, "cores": res["cores"], "kind": "port",
                         "sample": sample, "note": CPU_NOTE,
                         "per_env_redraw_value": redraw["env_steps_per_sec"]},
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
def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            d = json.loads(p.read_text())
            return d, "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1700.0}, "fallback (B200_PROFILING.md)"


def measured_peak_gbs():
    d, src = measured_peaks()
    return float(d["hbm_gbs"]), src


def int32_peak():
    """Measured integer-pipe issue rate (scripts/int32_peak.cu -> profiles/int32_peak.json), for the Threefry-bound kernels."""
    p = ROOT / "profiles" / "int32_peak.json"
    if p.exists():
        try:
            return json.loads(p.read_text())
        except Exception:
            pass
    return None


def traffic_for(kernel, key):
    """DRAM bytes per launch from the committed ncu capture of exactly this configuration, else None."""
    prof = ROOT / "profiles" / "traffic.json"
    try:
        return json.loads(prof.read_text()).get(kernel, {}).get(key)
    except Exception:
        return None


class Harness:
    """Timing plumbing shared by the legs: barrier + synchronize on both sides, CUDA events on the launching stream,
    max over ranks, SM clock spin-up, garbage collector parked."""

    def __init__(self, torch, dist, dev, world):
        self.torch, self.dist, self.dev, self.world = torch, dist, dev, world
        self.flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def flush_l2(self):
        self.flush_buf.fill_(1)

    def spin_up(self, seconds=0.4):
        """Keep the GPU busy long enough for the SM clock to leave its idle state before anything is timed."""
        t0 = time.perf_counter()
        while time.perf_counter() - t0 < seconds:
            for _ in range(20):
                self.flush_buf.fill_(0)
            self.torch.cuda.synchronize()

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def align_start(self):
        """Ranks leave an NCCL barrier tens of microseconds apart, and a rank that starts late makes every other rank's
        closing event (which waits for ALL ranks' rewards) that much later: after the barrier the ranks agree on a
        common wall-clock deadline (same host, same clock) and spin until it."""
        if self.world == 1:
            return
        t = self.torch.tensor([time.time_ns() + 400_000], dtype=self.torch.int64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        deadline = int(t.item())
        self.torch.cuda.synchronize()
        while time.time_ns() < deadline:
            pass

    def max_over_ranks(self, x):
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def timed(self, fn, steps, finish=None):
        """`steps` calls bracketed by barrier + synchronize; returns (device seconds between two CUDA events, host
        seconds spent enqueueing).  `finish` runs before the closing event (it orders the timed stream behind work
        the steps started elsewhere: the reward exchange)."""
        # a full (generation 2) Python garbage collection walks every object the imported packages hold: 40-500 ms,
        # wherever it happens to fall.  Collect now and park the survivors so that none falls inside the timed region.
        gc.collect()
        gc.freeze()
        torch = self.torch
        stream = torch.cuda.current_stream(self.dev)
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        self.align_start()
        start.record(stream)
        t0 = time.perf_counter()
        for i in range(steps):
            fn(i)
        host = time.perf_counter() - t0
        if finish is not None:
            finish()
        end.record(stream)
        self.barrier()
        return start.elapsed_time(end) * 1e-3, host


def vec_value_leg(h, make_env, acts_dev, K, W, replicas, sample_clocks=None, steady_steps=0):
    """Device-resident step loop: `replicas` independent envs round-robin.  Warm-up runs every replica through a
    whole episode (autoreset and the reward all-gather included) plus one untimed rehearsal of the K steps, so the
    timed K steps carry the steady one-in-`num_generations` reset rate on warm code paths."""
    envs = [make_env("device") for _ in range(replicas)]
    n_act = len(acts_dev)

    def step(i):
        envs[i % replicas].step(acts_dev[i % n_act])

    def finish():
        # The closing event waits for everything the timed steps STARTED: the crossover masks of following steps that
        # the library generates on its side stream, prefetched resets (env.join = bg_engine_join), and -- env-sharded
        # runs exchange the rewards asynchronously -- every rank's rewards of every replica's last episode.
        for env in envs:
            env.join()
            if hasattr(env, "wait_rewards"):
                env.wait_rewards()

    h.spin_up()
    # Stagger the replicas: replica r runs a few steps ahead, so that the replicas' episode ends (every 10th step) and
    # mask batches (every 8th) are spread over the round-robin the way independent env sets are in a long run -- not
    # all on consecutive iterations, which puts either all or none of them inside a K = 20 window.
    for r in range(replicas):
        for j in range(r * NUM_GENERATIONS // replicas):  # 0, 2, 5, 7 at 4 replicas
            envs[r].step(acts_dev[j % n_act])
    for i in range(max(W, replicas * (NUM_GENERATIONS + 1))):
        step(i)
    h.timed(step, K, finish)
    if sample_clocks is not None:
        sample_clocks.start()
    t_dev, t_host = h.timed(step, K, finish)
    clocks = sample_clocks.stop() if sample_clocks is not None else None
    steady = None
    if steady_steps and steady_steps > K:
        # the same loop over a window long enough to carry the steady-state share of everything that runs beside the
        # steps (a K = 20 window starts on crossover masks the rehearsal left ready; over a long run the mask kernel of
        # the FOLLOWING batches competes with the step kernels for issue slots all the time)
        ts_dev, ts_host = h.timed(step, steady_steps, finish)
        steady = (h.max_over_ranks(ts_dev), ts_host, steady_steps)
    return envs, h.max_over_ranks(t_dev), t_host, clocks, steady


def run_ours(args):
    import torch
    import torch.distributed as dist

    from breedgym_b200 import _lib
    from breedgym_b200.vector import ShardedVecBreedGym, VecBreedGym
    from breedgym_b200.vector.sharded import shard_range

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
    h = Harness(torch, dist, dev, world)
    if args.stream_priority != 0:
        # The step stream gets a higher priority than the library's side stream (which generates the crossover masks
        # of the FOLLOWING steps, one launch per batch of steps): whenever an SM has room, the block scheduler then
        # places the step kernel's CTAs first and the mask kernel only fills what is left (torch's default stream has
        # the lowest priority there is, so the side stream cannot be put below it).
        torch.cuda.set_stream(torch.cuda.Stream(device=dev, priority=args.stream_priority))

    germ, gmap = workload_inputs()
    lib = _lib.load()
    K, W = args.steps, max(args.warmup, 3)
    replicas = max(1, args.replicas)
    env_kw = dict(initial_population=germ, genetic_map=gmap, trait_names=["Yield"], individual_per_gen=N_IND,
                  num_generations=NUM_GENERATIONS, device=local_rank)

    def env_factory(total_envs):
        def make_env(info_device):
            if world > 1:  # one process per GPU, env-sharded; the reward all-gather goes through bg_allgather_f32
                env = ShardedVecBreedGym(total_envs=total_envs, info_device=info_device, async_rewards=True, **env_kw)
            else:
                env = VecBreedGym(num_envs=total_envs, info_device=info_device, **env_kw)
            env.reset(seed=7)
            return env
        return make_env

    def actions(count, n_act, seed):
        rng = np.random.default_rng(seed)
        host = [rng.integers(0, N_IND, (count, N_IND, 2), dtype=np.int32) for _ in range(n_act)]
        # host copies live in pinned memory (the e2e leg's inputs: the step copies them to the device as they are)
        return [torch.from_numpy(a).pin_memory() for a in host], [torch.from_numpy(a).to(dev) for a in host]

    total_envs = ENVS_PER_GPU * world
    begin, count = shard_range(total_envs, world, rank)
    acts_host, acts_dev = actions(count, 16, 1 + rank)

    # ---------------- value: device-resident inputs, pipelined, inputs larger than L2 ----------------
    sampler = ClockSampler(local_rank)
    launches0 = lib.bg_kernel_launches()
    envs, t_value, t_host, clocks, steady = vec_value_leg(h, env_factory(total_envs), acts_dev, K, W, replicas, sampler,
                                                          steady_steps=args.steady_steps)
    launches_total = lib.bg_kernel_launches() - launches0
    reward_exchange = None
    if world > 1:
        how = envs[0].collective
        reward_exchange = {"peer": "peer memory: the reward reduction stores into every rank's window over NVLink (bg_peer_*, no collective launch); "
                                   "the closing event of the timed region waits for every rank's rewards",
                           "native": "ncclAllGather through bg_allgather_f32 on its own stream",
                           "torch": "torch.distributed all_gather"}[how]
    # launches inside the timed region: counted over one more pass of K steps
    l0 = lib.bg_kernel_launches()
    for i in range(K):
        envs[i % replicas].step(acts_dev[i % 16])
    torch.cuda.synchronize()
    launches = lib.bg_kernel_launches() - l0

    # ---------------- per-kernel breakdown (same inputs, same flush policy) ----------------
    env = envs[0].env if world > 1 else envs[0]
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
    # the step = cross_gebv_fused (+ meiosis_masks on the side stream, one launch per 8 steps); blend_envs / gebv are
    # the two kernels of the unfused path (option fuse=0), timed for comparison
    bk = {"meiosis_masks": 0.0, "cross_gebv_fused": 0.0, "blend_envs": 0.0, "gebv": 0.0}
    reps = min(max(K, 20), 50)
    h.spin_up(0.2)
    # masks of `key` land in one of the engine's mask batches here: the timed bg_cross_gebv calls launch the fused kernel only
    _lib.check(lib.bg_cross_gebv(sim._engine, pop_words.data_ptr(), acts_dev[0].data_ptr(), out_words.data_ptr(), E, N_IND, N_IND,
                                 _lib.nptr(key), 0, 2, gebv_out.data_ptr(), sptr))
    for it in range(3 + reps):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(8)]
        h.flush_l2()
        ev[0].record(stream)
        _lib.check(lib.bg_meiosis_masks(sim._engine, mask.data_ptr(), 2 * N_IND, _lib.nptr(key), 0, 2, sptr))
        ev[1].record(stream)
        h.flush_l2()
        ev[2].record(stream)
        _lib.check(lib.bg_cross_gebv(sim._engine, pop_words.data_ptr(), acts_dev[it % 16].data_ptr(), out_words.data_ptr(), E,
                                     N_IND, N_IND, _lib.nptr(key), 0, 2, gebv_out.data_ptr(), sptr))
        ev[3].record(stream)
        h.flush_l2()
        ev[4].record(stream)
        _lib.check(lib.bg_blend_envs(sim._engine, pop_words.data_ptr(), acts_dev[it % 16].data_ptr(), mask.data_ptr(),
                                     None, out_words.data_ptr(), E, N_IND, N_IND, sptr))
        ev[5].record(stream)
        h.flush_l2()
        ev[6].record(stream)
        _lib.check(lib.bg_gebv(sim._engine, out_words.data_ptr(), E * N_IND, gebv_out.data_ptr(), sptr))
        ev[7].record(stream)
        torch.cuda.synchronize()
        if it >= 3:
            bk["meiosis_masks"] += ev[0].elapsed_time(ev[1])
            bk["cross_gebv_fused"] += ev[2].elapsed_time(ev[3])
            bk["blend_envs"] += ev[4].elapsed_time(ev[5])
            bk["gebv"] += ev[6].elapsed_time(ev[7])
    bk = {k: v / reps for k, v in bk.items()}  # ms per launch, each launch alone behind an L2 flush
    # The step kernel's average launch duration as the stream runs it: launches back to back between two events, over
    # the replicas' populations in turn (4 x 121 MB of inputs + outputs between two uses of the same buffers: larger
    # than L2, no flush kernel in between).  The isolated figure above additionally carries ~5 us of launch / event
    # latency per launch and cannot see the programmatic dependent launch of consecutive step kernels.
    srcs = [(e.env if world > 1 else e).populations.words for e in envs]
    outs = [torch.empty_like(w) for w in srcs]
    n_sets = len(srcs)
    nb2b = 10 * n_sets

    def b2b(n_launch):
        for i in range(n_launch):
            r = i % len(srcs)
            _lib.check(lib.bg_cross_gebv(sim._engine, srcs[r].data_ptr(), acts_dev[i % 16].data_ptr(), outs[r].data_ptr(), E, N_IND, N_IND,
                                         _lib.nptr(key), 0, 2, gebv_out.data_ptr(), sptr))

    b2b(2 * len(srcs))
    torch.cuda.synchronize()
    b2b_ms = []
    for _ in range(5):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        b2b(nb2b)
        ev1.record(stream)
        torch.cuda.synchronize()
        b2b_ms.append(ev0.elapsed_time(ev1) / nb2b)
    bk_isolated = dict(bk)
    bk["cross_gebv_fused"] = sorted(b2b_ms)[len(b2b_ms) // 2]
    del outs, srcs
    peak, peak_src = measured_peak_gbs()
    om = E * N_IND * N_MARKERS
    # algorithmic bytes: cross 0.75 B per offspring-marker; the fused kernel scores the offspring it
    # has just built, so its GEBV adds no bytes; the standalone GEBV reads 0.25 B per individual-marker
    alg = {"meiosis_masks": None, "cross_gebv_fused": B_ALG_CROSS * om, "blend_envs": B_ALG_CROSS * om, "gebv": B_ALG_GEBV * om}
    kernels = {}
    for k, ms in bk.items():
        kernels[k] = {"ms": ms, "algorithmic_bytes": alg[k],
                      "achieved_gbs": (alg[k] / (ms * 1e-3) / 1e9) if alg[k] else None}
        if k == "cross_gebv_fused":
            kernels[k]["ms_isolated_after_l2_flush"] = bk_isolated[k]
            kernels[k]["timing"] = (f"median of 5 runs of {nb2b} launches back to back between two CUDA events, over {n_sets} population "
                                    f"sets in turn (inputs larger than L2); the other kernels: one launch behind a 256 MiB L2 flush")
        if alg[k]:
            kernels[k]["frac_of_hbm_peak"] = kernels[k]["achieved_gbs"] / peak
    ip = int32_peak()
    if ip:  # Threefry-bound: 2 * rows * ceil(m / 2) blocks... reported as draws/s against the measured integer issue rate
        draws = 2 * N_IND * N_MARKERS
        kernels["meiosis_masks"]["gdraws_per_s"] = draws / (bk["meiosis_masks"] * 1e-3) / 1e9
        kernels["meiosis_masks"]["int_ops_per_draw"] = ip.get("threefry_int_ops_per_draw")
        if ip.get("int32_gops") and ip.get("threefry_int_ops_per_draw"):
            kernels["meiosis_masks"]["frac_of_int32_peak"] = (kernels["meiosis_masks"]["gdraws_per_s"] * ip["threefry_int_ops_per_draw"]
                                                              / ip["int32_gops"])
    dom = "cross_gebv_fused"  # the one kernel on the step's critical path
    roofline = {"kernel": dom, "bound": "hbm", "achieved": kernels[dom]["achieved_gbs"], "peak": peak, "unit": "GB/s",
                "frac": kernels[dom]["achieved_gbs"] / peak,
                "traffic": traffic_for(dom, f"E{E}_n{N_IND}_m{N_MARKERS}"), "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg[dom], "ms_per_launch": bk[dom],
                "ms_per_launch_isolated_after_l2_flush": bk_isolated[dom],
                "whole_step_frac": alg[dom] / (t_value / K) / 1e9 / peak}

    # ---------------- e2e: public API, host actions in, GEBV + rewards out ----------------
    del envs, env, sim, pop_words
    if args.skip_e2e:
        if rank == 0:
            print(json.dumps({"value": total_envs * K / t_value, "ms_per_step": 1e3 * t_value / K, "replicas": replicas,
                              "steady_ms_per_step": None if steady is None else 1e3 * steady[0] / steady[2],
                              "host_us_per_step": 1e6 * t_host / K, "kernels_ms": bk}), flush=True)
        return 0
    make_env = env_factory(total_envs)
    envs_h = [make_env("host") for _ in range(replicas)]

    def host_step(i):
        envs_h[i % replicas].step(acts_host[i % 16])

    def host_finish():  # as in the device-resident leg: the closing event waits for what the steps started on side streams
        for env in envs_h:
            env.join()

    h.spin_up()
    for r in range(replicas):  # staggered like the device-resident leg
        for j in range(r * NUM_GENERATIONS // replicas):
            envs_h[r].step(acts_host[j % 16])
    for i in range(max(W, replicas * (NUM_GENERATIONS + 1))):
        host_step(i)
    h.timed(host_step, min(K, 500), host_finish)  # untimed rehearsal, as above
    # timed in blocks (same total K): a block far slower than the others points at the box (clock state, a descheduled
    # host thread), not at the path; reported beside the total
    nblk = 4 if K >= 400 else 1
    blocks = []
    sampler_e2e = ClockSampler(local_rank)
    sampler_e2e.start()
    base = 0
    for b in range(nblk):
        kb = K // nblk + (1 if b < K % nblk else 0)
        blocks.append(h.timed(lambda i, base=base: host_step(base + i), kb, host_finish)[0] / kb)
        base += kb
    clocks_e2e = sampler_e2e.stop()
    t_e2e = h.max_over_ranks(sum(bt * (K // nblk + (1 if b < K % nblk else 0)) for b, bt in enumerate(blocks)))
    h2d = count * N_IND * 2 * 4
    d2h = count * N_IND * 4 + (count * 4) / NUM_GENERATIONS
    del envs_h

    # ---------------- BASELINE config[4]: 512 envs per GPU (weak) and 4096 envs in total (strong) ----------------
    c5 = None
    if not args.no_c5:
        c5 = {}
        k5 = min(K, 40)
        for name, tot in (("weak_512_per_gpu", 512 * world), ("strong_4096_total", 4096)):
            b5, cnt5 = shard_range(tot, world, rank)
            _, a5 = actions(cnt5, 4, 100 + rank)
            reps5 = 1 if 2 * cnt5 * N_IND * 2560 > (126 << 20) else 2  # one population pair already exceeds L2
            envs5, t5, th5, _, _ = vec_value_leg(h, env_factory(tot), a5, k5, 3, reps5)
            how5 = envs5[0].collective if world > 1 else None
            del envs5, a5
            torch.cuda.empty_cache()
            c5[name] = {"workload": f"C5 vector env: {tot} envs over {world} GPU(s) ({cnt5} per GPU) x {N_IND} x {N_MARKERS}, "
                                    + (f"reward exchange: {how5}" if world > 1 else "single GPU: no exchange"),
                        "envs_total": tot, "envs_per_gpu": cnt5, "n_gpus": world, "steps": k5,
                        "value": tot * k5 / t5, "unit": UNIT, "ms_per_step": 1e3 * t5 / k5, "host_us_per_step": 1e6 * th5 / k5,
                        "scaling": "weak" if name.startswith("weak") else "strong",
                        "frac_of_hbm_peak_whole_step": B_ALG_CROSS * cnt5 * N_IND * N_MARKERS / (t5 / k5) / 1e9 / peak}

    # ---------------- the other named shapes (N = 1 only) ----------------
    configs = None
    if world == 1 and not args.no_legs:
        from bench_legs import run_legs  # C1 / C3 / C4: scripts-level detail kept out of this file

        configs = run_legs(h, torch, lib, local_rank, peak, measured_peaks()[0], int32_peak(), cpu=not args.no_cpu_baseline)

    # ---------------- cpu baseline (rank 0, N=1 only) ----------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        res = cpu_steps(2, 1, envs=ENVS_PER_GPU, min_seconds=10.0)
        redraw = cpu_steps(1, 0, envs=8, shared_masks=False)
        cpu = {"value": res["env_steps_per_sec"], "unit": UNIT, "cores": res["cores"], "kind": "port",
               "sample": f"{res['steps']} full steps of {res['envs']} envs x {N_IND} x {N_MARKERS} "
                         f"(cross + GEBV + reward) in {res['seconds']:.1f} s, C oracle with OpenMP, shared-mask port "
                         f"(2n masks drawn once per step, as the reference's vmap does)",
               "per_env_redraw_value": redraw["env_steps_per_sec"], "note": CPU_NOTE}

    if rank == 0:
        value = total_envs * K / t_value
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": 1e3 * t_value / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32 bit planes (cross) + int64 fixed point (GEBV)", "data": "synthetic",
            "config": dict(config_dict(world, replicas), **({"reward_exchange": reward_exchange} if world > 1 else {})),
            "offspring_markers_per_sec": value * N_IND * N_MARKERS,
            "host_us_per_step": 1e6 * t_host / K,
            "steady_state": None if steady is None else {
                "steps": steady[2], "value": total_envs * steady[2] / steady[0], "unit": UNIT, "ms_per_step": 1e3 * steady[0] / steady[2],
                "host_us_per_step": 1e6 * steady[1] / steady[2],
                "note": "the same device-resident loop over a long window: carries the full share of the crossover-mask "
                        "generation (Threefry, integer-issue bound) that runs beside the step kernels"},
            "clocks": clocks,
            "e2e": {"value": total_envs * K / t_e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": "VecBreedGym.step(int32 actions in pinned host memory) -> numpy GEBV / rewards, one sync per step",
                    "us_per_step_by_block": [round(1e6 * bt, 1) for bt in blocks], "clocks": clocks_e2e},
            "gpu_launches": int(launches),
            "gpu_launches_note": f"our kernels in {K} steps (counted over a repeat of the timed loop; "
                                 f"{int(launches_total)} over warm-up + rehearsal + timed)",
            "roofline": roofline,
            "kernels": kernels,
            "cpu_baseline": cpu,
            "c5": c5,
            "configs": configs,
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
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-c5", action="store_true", help="skip the C5 (512 envs/GPU, 4096 envs total) records")
    ap.add_argument("--no-legs", action="store_true", help="skip the C1 / C3 / C4 legs")
    ap.add_argument("--replicas", type=int, default=REPLICAS, help="independent workload copies stepped round-robin")
    ap.add_argument("--skip-e2e", action="store_true", help="diagnostics: only the device-resident value")
    ap.add_argument("--steady-steps", type=int, default=2000,
                    help="length of the extra long-run window reported as steady_state (0: skip; skipped when --steps is longer)")
    ap.add_argument("--stream-priority", type=int, default=-1,
                    help="priority of the CUDA stream the steps run on (default -1: above the library's mask side stream; 0: torch's current stream)")
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
