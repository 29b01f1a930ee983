"""The other named shapes of BASELINE.json beside the vector env (bench.py's `configs` key, N = 1 only):

  C1  BreedGym sample-data shape: ONE env, 370 x 10 000 markers, 10 generations, through the Gym API
      (reference: README.md:48 / scripts/time_cross.py:11-42 time exactly `simulator.cross` on this shape)
  C3  wheat-scale genome (scripts/time_wheat.py:65-97): 1000 individuals x 100 002 markers x 21 chromosomes,
      RecombRate 1.5e-3, one trait with N(0,1) effects: unique-key cross + single-trait GEBV
  C4  synthetic large cross: 10 000 offspring of 1000 parents x 1 000 000 markers, 16 traits: unique-key cross +
      the 16-trait GEBV on the tensor cores

Every leg reports offspring-markers/s (BASELINE.json's cross metric), its kernels with the roofline that bounds each
(integer issue rate for the Threefry-bound unique-key cross, HBM / tensor pipe for the GEBV) and the C oracle timed on
the host cores on a bounded sample of the same workload.
"""
from __future__ import annotations

import os
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent


def _event_ms(torch, stream, fn, reps, flush):
    """Average device time of fn() over `reps` launches, L2 flushed before each (CUDA events on the launching stream)."""
    tot = 0.0
    for it in range(2 + reps):
        flush()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        fn()
        b.record(stream)
        torch.cuda.synchronize()
        if it >= 2:
            tot += a.elapsed_time(b)
    return tot / reps


def _int_roof(gdraws_per_s, ip):
    if not ip or not ip.get("int32_gops") or not ip.get("threefry_int_ops_per_draw"):
        return None
    return {"bound": "int32 issue (Threefry)", "achieved_gops": gdraws_per_s * ip["threefry_int_ops_per_draw"],
            "peak_gops": ip["int32_gops"], "frac": gdraws_per_s * ip["threefry_int_ops_per_draw"] / ip["int32_gops"],
            "int_ops_per_draw": ip["threefry_int_ops_per_draw"], "peak_source": "profiles/int32_peak.json (scripts/int32_peak.cu)"}


def _cpu_cross_gebv(pop_bool, pairs, r, eff, min_seconds=3.0):
    """C oracle: unique-key cross (E = 1) + GEBV of the offspring, all host cores; returns offspring-markers/s."""
    from oracle import c_oracle as co
    from oracle import jax_prng as jp

    co.set_threads(os.cpu_count() or 1)
    key = jp.key(7)
    n, m = pairs.shape[0], pop_bool.shape[1]
    tot, steps = 0.0, 0
    while tot < min_seconds or steps < 2:
        ks = jp.split(key, 2)
        key, k = ks[0], ks[1]
        t0 = time.perf_counter()
        off = co.cross_envs(pop_bool[None], pairs[None].astype(np.int32), r, k)
        g = co.gebv(off[0], eff)
        _ = float(g.mean())
        dt = time.perf_counter() - t0
        if steps > 0 or min_seconds == 0:  # first call: page faults of the output buffers
            tot += dt
        steps += 1
    timed = steps - 1 if min_seconds else steps
    return {"value": n * m * timed / tot, "unit": "offspring-markers/s", "cores": co.num_threads(), "kind": "port",
            "seconds": tot, "steps": timed}


def _random_words(torch, sim, rows, dev, seed):
    """Bernoulli(0.5) population generated directly as bit planes (padding bits zero, like a packed real population)."""
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    m, W = sim.n_markers, sim.words_per_row
    pop = torch.randint(-2**31, 2**31 - 1, (rows, 2, W), dtype=torch.int32, device=dev, generator=g)
    full, tail = m // 32, m % 32
    pop[..., full + (1 if tail else 0):] = 0
    if tail:
        pop[..., full] &= (1 << tail) - 1
    return pop


def leg_c1(h, torch, lib, device, peak, ip, cpu):
    from breedgym_b200 import _lib
    from breedgym_b200.breedgym import BreedGym

    n, m = 370, 10_000
    germ = np.random.default_rng(0).random((n, m, 2)) < 0.5
    gmap = ROOT / "breedgym_b200" / "data" / "small_genetic_map.txt"
    env = BreedGym(initial_population=germ, genetic_map=gmap, trait_names=["Yield"], device=device)
    rng = np.random.default_rng(1)
    acts = [rng.integers(0, n, (n, 2)) for _ in range(8)]

    def episode():
        env.reset(seed=7)
        for g in range(10):
            env.step(acts[g % 8])

    for _ in range(10):
        episode()
    torch.cuda.synchronize()
    eps = 50
    t0 = time.perf_counter()
    for _ in range(eps):
        episode()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    steps = 10 * eps
    # kernels of the step on their own
    sim = env.simulator
    pop = sim.as_packed(germ)
    a = torch.from_numpy(acts[0].astype(np.int32)).to(sim.device)
    out = sim._empty_words(n)
    gebv = torch.empty((n, 1), dtype=torch.float32, device=sim.device)
    key = np.array([0, 12345], dtype=np.uint32)
    stream = torch.cuda.current_stream(sim.device)
    sp = sim._stream()
    ms_cross = _event_ms(torch, stream, lambda: _lib.check(lib.bg_cross(sim._engine, pop.words.data_ptr(), a.data_ptr(), out.data_ptr(), 1, n, n,
                                                                        _lib.nptr(key), 0, 2, sp)), 30, h.flush_l2)
    ms_gebv = _event_ms(torch, stream, lambda: _lib.check(lib.bg_gebv(sim._engine, out.data_ptr(), n, gebv.data_ptr(), sp)), 30, h.flush_l2)
    om = n * m
    res = {
        "workload": "C1 single BreedGym env through the Gym API: 370 x 10000 markers, 10 generations, 370 random crosses per step, Yield",
        "metric": "offspring_markers_per_sec", "unit": "offspring-markers/s",
        "value": om * steps / dt, "env_steps_per_sec": steps / dt, "us_per_step": 1e6 * dt / steps,
        "api": "BreedGym.step(int array n x 2) -> (PackedPopulation, reward, ter, tru, {'GEBV': DataFrame}); one sync per step",
        "h2d_bytes_per_step": n * 2 * 4, "d2h_bytes_per_step": n * 4,
        "kernels": {
            "meiosis_rows (unique-key cross)": {"ms": ms_cross, "offspring_markers_per_sec": om / (ms_cross * 1e-3),
                                                "frac_of_hbm_peak": 0.75 * om / (ms_cross * 1e-3) / 1e9 / peak,
                                                "int_roofline": _int_roof(2 * om / (ms_cross * 1e-3) / 1e9, ip)},
            "gebv_tc2 (1 trait)": {"ms": ms_gebv, "frac_of_hbm_peak": 0.25 * om / (ms_gebv * 1e-3) / 1e9 / peak},
        },
        "roofline": {"bound": "host latency (a 3.7 M offspring-marker step is ~10 us of kernels behind ~80 us of launch / copy / sync)",
                     "kernel_us": 1e3 * (ms_cross + ms_gebv)},
    }
    if cpu:
        from oracle import chromax_ref as cr

        g = cr.read_genetic_map(gmap)
        c = _cpu_cross_gebv(germ, acts[0], cr.recombination_vector(g), cr.marker_effects(g, ["Yield"]), 3.0)
        c["sample"] = f"{c['steps']} steps of 370 crosses x 10000 markers (cross + GEBV) in {c['seconds']:.1f} s, C oracle with OpenMP"
        res["cpu_baseline"] = c
    return res


def _synthetic_sim(device, m, n_chr, n_traits, seed):
    import pandas as pd

    from breedgym_b200.simulator import Simulator

    rng = np.random.default_rng(seed)
    per = -(-m // n_chr)
    df = pd.DataFrame({"CHR.PHYS": (np.arange(m) // per).astype(np.int32), "RecombRate": np.full(m, 1.5e-3, dtype=np.float32)})
    for t in range(n_traits):
        df[f"trait{t}"] = rng.standard_normal(m).astype(np.float32)
    return Simulator(genetic_map=df, device=device, seed=7)


def _cross_gebv_leg(h, torch, lib, device, peak, peaks, ip, cpu, name, workload, m, n_chr, T, n_par, n_off, seeds, cpu_sample):
    from breedgym_b200 import _lib
    from breedgym_b200.population import PackedPopulation

    sim = _synthetic_sim(device, m, n_chr, T, seeds[0])
    dev = sim.device
    pop = _random_words(torch, sim, n_par, dev, seeds[1])
    pairs = np.random.default_rng(seeds[2]).integers(0, n_par, (n_off, 2)).astype(np.int32)
    a = torch.from_numpy(pairs).to(dev)
    out = sim._empty_words(n_off)
    gebv = torch.empty((n_off, T), dtype=torch.float32, device=dev)
    key = np.array([0, 12345], dtype=np.uint32)
    stream = torch.cuda.current_stream(dev)
    sp = sim._stream()
    reps = 5 if n_off * m > 2e9 else 20
    ms_cross = _event_ms(torch, stream, lambda: _lib.check(lib.bg_cross(sim._engine, pop.data_ptr(), a.data_ptr(), out.data_ptr(), 1, n_par, n_off,
                                                                        _lib.nptr(key), 0, 2, sp)), reps, h.flush_l2)
    ms_gebv = _event_ms(torch, stream, lambda: _lib.check(lib.bg_gebv(sim._engine, out.data_ptr(), n_off, gebv.data_ptr(), sp)), reps, h.flush_l2)

    def both():
        _lib.check(lib.bg_cross(sim._engine, pop.data_ptr(), a.data_ptr(), out.data_ptr(), 1, n_par, n_off, _lib.nptr(key), 0, 2, sp))
        _lib.check(lib.bg_gebv(sim._engine, out.data_ptr(), n_off, gebv.data_ptr(), sp))

    ms_step = _event_ms(torch, stream, both, reps, h.flush_l2)
    om = n_off * m
    kern = {
        "meiosis_rows (unique-key cross)": {"ms": ms_cross, "offspring_markers_per_sec": om / (ms_cross * 1e-3),
                                            "frac_of_hbm_peak": 0.75 * om / (ms_cross * 1e-3) / 1e9 / peak,
                                            "int_roofline": _int_roof(2 * om / (ms_cross * 1e-3) / 1e9, ip)},
    }
    gk = {"ms": ms_gebv, "frac_of_hbm_peak": 0.25 * om / (ms_gebv * 1e-3) / 1e9 / peak}
    if T > 1:
        digits = int(lib.bg_gebv_digits(sim._engine)) if hasattr(lib, "bg_gebv_digits") else 8
        flops = 2.0 * om * T
        int8_ops = 2.0 * om * T * digits  # what the tensor pipe executes: `digits` int8 digit columns per trait
        i8_peak = 2.0 * float(peaks.get("bf16_tflops", 1657.0))
        gk.update({"useful_tflops": flops / (ms_gebv * 1e-3) / 1e12, "int8_tops_executed": int8_ops / (ms_gebv * 1e-3) / 1e12,
                   "digits_per_effect": digits,
                   "tensor_roofline": {"bound": "tensor", "achieved": int8_ops / (ms_gebv * 1e-3) / 1e12, "peak": i8_peak, "unit": "TOP/s (int8)",
                                       "frac": int8_ops / (ms_gebv * 1e-3) / 1e12 / i8_peak,
                                       "peak_source": "2 x the measured dense bf16 rate of MEASURED_PEAKS.json (int8 = 2 x bf16 on sm_100)"}})
    kern[f"gebv_tc2 ({T} trait{'s' if T > 1 else ''})"] = gk
    res = {"workload": workload, "metric": "offspring_markers_per_sec", "unit": "offspring-markers/s",
           "value": om / (ms_step * 1e-3), "ms_per_step": ms_step, "kernels": kern,
           "roofline": {"kernel": "meiosis_rows_kernel", "bound": "hbm", "achieved": 0.75 * om / (ms_cross * 1e-3) / 1e9, "peak": peak,
                        "unit": "GB/s", "frac": 0.75 * om / (ms_cross * 1e-3) / 1e9 / peak,
                        "note": "small by construction: the unique-key cross is bound by the integer pipes (2 Threefry draws per "
                                "offspring-marker), see kernels[...].int_roofline"}}
    if cpu:
        # bounded sample: `cpu_sample` offspring whose parents are the first 128 individuals
        n_s, n_p = cpu_sample, min(128, n_par)
        host_pop = np.asarray(PackedPopulation(sim, pop[:n_p].contiguous()))
        ps = np.random.default_rng(seeds[2]).integers(0, n_p, (n_s, 2)).astype(np.int32)
        c = _cpu_cross_gebv(host_pop, ps, sim.recombination_vec, sim.GEBV_model.marker_effects, 3.0)
        c["sample"] = (f"{c['steps']} steps of {n_s} of the {n_off} crosses x {m} markers (cross + {T}-trait GEBV) in {c['seconds']:.1f} s, "
                       f"C oracle with OpenMP")
        res["cpu_baseline"] = c
    del pop, out, gebv, sim
    torch.cuda.empty_cache()
    return res


def run_legs(h, torch, lib, device, peak, peaks, ip, cpu=True):
    legs = {}
    legs["C1"] = leg_c1(h, torch, lib, device, peak, ip, cpu)
    legs["C3"] = _cross_gebv_leg(h, torch, lib, device, peak, peaks, ip, cpu, "C3",
                                 "C3 wheat-scale genome (time_wheat.py shape): 1000 x 100002 markers x 21 chromosomes, 1000 crosses, "
                                 "single-trait GEBV", 100_002, 21, 1, 1000, 1000, (2, 20, 3), 256)
    legs["C4"] = _cross_gebv_leg(h, torch, lib, device, peak, peaks, ip, cpu, "C4",
                                 "C4 synthetic large cross: 10000 offspring of 1000 parents x 1000000 markers, 16-trait GEBV on the "
                                 "tensor cores", 1_000_000, 10, 16, 1000, 10_000, (4, 40, 5), 128)
    return legs
