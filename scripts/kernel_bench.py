#!/usr/bin/env python
"""Per-kernel timings (CUDA events, L2 flushed between launches) for the BASELINE shapes.

    python scripts/kernel_bench.py [--config C2|C3|C4s|C5] [--reps 30]

Prints one JSON object per kernel: ms per launch, algorithmic bytes, fraction of the HBM peak.
"""
from __future__ import annotations

import argparse
import json
import sys
from pathlib import Path

import numpy as np
import pandas as pd
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from breedgym_b200 import _lib  # noqa: E402
from breedgym_b200.simulator import Simulator  # noqa: E402

CONFIGS = {
    # name: (envs, individuals, markers, traits)
    "C1": (1, 370, 10_000, 1),
    "C2": (64, 370, 10_000, 1),
    "C3": (1, 1000, 100_002, 1),
    "C4s": (1, 2000, 1_000_000, 16),   # C4 scaled to 2k offspring (same kernels, same row shape)
    "C4": (1, 10_000, 1_000_000, 16),  # BASELINE config 4 at full size: 10k offspring of 1000 parents x 1M markers, 16 traits
    "C5": (512, 370, 10_000, 1),       # one GPU's share of 4096 envs / 8
}


def peak():
    p = ROOT / "MEASURED_PEAKS.json"
    return float(json.loads(p.read_text())["hbm_gbs"]) if p.exists() else 6650.0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="C2")
    ap.add_argument("--reps", type=int, default=30)
    ap.add_argument("--layout", default="legacy")
    ap.add_argument("--no-flush", action="store_true")
    ap.add_argument("--no-sync", action="store_true", help="do not synchronise between timed launches")
    ap.add_argument("--only", default="", help="comma-separated kernel names to time (default: all)")
    args = ap.parse_args()
    E, n, m, T = CONFIGS[args.config]
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(0)
    if args.config in ("C1", "C2", "C5"):
        gmap = ROOT / "breedgym_b200" / "data" / "small_genetic_map.txt"
        sim = Simulator(genetic_map=gmap, trait_names=["Yield"], device=0, seed=0, rng_layout=args.layout)
    else:
        n_chr = 21 if args.config == "C3" else 10
        df = pd.DataFrame({"CHR.PHYS": np.arange(m) // (m // n_chr + 1), "RecombRate": np.full(m, 1.5e-3)})
        for t in range(T):
            df[f"t{t}"] = rng.standard_normal(m).astype(np.float32)
        sim = Simulator(genetic_map=df, device=0, seed=0, rng_layout=args.layout)
    lib = _lib.load()
    Wpad = sim.words_per_row
    n_src = n if args.config not in ("C4s", "C4") else 1000
    lead = (E, n_src) if E > 1 else (n_src,)
    pop = torch.randint(-2**31, 2**31 - 1, (*lead, 2, Wpad), dtype=torch.int32, device=dev)
    tail = m % 32
    if Wpad * 32 > m:  # keep padding bits zero like a real population
        pop[..., (m // 32) + (1 if tail else 0):] = 0
        if tail:
            pop[..., m // 32] &= (1 << tail) - 1
    acts = torch.randint(0, n_src, (E, n, 2) if E > 1 else (n, 2), dtype=torch.int32, device=dev)
    out = torch.empty(((E, n) if E > 1 else (n,)) + (2, Wpad), dtype=torch.int32, device=dev)
    mask = torch.empty((2 * n, Wpad), dtype=torch.int32, device=dev)
    gebv = torch.empty((E * n, T), dtype=torch.float32, device=dev)
    key = np.array([1, 2], dtype=np.uint32)
    sp = sim._stream()
    stream = torch.cuda.current_stream(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    lay, sch = sim._layout(), sim._schedule()
    om = E * n * m

    def time_it(fn):
        evs = []
        for it in range(3 + args.reps):
            if not args.no_flush:
                flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            fn()
            b.record(stream)
            if not args.no_sync:
                torch.cuda.synchronize()
            evs.append((a, b))
        torch.cuda.synchronize()
        return sum(a.elapsed_time(b) for a, b in evs[3:]) / args.reps

    pk = peak()
    res = {}

    only = {x for x in args.only.split(",") if x}

    def want(name):
        return not only or name in only

    def report(name, ms, alg_bytes=None, extra=None):
        r = {"kernel": name, "config": args.config, "ms": round(ms, 5)}
        if alg_bytes:
            r["achieved_gbs"] = round(alg_bytes / (ms * 1e-3) / 1e9, 1)
            r["frac_hbm_peak"] = round(r["achieved_gbs"] / pk, 4)
        if extra:
            r.update(extra)
        res[name] = r
        print(json.dumps(r), flush=True)

    if want("meiosis_masks"):
        ms = time_it(lambda: _lib.check(lib.bg_meiosis_masks(sim._engine, mask.data_ptr(), 2 * n, _lib.nptr(key), lay, sch, sp)))
        report("meiosis_masks", ms, None, {"gdraws_per_s": round(2 * n * m / (ms * 1e-3) / 1e9, 2)})
    else:
        _lib.check(lib.bg_meiosis_masks(sim._engine, mask.data_ptr(), 2 * n, _lib.nptr(key), lay, sch, sp))
    if E > 1 and want("blend_envs"):
        ms = time_it(lambda: _lib.check(lib.bg_blend_envs(sim._engine, pop.data_ptr(), acts.data_ptr(), mask.data_ptr(), None,
                                                         out.data_ptr(), E, n_src, n, sp)))
        report("blend_envs", ms, 0.75 * om)
    if want("cross_total"):
        ms = time_it(lambda: _lib.check(lib.bg_cross(sim._engine, pop.data_ptr(), acts.data_ptr(), out.data_ptr(), E, n_src, n,
                                                    _lib.nptr(key), lay, sch, sp)))
        report("cross_total", ms, 0.75 * om, {"offspring_markers_per_s": round(om / (ms * 1e-3) / 1e9, 2)})
    else:
        _lib.check(lib.bg_cross(sim._engine, pop.data_ptr(), acts.data_ptr(), out.data_ptr(), E, n_src, n, _lib.nptr(key), lay, sch, sp))
    if E > 1 and want("cross_gebv_fused"):
        fn = lambda: _lib.check(lib.bg_cross_gebv(sim._engine, pop.data_ptr(), acts.data_ptr(), out.data_ptr(), E, n_src, n,
                                                  _lib.nptr(key), lay, sch, gebv.data_ptr(), sp))
        fn()  # masks of this key land in a slot: the timed calls launch the fused kernel only
        ms = time_it(fn)
        report("cross_gebv_fused", ms, 0.75 * om, {"offspring_markers_per_s": round(om / (ms * 1e-3) / 1e9, 2)})
    for algo, name in ((1, "gebv_direct"), (2, "gebv_lut"), (3, "gebv_tcgen05_tmemA")):
        if algo == 1 and om > 5e8:
            continue
        if algo == 2 and (T > 4 or m > 200_000):
            continue
        if not want(name):
            continue
        ms = time_it(lambda: _lib.check(lib.bg_gebv_algo(sim._engine, out.data_ptr(), E * n, gebv.data_ptr(), algo, sp)))
        report(name, ms, 0.25 * om, {"tflops_int8_equiv": round(2 * om * 8 * T / (ms * 1e-3) / 1e12, 2)})
    return 0


if __name__ == "__main__":
    sys.exit(main())
