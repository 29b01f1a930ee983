#!/usr/bin/env python
"""Randomised shapes: the fused cross + GEBV step kernels (one CTA per (tile, K range) AND the persistent one with the
dynamic work queue) against the blend + GEBV kernels, bit for bit.

    python scripts/fuzz_fused.py [N_CASES] [SEED]

`run_cases` is also what tests/test_gpu_kernels.py::test_fuzz_fused_step_kernels calls.
"""
import sys
from pathlib import Path

import numpy as np
import pandas as pd
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from breedgym_b200 import _lib  # noqa: E402
from breedgym_b200.simulator import Simulator  # noqa: E402


def run_cases(n_cases: int, seed: int, verbose: bool = True) -> int:
    """Number of mismatching cases (0 = all three paths agree on every case)."""
    rng = np.random.default_rng(seed)
    dev = torch.device("cuda", 0)
    lib = _lib.load()
    bad = 0
    for case in range(n_cases):
        m = int(rng.choice([1, 31, 33, 127, 129, 1000, 4097, 10000, 33333, 100002]))
        T = int(rng.choice([1, 1, 1, 2, 3, 7, 16, 24]))
        E = int(rng.choice([2, 3, 5, 16, 31, 32, 33, 64, 100, 128, 200]))
        n_src = int(rng.integers(1, 60))
        n = int(rng.integers(1, 80))
        if E * n * (m / 32 + 32) * 8 > 2e9:
            continue
        df = pd.DataFrame({"CHR.PHYS": np.arange(m) // (m // 3 + 1), "RecombRate": rng.random(m) * 0.01})
        for t in range(T):
            df[f"t{t}"] = rng.standard_normal(m).astype(np.float32)
        sim = Simulator(genetic_map=df, device=0, seed=case, rng_layout=str(rng.choice(["legacy", "partitionable"])))
        W = sim.words_per_row
        pop = torch.randint(-2**31, 2**31 - 1, (E, n_src, 2, W), dtype=torch.int32, device=dev)
        full, tail = m // 32, m % 32
        pop[..., full + (1 if tail else 0):] = 0
        if tail:
            pop[..., full] &= (1 << tail) - 1
        acts = torch.from_numpy(rng.integers(-2, n_src + 2, (E, n, 2)).astype(np.int32)).to(dev)  # incl. wrap / clamp cases
        key = np.array(rng.integers(0, 2**32, 2), dtype=np.uint32)
        res = []
        for fuse, dyn in ((1, 0), (1, 1), (0, 0)):
            out = torch.zeros((E, n, 2, W), dtype=torch.int32, device=dev)
            gebv = torch.zeros((E, n, T), dtype=torch.float32, device=dev)
            sim.set_option("fuse", fuse)
            sim.set_option("fused_dyn", dyn)
            try:
                _lib.check(lib.bg_cross_gebv(sim._engine, pop.data_ptr(), acts.data_ptr(), out.data_ptr(), E, n_src, n, _lib.nptr(key),
                                             sim._layout(), sim._schedule(), gebv.data_ptr(), sim._stream()))
            finally:
                sim.set_option("fuse", 1)
                sim.set_option("fused_dyn", -1)
            torch.cuda.synchronize()
            res.append((out, gebv))
        ok = all(torch.equal(res[0][0], r[0]) and torch.equal(res[0][1], r[1]) for r in res[1:])
        bad += not ok
        if verbose:
            print(f"case {case:3d} m={m:6d} T={T:2d} E={E:3d} n_src={n_src:2d} n={n:2d} {'ok' if ok else 'MISMATCH'}", flush=True)
    return bad


if __name__ == "__main__":
    n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    bad = run_cases(n_cases, int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    print("fuzz:", "all ok" if bad == 0 else f"{bad} MISMATCHES")
    sys.exit(1 if bad else 0)
