"""Generate tests/golden/hotpath_golden.npz with the CPU oracle (oracle/).

No genotype-level golden vectors exist in the reference and
jax/chromax cannot be imported here, so these fixtures pin the ORACLE's output
(self-pins): they guard the oracle against regressions and give the CUDA path a
committed target that does not depend on oracle code at test time.

    python tests/golden/make_golden.py
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from oracle import chromax_ref as cr  # noqa: E402
from oracle import jax_prng as jp  # noqa: E402


def main():
    rng = np.random.default_rng(20260101)
    out = {}
    n_src, n, m, T = 6, 5, 77, 3
    pop = rng.random((n_src, m, 2)) < 0.5
    pairs = rng.integers(0, n_src, (n, 2)).astype(np.int32)
    r = (rng.random(m) * 0.2).astype(np.float32)
    r[[0, 19, 40, 64]] = 0.5
    r[[5, 6, 33]] = 0.0
    eff = rng.standard_normal((m, T)).astype(np.float32) * 3
    key = jp.key(1234)
    out.update(pop=pop, pairs=pairs, r=r, eff=eff, key=key)
    for lay in jp.LAYOUTS:
        for sch in cr.SCHEDULES:
            off = cr.cross(pop[pairs], r, key, 0.0, sch, lay)
            out[f"cross_{lay}_{sch}"] = off
            out[f"gebv_{lay}_{sch}"] = cr.gebv(off, eff)
        out[f"cross_mut_{lay}"] = cr.cross(pop[pairs], r, key, 0.05, "S2", lay)
        out[f"dh_{lay}"] = cr.double_haploid(pop, r, key, 3, 0.0, "S2", lay)
        # reset selection: E=3 envs, N=50 -> n=20, and one N large enough for two shuffle rounds
        rk, _, idx = cr.vec_reset(np.zeros((50, 1, 2), bool), 20, 3, jp.key(7), lay)
        out[f"reset_idx_{lay}"] = idx
        out[f"reset_key_{lay}"] = rk
        out[f"perm2000_{lay}"] = jp.permutation(jp.key(11), 2000, lay)
        out[f"split5_{lay}"] = jp.split(jp.key(99), 5, lay)
        out[f"bits9_{lay}"] = jp.random_bits(jp.key(99), 9, lay)
    np.savez_compressed(Path(__file__).with_name("hotpath_golden.npz"), **out)
    print("wrote", Path(__file__).with_name("hotpath_golden.npz"))


if __name__ == "__main__":
    main()
