#!/usr/bin/env python
"""Pin the oracle to REAL jax + chromax (run on any machine that has both; no GPU needed).

    pip install jax chromax
    python scripts/pin_with_chromax.py [--out tests/golden/chromax_pins.npz]

Neither package is installable on the build image, so the CPU oracle (`oracle/`) is pinned only to Random123 /
jax-documented PRNG vectors and to its own goldens.  This script closes that gap:

1. it runs the real `chromax.Simulator` / `chromax.functional` / `jax.random` on small seeded inputs -- cross,
   double_haploid, the recombination vector of an `RecombRate` map and of a `cM` map, `Simulator(seed).random_key`
   right after construction, `VecBreedGym.reset`-style `choice(replace=False)` indices, `select` ordering -- and
   writes the genotype-level results (with the jax / chromax versions) to an .npz that can be committed under
   tests/golden/ (tests/test_oracle.py::test_chromax_pins picks it up when present);
2. it replays the same inputs through this repo's oracle in every {rng_layout} x {key_schedule} combination and
   reports which one reproduces the real results bit for bit -- the defaults of `breedgym_b200.Simulator`
   (`rng_layout="legacy"`, `key_schedule="S2"`) should be set to that combination.
"""
import argparse
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=str(ROOT / "tests" / "golden" / "chromax_pins.npz"))
    args = ap.parse_args()
    try:
        import chromax
        import jax
        import jax.numpy as jnp
        from chromax import Simulator, functional
    except Exception as e:  # pragma: no cover - needs the real packages
        raise SystemExit(f"this script needs real jax + chromax ({e}); nothing written")
    import pandas as pd

    from oracle import chromax_ref as cr
    from oracle import jax_prng as jp

    data = ROOT / "breedgym_b200" / "data"
    gmap_r = data / "sample_with_r_genetic_map.txt"   # RecombRate column
    gmap_cm = data / "sample_genetic_map.txt"          # cM column only
    germ = np.load(data / "sample_geno.npy")
    out = {"jax_version": np.array(jax.__version__), "chromax_version": np.array(getattr(chromax, "__version__", "unknown")),
           "threefry_partitionable": np.array(bool(jax.config.jax_threefry_partitionable))}

    # ---- PRNG layout
    key = jax.random.key(1234)
    out["key_data"] = np.asarray(jax.random.key_data(key))
    out["split5"] = np.asarray(jax.random.key_data(jax.random.split(jax.random.key(99), 5)))
    out["bits9"] = np.asarray(jax.random.bits(jax.random.key(99), (9,), dtype=jnp.uint32))
    out["perm2000"] = np.asarray(jax.random.permutation(jax.random.key(11), 2000))

    # ---- chromax on the RecombRate map
    sim = Simulator(genetic_map=gmap_r, trait_names=["Yield"], seed=7)
    out["recombination_vec_r"] = np.asarray(sim.recombination_vec)
    out["random_key_after_init_seed7"] = np.asarray(jax.random.key_data(sim.random_key))
    rng = np.random.default_rng(20260101)
    pairs = rng.integers(0, len(germ), (12, 2))
    parents = jnp.asarray(germ)[pairs]
    sim.set_seed(3)
    out["pairs"] = pairs
    out["cross_seed3"] = np.asarray(sim.cross(parents))
    out["cross_seed3_second_call"] = np.asarray(sim.cross(parents))
    sim.set_seed(5)
    out["dh_seed5_n3"] = np.asarray(sim.double_haploid(jnp.asarray(germ[:6]), n_offspring=3))
    k = jax.random.key(42)
    out["functional_cross_key42"] = np.asarray(functional.cross(parents, sim.recombination_vec, k))
    out["gebv_model"] = np.asarray(sim.GEBV_model(jnp.asarray(germ)))
    sel, idx = sim.select(jnp.asarray(germ), k=10)
    out["select_idx_k10"] = np.asarray(idx)
    # ---- the cM map (Haldane conversion inside chromax)
    sim_cm = Simulator(genetic_map=gmap_cm, seed=7)
    out["recombination_vec_cm"] = np.asarray(sim_cm.recombination_vec)
    # ---- VecBreedGym.reset's draw: vmap(choice(replace=False)) over split(key, E + 1)[1:]
    keys = jax.random.split(jax.random.key(7), 4)
    out["reset_keys"] = np.asarray(jax.random.key_data(keys))
    out["reset_idx"] = np.stack([np.asarray(jax.random.choice(kk, 50, shape=(20,), replace=False)) for kk in keys[1:]])
    np.savez_compressed(args.out, **out)
    print("wrote", args.out)

    # ---- which oracle configuration reproduces the real thing?
    r = np.asarray(sim.recombination_vec)
    ok_r = np.array_equal(cr.recombination_vector(pd.read_table(gmap_r, sep="\t")), r)
    print(f"recombination vector (RecombRate map): oracle {'==' if ok_r else '!='} chromax")
    ok_cm = np.allclose(cr.recombination_vector(pd.read_table(gmap_cm, sep='\t')), np.asarray(sim_cm.recombination_vec), rtol=0, atol=0)
    print(f"recombination vector (cM map, Haldane): oracle {'==' if ok_cm else '!='} chromax")
    for layout in jp.LAYOUTS:
        lay_ok = (np.array_equal(jp.split(jp.key(99), 5, layout), out["split5"]) and np.array_equal(jp.random_bits(jp.key(99), 9, layout), out["bits9"])
                  and np.array_equal(jp.permutation(jp.key(11), 2000, layout), out["perm2000"]))
        print(f"layout {layout:14s}: split / bits / permutation {'match' if lay_ok else 'differ'}")
        for schedule in cr.SCHEDULES:
            got = cr.cross(germ[pairs], r, jp.key(42), 0.0, schedule, layout)
            print(f"  functional.cross, schedule {schedule}: {'BIT-EXACT' if np.array_equal(got, out['functional_cross_key42']) else 'differs'}")
            osim = cr.OracleSimulator(r, np.asarray(sim.GEBV_model.marker_effects), seed=3, schedule=schedule, layout=layout)
            osim.set_seed(3)
            print(f"  Simulator.cross after set_seed(3): {'BIT-EXACT' if np.array_equal(osim.cross(germ[pairs]), out['cross_seed3']) else 'differs'}")


if __name__ == "__main__":
    main()
