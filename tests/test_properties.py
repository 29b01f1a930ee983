"""Property tests (hypothesis, CPU only): the library's host-side PRNG / index math against the NumPy oracle, the two
oracles against each other, and algebraic properties of the restated algorithm, over generated inputs instead of
hand-picked ones."""
import numpy as np
import pytest
from hypothesis import given, settings
from hypothesis import strategies as st

from oracle import c_oracle as co
from oracle import chromax_ref as cr
from oracle import jax_prng as jp

LAYOUTS = st.sampled_from(["legacy", "partitionable"])
KEYS = st.tuples(st.integers(0, 2**32 - 1), st.integers(0, 2**32 - 1))
FAST = settings(max_examples=40, deadline=None, derandomize=True, database=None)


@FAST
@given(KEYS, st.integers(1, 300), LAYOUTS)
def test_random_bits_c_abi_equals_oracle(key, n, layout):
    from breedgym_b200 import _lib

    k = np.array(key, dtype=np.uint32)
    assert np.array_equal(_lib.random_bits(k, n, layout), jp.random_bits(k, n, layout))


@FAST
@given(KEYS, st.integers(1, 70), LAYOUTS)
def test_split_and_split_at_c_abi_equal_oracle(key, num, layout):
    from breedgym_b200 import _lib

    k = np.array(key, dtype=np.uint32)
    ref = jp.split(k, num, layout)
    assert np.array_equal(_lib.key_split(k, num, layout), ref)
    for i in {0, num // 2, num - 1}:
        assert np.array_equal(_lib.key_split_at(k, i, num, layout), ref[i])


@FAST
@given(KEYS, st.integers(1, 400), LAYOUTS)
def test_permutation_and_choice_host_equal_oracle(key, n, layout):
    from breedgym_b200 import jaxlike

    k = np.array(key, dtype=np.uint32)
    p = jaxlike.permutation(k, n, layout)
    assert np.array_equal(p, jp.permutation(k, n, layout))
    assert np.array_equal(np.sort(p), np.arange(n))
    d = max(1, n // 3)
    assert np.array_equal(jaxlike.choice_no_replace(k, n, d, layout), jp.choice_no_replace(k, n, d, layout))


@FAST
@given(st.lists(st.floats(0.0, 1.0, width=32), min_size=1, max_size=64))
def test_thresholds_are_the_integer_form_of_the_float_compare(rs):
    """u < r  <=>  (bits >> 9) < T(r) for the uniform jax builds from 23 random bits -- at the values around T."""
    from breedgym_b200 import _lib

    r = np.array(rs + [0.0, 1.0, 0.5, np.nextafter(np.float32(0.5), np.float32(1))], dtype=np.float32)
    T = _lib.thresholds(r)
    assert np.array_equal(T, jp.threshold_u32(r))
    for t, rv in zip(T, r):
        for b in {0, max(int(t) - 1, 0), min(int(t), 2**23 - 1), 2**23 - 1}:
            u = jp.bits_to_uniform(np.array([b << 9], dtype=np.uint32))[0]
            assert (u < rv) == (b < int(t))


@settings(max_examples=15, deadline=None, derandomize=True, database=None)
@given(st.integers(1, 90), st.integers(1, 6), st.integers(1, 5), st.integers(1, 3), st.integers(0, 2**31), LAYOUTS,
       st.sampled_from([("S1", 0.0), ("S2", 0.0), ("S2", 0.05)]))
def test_c_oracle_cross_equals_numpy_oracle(m, n_par, n_off, E, seed, layout, sched_mut):
    """The C restatement (one key for all envs, as the reference's vmap draws) == the NumPy restatement per env."""
    schedule, mutation = sched_mut
    rng = np.random.default_rng(seed)
    pops = rng.random((E, n_par, m, 2)) < 0.5
    acts = rng.integers(0, n_par, (E, n_off, 2))
    r = (rng.random(m) * 0.2).astype(np.float32)
    r[0] = 0.5
    key = jp.key(seed)
    got = co.cross_envs(pops, acts, r, key, mutation=mutation, schedule=schedule, layout=layout)
    shared = co.cross_envs(pops, acts, r, key, mutation=mutation, schedule=schedule, layout=layout, shared_masks=True)
    assert np.array_equal(got, shared)  # masks drawn once per step == masks re-drawn per env
    for e in range(E):
        ref = cr.cross(pops[e][acts[e]], r, key, mutation, schedule, layout)
        assert np.array_equal(got[e], ref)
        if mutation == 0.0:  # every offspring allele comes from the corresponding parent
            for i in range(n_off):
                for p in range(2):
                    par = pops[e][acts[e, i, p]]
                    assert np.all((ref[i, :, p] == par[:, 0]) | (ref[i, :, p] == par[:, 1]))


@FAST
@given(st.integers(1, 40), st.integers(1, 200), st.integers(1, 4), st.integers(0, 2**31))
def test_gebv_is_linear_in_the_effects_and_additive_over_haplotypes(n, m, T, seed):
    rng = np.random.default_rng(seed)
    pop = rng.random((n, m, 2)) < 0.4
    # (dyadic effects: w1 + 2 w2 is exact in float32, which the oracle rounds the effects to)
    w1, w2 = rng.integers(-64, 64, (m, T)) / 16.0, rng.integers(-64, 64, (m, T)) / 16.0
    g = cr.gebv(pop, w1 + 2 * w2)
    assert np.allclose(g, cr.gebv(pop, w1) + 2 * cr.gebv(pop, w2), rtol=1e-9, atol=1e-9)
    h0, h1 = pop.copy(), pop.copy()
    h0[..., 1] = False
    h1[..., 0] = False
    assert np.allclose(g, cr.gebv(h0, w1 + 2 * w2) + cr.gebv(h1, w1 + 2 * w2), rtol=1e-9, atol=1e-9)


@FAST
@given(st.lists(st.integers(0, 5), min_size=1, max_size=30), st.integers(1, 60))
def test_repeat_total_host_equals_oracle(reps, total):
    from breedgym_b200 import jaxlike

    x = np.arange(2 * len(reps)).reshape(len(reps), 2)
    reps = np.array(reps)
    if reps.sum() == 0:
        reps[-1] = 1
    assert np.array_equal(jaxlike.repeat_total(x, reps, total), jp.repeat_total(x, reps, total))
    assert len(jp.repeat_total(x, reps, total)) == total
