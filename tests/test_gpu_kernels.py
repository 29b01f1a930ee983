"""Parity of the CUDA kernels (through the C ABI) against the CPU oracle.  Needs a GPU.

Bar: bit-exact for genotypes / masks / indices; GEBV within 1e-5 relative of the float64
oracle (the kernels actually land within one float32 ulp, asserted too).
"""
import ctypes

import numpy as np
import pandas as pd
import pytest

from oracle import c_oracle as co
from oracle import chromax_ref as cr
from oracle import jax_prng as jp

pytestmark = pytest.mark.gpu

LAYOUTS = ("legacy", "partitionable")
SCHEDULES = ("S1", "S2")
GEBV_RTOL = 1e-5  # north_star tolerance (fp32)


def make_map(m, n_chr=3, T=1, seed=0, r_kind="random"):
    rng = np.random.default_rng(seed)
    chrom = np.sort(rng.integers(0, n_chr, m)) if m >= n_chr else np.zeros(m, int)
    if r_kind == "random":
        r = rng.random(m) * 0.2
        r[rng.random(m) < 0.2] = 0.0
    elif r_kind == "zero":
        r = np.zeros(m)
    elif r_kind == "half":
        r = np.full(m, 0.5)
    elif r_kind == "tiny":
        r = np.full(m, 3e-7)
    else:
        raise ValueError(r_kind)
    df = pd.DataFrame({"CHR.PHYS": chrom, "RecombRate": r})
    for t in range(T):
        df[f"trait{t}"] = (rng.standard_normal(m) * 3).astype(np.float32)
    return df


def make_sim(df, **kw):
    from breedgym_b200.simulator import Simulator

    sim = Simulator(genetic_map=df, seed=0, device=0, **kw)
    assert np.array_equal(sim.recombination_vec, cr.recombination_vector(df))
    return sim


def np_pack(pop):
    """bool[..., m, 2] -> uint32[..., 2, Wpad] with numpy only."""
    from breedgym_b200 import _lib

    m = pop.shape[-2]
    wpad = _lib.words_per_row(m)
    planes = np.moveaxis(pop, -1, -2)  # [..., 2, m]
    padded = np.zeros(planes.shape[:-1] + (wpad * 32,), dtype=np.uint8)
    padded[..., :m] = planes
    by = np.packbits(padded, axis=-1, bitorder="little")
    return by.view(np.uint32).reshape(planes.shape[:-1] + (wpad,))


def words_u32(p):
    return p.words.cpu().numpy().view(np.uint32)


@pytest.mark.parametrize("m", [1, 31, 32, 33, 77, 128, 129, 1000, 10000])
def test_pack_unpack_roundtrip(cuda_device, m):
    rng = np.random.default_rng(m)
    sim = make_sim(make_map(m))
    pop = rng.random((5, m, 2)) < 0.5
    p = sim.as_packed(pop)
    assert p.shape == (5, m, 2) and len(p) == 5
    assert np.array_equal(words_u32(p), np_pack(pop))
    assert np.array_equal(np.asarray(p), pop)
    pops = rng.random((2, 3, m, 2)) < 0.3
    pp = sim.as_packed(pops)
    assert pp.shape == (2, 3, m, 2)
    assert np.array_equal(np.asarray(pp), pops)
    # uint8 input with values other than 0/1 counts as True
    assert np.array_equal(np.asarray(sim.as_packed(pop.astype(np.uint8) * 7)), pop)


def oracle_masks(r, key, rows, schedule, layout):
    keys = jp.split(key, rows, layout)
    out = np.zeros((rows, len(r)), dtype=bool)
    for q in range(rows):
        k = jp.split(keys[q], 2, layout)[0] if schedule == "S2" else keys[q]
        sites = jp.uniform(k, len(r), layout) < r
        out[q] = np.bitwise_xor.accumulate(sites.astype(np.uint8)).astype(bool)
    return out


@pytest.mark.parametrize("layout", LAYOUTS)
@pytest.mark.parametrize("m,r_kind", [(1, "half"), (2, "half"), (31, "random"), (32, "random"), (33, "random"),
                                      (63, "half"), (64, "half"), (65, "random"), (1001, "random"),
                                      (4099, "half"), (10000, "random"), (100002, "tiny"), (70001, "random")])
def test_crossover_masks_bit_exact(cuda_device, layout, m, r_kind):
    import torch

    from breedgym_b200 import _lib

    df = make_map(m, r_kind=r_kind, seed=m)
    rows = 7
    for schedule in SCHEDULES:
        sim = make_sim(df, rng_layout=layout, key_schedule=schedule)
        key = jp.key(1000 + m)
        out = torch.zeros((rows, sim.words_per_row), dtype=torch.int32, device=cuda_device)
        _lib.check(_lib.load().bg_meiosis_masks(sim._engine, out.data_ptr(), rows, _lib.nptr(key), sim._layout(),
                                                sim._schedule(), sim._stream()))
        got = out.cpu().numpy().view(np.uint32)
        ref = oracle_masks(sim.recombination_vec, key, rows, schedule, layout)
        exp = np_pack(np.stack([ref, ref], axis=-1))[:, 0]
        assert np.array_equal(got, exp), (layout, schedule, m)


def test_cross_matches_golden_vectors(cuda_device, golden):
    g = golden
    m = g["pop"].shape[1]
    df = pd.DataFrame({"CHR.PHYS": np.zeros(m, int), "RecombRate": np.zeros(m)})
    for t in range(g["eff"].shape[1]):
        df[f"t{t}"] = g["eff"][:, t]
    for layout in LAYOUTS:
        for schedule in SCHEDULES:
            from breedgym_b200.simulator import Simulator

            sim = Simulator(genetic_map=df, device=0, rng_layout=layout, key_schedule=schedule)
            # install the golden recombination vector verbatim (bypasses the map conventions)
            from breedgym_b200 import _lib

            _lib.check(_lib.load().bg_engine_set_map(sim._engine, _lib.nptr(g["r"]), _lib.nptr(np.ascontiguousarray(g["eff"])),
                                                     m, g["eff"].shape[1], 0.0))
            pop = sim.as_packed(g["pop"])
            off = sim._cross_indexed(pop, g["pairs"], g["key"])
            assert np.array_equal(np.asarray(off), g[f"cross_{layout}_{schedule}"])
            gebv = sim.GEBV_model(off).cpu().numpy()
            assert np.allclose(gebv, g[f"gebv_{layout}_{schedule}"], rtol=GEBV_RTOL, atol=0)
        sim = Simulator(genetic_map=df, device=0, rng_layout=layout, mutation=0.05)
        _lib.check(_lib.load().bg_engine_set_map(sim._engine, _lib.nptr(g["r"]), _lib.nptr(np.ascontiguousarray(g["eff"])),
                                                 m, g["eff"].shape[1], 0.05))
        pop = sim.as_packed(g["pop"])
        assert np.array_equal(np.asarray(sim._cross_indexed(pop, g["pairs"], g["key"])), g[f"cross_mut_{layout}"])


@pytest.mark.parametrize("layout", LAYOUTS)
@pytest.mark.parametrize("schedule", SCHEDULES)
@pytest.mark.parametrize("m", [1, 33, 64, 1000, 10000, 100002])
def test_unique_key_cross_bit_exact(cuda_device, layout, schedule, m):
    rng = np.random.default_rng(m + 5)
    df = make_map(m, n_chr=7, seed=m)
    sim = make_sim(df, rng_layout=layout, key_schedule=schedule)
    n_src, n = 9, 6 if m > 20000 else 13
    pop = rng.random((n_src, m, 2)) < 0.5
    pairs = rng.integers(0, n_src, (n, 2))
    pairs[0] = (-1, n_src + 3)  # jnp indexing: negative wraps, too large clamps
    key = jp.key(77)
    got = np.asarray(sim._cross_indexed(sim.as_packed(pop), pairs, key))
    ref = co.cross_envs(pop[None], cr.normalize_index(pairs, n_src)[None], sim.recombination_vec, key, 0.0, schedule, layout)[0]
    assert got.shape == (n, m, 2)
    assert np.array_equal(got, ref)
    if m <= 1000:  # the NumPy restatement too (slower)
        assert np.array_equal(got, cr.cross(pop[cr.normalize_index(pairs, n_src)], sim.recombination_vec, key, 0.0, schedule, layout))


def test_simulator_cross_key_chain_and_dense_parents(cuda_device):
    """`Simulator.cross` advances the key like chromax (`random_key, k = split(random_key)`) and accepts
    both the lazy `population[action]` view and a materialised (n, 2, m, 2) array."""
    rng = np.random.default_rng(0)
    m = 500
    df = make_map(m, seed=1)
    sim = make_sim(df)
    osim = cr.OracleSimulator(sim.recombination_vec, np.ones((m, 1)), seed=0)
    pop = rng.random((8, m, 2)) < 0.5
    pairs = rng.integers(0, 8, (5, 2))
    packed = sim.as_packed(pop)
    for step in range(3):
        sim.set_seed(42 + step)
        osim.set_seed(42 + step)
        a = sim.cross(packed[pairs])
        b = sim.cross(pop[pairs])  # dense parents, next key in the chain
        assert np.array_equal(np.asarray(a), osim.cross(pop[pairs]))
        assert np.array_equal(np.asarray(b), osim.cross(pop[pairs]))
        assert np.array_equal(sim.random_key, osim.random_key)
    with pytest.raises(ValueError):
        sim.cross(np.zeros((3, 3, m, 2), bool))


@pytest.mark.parametrize("layout", LAYOUTS)
@pytest.mark.parametrize("m,mut", [(1000, 0.0), (10000, 0.0), (333, 0.02)])
def test_vector_env_cross_shares_masks_bit_exact(cuda_device, layout, m, mut):
    rng = np.random.default_rng(m)
    df = make_map(m, n_chr=4, seed=m)
    sim = make_sim(df, rng_layout=layout, mutation=mut)
    E, n_src, n = 5, 11, 9
    pops = rng.random((E, n_src, m, 2)) < 0.5
    acts = rng.integers(0, n_src, (E, n, 2))
    acts[1, 0] = (-2, n_src)  # wrap / clamp
    key = jp.key(31)
    got = np.asarray(sim._cross_indexed(sim.as_packed(pops), acts, key))
    ref = co.cross_envs(pops, cr.normalize_index(acts, n_src), sim.recombination_vec, key, mut, "S2", layout)
    assert np.array_equal(got, ref)


@pytest.mark.parametrize("layout", LAYOUTS)
def test_double_haploid_bit_exact(cuda_device, layout):
    rng = np.random.default_rng(3)
    m = 777
    df = make_map(m, seed=9)
    sim = make_sim(df, rng_layout=layout)
    osim = cr.OracleSimulator(sim.recombination_vec, np.ones((m, 1)), seed=5, layout=layout)
    pop = rng.random((6, m, 2)) < 0.5
    sim.set_seed(5)
    got = sim.double_haploid(pop, n_offspring=4)
    assert got.shape == (6, 4, m, 2)
    ref = osim.double_haploid(pop, 4)
    assert np.array_equal(np.asarray(got), ref)
    assert np.array_equal(np.asarray(got)[..., 0], np.asarray(got)[..., 1])  # homozygous lines
    one = sim.double_haploid(pop, n_offspring=1)
    assert one.shape == (6, m, 2) and np.array_equal(np.asarray(one), osim.double_haploid(pop, 1))
    # a batch of populations = vmap(double_haploid, in_axes=(None, 0)): ONE key for every env, one launch
    pops = rng.random((3, 5, m, 2)) < 0.5
    sim.set_seed(8)
    batch = np.asarray(sim.double_haploid(pops, n_offspring=2))
    assert batch.shape == (3, 5, 2, m, 2)
    for e in range(3):
        osim.set_seed(8)
        assert np.array_equal(batch[e], osim.double_haploid(pops[e], 2))


def test_double_haploid_matches_golden(cuda_device, golden):
    from breedgym_b200 import _lib
    from breedgym_b200.simulator import Simulator

    g = golden
    m = g["pop"].shape[1]
    df = pd.DataFrame({"CHR.PHYS": np.zeros(m, int), "RecombRate": np.zeros(m), "y": np.zeros(m, np.float32)})
    for layout in LAYOUTS:
        sim = Simulator(genetic_map=df, device=0, rng_layout=layout)
        _lib.check(_lib.load().bg_engine_set_map(sim._engine, _lib.nptr(g["r"]), _lib.nptr(np.zeros((m, 1), np.float32)), m, 1, 0.0))
        sim.random_key = g["key"]
        # feed the golden key directly as the cross key
        import torch

        pop = sim.as_packed(g["pop"])
        out = sim._empty_words(len(pop), 3)
        key = np.ascontiguousarray(g["key"], dtype=np.uint32)
        _lib.check(_lib.load().bg_double_haploid(sim._engine, pop.words.data_ptr(), out.data_ptr(), 1, len(pop), 3,
                                                 _lib.nptr(key), sim._layout(), sim._schedule(), sim._stream()))
        torch.cuda.synchronize()
        from breedgym_b200.population import PackedPopulation

        assert np.array_equal(np.asarray(PackedPopulation(sim, out)), g[f"dh_{layout}"])


@pytest.mark.parametrize("layout", LAYOUTS)
@pytest.mark.parametrize("N,n,E", [(50, 20, 3), (370, 370, 4), (370, 200, 2), (2000, 10, 2), (1, 1, 1)])
def test_reset_selection_indices_bit_exact(cuda_device, golden, layout, N, n, E):
    import torch

    from breedgym_b200 import _lib

    sim = make_sim(make_map(64), rng_layout=layout)
    key = jp.key(7)
    idx = torch.empty((E, n), dtype=torch.int32, device=cuda_device)
    _lib.check(_lib.load().bg_reset_indices(sim._engine, _lib.nptr(key), E, 0, E, N, n, sim._layout(), idx.data_ptr(), sim._stream()))
    _, _, ref = cr.vec_reset(np.zeros((N, 1, 2), bool), n, E, key, layout)
    assert np.array_equal(idx.cpu().numpy(), ref)
    if (N, n, E) == (50, 20, 3):
        assert np.array_equal(idx.cpu().numpy(), golden[f"reset_idx_{layout}"])
    # a shard computing envs [1, E) of E gets the same rows
    if E > 1:
        part = torch.empty((E - 1, n), dtype=torch.int32, device=cuda_device)
        _lib.check(_lib.load().bg_reset_indices(sim._engine, _lib.nptr(key), E, 1, E - 1, N, n, sim._layout(), part.data_ptr(), sim._stream()))
        assert np.array_equal(part.cpu().numpy(), ref[1:])


def gebv_algo(sim, packed, algo):
    import torch

    from breedgym_b200 import _lib

    w = packed.words.contiguous()
    rows = int(np.prod(w.shape[:-2]))
    out = torch.empty((rows, sim.GEBV_model.n_traits), dtype=torch.float32, device=w.device)
    _lib.check(_lib.load().bg_gebv_algo(sim._engine, w.data_ptr(), rows, out.data_ptr(), algo, sim._stream()))
    return out.cpu().numpy()


@pytest.mark.parametrize("m,T,rows", [(1, 1, 3), (33, 2, 5), (128, 1, 4), (129, 3, 300), (1000, 7, 64), (10000, 1, 700),
                                      (100002, 1, 40), (2500, 16, 130), (777, 32, 257)])
def test_gebv_matches_float64_oracle(cuda_device, m, T, rows):
    rng = np.random.default_rng(m * 7 + T)
    df = make_map(m, T=T, seed=m)
    sim = make_sim(df)
    pop = rng.random((rows, m, 2)) < rng.random((rows, 1, 1))  # varied allele frequencies
    packed = sim.as_packed(pop)
    eff = sim.GEBV_model.marker_effects
    ref = cr.gebv(pop, eff)
    direct, lut, tc = gebv_algo(sim, packed, 1), gebv_algo(sim, packed, 2), gebv_algo(sim, packed, 3)
    assert np.array_equal(direct, lut), "both CUDA-core kernels sum the same fixed-point integers"
    assert np.array_equal(direct, tc), "the tcgen05 int8 GEMM (operand in TMEM) reproduces the same integers"
    auto = sim.GEBV_model(packed).cpu().numpy()
    assert np.array_equal(auto, lut)
    # fixed point chosen per map (bg_gebv_digits): worst-case error 2^-25 * sum|w| (zero for all practical purposes
    # with 8 digits, which one- and two-trait maps always get), then ONE rounding to float32
    from breedgym_b200 import _lib
    D = _lib.load().bg_gebv_digits(sim._engine)
    assert 4 <= D <= 8 and (T > 2 or D == 8)
    ulp = np.spacing(np.abs(ref).astype(np.float32)) + 1e-30
    quant = 0.0 if D == 8 else 2.0 ** -25 * np.abs(eff).sum(axis=0)[None, :]
    assert np.all(np.abs(auto - ref) <= ulp + quant)
    assert np.allclose(auto, ref, rtol=GEBV_RTOL, atol=float(np.max(quant)))
    if D < 8:  # forcing 8 digits gives the correctly rounded result again, from all kernels
        sim8 = make_sim(df, engine_options={"gebv_digits": 8})
        p8 = sim8.as_packed(pop)
        a8 = gebv_algo(sim8, p8, 3)
        assert np.array_equal(a8, gebv_algo(sim8, p8, 2))
        assert np.all(np.abs(a8 - ref) <= ulp)
        assert np.allclose(a8, ref, rtol=GEBV_RTOL, atol=0)
    assert np.allclose(co.gebv(pop, eff), ref, rtol=1e-12)
    df_gebv = sim.GEBV(packed)
    assert list(df_gebv.columns) == sim.trait_names and df_gebv.shape == (rows, T)


def test_gebv_effect_scale_extremes_and_batch_dims(cuda_device):
    rng = np.random.default_rng(0)
    m = 700
    for scale in (1e-20, 1.0, 1e20):
        df = make_map(m, T=2, seed=1)
        df["trait0"] = (df["trait0"] * scale).astype(np.float32)
        sim = make_sim(df)
        pops = rng.random((3, 4, m, 2)) < 0.5
        out = sim.GEBV_model(pops)
        assert tuple(out.shape) == (3, 4, 2)
        assert np.allclose(out.cpu().numpy(), cr.gebv(pops, sim.GEBV_model.marker_effects), rtol=GEBV_RTOL, atol=0)
    # all-zero effects and an empty / full genome
    df = make_map(m, T=1, seed=2)
    df["trait0"] = 0.0
    sim = make_sim(df)
    assert np.all(sim.GEBV_model(pops).cpu().numpy() == 0)
    df = make_map(m, T=1, seed=2)
    sim = make_sim(df)
    ones = np.ones((2, m, 2), bool)
    assert np.allclose(sim.GEBV_model(ones).cpu().numpy()[:, 0], sim.GEBV_model.max * 0 + 2 * sim.GEBV_model.mean, rtol=1e-5)
    assert np.all(sim.GEBV_model(np.zeros((2, m, 2), bool)).cpu().numpy() == 0)


def test_gebv_linearity_property(cuda_device):
    """GEBV(offspring) is the sum of the two gametes' values; a DH line scores twice its gamete."""
    rng = np.random.default_rng(4)
    m = 3000
    sim = make_sim(make_map(m, T=2, seed=4))
    a = rng.random((16, m)) < 0.5
    b = rng.random((16, m)) < 0.5
    zero = np.zeros_like(a)
    ga = sim.GEBV_model(np.stack([a, zero], -1)).double()
    gb = sim.GEBV_model(np.stack([zero, b], -1)).double()
    gab = sim.GEBV_model(np.stack([a, b], -1)).double()
    assert np.allclose((ga + gb).cpu().numpy(), gab.cpu().numpy(), rtol=1e-6, atol=1e-6)
    gaa = sim.GEBV_model(np.stack([a, a], -1)).double()
    assert np.allclose((2 * ga).cpu().numpy(), gaa.cpu().numpy(), rtol=1e-6, atol=1e-6)


def test_reductions(cuda_device):
    import torch

    from breedgym_b200 import _lib

    sim = make_sim(make_map(64))
    x = torch.randn(6, 777, device=cuda_device)
    mx = torch.empty(6, device=cuda_device)
    mean = torch.empty(6, device=cuda_device)
    _lib.check(_lib.load().bg_reduce_max(sim._engine, x.data_ptr(), 6, 777, mx.data_ptr(), sim._stream()))
    _lib.check(_lib.load().bg_reduce_mean(sim._engine, x.data_ptr(), 6, 777, mean.data_ptr(), sim._stream()))
    assert torch.equal(mx, x.max(dim=1).values)
    assert np.allclose(mean.cpu().numpy(), x.double().mean(dim=1).cpu().numpy(), rtol=1e-6)


def test_errors_are_loud(cuda_device):
    from breedgym_b200.simulator import Simulator

    df = make_map(100)
    with pytest.raises(ValueError):
        Simulator(genetic_map=df, device=0, rng_layout="nope")
    with pytest.raises(ValueError):
        Simulator(genetic_map=df, device="cpu")
    sim = make_sim(df)
    with pytest.raises(ValueError):
        sim.as_packed(np.zeros((3, 99, 2), bool))
    s1 = make_sim(df, key_schedule="S1", mutation=0.1)
    with pytest.raises(ValueError):  # S1 predates the mutation key
        s1.cross(np.zeros((2, 2, 100, 2), bool))
    with pytest.raises(ValueError):
        sim.select(np.zeros((3, 100, 2), bool), k=5)


def test_full_size_vector_step_sampled_rows_and_properties(cuda_device):
    """BASELINE config C2 (64 envs x 370 x 10k, small_genetic_map): sampled gamete rows against the
    oracle plus size-independent properties."""
    from pathlib import Path

    from breedgym_b200.simulator import Simulator

    data = Path(__file__).resolve().parents[1] / "breedgym_b200" / "data"
    sim = Simulator(genetic_map=data / "small_genetic_map.txt", trait_names=["Yield"], device=0, seed=0)
    rng = np.random.default_rng(0)
    E, N, m = 64, 370, sim.n_markers
    germ = rng.random((N, m, 2)) < 0.5
    pops = np.broadcast_to(germ, (E, N, m, 2)).copy()
    for e in range(1, E):  # make envs differ
        pops[e] = germ[rng.permutation(N)]
    acts = rng.integers(0, N, (E, N, 2)).astype(np.int32)
    key = jp.key(7)
    packed = sim.as_packed(pops)
    out = sim._cross_indexed(packed, acts, key)
    got = np.asarray(out)
    # sampled (env, offspring, parent) rows vs the oracle
    keys = jp.split(key, 2 * N)
    for e, i, p in [(0, 0, 0), (0, 369, 1), (17, 123, 0), (63, 369, 1), (31, 5, 1), (40, 200, 0)]:
        ref = cr.meiosis(pops[e, acts[e, i, p]], sim.recombination_vec, keys[2 * i + p])
        assert np.array_equal(got[e, i, :, p], ref)
    # every offspring allele comes from one of the two haplotypes of its parent at that marker
    e = 9
    par = pops[e][acts[e]]  # (N, 2, m, 2)
    child = got[e]
    for p in range(2):
        ok = (child[:, :, p] == par[:, p, :, 0]) | (child[:, :, p] == par[:, p, :, 1])
        assert ok.all()
    # masks are shared: envs with the same parents give the same offspring
    acts2 = np.tile(acts[:1], (E, 1, 1))
    same = sim._cross_indexed(sim.as_packed(np.broadcast_to(germ, (E, N, m, 2)).copy()), acts2, key)
    w = words_u32(same)
    assert np.array_equal(w[0], w[1]) and np.array_equal(w[0], w[63])
    # GEBV of the step output against the float64 oracle on a sample of envs
    gebv = sim.GEBV_model(out).cpu().numpy()
    for e in (0, 33, 63):
        assert np.allclose(gebv[e], cr.gebv(got[e], sim.GEBV_model.marker_effects), rtol=GEBV_RTOL, atol=0)


@pytest.mark.parametrize("m,T,E,n_src,n", [(10000, 1, 5, 37, 41), (333, 3, 3, 20, 130), (100002, 1, 2, 9, 7), (4099, 16, 4, 12, 64),
                                           (10000, 1, 64, 20, 25), (1000, 2, 32, 10, 13), (777, 1, 40, 7, 9), (2049, 1, 128, 5, 3),
                                           (4097, 4, 17, 6, 33), (300, 24, 2, 6, 10), (257, 32, 3, 5, 9)])
def test_fused_cross_gebv_equals_cross_then_gebv(cuda_device, m, T, E, n_src, n):
    """bg_cross_gebv (one fused kernel for E > 1) == bg_cross followed by bg_gebv, bit for bit, and == the oracle."""
    import torch

    from breedgym_b200 import _lib

    rng = np.random.default_rng(m + T)
    sim = make_sim(make_map(m, n_chr=5, T=T, seed=m))
    pops = rng.random((E, n_src, m, 2)) < 0.5
    acts = rng.integers(0, n_src, (E, n, 2)).astype(np.int32)
    acts[0, 0] = (-1, n_src + 5)
    key = jp.key(123)
    packed = sim.as_packed(pops)
    ref_pop = sim._cross_indexed(packed, acts, key)
    ref_gebv = sim.GEBV_model(ref_pop).cpu().numpy()
    out = sim._empty_words(E, n)
    gebv = torch.empty((E, n, T), dtype=torch.float32, device=cuda_device)
    a = torch.from_numpy(acts).to(cuda_device)
    import os

    # the fused step kernel with one CTA per (tile, K range) (fused_dyn=0), the persistent one with the dynamic work
    # queue (fused_dyn=1: the kernel launches with many tiles take), and blend + GEBV kernels (fuse=0)
    for fuse, dyn in ((1, 0), (1, 1), (0, 0)):
        out.zero_()
        gebv.zero_()
        sim.set_option("fuse", fuse)
        sim.set_option("fused_dyn", dyn)
        try:
            for _ in range(2):  # twice: the persistent kernel's work counter must come back to zero
                _lib.check(_lib.load().bg_cross_gebv(sim._engine, packed.words.data_ptr(), a.data_ptr(), out.data_ptr(), E, n_src, n,
                                                     _lib.nptr(key), sim._layout(), sim._schedule(), gebv.data_ptr(), sim._stream()))
        finally:
            sim.set_option("fuse", 1)
            sim.set_option("fused_dyn", -1)
        assert np.array_equal(out.cpu().numpy(), ref_pop.words.cpu().numpy()), (fuse, dyn)
        assert np.array_equal(gebv.cpu().numpy(), ref_gebv), (fuse, dyn)
    oref = co.cross_envs(pops, cr.normalize_index(acts, n_src), sim.recombination_vec, key)
    from breedgym_b200.population import PackedPopulation

    assert np.array_equal(np.asarray(PackedPopulation(sim, out)), oref)
    assert np.allclose(gebv.cpu().numpy(), cr.gebv(oref, sim.GEBV_model.marker_effects), rtol=GEBV_RTOL, atol=0)


def test_wheat_scale_c3_cross_and_gebv_sampled_rows(cuda_device):
    """BASELINE config C3 (time_wheat.py shape): 1000 individuals x 100 002 markers x 21 chromosomes, one trait.
    Unique-key cross at full size; sampled gamete rows and GEBV rows against the oracle."""
    rng = np.random.default_rng(3)
    m, n = 100_002, 1000
    df = pd.DataFrame({"CHR.PHYS": np.arange(m) // 4762, "RecombRate": np.full(m, 1.5e-3),
                       "Yield": rng.standard_normal(m).astype(np.float32)})
    sim = make_sim(df)
    assert len(sim.chr_lens) == 21
    pop = rng.random((n, m, 2)) < 0.5
    pairs = rng.integers(0, n, (n, 2))
    key = jp.key(7)
    out = sim._cross_indexed(sim.as_packed(pop), pairs, key)
    got = np.asarray(out)
    keys = jp.split(key, 2 * n)
    for i, p in [(0, 0), (0, 1), (499, 1), (999, 0), (999, 1)]:
        assert np.array_equal(got[i, :, p], cr.meiosis(pop[pairs[i, p]], sim.recombination_vec, keys[2 * i + p]))
    gebv = sim.GEBV_model(out).cpu().numpy()
    rows = [0, 1, 500, 999]
    assert np.allclose(gebv[rows], cr.gebv(got[rows], sim.GEBV_model.marker_effects), rtol=GEBV_RTOL, atol=0)
    # roughly 1.5e-3 * m + chromosome starts crossovers per gamete
    switches = []
    for i in range(20):  # where the parent's haplotypes differ the gamete reveals its source; count the switches
        par = pop[pairs[i, 0]]
        informative = par[:, 0] != par[:, 1]
        src = (got[i, :, 0] == par[:, 1])[informative]
        switches.append(np.count_nonzero(src[1:] != src[:-1]))
    assert 80 < np.mean(switches) < 250  # expectation: 1.5e-3 * m + 21 chromosome starts / 2 ~ 160


def test_million_marker_multi_trait_c4_shape(cuda_device):
    """BASELINE config C4 row shape: 1 M markers, 16 traits (tensor-core GEBV), a reduced number of offspring
    (the kernels are row-parallel: every gamete row is an independent function of (key, row))."""
    import torch

    rng = np.random.default_rng(4)
    m, n_par, n_off, T = 1_000_000, 6, 24, 16
    df = pd.DataFrame({"CHR.PHYS": np.arange(m) // 100_000, "RecombRate": np.full(m, 1.5e-3)})
    for t in range(T):
        df[f"t{t}"] = rng.standard_normal(m).astype(np.float32)
    sim = make_sim(df)
    pop = rng.random((n_par, m, 2)) < 0.5
    pairs = rng.integers(0, n_par, (n_off, 2))
    key = jp.key(11)
    out = sim._cross_indexed(sim.as_packed(pop), pairs, key)
    got = np.asarray(out)
    keys = jp.split(key, 2 * n_off)
    for i, p in [(0, 0), (23, 1), (11, 0)]:
        assert np.array_equal(got[i, :, p], cr.meiosis(pop[pairs[i, p]], sim.recombination_vec, keys[2 * i + p]))
    gebv = sim.GEBV_model(out).cpu().numpy()
    assert gebv.shape == (n_off, T)
    assert np.allclose(gebv, cr.gebv(got, sim.GEBV_model.marker_effects), rtol=GEBV_RTOL, atol=0)
    assert np.array_equal(gebv, gebv_algo(sim, out, 2))  # the CUDA-core LUT kernel sums the same integers
    torch.cuda.synchronize()


@pytest.mark.parametrize("m,T,E,n", [(10000, 1, 64, 370), (10000, 1, 512, 370), (4000, 2, 200, 90)])
def test_fused_step_kernel_full_size_equals_two_kernel_path(cuda_device, m, T, E, n):
    """BASELINE configs C2 (64 envs x 370 x 10 000) and one GPU's share of C5 (512 envs): the fused cross + GEBV kernel
    against the blend + GEBV kernels (themselves checked against the oracle above), bit for bit, on random packed
    populations generated on the GPU; plus the property that every offspring allele comes from its parent."""
    import os

    import torch

    from breedgym_b200 import _lib

    sim = make_sim(make_map(m, n_chr=10, T=T, seed=5))
    W = sim.words_per_row
    g = torch.Generator(device=cuda_device).manual_seed(E + n)
    pop = torch.randint(-2**31, 2**31 - 1, (E, n, 2, W), dtype=torch.int32, device=cuda_device, generator=g)
    full, tail = m // 32, m % 32
    pop[..., full + (1 if tail else 0):] = 0  # padding bits are zero in a real population
    if tail:
        pop[..., full] &= (1 << tail) - 1
    acts = torch.randint(0, n, (E, n, 2), dtype=torch.int32, device=cuda_device, generator=g)
    key = jp.key(99)
    outs, gebvs = [], []
    for fuse, dyn in ((1, 0), (0, 0), (1, 1)):  # one CTA per (tile, K range); blend + GEBV; persistent with the work queue
        out = torch.zeros((E, n, 2, W), dtype=torch.int32, device=cuda_device)
        gebv = torch.zeros((E, n, T), dtype=torch.float32, device=cuda_device)
        sim.set_option("fuse", fuse)
        sim.set_option("fused_dyn", dyn)
        try:
            _lib.check(_lib.load().bg_cross_gebv(sim._engine, pop.data_ptr(), acts.data_ptr(), out.data_ptr(), E, n, n,
                                                 _lib.nptr(key), sim._layout(), sim._schedule(), gebv.data_ptr(), sim._stream()))
        finally:
            sim.set_option("fuse", 1)
            sim.set_option("fused_dyn", -1)
        outs.append(out)
        gebvs.append(gebv)
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
    assert torch.equal(gebvs[0], gebvs[1]) and torch.equal(gebvs[0], gebvs[2])
    # allele provenance on one env: offspring plane p only holds bits present in one of parent p's two planes
    e = E // 3
    par = pop[e][acts[e].long()]  # [n, 2 (which parent), 2 (plane), W]
    child = outs[0][e]            # [n, 2 (plane = which parent's gamete), W]
    for p in range(2):
        c, h0, h1 = child[:, p], par[:, p, 0], par[:, p, 1]
        assert bool(torch.all((c & ~(h0 | h1)) == 0)) and bool(torch.all((~c & (h0 & h1)) == 0))


def test_fuzz_fused_step_kernels(cuda_device):
    """scripts/fuzz_fused.py: randomised shapes (1 ... 100 002 markers, 1 ... 24 traits, 2 ... 200 envs, ragged tiles, index wrap /
    clamp, both PRNG layouts) -- both fused step kernels == the blend + GEBV kernels, bit for bit."""
    import importlib.util
    from pathlib import Path

    spec = importlib.util.spec_from_file_location("fuzz_fused", Path(__file__).resolve().parents[1] / "scripts" / "fuzz_fused.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert mod.run_cases(12, 20261018, verbose=False) == 0
