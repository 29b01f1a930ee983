"""Parity at BASELINE.json's full sizes, through the C ABI, against the CPU oracle on ALL rows where the oracle
finishes in seconds, and through sampled rows + an independent kernel where it does not (C4), plus the wrapper /
multi-device surface the reference tests (tests/test_vec.py:72-112, 186-204; breeding_programs_env.py)."""
from pathlib import Path

import numpy as np
import pandas as pd
import pytest

from oracle import c_oracle as co
from oracle import chromax_ref as cr
from oracle import jax_prng as jp

pytestmark = pytest.mark.gpu

ROOT = Path(__file__).resolve().parents[1]
DATA = ROOT / "breedgym_b200" / "data"
GENOME = DATA / "sample_geno.npy"
GMAP = DATA / "sample_with_r_genetic_map.txt"
RTOL = 1e-5


def gym():
    from breedgym_b200 import gym_compat

    return gym_compat


@pytest.mark.parametrize("E,fused_dyn", [(64, 0), (64, 1), (512, 0), (512, 1)])
def test_fused_step_kernel_every_row_equals_c_oracle(cuda_device, E, fused_dyn):
    """BASELINE C2 (64 envs x 370 x 10 000) and one GPU's share of C5 (512 envs): BOTH fused cross + GEBV kernels (one
    CTA per (tile, K range); persistent with the dynamic work queue -- the library's own choice is the former at 64 envs
    and the latter at 512), called through bg_cross_gebv, against the C oracle on EVERY offspring row and every GEBV
    (the oracle draws the 2n masks once and is run per 64-env chunk with the same key, so the host never holds more
    than one chunk)."""
    import torch

    from breedgym_b200 import _lib
    from breedgym_b200.simulator import Simulator

    sim = Simulator(genetic_map=DATA / "small_genetic_map.txt", trait_names=["Yield"], device=0, seed=0,
                    engine_options={"fused_dyn": fused_dyn})
    N, m, CH = 370, sim.n_markers, 64
    rng = np.random.default_rng(E)
    acts = rng.integers(0, N, (E, N, 2)).astype(np.int32)
    words = sim._empty_words(E, N)
    chunks = []
    for c in range(E // CH):
        pc = rng.integers(0, 2, (CH, N, m, 2), dtype=np.uint8).view(np.bool_)
        words[c * CH:(c + 1) * CH] = sim.as_packed(pc).words
        chunks.append(pc if E == CH else None)
    out = sim._empty_words(E, N)
    gebv = torch.empty((E, N, 1), dtype=torch.float32, device=cuda_device)
    key = jp.key(7)
    a = torch.from_numpy(acts).to(cuda_device)
    assert sim._engine and _lib.load().bg_gebv_digits(sim._engine) == 8
    _lib.check(_lib.load().bg_cross_gebv(sim._engine, words.data_ptr(), a.data_ptr(), out.data_ptr(), E, N, N, _lib.nptr(key),
                                         sim._layout(), sim._schedule(), gebv.data_ptr(), sim._stream()))
    torch.cuda.synchronize()
    g = gebv.cpu().numpy()
    eff = sim.GEBV_model.marker_effects
    rng = np.random.default_rng(E)
    rng.integers(0, N, (E, N, 2))  # replay the stream: the chunks are regenerated instead of kept (512 envs = 3.8 GB)
    for c in range(E // CH):
        pc = rng.integers(0, 2, (CH, N, m, 2), dtype=np.uint8).view(np.bool_)
        ref = co.cross_envs(pc, acts[c * CH:(c + 1) * CH], sim.recombination_vec, key, shared_masks=True)
        ref_words = sim.as_packed(ref).words
        assert torch.equal(out[c * CH:(c + 1) * CH], ref_words), f"offspring differ from the C oracle in env chunk {c}"
        rg = co.gebv(ref, eff)
        assert np.allclose(g[c * CH:(c + 1) * CH], rg, rtol=RTOL, atol=0)
        assert np.all(np.abs(g[c * CH:(c + 1) * CH] - rg) <= np.spacing(np.abs(rg).astype(np.float32)))  # <= 1 ulp


def test_c4_full_size_cross_and_16_trait_gebv(cuda_device):
    """BASELINE C4 at FULL size: 10 000 offspring of 1000 parents x 1 000 000 markers, 16 traits.  The 20 GB of
    offspring cannot visit the host, so: 32 sampled gamete rows (16 offspring x 2) against the NumPy oracle's meiosis
    with the row's own key of split(key, 20 000), and ALL 160 000 GEBVs of the tcgen05 kernel against the CUDA-core LUT kernel
    (same fixed-point integers, different machinery) with the sampled rows also against the float64 oracle."""
    import torch

    from breedgym_b200 import _lib
    from breedgym_b200.population import PackedPopulation
    from breedgym_b200.simulator import Simulator

    m, n_par, n_off, T = 1_000_000, 1000, 10_000, 16
    rng = np.random.default_rng(4)
    df = pd.DataFrame({"CHR.PHYS": np.arange(m) // 100_000, "RecombRate": np.full(m, 1.5e-3, dtype=np.float32)})
    for t in range(T):
        df[f"t{t}"] = rng.standard_normal(m).astype(np.float32)
    sim = Simulator(genetic_map=df, seed=0, device=0)
    lib = _lib.load()
    D = lib.bg_gebv_digits(sim._engine)
    W = sim.words_per_row
    gen = torch.Generator(device=cuda_device)
    gen.manual_seed(40)
    pop = torch.randint(-2**31, 2**31 - 1, (n_par, 2, W), dtype=torch.int32, device=cuda_device, generator=gen)
    pop[..., m // 32:] = 0  # 1 000 000 = 31250 * 32: no partial word; the padding words are zero
    pairs = rng.integers(0, n_par, (n_off, 2)).astype(np.int32)
    key = jp.key(11)
    out = sim._cross_indexed(PackedPopulation(sim, pop), pairs, key)
    assert out.words.shape == (n_off, 2, W)
    keys = jp.split(key, 2 * n_off)
    sample = rng.choice(n_off, 16, replace=False)
    got = np.asarray(PackedPopulation(sim, out.words[torch.from_numpy(sample).to(cuda_device)].contiguous()))  # [16, m, 2]
    need = np.unique(pairs[sample])
    host_par = {int(a): np.asarray(PackedPopulation(sim, pop[int(a):int(a) + 1].contiguous()))[0] for a in need}
    for s, i in enumerate(sample):
        for p in range(2):  # 32 gamete rows
            ref = cr.meiosis(host_par[int(pairs[i, p])], sim.recombination_vec, keys[2 * i + p])
            assert np.array_equal(got[s, :, p], ref), f"offspring {i}, gamete {p} differs from the oracle"
    # GEBV: every value, tensor cores vs the LUT kernel; sampled rows vs float64
    gebv = torch.empty((n_off, T), dtype=torch.float32, device=cuda_device)
    lut = torch.empty_like(gebv)
    _lib.check(lib.bg_gebv_algo(sim._engine, out.words.data_ptr(), n_off, gebv.data_ptr(), 3, sim._stream()))
    _lib.check(lib.bg_gebv_algo(sim._engine, out.words.data_ptr(), n_off, lut.data_ptr(), 2, sim._stream()))
    torch.cuda.synchronize()
    assert torch.equal(gebv, lut), "tcgen05 GEBV and LUT GEBV must produce the same bits on all 160 000 values"
    eff = sim.GEBV_model.marker_effects
    ref = cr.gebv(got, eff)
    quant = 0.0 if D == 8 else 2.0 ** -25 * np.abs(eff).sum(axis=0)[None, :]
    mine = gebv[torch.from_numpy(sample).to(cuda_device)].cpu().numpy()
    assert np.all(np.abs(mine - ref) <= np.spacing(np.abs(ref).astype(np.float32)) + quant)
    assert np.allclose(mine, ref, rtol=RTOL, atol=float(np.max(quant)))
    assert np.allclose(co.gebv(got, eff), ref, rtol=1e-12)


def _pair_scores_reference(scores, n):
    """breedgym/vector/vec_wrappers.py:100-112 per env with the oracle's jax-like helpers."""
    out = []
    for e in range(scores.shape[0]):
        v, i = jp.top_k(scores[e].astype(np.float32).reshape(-1), n)
        x = v.astype(np.float32)
        ex = np.exp(x - x.max())
        sm = ex / ex.sum(dtype=np.float32)
        reps = np.ceil(sm * np.float32(n)).astype(np.int32)
        out.append(jp.repeat_total(np.stack((i // scores.shape[-1], i % scores.shape[-1]), 1), reps, n))
    return np.stack(out).astype(np.int32)


def test_pair_scores_trajectory_matches_oracle_replay(cuda_device):
    """The reference's PairScores scenario (tests/test_vec.py:186-204: GEBV outer-sum policy, seed 7) as a whole
    trajectory against an oracle replay: low-level actions, populations, GEBVs and rewards."""
    num_envs, n = 4, 60
    env = gym().make("PairScores", num_envs=num_envs, initial_population=GENOME, genetic_map=GMAP, individual_per_gen=n,
                     trait_names=["Yield"], num_generations=5)
    germ = np.load(GENOME)
    _, infos = env.reset(seed=7)
    sim = env.simulator
    osim = cr.OracleSimulator(sim.recombination_vec, sim.GEBV_model.marker_effects, seed=7)
    _, opops, _ = cr.vec_reset(germ, n, num_envs, jp.key(7))
    for step in range(5):
        gebvs = np.asarray(infos["GEBV"]).squeeze(-1)
        scores = gebvs[:, :, None] + gebvs[:, None, :]  # np.add.outer per env
        pop, rews, _, tru, infos = env.step(scores)
        ref_pairs = _pair_scores_reference(scores, n)
        assert np.array_equal(infos["low_level_actions"], ref_pairs), f"step {step}: pairs differ from lax.top_k/softmax/repeat"
        opops = cr.vec_step(osim, opops, ref_pairs)
        g = cr.gebv(opops, osim.effects)
        assert np.allclose(infos["GEBV"], g, rtol=RTOL, atol=0)
        if step < 4:
            assert np.array_equal(np.asarray(pop), opops)
    assert np.all(tru) and np.allclose(rews, g.max(axis=(1, 2)), rtol=RTOL, atol=0)


def test_wheat_breedgym_trajectory_matches_oracle_replay(cuda_device):
    """WheatBreedGym (breedgym/vector/breeding_programs_env.py:51-72) step by step against an oracle replay: pair
    conversion, shared-key cross, shared-key double haploids (one launch for all envs), best k per line, global
    selection, infos and reward."""
    from breedgym_b200.vector import VecBreedGym, WheatBreedGym

    E, n, n_lines, plants, k_line = 3, 24, 10, 6, 3
    env = WheatBreedGym(VecBreedGym(num_envs=E, initial_population=GENOME, genetic_map=GMAP, individual_per_gen=n, autoreset=False,
                                    trait_names=["Yield"], num_generations=3),
                        n_lines=n_lines, plant_per_line=plants, k_per_line=k_line)
    germ = np.load(GENOME)
    _, infos = env.reset(seed=3)
    sim = env.simulator
    osim = cr.OracleSimulator(sim.recombination_vec, sim.GEBV_model.marker_effects, seed=3)
    _, opops, _ = cr.vec_reset(germ, n, E, jp.key(3))
    rng = np.random.default_rng(1)
    for step in range(3):
        scores = rng.standard_normal((E, n_lines, n_lines)).astype(np.float32)
        pop, rews, _, tru, infos = env.step(scores)
        pairs = _pair_scores_reference(scores, n_lines)
        lines = cr.vec_step(osim, opops, pairs)                       # [E, n_lines, m, 2], one key for all envs
        kdh = osim.next_cross_key()                                    # one key for all envs' double haploids
        nxt = []
        for e in range(E):
            dh = cr.double_haploid(lines[e], osim.r, kdh, plants, 0.0, osim.schedule, osim.layout)  # [n_lines, plants, m, 2]
            kept = []
            for ln in range(n_lines):
                v = cr.gebv(dh[ln], osim.effects).sum(-1).astype(np.float32)
                kept.append(dh[ln][jp.top_k(v, k_line)[1]])
            kept = np.concatenate(kept)                                # [n_lines * k_line, m, 2]
            v = cr.gebv(kept, osim.effects).sum(-1).astype(np.float32)
            nxt.append(kept[jp.top_k(v, n)[1]])
        opops = np.stack(nxt)
        assert np.array_equal(np.asarray(pop), opops), f"step {step}: populations differ from the oracle replay"
        g = cr.gebv(opops, osim.effects)
        assert np.allclose(infos["GEBV"], g, rtol=RTOL, atol=0)
    assert np.all(tru) and np.allclose(rews, g.max(axis=(1, 2)), rtol=RTOL, atol=0)


def test_distributed_breedgym_contract(cuda_device):
    """DistributedBreedGym (breedgym/vector/vec_env.py:150-236; tests/test_vec.py:72-112) with the devices at hand
    (two shards on cuda:0 when the box has one GPU): shapes, spaces, `_GEBV` info masks, and shard i == a VecBreedGym
    seeded with seed + i (AsyncVectorEnv's seeding convention)."""
    import torch

    from breedgym_b200.vector import DistributedBreedGym, VecBreedGym

    n_dev = torch.cuda.device_count()
    devices = list(range(n_dev)) if n_dev > 1 else [0, 0]
    envs_per_device, n = 3, 50
    env = DistributedBreedGym(envs_per_device=envs_per_device, devices=devices, initial_population=GENOME, genetic_map=GMAP,
                              individual_per_gen=n, trait_names=["Yield"])
    num_envs = envs_per_device * len(devices)
    assert env.num_envs == num_envs
    assert env.observation_space.shape[0] == num_envs and env.action_space.shape[0] == num_envs
    pop, infos = env.reset(seed=11)
    assert pop.shape == (num_envs, n, 1000, 2)
    assert infos["GEBV"].shape == (num_envs, n, 1) and infos["_GEBV"].shape == (num_envs,) and infos["_GEBV"].all()
    actions = np.random.default_rng(0).integers(0, n, size=(num_envs, n, 2))
    new_pop, reward, terminated, truncated, infos = env.step(actions)
    assert new_pop.shape == (num_envs, n, 1000, 2)
    assert reward.shape == terminated.shape == truncated.shape == (num_envs,)
    assert np.all(~terminated) and np.all(~truncated)
    assert isinstance(infos, dict) and len(infos["GEBV"]) == num_envs and infos["_GEBV"].all()
    for info in infos["GEBV"]:
        assert info.shape == (n, 1)
    # shard i is an independent VecBreedGym seeded with seed + i
    host = np.asarray(new_pop)
    for i, d in enumerate(devices):
        ref = VecBreedGym(num_envs=envs_per_device, initial_population=GENOME, genetic_map=GMAP, individual_per_gen=n,
                          trait_names=["Yield"], autoreset=False, device=d)
        ref.reset(seed=11 + i)
        sl = slice(i * envs_per_device, (i + 1) * envs_per_device)
        rp, rr, _, _, ri = ref.step(actions[sl])
        assert np.array_equal(host[sl], np.asarray(rp)) and np.array_equal(infos["GEBV"][sl], ri["GEBV"])
    env.close()


def test_txt_population_loader(cuda_device, tmp_path):
    """`Simulator.load_population` on a text file (np.loadtxt(bool) reshaped (n, m, 2)) and on .npy."""
    from breedgym_b200.simulator import Simulator

    sim = Simulator(genetic_map=GMAP, device=0, seed=0)
    germ = np.load(GENOME)[:7]
    txt = tmp_path / "pop.txt"
    np.savetxt(txt, germ.reshape(len(germ), -1).astype(np.int8), fmt="%d")
    assert np.array_equal(np.asarray(sim.load_population(txt)), germ)
    npy = tmp_path / "pop.npy"
    sim.save_population(sim.as_packed(germ), npy)
    assert np.array_equal(np.asarray(sim.load_population(npy)), germ)
    env = gym().make("breedgym:BreedGym", initial_population=txt, genetic_map=GMAP)
    pop, _ = env.reset(seed=0)
    assert np.array_equal(np.asarray(pop), germ)
