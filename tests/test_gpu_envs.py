"""Environment-level tests on the GPU: the reference's own test-suite scenarios
(tests/test_env.py, tests/test_vec.py, tests/test_wrappers.py of younik/breedgym) re-run against
this implementation, with the un-evaluable golden rewards replaced by whole-trajectory parity
against the CPU oracle (same seeds, same actions)."""
from pathlib import Path

import numpy as np
import pytest

from oracle import chromax_ref as cr
from oracle import jax_prng as jp

pytestmark = pytest.mark.gpu

DATA = Path(__file__).resolve().parents[1] / "breedgym_b200" / "data"
GENOME = DATA / "sample_geno.npy"  # (200, 1000, 2)
GMAP = DATA / "sample_with_r_genetic_map.txt"  # 1000 markers, trait Yield
RTOL = 1e-5


def gym():
    from breedgym_b200 import gym_compat

    return gym_compat


def oracle_sim(sim, seed):
    return cr.OracleSimulator(sim.recombination_vec, sim.GEBV_model.marker_effects, seed=seed,
                              mutation=sim.mutation, schedule=sim.key_schedule, layout=sim.rng_layout)


# ---- tests/test_env.py ------------------------------------------------------------------
def test_reset_population(cuda_device):
    env = gym().make("breedgym:BreedGym", initial_population=GENOME, genetic_map=GMAP)
    pop, _ = env.reset()
    init_pop = np.copy(pop)
    env.step(np.asarray(env.action_space.sample()) % len(pop))
    pop, _ = env.reset()
    assert np.all(init_pop == pop)
    assert np.array_equal(init_pop, np.load(GENOME))


@pytest.mark.parametrize("n", [1, 5, 10])
def test_num_progenies(cuda_device, n):
    env = gym().make("breedgym:BreedGym", initial_population=GENOME, genetic_map=GMAP)
    pop, _ = env.reset()
    action = np.random.randint(len(pop), size=(n, 2))
    obs, _, _, _, info = env.step(action)
    assert len(env.population) == n and obs.shape == (n, 1000, 2)
    assert env.observation_space.shape == (n, 1000, 2)
    assert info["GEBV"].shape == (n, 1)


def test_caching(cuda_device):
    env = gym().make("breedgym:BreedGym", initial_population=GENOME, genetic_map=GMAP)
    env.reset()
    GEBV = env.unwrapped.GEBV
    GEBV_copy = np.copy(GEBV)
    GEBV2 = env.unwrapped.GEBV
    assert id(GEBV) == id(GEBV2)
    assert np.all(GEBV_copy == GEBV2)
    corrcoef = env.corrcoef
    corrcoef_copy = np.copy(corrcoef)
    corrcoef2 = env.corrcoef
    assert id(corrcoef) == id(corrcoef2)
    assert np.all(corrcoef_copy == corrcoef2)
    env.step(np.array([[1, 3], [4, 2]]))
    assert id(corrcoef) != id(env.corrcoef)
    assert id(GEBV) != id(env.unwrapped.GEBV)


def test_reward_shaping(cuda_device):
    env = gym().make("breedgym:BreedGym", initial_population=GENOME, genetic_map=GMAP, reward_shaping=False)
    pop, _ = env.reset()
    for _ in range(9):
        action = np.asarray(env.action_space.sample()) % len(pop)
        pop, reward, _, truncated, _ = env.step(action)
        assert reward == 0
        assert not truncated
    action = np.asarray(env.action_space.sample()) % len(pop)
    _, reward, _, truncated, _ = env.step(action)
    assert reward != 0
    assert truncated
    env2 = gym().make("breedgym:BreedGym", initial_population=GENOME, genetic_map=GMAP, reward_shaping=True)
    pop, _ = env2.reset()
    action = np.asarray(env2.action_space.sample()) % len(pop)
    _, reward, _, _, _ = env2.step(action)
    assert reward != 0


@pytest.mark.parametrize("layout", ["legacy", "partitionable"])
def test_deterministic_trajectory_matches_oracle(cuda_device, layout):
    """Reference tests/test_env.py:103-119 (seed 7, fixed 10-pair action x 10 generations); the golden
    reward there needs chromax.sample_data, so the whole trajectory is checked against the oracle."""
    env = gym().make("breedgym:BreedGym", initial_population=GENOME, genetic_map=GMAP, reward_shaping=False,
                     rng_layout=layout)
    env.reset(seed=7)
    osim = oracle_sim(env.simulator, 7)
    opop = np.load(GENOME)
    action = np.array([[1, 2], [1, 5], [1, 7], [2, 5], [2, 9], [4, 7], [4, 8], [5, 9], [6, 8], [6, 9]])
    for _ in range(10):
        obs, r, _, tru, info = env.step(action)
        opop = osim.cross(opop[action])
        assert np.array_equal(np.asarray(obs), opop)
        assert np.allclose(info["GEBV"].to_numpy(), cr.gebv(opop, osim.effects), rtol=RTOL, atol=0)
    assert tru
    assert abs(r - np.mean(cr.gebv(opop, osim.effects))) <= RTOL * abs(r)
    # same seed -> same trajectory
    env.reset(seed=7)
    for _ in range(10):
        obs2, r2, _, _, _ = env.step(action)
    assert r2 == r and np.array_equal(np.asarray(obs2), opop)


def test_reset_subset_uses_np_random(cuda_device):
    env = gym().make("breedgym:BreedGym", initial_population=GENOME, genetic_map=GMAP)
    pop, _ = env.reset(seed=3, options={"n_individuals": 17})
    sel = np.random.Generator(np.random.PCG64(np.random.SeedSequence(3))).choice(200, 17, replace=False)
    assert np.array_equal(np.asarray(pop), np.load(GENOME)[sel])


# ---- tests/test_vec.py --------------------------------------------------------------------
def test_vec(cuda_device):
    num_envs, n = 8, 200
    env = gym().make("VecBreedGym", num_envs=num_envs, initial_population=GENOME, genetic_map=GMAP,
                     individual_per_gen=n)
    pop, _ = env.reset()
    expected_shape = (num_envs, n, env.simulator.n_markers, 2)
    assert pop.shape == expected_shape
    actions = np.random.randint(0, n, size=(num_envs, n, 2))
    new_pop, reward, terminated, truncated, infos = env.step(actions)
    assert new_pop.shape == expected_shape
    assert reward.shape == (num_envs,)
    assert np.all(~terminated)
    assert np.all(~truncated)
    assert isinstance(infos, dict)
    assert len(infos["GEBV"]) == num_envs
    for info in infos["GEBV"]:
        assert info.shape == (n, 1)


def test_vec_multi_trait_info_shape(cuda_device):
    env = gym().make("VecBreedGym", num_envs=2, initial_population=np.random.rand(30, 9839, 2) < 0.5,
                     genetic_map=DATA / "wheat_genetic_map.csv", individual_per_gen=20)
    _, infos = env.reset(seed=0)
    assert infos["GEBV"].shape == (2, 20, 7)
    assert len(env.simulator.chr_lens) == 21


@pytest.mark.parametrize("layout", ["legacy", "partitionable"])
def test_vec_deterministic_trajectory_matches_oracle(cuda_device, layout):
    """Reference tests/test_vec.py:115-133: E=4, n=200, seed 7, 20 random steps (spans an autoreset)."""
    num_envs, n = 4, 200
    env = gym().make("VecBreedGym", num_envs=num_envs, initial_population=GENOME, genetic_map=GMAP,
                     individual_per_gen=n, rng_layout=layout)
    germ = np.load(GENOME)
    np.random.seed(seed=7)
    pop, infos = env.reset(seed=7)
    osim = oracle_sim(env.simulator, 7)
    okey, opops, _ = cr.vec_reset(germ, n, num_envs, jp.key(7), layout)
    assert np.array_equal(np.asarray(pop), opops)
    assert np.allclose(infos["GEBV"], cr.gebv(opops, osim.effects), rtol=RTOL, atol=0)
    for step in range(20):
        action = np.random.randint(len(pop), size=(num_envs, n, 2))  # len(pop) == num_envs, as in the reference
        pop, rews, ter, tru, infos = env.step(action)
        opops = cr.vec_step(osim, opops, action)
        g = cr.gebv(opops, osim.effects)
        assert np.allclose(infos["GEBV"], g, rtol=RTOL, atol=0)
        if step % 10 == 9:
            assert np.all(tru)
            assert np.allclose(rews, g.max(axis=(1, 2)), rtol=RTOL, atol=0)
            okey, opops, _ = cr.vec_reset(germ, n, num_envs, okey, layout)  # autoreset
            if step == 9:  # the autoreset's infos are fetched lazily (host mode): read them once, leave them unread once
                held = env.reset_infos
                assert "GEBV" in held and np.allclose(held["GEBV"], cr.gebv(opops, osim.effects), rtol=RTOL, atol=0)
                assert held["GEBV"].shape == (num_envs, n, 1) and list(held.keys()) == ["GEBV"]
        else:
            assert not np.any(tru) and np.all(rews == 0)
        assert np.array_equal(np.asarray(pop), opops)
    assert np.array_equal(env.random_key, okey)


def test_selection_vec_and_gebv_policy_matches_oracle(cuda_device):
    """Reference tests/test_vec.py:43-69,136-154 with the oracle replaying the wrapper's index math."""
    num_envs, n = 4, 200
    env = gym().make("SelectionScores", k=10, num_envs=num_envs, initial_population=GENOME, genetic_map=GMAP,
                     individual_per_gen=n, trait_names=["Yield"])
    pop, infos = env.reset(seed=7)
    assert pop.shape == (num_envs, n, 1000, 2)
    germ = np.load(GENOME)
    osim = oracle_sim(env.simulator, 7)
    okey, opops, _ = cr.vec_reset(germ, n, num_envs, jp.key(7), "legacy")
    for _ in range(10):
        scores = infos["GEBV"].squeeze()
        _, rews, _, tru, infos = env.step(scores)
        keys = jp.split(okey, num_envs + 1)
        okey = keys[0]
        acts = []
        for e in range(num_envs):
            _, best = jp.top_k(cr.gebv(opops[e], osim.effects)[:, 0].astype(np.float32), 10)
            d = cr.diallel_indices(best)
            sel = jp.choice_no_replace(keys[1 + e], len(d), 45)
            acts.append(jp.repeat_total(d[sel], int(np.ceil(n / 45)), n))
        opops = cr.vec_step(osim, opops, np.stack(acts))
        for info in infos["GEBV"]:
            assert info.shape == (n, 1)
    assert np.all(tru)
    assert np.allclose(rews, cr.gebv(opops, osim.effects).max(axis=(1, 2)), rtol=RTOL, atol=0)
    assert rews.min() > cr.gebv(germ, osim.effects).mean()  # selection improves the population


def test_vec_wrapper_n_crosses(cuda_device):
    from breedgym_b200.vector import SelectionScores, VecBreedGym

    env = VecBreedGym(num_envs=4, initial_population=GENOME, genetic_map=GMAP, individual_per_gen=200,
                      trait_names=["Yield"])
    wrap_env = SelectionScores(env, k=10, n_crosses=20)
    _, infos = wrap_env.reset(seed=7)
    pop, _, _, _, _ = wrap_env.step(infos["GEBV"].squeeze())
    assert pop.shape[1] == 200
    for k, nc in ((100, 201), (2, 10), (1, 1), (500, 10)):
        with pytest.raises(ValueError):
            SelectionScores(env, k=k, n_crosses=nc)


def test_vec_pair_score_and_ravel_index(cuda_device):
    from breedgym_b200.vector import PairScores, RavelIndex, VecBreedGym

    num_envs, n = 3, 60
    env = gym().make("PairScores", num_envs=num_envs, initial_population=GENOME, genetic_map=GMAP, individual_per_gen=n)
    _, infos = env.reset(seed=7)
    for _ in range(10):
        gebvs = infos["GEBV"].squeeze()
        scores = np.stack([np.add.outer(g, g) for g in gebvs])
        _, rews, _, tru, infos = env.step(scores)
        assert infos["low_level_actions"].shape == (num_envs, n, 2)
    assert np.all(tru) and rews.shape == (num_envs,) and np.all(rews != 0)

    base = PairScores(VecBreedGym(num_envs=2, initial_population=GENOME, genetic_map=GMAP, individual_per_gen=n))
    rav = RavelIndex(base)
    rav.reset(seed=1)
    flat = np.random.randint(0, n * n, size=(2, n))
    pop, _, _, _, _ = rav.step(flat)
    assert pop.shape == (2, n, 1000, 2)
    assert np.array_equal(rav._convert_actions(flat), np.stack([flat // n, flat % n], axis=-1))


def test_sharded_reset_and_step_equal_the_unsharded_env(cuda_device):
    """Multi-GPU contract checked on one GPU: two shards [0,3) and [3,5) of a 5-env run reproduce the
    5-env VecBreedGym exactly (replicated constants, shared cross key, sliced reset keys)."""
    from breedgym_b200.vector import VecBreedGym

    n = 50
    kw = dict(initial_population=GENOME, genetic_map=GMAP, individual_per_gen=n)
    full = VecBreedGym(num_envs=5, **kw)
    a = VecBreedGym(num_envs=3, env_shard=(0, 5), **kw)
    b = VecBreedGym(num_envs=2, env_shard=(3, 5), **kw)
    pf, _ = full.reset(seed=11)
    pa, _ = a.reset(seed=11)
    pb, _ = b.reset(seed=11)
    assert np.array_equal(np.asarray(pf), np.concatenate([np.asarray(pa), np.asarray(pb)]))
    rng = np.random.default_rng(0)
    for _ in range(12):  # crosses an autoreset
        act = rng.integers(0, n, (5, n, 2))
        pf, rf, _, _, _ = full.step(act)
        pa, ra, _, _, _ = a.step(act[:3])
        pb, rb, _, _, _ = b.step(act[3:])
        assert np.array_equal(np.asarray(pf), np.concatenate([np.asarray(pa), np.asarray(pb)]))
        assert np.array_equal(rf, np.concatenate([ra, rb]))
    assert np.array_equal(full.random_key, a.random_key) and np.array_equal(full.random_key, b.random_key)


# ---- tests/test_wrappers.py -----------------------------------------------------------------
def _check_obs(obs, n):
    assert len(obs["GEBV"]) == n and len(obs["corrcoef"]) == n
    assert np.all(obs["corrcoef"] >= -1) and np.all(obs["corrcoef"] <= 1)


def test_simplified_env(cuda_device):
    n = 200
    env = gym().make("breedgym:SimplifiedBreedGym", individual_per_gen=n, initial_population=GENOME, genetic_map=GMAP)
    obs, _ = env.reset()
    _check_obs(obs, n)
    for action in ({"n_bests": 10, "n_crosses": 20}, {"n_bests": 21, "n_crosses": 200}, {"n_bests": 2, "n_crosses": 1}):
        obs, _, _, _, _ = env.step(action)
        _check_obs(obs, n)
    for bad in ({"n_bests": 100, "n_crosses": 201}, {"n_bests": 2, "n_crosses": 10}, {"n_bests": 1, "n_crosses": 1},
                {"n_bests": 500, "n_crosses": 10}):
        with pytest.raises(Exception):
            env.step(bad)


def test_kbest_env_and_policy_matches_oracle(cuda_device):
    n = 200
    env = gym().make("breedgym:KBestBreedGym", individual_per_gen=n, initial_population=GENOME, genetic_map=GMAP,
                     trait_names=["Yield"])
    obs, _ = env.reset(seed=7)
    _check_obs(obs, n)
    germ = np.load(GENOME)
    osim = oracle_sim(env.simulator, 7)
    rng = np.random.Generator(np.random.PCG64(np.random.SeedSequence(7)))
    opop = germ[rng.choice(200, n, replace=False)]
    assert np.array_equal(np.asarray(env.population), opop)
    for _ in range(10):
        obs, r, _, tru, _ = env.step(10)
        _check_obs(obs, n)
        _, best = jp.top_k(cr.gebv(opop, osim.effects)[:, 0].astype(np.float32), 10)
        sel = opop[best]
        pairs = cr.diallel_indices(np.arange(10))
        chosen = rng.choice(len(pairs), 45, replace=False)
        act = np.repeat(pairs[chosen], int(np.ceil(n / 45)), axis=0)[:n]
        opop = osim.cross(sel[act])
        assert np.array_equal(np.asarray(env.population), opop)
        assert np.allclose(obs["corrcoef"], cr.simplified_correlation(opop), rtol=1e-4, atol=1e-5)
    assert tru and abs(r - cr.gebv(opop, osim.effects).mean()) <= RTOL * abs(r)
    for bad in (1, 21):
        with pytest.raises(Exception):
            env.step(bad)


def test_corrcoef_and_select_match_oracle(cuda_device):
    from breedgym_b200.simulator import Simulator

    sim = Simulator(genetic_map=GMAP, device=0, seed=0)
    germ = np.load(GENOME)
    pop = sim.as_packed(germ)
    assert np.allclose(sim.corrcoef(pop), cr.corrcoef(germ), rtol=1e-4, atol=1e-5)
    sel, idx = sim.select(pop, k=12)
    vals = cr.gebv(germ, sim.GEBV_model.marker_effects).sum(-1).astype(np.float32)
    _, ref_idx = jp.top_k(vals, 12)
    assert np.array_equal(idx, ref_idx) and np.array_equal(np.asarray(sel), germ[ref_idx])
    assert np.array_equal(sim._diallel_indices(np.array([4, 2, 9])), cr.diallel_indices(np.array([4, 2, 9])))


def test_wheat_breedgym_shapes(cuda_device):
    from breedgym_b200.vector import VecBreedGym, WheatBreedGym

    n = 40
    env = WheatBreedGym(VecBreedGym(num_envs=2, initial_population=GENOME, genetic_map=GMAP, individual_per_gen=n,
                                    autoreset=False),
                        n_lines=12, plant_per_line=10, k_per_line=5)
    _, infos = env.reset(seed=0)
    for _ in range(10):
        pop, rews, _, tru, infos = env.step(np.random.rand(2, 12, 12))
        assert pop.shape == (2, n, 1000, 2) and infos["GEBV"].shape == (2, n, 1)
    assert np.all(tru) and np.all(rews != 0)
    got = np.asarray(pop)
    assert np.array_equal(got[..., 0], got[..., 1])  # double haploids are homozygous


def test_vec_env_edge_shapes_and_reward_shaping(cuda_device):
    """One env (unique-key kernel path inside bg_vec_step), offspring count different from the parent count,
    reward shaping on every step, device-resident infos, and bad action shapes."""
    import torch

    from breedgym_b200.vector import VecBreedGym

    germ = np.load(GENOME)
    for E in (1, 3):
        env = VecBreedGym(num_envs=E, initial_population=GENOME, genetic_map=GMAP, individual_per_gen=30,
                          reward_shaping=True, num_generations=3)
        pop, _ = env.reset(seed=5)
        osim = oracle_sim(env.simulator, 5)
        _, opops, _ = cr.vec_reset(germ, 30, E, jp.key(5))
        rng = np.random.default_rng(E)
        for n_off in (30, 17, 40):  # the population size follows the number of crosses
            act = rng.integers(0, opops.shape[1], (E, n_off, 2))
            pop, rews, ter, tru, infos = env.step(act)
            opops = cr.vec_step(osim, opops, act)
            g = cr.gebv(opops, osim.effects)
            assert infos["GEBV"].shape == (E, n_off, 1)
            assert np.allclose(rews, g.max(axis=(1, 2)), rtol=RTOL, atol=0)  # shaping: a reward on every step
            if not tru[0]:
                assert np.array_equal(np.asarray(pop), opops)
        assert tru[0] and not ter[0]
        with pytest.raises(ValueError):
            env.step(np.zeros((E + 1, 30, 2), dtype=np.int64))
        with pytest.raises(ValueError):
            env.step(np.zeros((E, 30, 3), dtype=np.int64))
    dev_env = VecBreedGym(num_envs=2, initial_population=GENOME, genetic_map=GMAP, individual_per_gen=20, info_device="device")
    dev_env.reset(seed=1)
    _, rews, _, _, infos = dev_env.step(torch.randint(0, 20, (2, 20, 2), device="cuda"))
    assert isinstance(infos["GEBV"], torch.Tensor) and infos["GEBV"].is_cuda and isinstance(rews, torch.Tensor)


def test_empty_cross_and_single_offspring(cuda_device):
    from breedgym_b200.simulator import Simulator

    sim = Simulator(genetic_map=GMAP, device=0, seed=0)
    pop = sim.as_packed(np.load(GENOME))
    none = sim.cross(pop[np.zeros((0, 2), dtype=np.int64)])
    assert none.shape == (0, 1000, 2) and sim.GEBV(none).shape == (0, 1)
    one = sim.cross(pop[np.array([[3, 3]])])
    assert one.shape == (1, 1000, 2)
    # selfing a fully homozygous parent returns the parent
    homo = np.repeat(np.load(GENOME)[:1, :, :1], 2, axis=2)
    kid = sim.cross(sim.as_packed(homo)[np.array([[0, 0]])])
    assert np.array_equal(np.asarray(kid), homo)


def test_peer_reward_exchange_two_shards_in_one_process(cuda_device):
    """The reward exchange over peer memory (csrc/peer.cu) with both "ranks" on this GPU: shards [0,3) and [3,5) of a
    5-env run publish their rewards from inside bg_vec_step's reduction into both windows; every shard then reads the
    5-env VecBreedGym's rewards (device mode: ragged window; three episode ends, so both window halves are re-used
    and the flow control runs), and the standalone publish gives the same."""
    import torch

    from breedgym_b200.vector import ShardedVecBreedGym, VecBreedGym

    n = 50
    kw = dict(initial_population=GENOME, genetic_map=GMAP, individual_per_gen=n, num_generations=3, info_device="device")
    full = VecBreedGym(num_envs=5, **kw)
    shards = [ShardedVecBreedGym(total_envs=5, world=2, rank=r, async_rewards=True, **kw) for r in range(2)]
    ShardedVecBreedGym.connect_local(shards)
    for s in shards:
        assert s.collective == "peer"
        s._peer.set_timeout_ms(2000)
        s.reset(seed=11)
    full.reset(seed=11)
    rng = np.random.default_rng(0)
    dev = full.device
    n_rewards = 0
    for step in range(10):
        act = torch.from_numpy(rng.integers(0, n, (5, n, 2)).astype(np.int32)).to(dev)
        pf, rf, _, tf, _ = full.step(act)
        outs = [s.step(act[s.local_slice()].contiguous()) for s in shards]
        for s in shards:
            s.wait_rewards()
        assert np.array_equal(np.asarray(pf), np.concatenate([np.asarray(o[0]) for o in outs]))
        for o in outs:
            assert o[1].shape == (5,) and torch.equal(o[1], rf), f"step {step}"
            assert np.array_equal(o[3], tf)
        n_rewards += bool(tf[0])
    assert n_rewards == 3
    # rewards that did not come out of a step: publish + wait
    vals = [torch.arange(3, dtype=torch.float32, device=dev) + 10, torch.arange(2, dtype=torch.float32, device=dev) + 20]
    got = [s.gather_rewards(v) for s, v in zip(shards, vals)]
    for s in shards:
        s.wait_rewards()
    got = [s._window() for s in shards]
    for g in got:
        assert torch.equal(g, torch.cat(vals))
    assert all(s._peer.timeouts() == 0 for s in shards)
    for s in shards:
        s.close()


def test_sharded_env_over_nccl_matches_unsharded(cuda_device):
    """Two ranks on two GPUs (skipped on a single-GPU box): scripts/multi_gpu_check.py under torchrun."""
    import subprocess
    import sys

    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = Path(__file__).resolve().parents[1]
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", str(root / "scripts" / "multi_gpu_check.py")],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "multi-gpu check ok" in r.stdout


def test_reset_infos_gathered_from_germplasm_equal_rescored(cuda_device):
    """VecBreedGym.reset gathers the reset infos from the germplasm's GEBVs (a GEBV is a function of the individual
    alone); re-scoring the drawn populations (reuse_germplasm_gebv=False, what vec_env.py:130 does) gives the same bits."""
    kw = dict(num_envs=5, initial_population=GENOME, genetic_map=GMAP, individual_per_gen=37)
    env_a = gym().make("VecBreedGym", **kw)
    pop_a, infos_a = env_a.reset(seed=11)
    env_b = gym().make("VecBreedGym", **kw)
    env_b.reuse_germplasm_gebv = False
    pop_b, infos_b = env_b.reset(seed=11)
    assert env_a._germ_gebv is not None and env_b._germ_gebv is None
    assert np.array_equal(np.asarray(pop_a), np.asarray(pop_b))
    assert np.array_equal(infos_a["GEBV"], infos_b["GEBV"])
    assert np.array_equal(infos_a["GEBV"], env_a.get_info()["GEBV"])


@pytest.mark.parametrize("k,n_crosses,layout", [(10, None, "legacy"), (10, 20, "legacy"), (7, 3, "partitionable"), (20, 190, "legacy")])
def test_selection_scores_device_index_math_equals_host(cuda_device, k, n_crosses, layout):
    """SelectionScores' GPU translation (stable sort, bg_reset_indices permutation, gathers, repeat) == the NumPy one
    (jaxlike: lax.top_k, random.choice(replace=False), repeat), including tied scores."""
    from breedgym_b200 import _lib
    from breedgym_b200.vector import SelectionScores, VecBreedGym

    num_envs, n = 6, 200
    env = SelectionScores(VecBreedGym(num_envs=num_envs, initial_population=GENOME, genetic_map=GMAP, individual_per_gen=n,
                                      rng_layout=layout), k=k, n_crosses=n_crosses)
    env.reset(seed=3)
    rng = np.random.default_rng(k)
    for trial in range(3):
        scores = rng.standard_normal((num_envs, n)).astype(np.float32)
        scores[:, rng.integers(0, n, 40)] = 0.25  # ties: the lower index wins
        key = np.array(rng.integers(0, 2**32, 2), dtype=np.uint32)
        keys = _lib.key_split(key, num_envs + 1, layout)
        host = env._convert_actions(scores, keys[1:])
        dev = env._convert_actions_device(scores, key).cpu().numpy()
        assert dev.shape == (num_envs, n, 2) and dev.dtype == np.int32
        assert np.array_equal(dev, host)


def test_vec_reseed_mid_episode_discards_lookahead_masks(cuda_device):
    """The masks of the next two steps are generated ahead of time from the key chain; `reset(seed=...)` in the middle of
    an episode changes the chain, so those masks must be dropped (slot lookup is by key) and the trajectory must equal a
    fresh env's under the new seed."""
    num_envs, n = 33, 50  # 33 envs: the fused step kernel with mask rows staged in shared memory
    env = gym().make("VecBreedGym", num_envs=num_envs, initial_population=GENOME, genetic_map=GMAP, individual_per_gen=n)
    germ = np.load(GENOME)
    rng = np.random.default_rng(5)
    env.reset(seed=1)
    for _ in range(3):
        env.step(rng.integers(0, n, (num_envs, n, 2)))
    pop, infos = env.reset(seed=9)
    osim = oracle_sim(env.simulator, 9)
    _, opops, _ = cr.vec_reset(germ, n, num_envs, jp.key(9), "legacy")
    assert np.array_equal(np.asarray(pop), opops)
    for _ in range(4):
        action = rng.integers(0, n, (num_envs, n, 2))
        pop, rews, ter, tru, infos = env.step(action)
        opops = cr.vec_step(osim, opops, action)
        assert np.array_equal(np.asarray(pop), opops)
        assert np.allclose(infos["GEBV"], cr.gebv(opops, osim.effects), rtol=RTOL, atol=0)


def test_vec_device_mode_ring_autoreset_matches_oracle(cuda_device):
    """info_device="device": observations / infos cycle through the preallocated ring (no allocation per step); the
    trajectory across two autoresets and a reseed in the middle of an episode (the mask lookahead then misses and
    regenerates) must equal the oracle's, and a handle must stay valid for one further step."""
    import torch

    num_envs, n = 3, 60
    env = gym().make("VecBreedGym", num_envs=num_envs, initial_population=GENOME, genetic_map=GMAP, individual_per_gen=n,
                     num_generations=3, info_device="device")
    germ = np.load(GENOME)
    rng = np.random.default_rng(2)
    for seed in (7, 8):
        pop, infos = env.reset(seed=seed)
        osim = oracle_sim(env.simulator, seed)
        okey, opops, _ = cr.vec_reset(germ, n, num_envs, jp.key(seed), "legacy")
        assert np.array_equal(np.asarray(pop), opops)
        for step in range(7):
            action = rng.integers(0, n, (num_envs, n, 2))
            prev_words, prev_ref = pop.words, opops
            pop, rews, ter, tru, infos = env.step(torch.from_numpy(action).to(cuda_device))
            # the previous observation's buffer is still intact one step later (ring of 3)
            from breedgym_b200.population import PackedPopulation
            assert np.array_equal(np.asarray(PackedPopulation(env.simulator, prev_words)), prev_ref)
            opops = cr.vec_step(osim, opops, action)
            assert np.allclose(infos["GEBV"].cpu().numpy(), cr.gebv(opops, osim.effects), rtol=RTOL, atol=0)
            if step % 3 == 2:  # autoreset: the returned population is the next episode's first
                assert bool(np.all(tru))
                okey, opops, _ = cr.vec_reset(germ, n, num_envs, okey, "legacy")
                assert np.allclose(env.reset_infos["GEBV"].cpu().numpy(), cr.gebv(opops, osim.effects), rtol=RTOL, atol=0)
            assert np.array_equal(np.asarray(pop), opops)


def test_reference_parent_gather_idiom_on_packed_populations(cuda_device):
    """`populations[arange_envs, actions]` (breedgym/vector/vec_env.py:89-90): a lazy per-env parents view whose
    materialisation equals numpy fancy indexing and whose cross equals `env.cross(actions)`; other tuple indices are
    plain numpy multi-axis indexing (not parent pairs)."""
    from breedgym_b200.population import ParentsView

    E, n = 3, 20
    env = gym().make("VecBreedGym", num_envs=E, initial_population=GENOME, genetic_map=GMAP, individual_per_gen=n)
    pops, _ = env.reset(seed=4)
    host = np.asarray(pops)
    rng = np.random.default_rng(0)
    actions = rng.integers(0, n, (E, 7, 2))
    arange_envs = np.arange(E)[:, None, None]
    parents = pops[arange_envs, actions]
    assert isinstance(parents, ParentsView) and parents.shape == (E, 7, 2, 1000, 2)
    assert np.array_equal(np.asarray(parents), host[arange_envs, actions])
    env.simulator.set_seed(5)
    a = np.asarray(env.cross(parents))
    env.simulator.set_seed(5)
    b = np.asarray(env.cross(actions))
    assert a.shape == (E, 7, 1000, 2) and np.array_equal(a, b)
    # a homogeneous tuple of index arrays is numpy multi-axis indexing, not a parents view
    rows, cols = np.array([0, 2]), np.array([1, 3])
    assert np.array_equal(pops[(rows, cols)], host[(rows, cols)])
    assert np.array_equal(pops[1, 2:4], host[1, 2:4])
    single = pops[0]
    assert np.array_equal(np.asarray(single[[3, 1, 2]]), host[0][[3, 1, 2]])


@pytest.mark.parametrize("rows,length,k", [(3, 144, 12), (5, 1000, 1), (2, 137, 137), (4, 136900, 370), (1, 70000, 1024)])
def test_topk_kernel_matches_lax_top_k_semantics(cuda_device, rows, length, k):
    """bg_topk (radix select + bitonic sort) == jax.lax.top_k: descending, ties -> lower index, -0.0 == +0.0; including
    rows made of few distinct values (ties at the threshold) and an all-equal row."""
    import torch

    from breedgym_b200 import _lib
    from breedgym_b200.simulator import Simulator

    sim = Simulator(genetic_map=GMAP, device=0, seed=0)
    rng = np.random.default_rng(rows * 7 + k)
    x = rng.standard_normal((rows, length)).astype(np.float32)
    x[0] = np.round(x[0] * 2) / 2          # heavy ties
    if rows > 1:
        x[1] = 0.0                          # all equal: the k lowest indices, in order
        x[1, ::7] = -0.0
    if rows > 2:
        x[2, : length // 2] = -np.abs(x[2, : length // 2])  # negatives
    xs = torch.from_numpy(x).to(cuda_device)
    vals = torch.empty((rows, k), dtype=torch.float32, device=cuda_device)
    idx = torch.empty((rows, k), dtype=torch.int32, device=cuda_device)
    _lib.check(_lib.load().bg_topk(sim._engine, xs.data_ptr(), rows, length, k, vals.data_ptr(), idx.data_ptr(), sim._stream()))
    torch.cuda.synchronize()
    for r in range(rows):
        rv, ri = jp.top_k(x[r] + np.float32(0.0), k)
        assert np.array_equal(idx[r].cpu().numpy(), ri), f"row {r}: indices differ from lax.top_k"
        assert np.array_equal(vals[r].cpu().numpy(), rv)


@pytest.mark.parametrize("E,k,n", [(4, 60, 60), (3, 370, 370), (2, 1024, 1500), (5, 1, 9), (2, 7, 10)])
def test_pairs_from_topk_kernel_matches_softmax_repeat(cuda_device, E, k, n):
    """bg_pairs_from_topk (csrc/pairs.cu) == PairScores._convert_actions after the top-k (vec_wrappers.py:100-112):
    ceil(softmax(values) * k) offspring per pair, `jnp.repeat(..., total_repeat_length=k)`, (flat // n, flat % n);
    including a flat row (every pair once), one dominant pair (everything goes to it) and heavy ties."""
    import torch

    from breedgym_b200 import _lib
    from breedgym_b200.simulator import Simulator

    sim = Simulator(genetic_map=GMAP, device=0, seed=0)
    rng = np.random.default_rng(E * 1000 + k)
    vals = -np.sort(-rng.standard_normal((E, k)).astype(np.float32) * 3, axis=1)  # descending, as bg_topk returns them
    vals[0] = 1.25                                                                  # flat: softmax = 1 / k exactly
    if E > 1:
        vals[1, 0] = vals[1, 0] + 60.0                                              # one pair takes every slot
    if E > 2:
        vals[2] = -np.sort(-np.round(vals[2]))                                      # ties
    idx = np.stack([rng.permutation(n * n)[:k] for _ in range(E)]).astype(np.int32)
    out = torch.empty((E, k, 2), dtype=torch.int32, device=cuda_device)
    v_d, i_d = torch.from_numpy(vals).to(cuda_device), torch.from_numpy(idx).to(cuda_device)
    _lib.check(_lib.load().bg_pairs_from_topk(sim._engine, v_d.data_ptr(), i_d.data_ptr(), E, k, n, out.data_ptr(), sim._stream()))
    got = out.cpu().numpy()
    for e in range(E):
        ex = np.exp(vals[e] - vals[e].max())
        sm = ex / ex.sum(dtype=np.float32)
        reps = np.ceil(sm * np.float32(k)).astype(np.int32)
        # (a count within one float32 ulp of an integer may round either way between exp implementations: compare the
        #  slots whose owner does not depend on such a count)
        safe = np.abs(sm.astype(np.float64) * k - np.round(sm.astype(np.float64) * k)) > 1e-4
        ref = jp.repeat_total(np.stack((idx[e] // n, idx[e] % n), 1), reps, k)
        if safe.all() or e == 0:
            assert np.array_equal(got[e], ref), f"env {e}"
        else:
            assert (got[e] == ref).all(axis=1).mean() > 0.9
    assert np.array_equal(got[0], np.stack((idx[0] // n, idx[0] % n), 1))           # flat row: every pair exactly once
    if E > 1:
        assert np.all(got[1] == np.array([idx[1, 0] // n, idx[1, 0] % n]))


@pytest.mark.parametrize("E,k,nc,n", [(3, 10, 20, 200), (2, 5, 10, 7), (4, 30, 435, 500), (1, 2, 1, 4)])
def test_diallel_pairs_kernel_matches_triu_repeat(cuda_device, E, k, nc, n):
    """bg_diallel_pairs (csrc/pairs.cu) == SelectionScores._convert_actions after top-k and choice
    (vec_wrappers.py:60-78): `_diallel_indices(best)[perm]`, each repeated ceil(n / nc) times, total_repeat_length n."""
    import torch

    from breedgym_b200 import _lib
    from breedgym_b200.simulator import Simulator

    sim = Simulator(genetic_map=GMAP, device=0, seed=0)
    rng = np.random.default_rng(E + 10 * k)
    best = np.stack([rng.permutation(1000)[:k] for _ in range(E)]).astype(np.int32)
    n_pairs = k * (k - 1) // 2
    perm = np.stack([rng.permutation(n_pairs)[:nc] for _ in range(E)]).astype(np.int32)
    out = torch.empty((E, n, 2), dtype=torch.int32, device=cuda_device)
    b_d, p_d = torch.from_numpy(best).to(cuda_device), torch.from_numpy(perm).to(cuda_device)
    _lib.check(_lib.load().bg_diallel_pairs(sim._engine, b_d.data_ptr(), p_d.data_ptr(), E, k, nc, n, out.data_ptr(), sim._stream()))
    got = out.cpu().numpy()
    ia, ib = np.triu_indices(k, k=1)
    rep = -(-n // nc)
    for e in range(E):
        chosen = np.stack([best[e][ia], best[e][ib]], axis=1)[perm[e]]
        ref = jp.repeat_total(chosen, np.full(nc, rep, dtype=np.int32), n)
        assert np.array_equal(got[e], ref), f"env {e}"


def test_phenotype_gxe_matches_oracle(cuda_device):
    """Simulator.phenotype / create_environments / GxE_model (chromax, scripts/time_wheat.py:17-50) against the oracle:
    GxE effects from the constructor's split key, environments from the key chain, mean over environments of
    GEBV + env * GxE in float64 (rtol 1e-5), and selection by `phenotype_index`."""
    from breedgym_b200.simulator import Simulator
    from breedgym_b200.utils.index_functions import phenotype_index

    h2 = np.array([0.3], dtype=np.float32)
    sim = Simulator(genetic_map=GMAP, trait_names=["Yield"], seed=11, device=0, h2=h2)
    pop = np.load(GENOME)[:64]
    key0 = jp.key(11)
    halves = jp.split(key0, 2)
    gxe = cr.gxe_effects(sim.GEBV_model.marker_effects, halves[1], h2)
    assert np.array_equal(sim.GxE_model.marker_effects, gxe)
    nxt = jp.split(halves[0], 2)                       # create_environments: random_key, k = split(random_key)
    envs = sim.create_environments(5)
    assert np.array_equal(envs, jp.normal(nxt[1], 5))
    assert np.array_equal(sim.random_key, nxt[0])
    got = sim.phenotype(pop, environments=envs).cpu().numpy()
    ref = cr.phenotype(pop, sim.GEBV_model.marker_effects, gxe, envs)
    assert got.shape == (64, 1) and np.allclose(got, ref, rtol=RTOL, atol=1e-4 * np.abs(ref).max())
    with pytest.raises(ValueError):
        sim.phenotype(pop, num_environments=2, environments=envs)
    one = sim.phenotype(pop)                            # draws one environment from the key chain
    e1 = jp.normal(jp.split(nxt[0], 2)[1], 1)
    assert np.allclose(one.cpu().numpy(), cr.phenotype(pop, sim.GEBV_model.marker_effects, gxe, e1), rtol=RTOL,
                       atol=1e-4 * np.abs(ref).max())
    sel, idx = sim.select(pop, 7, phenotype_index(sim, envs))
    assert np.array_equal(idx, jp.top_k(got[:, 0], 7)[1])
    assert np.array_equal(np.asarray(sel), pop[idx])


def test_pinned_host_actions_equal_numpy_actions(cuda_device):
    """A pinned int32 host tensor is copied to the device as it is (no staging copy): same trajectory as numpy actions,
    in host and in device mode."""
    import torch

    from breedgym_b200.vector import VecBreedGym

    n = 40
    for mode in ("host", "device"):
        kw = dict(num_envs=3, initial_population=GENOME, genetic_map=GMAP, individual_per_gen=n, num_generations=3, info_device=mode)
        a, b = VecBreedGym(**kw), VecBreedGym(**kw)
        a.reset(seed=4)
        b.reset(seed=4)
        rng = np.random.default_rng(1)
        for _ in range(5):
            act = rng.integers(0, n, (3, n, 2)).astype(np.int32)
            pa, ra, _, ta, ia = a.step(act)
            pinned = torch.from_numpy(act.copy()).pin_memory()
            pb, rb, _, tb, ib = b.step(pinned)
            torch.cuda.synchronize()
            assert np.array_equal(np.asarray(pa), np.asarray(pb))
            ga, gb = ia["GEBV"], ib["GEBV"]
            assert np.array_equal(ga.cpu().numpy() if hasattr(ga, "cpu") else ga, gb.cpu().numpy() if hasattr(gb, "cpu") else gb)
            assert np.array_equal(ra.cpu().numpy() if hasattr(ra, "cpu") else ra, rb.cpu().numpy() if hasattr(rb, "cpu") else rb)
            assert np.array_equal(ta, tb)
