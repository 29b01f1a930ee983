"""The reference's own golden rewards (the only results its test-suite pins at the chromax boundary):

    tests/test_env.py:103-119        BreedGym, 10 fixed crosses x 10 generations, seed 7          -> 1.8197979
    tests/test_wrappers.py:73-88     KBestBreedGym, k = 10, 10 generations, seed 7                 -> 18.826467
    tests/test_vec.py:115-133        VecBreedGym, 4 envs x 200, 20 random steps, seed 7            -> 4 rewards
    tests/test_vec.py:136-154        SelectionScores(k=10), GEBV policy, 10 steps, seed 7          -> 4 rewards
    tests/test_vec.py:186-204        PairScores, GEBV outer-sum policy, 10 steps, seed 7           -> 4 rewards

All five need `chromax.sample_data` (`genome.npy`, `genetic_map.txt`), which is neither vendored in the reference nor
installable on this image, so they SKIP until the two files are available -- through an importable
`chromax`, `$CHROMAX_SAMPLE_DATA`, or `tests/golden/chromax_sample_data/` -- and go live the moment they are.  They were
produced by real jax + chromax, so a pass pins the oracle's reading of the PRNG layout and key schedule
(`rng_layout` / `key_schedule`); scripts/pin_with_chromax.py generates genotype-level fixtures on a machine that has both.
"""
import os
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = Path(__file__).resolve().parents[1]


def _sample_data():
    try:
        from chromax.sample_data import genetic_map, genome  # type: ignore

        return Path(genome), Path(genetic_map)
    except Exception:
        pass
    for d in (os.environ.get("CHROMAX_SAMPLE_DATA"), ROOT / "tests" / "golden" / "chromax_sample_data"):
        if d and (Path(d) / "genome.npy").exists() and (Path(d) / "genetic_map.txt").exists():
            return Path(d) / "genome.npy", Path(d) / "genetic_map.txt"
    return None


@pytest.fixture(scope="module")
def sample(cuda_device):
    found = _sample_data()
    if found is None:
        pytest.skip("chromax.sample_data (genome.npy, genetic_map.txt) is not available: the reference's golden rewards "
                    "cannot be evaluated (set CHROMAX_SAMPLE_DATA or drop the files into tests/golden/chromax_sample_data/)")
    return found


def gym():
    from breedgym_b200 import gym_compat

    return gym_compat


def test_env_deterministic_golden(sample):
    genome, genetic_map = sample
    env = gym().make("breedgym:BreedGym", initial_population=genome, genetic_map=genetic_map, reward_shaping=False)
    env.reset(seed=7)
    action = np.array([[1, 2], [1, 5], [1, 7], [2, 5], [2, 9], [4, 7], [4, 8], [5, 9], [6, 8], [6, 9]])
    for _ in range(10):
        _, r, _, _, _ = env.step(action)
    assert abs(r - 1.8197979) < 1e-6


def test_kbest_gebv_policy_golden(sample):
    genome, genetic_map = sample
    env = gym().make("breedgym:KBestBreedGym", individual_per_gen=200, initial_population=genome, genetic_map=genetic_map,
                     trait_names=["Yield"])
    env.reset(seed=7)
    for _ in range(10):
        _, r, _, _, _ = env.step(10)
    assert abs(r - 18.826467) < 1e-5


def test_vec_deterministic_golden(sample):
    genome, genetic_map = sample
    num_envs, n = 4, 200
    env = gym().make("VecBreedGym", num_envs=num_envs, initial_population=genome, genetic_map=genetic_map, individual_per_gen=n)
    np.random.seed(seed=7)
    pop, _ = env.reset(seed=7)
    for _ in range(20):
        action = np.random.randint(len(pop), size=(num_envs, n, 2))
        pop, rews, _, _, _ = env.step(action)
    assert np.allclose(rews, np.array([10.844662, 8.436224, 8.759013, 9.1480465]))


def test_vec_gebv_policy_golden(sample):
    genome, genetic_map = sample
    env = gym().make("SelectionScores", k=10, num_envs=4, initial_population=genome, genetic_map=genetic_map,
                     individual_per_gen=200, trait_names=["Yield"])
    _, infos = env.reset(seed=7)
    for _ in range(10):
        _, rews, _, _, infos = env.step(np.asarray(infos["GEBV"]).squeeze())
    assert np.allclose(rews, np.array([18.514063, 19.428415, 19.090204, 19.841211]))


def test_vec_pair_score_golden(sample):
    genome, genetic_map = sample
    env = gym().make("PairScores", num_envs=4, initial_population=genome, genetic_map=genetic_map, individual_per_gen=200)
    _, infos = env.reset(seed=7)
    for _ in range(10):
        gebvs = np.asarray(infos["GEBV"]).squeeze()
        _, rews, _, _, infos = env.step(np.add.outer(gebvs, gebvs))
    assert np.allclose(rews, np.array([8.650201, 9.414795, 8.115064, 10.0190525]))
