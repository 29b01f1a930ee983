"""CPU-side checks of the product: the C-ABI library loads and exports the header's symbols,
the host key chain / index math match the oracle, the Gymnasium shim and the sharding plumbing
behave.  No kernel is launched here (no GPU in the build container)."""
import ctypes
import os
import re
import socket
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]


def test_library_exports_every_declared_symbol():
    from breedgym_b200 import _lib

    header = (ROOT / "include" / "breedgym_b200.h").read_text()
    declared = set(re.findall(r"\b(bg_[a-z0-9_]+)\s*\(", header))
    declared.discard("bg_engine")
    assert declared, "no declarations parsed"
    lib = _lib.load()
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/breedgym_b200.h but not exported"
    assert declared == set(_lib._SIGNATURES), "ctypes signatures out of sync with the header"
    assert lib.bg_version() == 210


def test_gpu_calls_fail_loudly_without_a_device():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from breedgym_b200 import _lib

    eng = ctypes.c_void_p()
    rc = _lib.load().bg_engine_create(0, ctypes.byref(eng))
    assert rc < 0
    with pytest.raises(RuntimeError):
        _lib.check(rc)


def test_host_key_chain_matches_oracle(golden):
    from breedgym_b200 import _lib
    from oracle import jax_prng as jp

    for layout in ("legacy", "partitionable"):
        assert np.array_equal(_lib.key_split(_lib.key_data(99), 5, layout), golden[f"split5_{layout}"])
        assert np.array_equal(_lib.random_bits(_lib.key_data(99), 9, layout), golden[f"bits9_{layout}"])
        for n in (1, 2, 3, 64, 777):
            assert np.array_equal(_lib.random_bits(_lib.key_data(5), n, layout), jp.random_bits(jp.key(5), n, layout))
            assert np.array_equal(_lib.key_split(_lib.key_data(5), n, layout), jp.split(jp.key(5), n, layout))
        for q in (0, 3, 64):
            assert np.array_equal(_lib.key_split_at(_lib.key_data(5), q, 65, layout), jp.split(jp.key(5), 65, layout)[q])
        st = _lib.key_data(21).copy()
        out = np.zeros(6, dtype=np.uint32)
        _lib.check(_lib.load().bg_key_chain_next(_lib.nptr(st), _lib.LAYOUT_ID[layout], _lib.nptr(out)))
        ks = jp.split(jp.key(21), 2, layout)
        assert np.array_equal(st, ks[0]) and np.array_equal(out[:2], ks[1])
        ks2 = jp.split(ks[0], 2, layout)
        assert np.array_equal(out[2:4], ks2[1])                          # the key of the next call
        assert np.array_equal(out[4:], jp.split(ks2[0], 2, layout)[1])   # and of the one after it
    assert _lib.key_data((3 << 32) + 4).tolist() == [3, 4]
    out = (ctypes.c_uint32 * 2)()
    _lib.load().bg_threefry2x32(0x13198A2E, 0x03707344, 0x243F6A88, 0x85A308D3, out)
    assert list(out) == [0xC4923A9C, 0x483DF7A0]


def test_thresholds_match_float_compare():
    from breedgym_b200 import _lib
    from oracle import jax_prng as jp

    r = np.array([0.0, -1.0, np.nan, 1e-9, 2.0**-23, 1.5e-3, 0.1, 0.5, 0.9999999, 1.0, 7.0], dtype=np.float32)
    t = _lib.thresholds(r)
    assert np.array_equal(t, jp.threshold_u32(r))
    assert t[0] == 0 and t[1] == 0 and t[2] == 0 and t[7] == 1 << 22 and t[9] == 1 << 23 and t[10] == 1 << 23
    bits = np.random.default_rng(0).integers(0, 2**32, 50000, dtype=np.uint64).astype(np.uint32)
    u = jp.bits_to_uniform(bits)
    for ri, ti in zip(r, t):
        assert np.array_equal(u < ri, (bits >> np.uint32(9)) < ti)


def test_words_per_row():
    from breedgym_b200 import _lib

    assert [_lib.words_per_row(m) for m in (1, 32, 33, 128, 129, 10000, 1000000)] == [32, 32, 32, 32, 32, 320, 31264]


def test_jaxlike_matches_oracle(golden):
    from breedgym_b200 import _lib, jaxlike
    from oracle import jax_prng as jp

    for layout in ("legacy", "partitionable"):
        assert np.array_equal(jaxlike.permutation(_lib.key_data(11), 2000, layout), golden[f"perm2000_{layout}"])
        for n, k in ((45, 20), (10, 10), (1, 1)):
            assert np.array_equal(jaxlike.choice_no_replace(_lib.key_data(3), n, k, layout),
                                  jp.choice_no_replace(jp.key(3), n, k, layout))
        keys = jp.split(jp.key(8), 4, layout)
        for n in (45, 2000):  # one and two shuffle rounds
            batch = jaxlike.permutation_batch(keys, n, layout)
            for e in range(4):
                assert np.array_equal(batch[e], jp.permutation(keys[e], n, layout))
    with pytest.raises(ValueError):
        jaxlike.choice_no_replace(_lib.key_data(3), 1, 10)
    x = np.array([[1.0, 3.0, 3.0, 2.0], [0.0, -1.0, 5.0, 5.0]])
    v, i = jaxlike.top_k(x, 2)
    assert i.tolist() == [[1, 2], [2, 3]]
    a = np.arange(6).reshape(3, 2)
    assert np.array_equal(jaxlike.repeat_total(a, np.array([0, 3, 1]), 4), jp.repeat_total(a, np.array([0, 3, 1]), 4))
    assert np.array_equal(jaxlike.repeat_total(a, 1, 5), jp.repeat_total(a, 1, 5))
    s = jaxlike.softmax_f32(np.array([1.0, 2.0, 3.0]))
    assert s.dtype == np.float32 and abs(s.sum() - 1) < 1e-6


def test_pair_scores_conversion_matches_per_env_composition():
    """PairScores' batched conversion (torch; runs on the GPU in production) == top_k / softmax / repeat per env."""
    import torch

    from breedgym_b200 import jaxlike
    from breedgym_b200.vector.vec_wrappers import _pairs_from_scores
    from oracle import jax_prng as jp

    rng = np.random.default_rng(0)
    n = 12
    sc = rng.standard_normal((3, n, n)).astype(np.float32)
    sc[1, 2, 3] = sc[1, 0, 0] = sc[1].max() + 1  # a tie for the top pair: lower flat index first
    sc[2, 5:9, :] = 0.0
    sc[2, 6, 1:4] = -0.0  # -0.0 == +0.0: still ordered by flat index
    sc[2, 0, :] = -1.5    # negative ties
    got = _pairs_from_scores(torch.from_numpy(sc), n, "cpu").numpy()
    for e in range(3):
        v, i = jp.top_k(sc[e].reshape(-1), n)
        reps = np.ceil(jaxlike.softmax_f32(v) * np.float32(n)).astype(np.int32)
        assert np.array_equal(got[e], jp.repeat_total(np.stack((i // n, i % n), 1), reps, n))


def test_gym_shim_semantics():
    from breedgym_b200 import gym_compat as gc

    if gc.HAVE_GYMNASIUM:
        pytest.skip("real gymnasium installed")
    seq = gc.spaces.Sequence(gc.spaces.Tuple((gc.spaces.Discrete(5), gc.spaces.Discrete(5))))
    seq.seed(0)
    lens = []
    for _ in range(200):
        s = seq.sample()
        lens.append(len(s))
        a = np.asarray(s)
        assert a.ndim == 2 and a.shape[1] == 2 and a.min() >= 0 and a.max() < 5
    assert min(lens) >= 1 and 2.5 < np.mean(lens) < 6  # Geometric(0.25): mean 4
    d = gc.spaces.Discrete(3, start=2)
    assert all(d.sample() in (2, 3, 4) for _ in range(50)) and d.contains(4) and not d.contains(5)

    class E(gc.Env):
        pass

    e = E()
    e.reset(seed=5)
    a = e.np_random.integers(1000)
    ref = np.random.Generator(np.random.PCG64(np.random.SeedSequence(5))).integers(1000)
    assert a == ref

    class W(gc.Wrapper):
        pass

    e.population = "x"
    e.observation_space = gc.spaces.Box(0, 1, (2,))
    w = W(e)
    assert w.population == "x" and w.unwrapped is e and w.observation_space is e.observation_space
    w.observation_space = gc.spaces.Box(0, 1, (3,))
    assert w.observation_space.shape == (3,) and e.observation_space.shape == (2,)
    with pytest.raises(KeyError):
        gc.make("NoSuchEnv")


def test_registered_ids_and_alias_package():
    import breedgym  # noqa: F401  (alias package)
    from breedgym.vector import PairScores, RavelIndex, SelectionScores, VecBreedGym, WheatBreedGym  # noqa: F401
    from breedgym_b200 import gym_compat as gc

    if not gc.HAVE_GYMNASIUM:
        for i in ("BreedGym", "SimplifiedBreedGym", "KBestBreedGym", "VecBreedGym", "SelectionScores", "PairScores"):
            assert i in gc._REGISTRY


def test_shard_ranges_cover_and_balance():
    from breedgym_b200.vector.sharded import shard_counts, shard_range

    for total, world in ((4096, 8), (64, 8), (10, 4), (3, 8), (0, 2)):
        spans = [shard_range(total, world, r) for r in range(world)]
        assert spans[0][0] == 0 and sum(c for _, c in spans) == total
        for (b0, c0), (b1, _) in zip(spans, spans[1:]):
            assert b0 + c0 == b1
        counts = shard_counts(total, world)
        assert max(counts) - min(counts) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _gloo_worker(rank, world, port, total, q):
    import torch
    import torch.distributed as dist

    from breedgym_b200.vector.sharded import allgather_rewards, shard_counts, shard_range

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    begin, count = shard_range(total, world, rank)
    local = torch.arange(begin, begin + count, dtype=torch.float32) * 1.5
    full = allgather_rewards(local, shard_counts(total, world))
    q.put((rank, full.tolist()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("total", [8, 7])
def test_reward_allgather_world2_gloo(total):
    """The only collective on the path: per-env rewards all-gathered across shards (equal and ragged)."""
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expect = [1.5 * i for i in range(total)]
    for _, full in got:
        assert full == expect


def test_bench_reference_arm_prints_one_json_line():
    """bench.py contract: ONE JSON line on stdout (library chatter goes to stderr); the reference arm runs without a GPU."""
    import json
    import subprocess
    import sys
    from pathlib import Path

    root = Path(__file__).resolve().parents[1]
    r = subprocess.run([sys.executable, str(root / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=300, cwd=str(root))
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "env_steps_per_sec" and d["unit"] == "env-steps/s"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["n_gpus"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["config"]["workload"].startswith("C2 vector env")
