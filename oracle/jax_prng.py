"""NumPy restatement of the parts of `jax.random` the reference path uses.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The reference draws every random number through `jax.random` (call sites:
`breedgym/vector/vec_env.py:115,120` -> key/split, `vec_env.py:22-27` ->
choice(replace=False), `breedgym/vector/vec_wrappers.py:70-72,82` -> choice /
split, and inside chromax `uniform`/`split`).  JAX is a third-party dependency
absent from /root/reference and un-pinned (`pyproject.toml:24`); its published
algorithm (jax/_src/prng.py: threefry_2x32, threefry_split,
threefry_random_bits; jax/_src/random.py: _uniform, _shuffle, choice) is
restated here.  Two bit layouts exist:

  * "legacy"        jax_threefry_partitionable=False (default before jax 0.5.0;
                    the window the reference's gymnasium-0.29 pin implies)
  * "partitionable" jax_threefry_partitionable=True  (default from jax 0.5.0)
"""
from __future__ import annotations

import math

import numpy as np

LAYOUTS = ("legacy", "partitionable")

_ROT = ((13, 15, 26, 6), (17, 29, 16, 24))
_U32 = np.uint32


def _rotl(x, r):
    return (x << _U32(r)) | (x >> _U32(32 - r))


def threefry2x32(k0, k1, x0, x1):
    """Threefry-2x32, 20 rounds (Random123).  All args uint32 scalars/arrays.

    Follows jax/_src/prng.py `_threefry2x32_lowering` / `threefry2x32_p`:
    key schedule ks = [k0, k1, k0^k1^0x1BD11BDA]; five groups of four rounds
    with rotation sets (13,15,26,6)/(17,29,16,24) alternating; after group g
    inject ks[(g+1)%3], ks[(g+2)%3] + (g+1).
    """
    with np.errstate(over="ignore"):
        k0 = np.asarray(k0, dtype=_U32)
        k1 = np.asarray(k1, dtype=_U32)
        x0 = np.array(x0, dtype=_U32, copy=True)
        x1 = np.array(x1, dtype=_U32, copy=True)
        ks = (k0, k1, k0 ^ k1 ^ _U32(0x1BD11BDA))
        x0 = x0 + ks[0]
        x1 = x1 + ks[1]
        for g in range(5):
            for r in _ROT[g % 2]:
                x0 = x0 + x1
                x1 = _rotl(x1, r)
                x1 = x1 ^ x0
            x0 = x0 + ks[(g + 1) % 3]
            x1 = x1 + ks[(g + 2) % 3] + _U32(g + 1)
    return x0, x1


def key(seed: int) -> np.ndarray:
    """`jax.random.key(seed)` / `PRNGKey(seed)` raw data: (hi32, lo32)."""
    seed = int(seed)
    if seed < 0:
        seed += 1 << 64
    return np.array([(seed >> 32) & 0xFFFFFFFF, seed & 0xFFFFFFFF], dtype=_U32)


def random_bits(k, n: int, layout: str = "legacy") -> np.ndarray:
    """`jax.random.bits(key, (n,), uint32)` -- threefry_random_bits, 32-bit."""
    k = np.asarray(k, dtype=_U32)
    n = int(n)
    if n == 0:
        return np.zeros(0, dtype=_U32)
    if layout == "legacy":
        # iota(n); odd n is padded with ONE zero counter; first half of the
        # counters goes to x0, second half to x1; outputs are concatenated.
        c = np.arange(n, dtype=_U32)
        if n % 2:
            c = np.concatenate([c, np.zeros(1, dtype=_U32)])
        h = c.size // 2
        a, b = threefry2x32(k[0], k[1], c[:h], c[h:])
        return np.concatenate([a, b])[:n]
    if layout == "partitionable":
        # one block per element: counter = 64-bit flat index (hi, lo); out = x0^x1
        idx = np.arange(n, dtype=np.uint64)
        hi = (idx >> np.uint64(32)).astype(_U32)
        lo = (idx & np.uint64(0xFFFFFFFF)).astype(_U32)
        a, b = threefry2x32(k[0], k[1], hi, lo)
        return a ^ b
    raise ValueError(layout)


def split(k, num: int = 2, layout: str = "legacy") -> np.ndarray:
    """`jax.random.split(key, num)` raw key data, shape (num, 2)."""
    k = np.asarray(k, dtype=_U32)
    num = int(num)
    if layout == "legacy":
        return random_bits(k, 2 * num, "legacy").reshape(num, 2)
    if layout == "partitionable":
        idx = np.arange(num, dtype=np.uint64)
        hi = (idx >> np.uint64(32)).astype(_U32)
        lo = (idx & np.uint64(0xFFFFFFFF)).astype(_U32)
        a, b = threefry2x32(k[0], k[1], hi, lo)
        return np.stack([a, b], axis=1)
    raise ValueError(layout)


def bits_to_uniform(bits: np.ndarray) -> np.ndarray:
    """jax/_src/random.py `_uniform` for float32, minval=0, maxval=1."""
    f = ((bits >> _U32(9)) | _U32(0x3F800000)).view(np.float32) - np.float32(1.0)
    return np.maximum(np.float32(0.0), f)


def uniform(k, n: int, layout: str = "legacy") -> np.ndarray:
    """`jax.random.uniform(key, (n,), float32)`."""
    return bits_to_uniform(random_bits(k, n, layout))


def erf_inv_f32(x: np.ndarray) -> np.ndarray:
    """`lax.erf_inv` in float32 as XLA expands it (M. Giles, "Approximating the erfinv function": two degree-8
    polynomials in w = -log1p(-x*x)); evaluated here in float32 Horner form."""
    x = np.asarray(x, dtype=np.float32)
    w = (-np.log1p((-x * x).astype(np.float32))).astype(np.float32)
    lt = w < np.float32(5.0)
    w = np.where(lt, w - np.float32(2.5), np.sqrt(w, dtype=np.float32) - np.float32(3.0)).astype(np.float32)
    c_lt = [2.81022636e-08, 3.43273939e-07, -3.5233877e-06, -4.39150654e-06, 0.00021858087, -0.00125372503, -0.00417768164,
            0.246640727, 1.50140941]
    c_ge = [-0.000200214257, 0.000100950558, 0.00134934322, -0.00367342844, 0.00573950773, -0.0076224613, 0.00943887047,
            1.00167406, 2.83297682]
    p = np.where(lt, np.float32(c_lt[0]), np.float32(c_ge[0])).astype(np.float32)
    for a, b in zip(c_lt[1:], c_ge[1:]):
        p = (np.where(lt, np.float32(a), np.float32(b)) + p * w).astype(np.float32)
    out = (p * x).astype(np.float32)
    return np.where(np.abs(x) == 1, np.copysign(np.float32(np.inf), x), out).astype(np.float32)


def normal(k, n: int, layout: str = "legacy") -> np.ndarray:
    """`jax.random.normal(key, (n,), float32)`: sqrt(2) * erf_inv(uniform(key, minval=nextafter(-1, 0), maxval=1))."""
    lo = np.nextafter(np.float32(-1.0), np.float32(0.0))
    f = bits_to_uniform(random_bits(k, n, layout))  # [0, 1)
    u = np.maximum(lo, (f * (np.float32(1.0) - lo) + lo).astype(np.float32))
    return (np.float32(np.sqrt(2)) * erf_inv_f32(u)).astype(np.float32)


def shuffle_rounds(size: int) -> int:
    """Number of sort rounds in jax/_src/random.py `_shuffle`."""
    uint32max = np.iinfo(np.uint32).max
    return int(np.ceil(3 * np.log(max(1, size)) / np.log(uint32max)))


def permutation(k, n: int, layout: str = "legacy") -> np.ndarray:
    """`jax.random.permutation(key, n)`: repeated stable sort by fresh keys."""
    k = np.asarray(k, dtype=_U32)
    x = np.arange(n, dtype=np.int64)
    for _ in range(shuffle_rounds(n)):
        ks = split(k, 2, layout)
        k, sub = ks[0], ks[1]
        sort_keys = random_bits(sub, n, layout)
        x = x[np.argsort(sort_keys, kind="stable")]
    return x


def choice_no_replace(k, n_inputs: int, n_draws: int, layout: str = "legacy") -> np.ndarray:
    """`jax.random.choice(key, n_inputs, (n_draws,), replace=False)` indices."""
    if n_draws > n_inputs:
        raise ValueError("Cannot take a larger sample than population when 'replace=False'")
    return permutation(k, n_inputs, layout)[:n_draws]


def top_k(x: np.ndarray, k: int):
    """`jax.lax.top_k`: descending values, ties -> lower index first."""
    x = np.asarray(x)
    order = np.argsort(-x, kind="stable")[:k]
    return x[order], order


def repeat_total(x: np.ndarray, repeats, total: int) -> np.ndarray:
    """`jnp.repeat(x, repeats, axis=0, total_repeat_length=total)`.

    Truncates when the repeats overflow; pads by repeating the final entry
    (jnp docs) when they fall short.
    """
    x = np.asarray(x)
    out = np.repeat(x, repeats, axis=0)
    if len(out) >= total:
        return out[:total]
    pad = np.repeat(x[-1:], total - len(out), axis=0)
    return np.concatenate([out, pad], axis=0)


def threshold_u32(r) -> np.ndarray:
    """Integer form of `uniform < r`:  u < r  <=>  (bits >> 9) < T(r).

    u = (bits>>9) * 2^-23 exactly, so T = clamp(ceil(r * 2^23), 0, 2^23).
    Used by tests to cross-check the kernel's threshold table; the oracle's
    meiosis itself compares floats exactly as JAX does.
    """
    r = np.asarray(r, dtype=np.float32).astype(np.float64)
    t = np.ceil(r * float(1 << 23))
    t = np.where(np.isnan(t), 0.0, t)
    return np.clip(t, 0, float(1 << 23)).astype(np.uint32)
