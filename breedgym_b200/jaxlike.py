"""Host-side index math with `jax.random` / `jax.lax` semantics.

The action wrappers of the reference translate scores into (parent, parent)
index pairs with `jax.lax.top_k`, `jax.random.choice(replace=False)`,
`jnp.repeat(total_repeat_length=...)` and `jax.nn.softmax`
(breedgym/vector/vec_wrappers.py:61-79,100-112, breedgym/wrappers.py:80-85).
These are O(n) .. O(n^2) integer/float ops on a few hundred elements per env;
they run on the host with NumPy, drawing random bits from the library's
Threefry (`bg_random_bits`, `bg_key_split`) so the streams match jax's.
"""
from __future__ import annotations

import numpy as np

from . import _lib


def shuffle_rounds(size: int) -> int:
    return int(np.ceil(3 * np.log(max(1, size)) / np.log(np.iinfo(np.uint32).max)))


def permutation(key: np.ndarray, n: int, layout: str = "legacy") -> np.ndarray:
    """`jax.random.permutation(key, n)`: repeated stable sort by fresh 32-bit keys."""
    x = np.arange(n, dtype=np.int64)
    for _ in range(shuffle_rounds(n)):
        ks = _lib.key_split(key, 2, layout)
        key, sub = ks[0], ks[1]
        x = x[np.argsort(_lib.random_bits(sub, n, layout), kind="stable")]
    return x


def permutation_batch(keys: np.ndarray, n: int, layout: str = "legacy") -> np.ndarray:
    """`vmap(jax.random.permutation)(keys, n)` for keys `[E, 2]` -> int64 `[E, n]`, vectorised over E."""
    keys = np.ascontiguousarray(keys, dtype=np.uint32)
    E, rounds = keys.shape[0], shuffle_rounds(n)
    sort_keys = np.empty((rounds, E, n), dtype=np.uint32)
    _lib.check(_lib.load().bg_shuffle_sort_keys(_lib.nptr(keys), E, n, _lib.LAYOUT_ID[layout], rounds, _lib.nptr(sort_keys)))
    x = np.broadcast_to(np.arange(n, dtype=np.int64), (E, n))
    for r in range(rounds):
        x = np.take_along_axis(x, np.argsort(sort_keys[r], axis=1, kind="stable"), axis=1)
    return np.ascontiguousarray(x)


def choice_no_replace(key: np.ndarray, n_inputs: int, n_draws: int, layout: str = "legacy") -> np.ndarray:
    if n_draws > n_inputs:
        raise ValueError("Cannot take a larger sample than population when 'replace=False'")
    return permutation(key, n_inputs, layout)[:n_draws]


def erf_inv_f32(x: np.ndarray) -> np.ndarray:
    """`lax.erf_inv` in float32 (XLA's expansion: Giles' two degree-8 polynomials in w = -log1p(-x^2))."""
    x = np.asarray(x, dtype=np.float32)
    w = -np.log1p(-(x * x))
    small = w < np.float32(5.0)
    w = np.where(small, w - np.float32(2.5), np.sqrt(w) - np.float32(3.0)).astype(np.float32)
    lo = (2.81022636e-08, 3.43273939e-07, -3.5233877e-06, -4.39150654e-06, 0.00021858087, -0.00125372503, -0.00417768164,
          0.246640727, 1.50140941)
    hi = (-0.000200214257, 0.000100950558, 0.00134934322, -0.00367342844, 0.00573950773, -0.0076224613, 0.00943887047,
          1.00167406, 2.83297682)
    p = np.where(small, np.float32(lo[0]), np.float32(hi[0])).astype(np.float32)
    for a, b in zip(lo[1:], hi[1:]):
        p = (np.where(small, np.float32(a), np.float32(b)) + p * w).astype(np.float32)
    out = (p * x).astype(np.float32)
    return np.where(np.abs(x) == 1, np.copysign(np.float32(np.inf), x), out).astype(np.float32)


def normal(key: np.ndarray, n: int, layout: str = "legacy") -> np.ndarray:
    """`jax.random.normal(key, (n,))` in float32: sqrt(2) * erf_inv(u), u uniform on (-1, 1) from the key's bits."""
    bits = _lib.random_bits(key, n, layout)
    f = ((bits >> np.uint32(9)) | np.uint32(0x3F800000)).view(np.float32) - np.float32(1.0)
    lo = np.nextafter(np.float32(-1.0), np.float32(0.0))
    u = np.maximum(lo, (f * (np.float32(1.0) - lo) + lo).astype(np.float32))
    return (np.float32(np.sqrt(2)) * erf_inv_f32(u)).astype(np.float32)


def top_k(x: np.ndarray, k: int):
    """`jax.lax.top_k` along the last axis: descending, ties -> lower index."""
    x = np.asarray(x)
    order = np.argsort(-x, axis=-1, kind="stable")[..., :k]
    return np.take_along_axis(x, order, axis=-1), order


def repeat_total(x: np.ndarray, repeats, total: int) -> np.ndarray:
    """`jnp.repeat(x, repeats, axis=0, total_repeat_length=total)`: truncate or pad with the last row."""
    x = np.asarray(x)
    out = np.repeat(x, repeats, axis=0)
    if len(out) >= total:
        return out[:total]
    return np.concatenate([out, np.repeat(x[-1:], total - len(out), axis=0)], axis=0)


def softmax_f32(x: np.ndarray) -> np.ndarray:
    """`jax.nn.softmax` in float32."""
    x = np.asarray(x, dtype=np.float32)
    e = np.exp(x - np.max(x, axis=-1, keepdims=True))
    return e / np.sum(e, axis=-1, keepdims=True, dtype=np.float32)
