"""ctypes binding of libbreedgym_b200.so (the C ABI in include/breedgym_b200.h).

There is no CPU fallback: if the shared library is missing or a GPU call fails,
an exception is raised.
"""
from __future__ import annotations

import ctypes
from ctypes import POINTER, c_char_p, c_float, c_int, c_int32, c_int64, c_uint32, c_void_p
from pathlib import Path

import numpy as np

import os

# BG_LIB_PATH: load an experimental build of the same C ABI instead (kernel tuning only)
LIB_PATH = Path(os.environ.get("BG_LIB_PATH") or Path(__file__).resolve().parent / "libbreedgym_b200.so")

LAYOUT_ID = {"legacy": 0, "partitionable": 1}
SCHEDULE_ID = {"S1": 1, "S2": 2}
PEER_HANDLE_BYTES = 96  # BG_PEER_HANDLE_BYTES

_ERRORS = {-1: ValueError, -2: RuntimeError, -3: MemoryError, -4: ValueError, -5: RuntimeError}

# name -> (restype, argtypes); every symbol include/breedgym_b200.h declares
_SIGNATURES = {
    "bg_version": (c_int, []),
    "bg_last_error": (c_char_p, []),
    "bg_kernel_launches": (c_int64, []),
    "bg_blend_envs": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_void_p]),
    "bg_threefry2x32": (None, [c_uint32, c_uint32, c_uint32, c_uint32, c_void_p]),
    "bg_key_split": (c_int, [c_void_p, c_int64, c_int, c_void_p]),
    "bg_key_split_at": (c_int, [c_void_p, c_int64, c_int64, c_int, c_void_p]),
    "bg_random_bits": (c_int, [c_void_p, c_int64, c_int, c_void_p]),
    "bg_shuffle_sort_keys": (c_int, [c_void_p, c_int64, c_int64, c_int, c_int, c_void_p]),
    "bg_key_chain_next": (c_int, [c_void_p, c_int, c_void_p]),
    "bg_thresholds": (c_int, [c_void_p, c_int64, c_void_p]),
    "bg_words_per_row": (c_int64, [c_int64]),
    "bg_engine_create": (c_int, [c_int, POINTER(c_void_p)]),
    "bg_engine_destroy": (c_int, [c_void_p]),
    "bg_engine_set_map": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_float]),
    "bg_pack": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "bg_unpack": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "bg_gather_individuals": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64, c_void_p]),
    "bg_cross": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_void_p, c_int, c_int, c_void_p]),
    "bg_cross_gebv": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_void_p, c_int, c_int, c_void_p,
                              c_void_p]),
    "bg_double_haploid": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_void_p, c_int, c_int, c_void_p]),
    "bg_meiosis_masks": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_int, c_int, c_void_p]),
    "bg_gebv": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p]),
    "bg_gebv_algo": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_int, c_void_p]),
    "bg_gebv_digits": (c_int, [c_void_p]),
    "bg_reduce_max": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_void_p]),
    "bg_reduce_mean": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_void_p]),
    "bg_topk": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int32, c_void_p, c_void_p, c_void_p]),
    "bg_pairs_from_topk": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_int64, c_void_p, c_void_p]),
    "bg_diallel_pairs": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_int32, c_int64, c_void_p, c_void_p]),
    "bg_reset_indices": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64, c_int64, c_int, c_void_p, c_void_p]),
    "bg_vec_reset": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int64, c_int64, c_int, c_void_p, c_void_p,
                             c_void_p, c_void_p, c_void_p, c_void_p]),
    "bg_vec_reset_prefetch": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int64, c_int64, c_int, c_void_p, c_void_p,
                                      c_void_p, c_void_p, c_void_p]),
    "bg_vec_reset_adopt": (c_int, [c_void_p, c_void_p]),
    "bg_engine_join": (c_int, [c_void_p, c_void_p]),
    "bg_vec_step": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_void_p,
                            c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "bg_engine_set_option": (c_int, [c_void_p, c_char_p, c_int64]),
    "bg_comm_unique_id": (c_int, [c_void_p]),
    "bg_comm_create": (c_int, [c_void_p, c_void_p, c_int, c_int, POINTER(c_void_p)]),
    "bg_comm_destroy": (c_int, [c_void_p]),
    "bg_allgather_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "bg_peer_create": (c_int, [c_void_p, c_int, c_int, c_int64, c_int64, POINTER(c_void_p)]),
    "bg_peer_handle": (c_int, [c_void_p, c_void_p]),
    "bg_peer_connect": (c_int, [c_void_p, c_void_p]),
    "bg_engine_set_peer": (c_int, [c_void_p, c_void_p]),
    "bg_peer_publish_f32": (c_int, [c_void_p, c_void_p, c_int64, c_void_p]),
    "bg_peer_wait": (c_int, [c_void_p, c_void_p]),
    "bg_peer_epoch": (c_int64, [c_void_p]),
    "bg_peer_result": (c_void_p, [c_void_p, c_int]),
    "bg_peer_set_timeout_ms": (c_int, [c_void_p, c_int64]),
    "bg_peer_timeouts": (c_int64, [c_void_p]),
    "bg_peer_destroy": (c_int, [c_void_p]),
}

_lib = None


def load() -> ctypes.CDLL:
    """Load the shared library (built in-tree by `python -m breedgym_b200.build`)."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -m breedgym_b200.build` "
                "(breedgym_b200 has no CPU fallback)")
        lib = ctypes.CDLL(str(LIB_PATH))
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc: int):
    if rc != 0:
        msg = load().bg_last_error().decode("utf-8", "replace")
        raise _ERRORS.get(rc, RuntimeError)(f"breedgym_b200: {msg} (code {rc})")


def nptr(a: np.ndarray):
    return a.ctypes.data_as(c_void_p)


# ---- host-side key chain -------------------------------------------------------
def key_data(seed: int) -> np.ndarray:
    """Raw data of `jax.random.key(seed)`: (hi32, lo32)."""
    seed = int(seed)
    if seed < 0:
        seed += 1 << 64
    return np.array([(seed >> 32) & 0xFFFFFFFF, seed & 0xFFFFFFFF], dtype=np.uint32)


def key_split(key: np.ndarray, num: int = 2, layout: str = "legacy") -> np.ndarray:
    key = np.ascontiguousarray(key, dtype=np.uint32)
    out = np.empty((int(num), 2), dtype=np.uint32)
    check(load().bg_key_split(nptr(key), int(num), LAYOUT_ID[layout], nptr(out)))
    return out


def key_split_at(key: np.ndarray, index: int, num: int, layout: str = "legacy") -> np.ndarray:
    key = np.ascontiguousarray(key, dtype=np.uint32)
    out = np.empty(2, dtype=np.uint32)
    check(load().bg_key_split_at(nptr(key), int(index), int(num), LAYOUT_ID[layout], nptr(out)))
    return out


def random_bits(key: np.ndarray, n: int, layout: str = "legacy") -> np.ndarray:
    key = np.ascontiguousarray(key, dtype=np.uint32)
    out = np.empty(int(n), dtype=np.uint32)
    check(load().bg_random_bits(nptr(key), int(n), LAYOUT_ID[layout], nptr(out)))
    return out


def thresholds(r: np.ndarray) -> np.ndarray:
    r = np.ascontiguousarray(r, dtype=np.float32)
    out = np.empty(r.shape[0], dtype=np.uint32)
    check(load().bg_thresholds(nptr(r), r.shape[0], nptr(out)))
    return out


def words_per_row(m: int) -> int:
    return int(load().bg_words_per_row(int(m)))
