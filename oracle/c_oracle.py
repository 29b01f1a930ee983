"""ctypes loader for the C restatement (oracle/csrc/oracle.c).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Built by `oracle/build.py`
(also invoked from `__graft_entry__.build()`); the .so is git-ignored and
travels to the GPU box with the snapshot.
"""
from __future__ import annotations

import ctypes
import os
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "liboracle.so"
LAYOUT_ID = {"legacy": 0, "partitionable": 1}
SCHEDULE_ID = {"S1": 1, "S2": 2}

_lib = None


def load():
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            from . import build as _b

            _b.build()
        _lib = ctypes.CDLL(str(LIB_PATH))
        _lib.orc_cross_envs.restype = ctypes.c_int
        _lib.orc_cross_envs_shared.restype = ctypes.c_int
        _lib.orc_num_threads.restype = ctypes.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def num_threads() -> int:
    return int(load().orc_num_threads())


def threefry2x32(k0, k1, x0, x1):
    o0, o1 = ctypes.c_uint32(), ctypes.c_uint32()
    load().orc_threefry2x32(ctypes.c_uint32(k0), ctypes.c_uint32(k1), ctypes.c_uint32(x0), ctypes.c_uint32(x1),
                            ctypes.byref(o0), ctypes.byref(o1))
    return o0.value, o1.value


def random_bits(key, n, layout="legacy"):
    key = np.ascontiguousarray(key, dtype=np.uint32)
    out = np.empty(n, dtype=np.uint32)
    load().orc_random_bits(_p(key), ctypes.c_int64(n), ctypes.c_int(LAYOUT_ID[layout]), _p(out))
    return out


def split(key, num, layout="legacy"):
    key = np.ascontiguousarray(key, dtype=np.uint32)
    out = np.empty((num, 2), dtype=np.uint32)
    load().orc_split(_p(key), ctypes.c_int64(num), ctypes.c_int(LAYOUT_ID[layout]), _p(out))
    return out


def cross_envs(pops, actions, r, cross_key, mutation=0.0, schedule="S2", layout="legacy", shared_masks=False):
    """pops bool[E,N,m,2], actions int[E,n,2] -> bool[E,n,m,2]; one key for all envs.
    shared_masks: draw the 2n crossover masks once and reuse them for every env (how the reference's vmap runs)
    instead of re-drawing per (env, gamete); same results."""
    pops = np.ascontiguousarray(pops, dtype=np.bool_)
    actions = np.ascontiguousarray(actions, dtype=np.int32)
    r = np.ascontiguousarray(r, dtype=np.float32)
    cross_key = np.ascontiguousarray(cross_key, dtype=np.uint32)
    E, N, m, _ = pops.shape
    n = actions.shape[1]
    out = np.empty((E, n, m, 2), dtype=np.bool_)
    fn = load().orc_cross_envs_shared if shared_masks else load().orc_cross_envs
    rc = fn(_p(pops), _p(actions), _p(r), ctypes.c_int64(E), ctypes.c_int64(N),
                               ctypes.c_int64(n), ctypes.c_int64(m), _p(cross_key), ctypes.c_float(mutation),
                               ctypes.c_int(SCHEDULE_ID[schedule]), ctypes.c_int(LAYOUT_ID[layout]), _p(out))
    if rc != 0:
        raise MemoryError("orc_cross_envs")
    return out


def gebv(pop, effects):
    """pop bool[...,m,2], effects f32[m,T] -> f64[...,T]."""
    pop = np.ascontiguousarray(pop, dtype=np.bool_)
    effects = np.ascontiguousarray(effects, dtype=np.float32)
    if effects.ndim == 1:
        effects = effects[:, None]
    m, T = effects.shape
    lead = pop.shape[:-2]
    rows = int(np.prod(lead)) if lead else 1
    out = np.empty((rows, T), dtype=np.float64)
    load().orc_gebv(_p(pop), _p(effects), ctypes.c_int64(rows), ctypes.c_int64(m), ctypes.c_int64(T), _p(out))
    return out.reshape(*lead, T)


def set_threads(n: int):
    """Use n OpenMP threads (torchrun exports OMP_NUM_THREADS=1, which would hide the host's cores)."""
    load().orc_set_threads(ctypes.c_int(int(n)))
