#!/usr/bin/env python
"""Dump the clock64() pipeline trace of one CTA of the fused step kernel (library built with -DXG_TRACE=1).

    BG_LIB_PATH=breedgym_b200/_variants/tr.so python scripts/fused_trace.py
"""
import ctypes
import os
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from breedgym_b200 import _lib  # noqa: E402
from breedgym_b200.simulator import Simulator  # noqa: E402

E, n, m = 64, 370, 10_000
sim = Simulator(genetic_map=ROOT / "breedgym_b200" / "data" / "small_genetic_map.txt", trait_names=["Yield"], device=0, seed=0)
lib = _lib.load()
dev = torch.device("cuda", 0)
W = sim.words_per_row
pop = torch.randint(-2**31, 2**31 - 1, (E, n, 2, W), dtype=torch.int32, device=dev)
pop[..., 313:] = 0
acts = torch.randint(0, n, (E, n, 2), dtype=torch.int32, device=dev)
out = torch.empty((E, n, 2, W), dtype=torch.int32, device=dev)
gebv = torch.empty((E * n, 1), dtype=torch.float32, device=dev)
key = np.array([1, 2], dtype=np.uint32)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(5):
    flush.fill_(1)
    _lib.check(lib.bg_cross_gebv(sim._engine, pop.data_ptr(), acts.data_ptr(), out.data_ptr(), E, n, n, _lib.nptr(key),
                                 sim._layout(), sim._schedule(), gebv.data_ptr(), sim._stream()))
    torch.cuda.synchronize()
raw = ctypes.CDLL(str(_lib.LIB_PATH))
buf = np.zeros(16 * 64, dtype=np.int64)
raw.bg_debug_read_trace.argtypes = [ctypes.c_void_p, ctypes.c_int]
assert raw.bg_debug_read_trace(buf.ctypes.data, buf.size) == 0
t = buf.reshape(16, 64)
t0 = t[15, 0]
names = ["st done warp0 (g0)", "st done warp1", "raw_full g0", "raw_full g1", "st done warp2", "4 MMAs issued", "fields done (g0 steps)",
         "a_empty seen", "tmem st done", "st done warp3", "B half landed", "A full", "MMA issued", "gather issued", "stage_done seen", "cta"]
for s in range(16):
    vals = [int(v - t0) for v in t[s] if v != 0]
    print(f"{s:2d} {names[s]:24s}", " ".join(f"{v:6d}" for v in vals[:32]))
