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
reader = raw.bg_debug_read_trace_dyn if os.environ.get("BG_OPT_FUSED_DYN") == "1" else raw.bg_debug_read_trace
reader.argtypes = [ctypes.c_void_p, ctypes.c_int]
assert reader(buf.ctypes.data, buf.size) == 0
t = buf.reshape(16, 64)
t0 = t[15, 0]
names = ["tmem st published w0", "w1", "raw_full g0", "raw_full g1", "w2", "8 MMAs issued (pair)", "fields done (step)",
         "a_empty seen (step)", "tmem st done (step)", "w3", "digit pair landed", "A pair full", "commits issued", "gather issued (stage)",
         "stage_done seen (stage)", "cta start/roles done/acc done/epilogue done"]
if os.environ.get("BG_OPT_FUSED_DYN") == "1":  # the persistent kernel (cross_gebv_dyn.cu) stamps other events
    names = ["parents in registers (@stage)", "acc seen / epilogue done (2 per item)", "raw_full g0 (stage)", "raw_full g1", "loaders start item",
             "expanders g0 start item", "MMA starts item", "expanders g1 start item", "tmem st done (step)", "next item known (item)",
             "digit pair landed", "A pair full", "commits issued", "gather issued (stage)", "stage_done seen (stage)", "cta start/-/-/end"]
for s in range(16):
    vals = [(i, int(v - t0)) for i, v in enumerate(t[s]) if v != 0]
    print(f"{s:2d} {names[s]:36s}", " ".join(f"{i}:{v}" for i, v in vals[:40]))
