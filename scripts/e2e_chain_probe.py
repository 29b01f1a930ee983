#!/usr/bin/env python
"""GPU timeline of the host-facing step's chain (H2D of the actions -> step kernel -> D2H of the GEBVs), CUDA events."""
import sys
import time
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
pops = [torch.randint(-2**31, 2**31 - 1, (E, n, 2, W), dtype=torch.int32, device=dev) for _ in range(4)]
for p in pops:
    p[..., 313:] = 0
outs = [torch.empty_like(p) for p in pops]
act_pin = torch.randint(0, n, (E, n, 2), dtype=torch.int32).pin_memory()
act_dev = torch.empty((E, n, 2), dtype=torch.int32, device=dev)
gebv_dev = torch.empty((E, n, 1), dtype=torch.float32, device=dev)
gebv_pin = torch.empty((E, n, 1), dtype=torch.float32).pin_memory()
key = np.array([1, 2], dtype=np.uint32)
st = torch.cuda.current_stream()
acc = np.zeros(4)
N = 400
for it in range(100 + N):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    t0 = time.perf_counter()
    ev[0].record(st)
    act_dev.copy_(act_pin, non_blocking=True)
    ev[1].record(st)
    _lib.check(lib.bg_cross_gebv(sim._engine, pops[it % 4].data_ptr(), act_dev.data_ptr(), outs[it % 4].data_ptr(), E, n, n,
                                 _lib.nptr(key), sim._layout(), sim._schedule(), gebv_dev.data_ptr(), sim._stream()))
    ev[2].record(st)
    gebv_pin.copy_(gebv_dev, non_blocking=True)
    ev[3].record(st)
    st.synchronize()
    t1 = time.perf_counter()
    if it >= 100:
        acc += [ev[0].elapsed_time(ev[1]) * 1e3, ev[1].elapsed_time(ev[2]) * 1e3, ev[2].elapsed_time(ev[3]) * 1e3, (t1 - t0) * 1e6]
acc /= N
print(f"GPU timeline: H2D {acc[0]:.1f} us, step kernel {acc[1]:.1f} us, D2H {acc[2]:.1f} us; host wall for the chain {acc[3]:.1f} us")
