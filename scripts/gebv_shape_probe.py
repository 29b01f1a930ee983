import sys, numpy as np, pandas as pd, torch
sys.path.insert(0, '/root/repo')
from breedgym_b200.simulator import Simulator
from breedgym_b200 import _lib
lib = _lib.load()
def run(m, T, rows, seed=0):
    rng = np.random.default_rng(seed)
    df = pd.DataFrame({"CHR.PHYS": np.arange(m) // max(1, m // 3 + 1), "RecombRate": np.full(m, 1.5e-3, dtype=np.float32)})
    for t in range(T):
        df[f"t{t}"] = rng.standard_normal(m).astype(np.float32)
    outs = {}
    for shape in (1, 2):
        sim = Simulator(genetic_map=df, seed=0, device=0, engine_options={"gebv_shape": shape})
        W = sim.words_per_row
        g = torch.Generator(device="cuda"); g.manual_seed(1)
        pop = torch.randint(-2**31, 2**31 - 1, (rows, 2, W), dtype=torch.int32, device="cuda", generator=g)
        full, tail = m // 32, m % 32
        pop[..., full + (1 if tail else 0):] = 0
        if tail: pop[..., full] &= (1 << tail) - 1
        o = torch.empty((rows, T), dtype=torch.float32, device="cuda")
        _lib.check(lib.bg_gebv_algo(sim._engine, pop.data_ptr(), rows, o.data_ptr(), 3, sim._stream()))
        if shape == 1:
            l = torch.empty_like(o)
            _lib.check(lib.bg_gebv_algo(sim._engine, pop.data_ptr(), rows, l.data_ptr(), 2, sim._stream()))
            outs["lut"] = l
        torch.cuda.synchronize()
        outs[shape] = o
    bad = (outs[1] != outs[2])
    print(f"m={m} T={T} rows={rows} D={lib.bg_gebv_digits(sim._engine)}: short==lut {bool(torch.equal(outs[1], outs['lut']))} long==lut {bool(torch.equal(outs[2], outs['lut']))} mismatches {int(bad.sum())} rows_bad {bad.any(1).nonzero().flatten()[:8].tolist()} cols_bad {bad.any(0).nonzero().flatten()[:16].tolist()}")
import os
CASES = [(1000000, 16, 256), (1000000, 8, 256), (1000000, 16, 1000)] if os.environ.get("PROBE_SHORT") else [(4096, 1, 128), (4096, 1, 300), (40000, 1, 300), (40000, 3, 300), (40000, 16, 300), (200000, 16, 128), (1000000, 16, 256), (1000000, 1, 256), (1000000, 16, 1000)]
for rep in range(int(os.environ.get("PROBE_REPS", "1"))):
    for m, T, rows in CASES:
        run(m, T, rows, seed=rep)
