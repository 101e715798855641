#!/usr/bin/env python
"""Developer timing of the table build (config 4) and the element-wise kernels for one library build.
Usage: [NOA_DCS_LIB=<lib>] [NOA_DCS_TABLE_LAUNCH=split|combined] python tools/table_perf.py [--check]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from noa_b200 import dcs, grids, STANDARD_ROCK, MUON_MASS

def t(fn, reps=8, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return min(ts), float(np.median(ts))

out = {"lib": os.environ.get("NOA_DCS_LIB", "default"), "launch": os.environ.get("NOA_DCS_TABLE_LAUNCH", "default")}
Kt = torch.from_numpy(grids.table_energies(10000)).cuda()
d = torch.zeros((4, 10000), dtype=torch.float64, device="cuda"); c = torch.zeros_like(d)
for mp in (1000, 180):
    out[f"table{mp}"] = t(lambda: dcs.cuda.tables(Kt, 0.05, STANDARD_ROCK, MUON_MASS, mp, out=(d, c)))[0]
    for pr in dcs.PROCESSES:
        out[f"table{mp}/{pr.name[:5]}"] = t(lambda: dcs.cuda.tables(Kt, 0.05, STANDARD_ROCK, MUON_MASS, mp, processes=(pr,), out=(d, c)), reps=4, warm=1)[0]
# an eighth of the rows (what one rank of eight builds)
K8 = Kt[::8].contiguous()
d8 = torch.zeros((4, K8.numel()), dtype=torch.float64, device="cuda"); c8 = torch.zeros_like(d8)
out["table1000_eighth"] = t(lambda: dcs.cuda.tables(K8, 0.05, STANDARD_ROCK, MUON_MASS, 1000, out=(d8, c8)))[0]
n = 1 << 24
K, q = grids.set_b(n); Kd, qd = torch.from_numpy(K).cuda(), torch.from_numpy(q).cuda()
r = torch.empty_like(Kd)
for pr in dcs.PROCESSES:
    m = n if pr.index in (0, 3) else n // 4
    ms = t(lambda: dcs.vmap(pr)(r[:m], Kd[:m], qd[:m], STANDARD_ROCK, MUON_MASS), reps=6)[0]
    out[f"vmap/{pr.name[:5]}"] = ms
    out[f"vmap/{pr.name[:5]}_Gevals"] = m / ms / 1e6
if "--check" in sys.argv:
    import oracle
    port = oracle.load_port()
    dcs.cuda.tables(Kt, 0.05, STANDARD_ROCK, MUON_MASS, 1000, out=(d, c))
    idx = np.arange(0, 10000, 97)
    Ks = grids.table_energies(10000)[idx]
    bad = 0
    for pr in dcs.PROCESSES:
        for ig, got in ((0, d), (1, c)):
            want = port.vmap_integral(pr.index, ig, Ks, 0.05, 1000, tuple(STANDARD_ROCK), MUON_MASS, threads=16)
            bad += int((got[pr.index].cpu().numpy()[idx] != want).sum())
    out["table_mismatches"] = bad
print(json.dumps(out), flush=True)
