#!/usr/bin/env python
"""Flat (workspace) form of the table build against the row-per-CTA form: config 4 and its 1/W
shares (what one rank of W builds), per process, with a bit comparison of the two.
Usage: [NOA_DCS_LIB=<lib>] python tools/flat_perf.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from noa_b200 import dcs, grids, STANDARD_ROCK, MUON_MASS

def t(fn, reps=8, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return min(ts)

out = {"lib": os.environ.get("NOA_DCS_LIB", "default")}
Kall = torch.from_numpy(grids.table_energies(10000)).cuda()
for W in (1, 2, 4, 8):
    Kt = Kall[::W].contiguous(); n = Kt.numel()
    for mp in ((1000, 180) if W == 1 else (1000,)):
        res = {}
        tabs = {}
        for flat in (False, True):
            d = torch.zeros((4, n), dtype=torch.float64, device="cuda"); c = torch.zeros_like(d)
            key = "flat" if flat else "rows"
            res[key] = t(lambda: dcs.cuda.tables(Kt, 0.05, STANDARD_ROCK, MUON_MASS, mp, out=(d, c), flat=flat))
            tabs[key] = (d.clone(), c.clone())
            if W in (1, 8) and mp == 1000:
                for pr in dcs.PROCESSES:
                    res[f"{key}/{pr.name[:5]}"] = t(lambda: dcs.cuda.tables(Kt, 0.05, STANDARD_ROCK, MUON_MASS, mp, processes=(pr,), out=(d, c), flat=flat), reps=4, warm=1)
        res["equal"] = bool(torch.equal(tabs["flat"][0], tabs["rows"][0]) and torch.equal(tabs["flat"][1], tabs["rows"][1]))
        res["nan_free"] = bool(torch.isfinite(tabs["flat"][0]).all())
        out[f"1/{W} mp{mp}"] = res
print(json.dumps(out), flush=True)
