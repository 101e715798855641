#!/usr/bin/env python
"""Developer sweep of launch bounds: times only the kernel family a variant library changes.
Usage: NOA_DCS_LIB=<lib> python tools/bounds_sweep.py <family: pair|photo|stream|all|table>"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from noa_b200 import dcs, grids, STANDARD_ROCK, MUON_MASS, WATER

fam = sys.argv[1]
def t(fn, reps=12):
    for _ in range(3): fn()
    torch.cuda.synchronize(); best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); best = min(best, a.elapsed_time(b))
    return best
n = 1 << 24 if fam == "stream" else 1 << 22
K, q = grids.set_b(n); Kd, qd = torch.from_numpy(K).cuda(), torch.from_numpy(q).cuda()
r = torch.empty_like(Kd)
if fam == "pair": print("pair %.4f ms" % t(lambda: dcs.vmap(dcs.pair_production)(r, Kd, qd, STANDARD_ROCK, MUON_MASS)))
if fam == "photo": print("photo %.4f ms" % t(lambda: dcs.vmap(dcs.photonuclear)(r, Kd, qd, STANDARD_ROCK, MUON_MASS)))
if fam == "stream":
    print("brems %.4f ms" % t(lambda: dcs.vmap(dcs.bremsstrahlung)(r, Kd, qd, STANDARD_ROCK, MUON_MASS)))
    print("ion   %.4f ms" % t(lambda: dcs.vmap(dcs.ionisation)(r, Kd, qd, STANDARD_ROCK, MUON_MASS)))
if fam == "all":
    r4 = torch.empty((4, n), dtype=torch.float64, device="cuda")
    print("all4  %.4f ms" % t(lambda: dcs.cuda.vmap_all(r4, Kd, qd, STANDARD_ROCK, MUON_MASS), reps=6))
    print("water %.4f ms" % t(lambda: dcs.cuda.vmap_material(r4, Kd, qd, WATER, MUON_MASS), reps=6))
if fam == "table":
    Kt = torch.from_numpy(grids.table_energies(10000)).cuda()
    print("table1000 %.4f ms" % t(lambda: dcs.cuda.tables(Kt, 0.05, STANDARD_ROCK, MUON_MASS, 1000), reps=5))
    print("table180  %.4f ms" % t(lambda: dcs.cuda.tables(Kt, 0.05, STANDARD_ROCK, MUON_MASS, 180), reps=5))
