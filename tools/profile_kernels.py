#!/usr/bin/env python
"""Launch each kernel of the path twice at its bench size (for ncu --set full captures)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from noa_b200 import dcs, grids, STANDARD_ROCK, MUON_MASS

which = sys.argv[1].split(",") if len(sys.argv) > 1 else ["brems", "pair", "photo", "ion", "table"]
n22, n24 = 1 << 22, 1 << 24
K, q = grids.set_b(n24)
Kd, qd = torch.from_numpy(K).cuda(), torch.from_numpy(q).cuda()
r = torch.empty_like(Kd)
Kt = torch.from_numpy(grids.table_energies(10000)).cuda()
for rep in range(2):
    if "brems" in which: dcs.vmap(dcs.bremsstrahlung)(r, Kd, qd, STANDARD_ROCK, MUON_MASS)
    if "pair" in which: dcs.vmap(dcs.pair_production)(r[:n22], Kd[:n22], qd[:n22], STANDARD_ROCK, MUON_MASS)
    if "photo" in which: dcs.vmap(dcs.photonuclear)(r[:n22], Kd[:n22], qd[:n22], STANDARD_ROCK, MUON_MASS)
    if "ion" in which: dcs.vmap(dcs.ionisation)(r, Kd, qd, STANDARD_ROCK, MUON_MASS)
    if "table" in which: dcs.cuda.tables(Kt, 0.05, STANDARD_ROCK, MUON_MASS, 1000)
torch.cuda.synchronize()
print("done")
