#!/usr/bin/env python
"""Small invocation of every kernel family for compute-sanitizer (memcheck / racecheck):
  compute-sanitizer --tool memcheck python tools/sanitize_small.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from noa_b200 import dcs, grids, STANDARD_ROCK, MUON_MASS, WATER

K, q = grids.set_a(3001)
Kd, qd = torch.from_numpy(K).cuda(), torch.from_numpy(q).cuda()
for pr in dcs.PROCESSES:
    dcs.map(pr)(Kd, qd, STANDARD_ROCK, MUON_MASS)
dcs.cuda.map_all(Kd, qd, STANDARD_ROCK, MUON_MASS)
dcs.cuda.map_material(Kd, qd, WATER, MUON_MASS)
Kt = torch.from_numpy(grids.table_energies(24, -2.0, 6.0)).cuda()
dcs.cuda.tables(Kt, dcs.X_FRACTION, STANDARD_ROCK, MUON_MASS, 180)
dcs.cuda.tables(Kt, dcs.X_FRACTION, STANDARD_ROCK, MUON_MASS, 1000)
r = torch.zeros_like(Kt)
dcs.soft_scattering(r, Kt, STANDARD_ROCK, MUON_MASS)
torch.cuda.synchronize()
print("done")
