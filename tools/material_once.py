#!/usr/bin/env python
"""One per-material table assembly (water, 10^4 energies, 180 nodes) for an ncu launch list."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from noa_b200 import dcs, grids, WATER, MUON_MASS
Kt = torch.from_numpy(grids.table_energies(10000)).cuda()
for _ in range(2):
    dcs.cuda.material_assembly(Kt, 0.05, WATER, MUON_MASS, 180)
    torch.cuda.synchronize()
