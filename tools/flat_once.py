#!/usr/bin/env python
"""One config-4 table build per form (for an ncu launch list). Usage: python tools/flat_once.py [W]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from noa_b200 import dcs, grids, STANDARD_ROCK, MUON_MASS
W = int(sys.argv[1]) if len(sys.argv) > 1 else 1
Kt = torch.from_numpy(grids.table_energies(10000)).cuda()[::W].contiguous()
n = Kt.numel()
d = torch.zeros((4, n), dtype=torch.float64, device="cuda"); c = torch.zeros_like(d)
for flat in (False, True, False, True):
    dcs.cuda.tables(Kt, 0.05, STANDARD_ROCK, MUON_MASS, 1000, out=(d, c), flat=flat)
    torch.cuda.synchronize()
