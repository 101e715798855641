#!/usr/bin/env python
"""Throughput of the element-wise kernels against resident CTAs per SM (noa_dcs_set_max_blocks_per_sm):
separates latency-bound (linear in warps) from throughput-bound (flat) behaviour."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from noa_b200 import dcs, grids, _lib, STANDARD_ROCK, MUON_MASS
lib = _lib.require_device()
n = 1 << 22
K, q = grids.set_b(n)
Kd, qd = torch.from_numpy(K).cuda(), torch.from_numpy(q).cuda()
r = torch.empty_like(Kd)
for pr in (dcs.pair_production, dcs.photonuclear, dcs.bremsstrahlung):
    row = {}
    for per_sm in (1, 2, 3, 4, 5):
        lib.noa_dcs_set_max_blocks_per_sm(per_sm)
        best = 1e9
        for _ in range(4):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); dcs.vmap(pr)(r, Kd, qd, STANDARD_ROCK, MUON_MASS); b.record(); torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b))
        row[f"{per_sm}_cta_per_sm_({2 * per_sm}_warps_per_scheduler)"] = round(n / best / 1e6, 3)
    lib.noa_dcs_set_max_blocks_per_sm(0)
    print(json.dumps({pr.name + "_Gevals_s": row}), flush=True)
