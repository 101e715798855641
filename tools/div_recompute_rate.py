#!/usr/bin/env python
"""Developer check: how often the folded division checks (csrc/fdiv.cuh) force a second, plain-IEEE
evaluation, per process, on the synthetic sets A (broad) and B (in range)."""
import ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from noa_b200 import dcs, grids, _lib, STANDARD_ROCK, MUON_MASS

lib = _lib.require_device()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 22
cnt = ctypes.c_int64(0)
for name, gen in (("A", grids.set_a), ("B", grids.set_b)):
    K, q = gen(n)
    Kd, qd = torch.from_numpy(K).cuda(), torch.from_numpy(q).cuda()
    r = torch.empty_like(Kd)
    for pr in dcs.PROCESSES:
        _lib.check(lib.noa_dcs_div_recomputes(ctypes.byref(cnt), 1))
        dcs.vmap(pr)(r, Kd, qd, STANDARD_ROCK, MUON_MASS)
        _lib.check(lib.noa_dcs_div_recomputes(ctypes.byref(cnt), 1))
        print(json.dumps({"set": name, "process": pr.name, "n": n, "recomputed": cnt.value,
                          "fraction": cnt.value / n}), flush=True)
