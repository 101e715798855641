#!/usr/bin/env python
"""Flat form through the C ABI with a preallocated workspace (no allocator in the timed region)."""
import ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from noa_b200 import _lib, grids, MUON_MASS
lib = _lib.require_device(); vp = ctypes.c_void_p
out = {}
Kall = torch.from_numpy(grids.table_energies(10000)).cuda()
for W in (1, 8):
    Kt = Kall[::W].contiguous(); n = Kt.numel()
    need = int(lib.noa_dcs_table_workspace_doubles(n, 1000))
    ws = torch.empty(need, dtype=torch.float64, device="cuda")
    d = torch.zeros((4, n), dtype=torch.float64, device="cuda"); c = torch.zeros_like(d)
    st = vp(torch.cuda.current_stream().cuda_stream)
    for mask in (15, 2, 4, 1, 8):
        for name, nws in (("flat", need), ("rows", 0)):
            def build():
                _lib.check(lib.noa_dcs_table_ws_f64(mask, vp(Kt.data_ptr()), n, 0.05, 1000, 22., 0.1364e-6, 11, MUON_MASS,
                                                    vp(d.data_ptr()), vp(c.data_ptr()), vp(ws.data_ptr()), nws, st))
            for _ in range(3): build()
            torch.cuda.synchronize(); ts = []
            for _ in range(8):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); build(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
            out[f"1/{W} mask{mask} {name}"] = round(min(ts), 4)
print(json.dumps(out), flush=True)
