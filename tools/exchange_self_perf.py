#!/usr/bin/env python
"""Single-GPU timing of the EXCHANGE form of the table build on a 1/W share of config 4's rows:
noa_dcs_table_exchange_f64 with this GPU as its only peer, against dcs.cuda.tables on the same rows.
Shows what one rank of W pays for its compute and for the exchange machinery (fences, flags)
without needing W GPUs.  Usage: python tools/exchange_self_perf.py"""
import ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from noa_b200 import _lib, dcs, grids, STANDARD_ROCK, MUON_MASS

lib = _lib.require_device()
vp = ctypes.c_void_p
K = grids.table_energies(10000)
out = {"launch": os.environ.get("NOA_DCS_TABLE_LAUNCH", "default")}
for W in (1, 2, 4, 8):
    Kl = torch.from_numpy(np.ascontiguousarray(K[0::W])).cuda()
    n = Kl.numel()
    table = torch.zeros((2, 4, n), dtype=torch.float64, device="cuda")
    flags = torch.zeros(16, dtype=torch.int32, device="cuda")
    sync = torch.zeros(8, dtype=torch.int32, device="cuda")
    dl = (vp * 1)(table.data_ptr()); cl = (vp * 1)(table.data_ptr() + 4 * n * 8); fl = (vp * 1)(flags.data_ptr())
    scratch = torch.empty(int(lib.noa_dcs_table_workspace_doubles(n, 1000)), dtype=torch.float64, device="cuda")
    epoch = [0]
    def build():
        epoch[0] += 1
        _lib.check(lib.noa_dcs_table_exchange_f64(15, vp(Kl.data_ptr()), n, 0.05, 1000, 22., 0.1364e-6, 11, MUON_MASS,
                   1, 0, dl, cl, fl, None, None, None, vp(sync.data_ptr()), vp(scratch.data_ptr()), scratch.numel(), epoch[0], n, 0, 1, 10.0,
                   vp(torch.cuda.current_stream().cuda_stream)))
    for _ in range(3): build()
    torch.cuda.synchronize(); ts = []
    for _ in range(8):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); build(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    d, c = dcs.cuda.tables(Kl, 0.05, STANDARD_ROCK, MUON_MASS, 1000)
    ok = bool(torch.equal(table[0], d) and torch.equal(table[1], c))
    ts2 = []
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); dcs.cuda.tables(Kl, 0.05, STANDARD_ROCK, MUON_MASS, 1000, out=(d, c)); b.record(); torch.cuda.synchronize(); ts2.append(a.elapsed_time(b))
    out[f"1/{W}"] = {"persistent_ms": min(ts), "cta_per_item_ms": min(ts2), "ideal_ms": None, "equal": ok}
full = out["1/1"]["cta_per_item_ms"]
for W in (1, 2, 4, 8):
    out[f"1/{W}"]["ideal_ms"] = full / W
print(json.dumps(out), flush=True)
