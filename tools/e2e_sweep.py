#!/usr/bin/env python
"""Developer sweep of the host-buffer pipeline (noa_dcs_vmap_host_f64): chunk size x slots, plus the
raw pinned H2D / D2H copy rates the pipeline is bounded by."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from noa_b200 import dcs, grids, STANDARD_ROCK, MUON_MASS

n = 1 << 22
K, q = grids.set_b(n)
Kh, qh = torch.from_numpy(K).pin_memory(), torch.from_numpy(q).pin_memory()
outh = torch.empty(n, dtype=torch.float64).pin_memory()
dev = torch.empty(2 * n, dtype=torch.float64, device="cuda")
def wall(fn, reps=10):
    fn(); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / reps
t = wall(lambda: (dev[:n].copy_(Kh, non_blocking=True), dev[n:].copy_(qh, non_blocking=True)))
print(json.dumps({"h2d_GBps": 2 * n * 8 / t / 1e9, "ms": t * 1e3}))
t = wall(lambda: outh.copy_(dev[:n], non_blocking=True))
print(json.dumps({"d2h_GBps": n * 8 / t / 1e9, "ms": t * 1e3}))
Kd, qd = dev[:n], dev[n:]
r = torch.empty(n, dtype=torch.float64, device="cuda")
t = wall(lambda: dcs.vmap(dcs.pair_production)(r, Kd, qd, STANDARD_ROCK, MUON_MASS))
print(json.dumps({"kernel_ms": t * 1e3}))
for chunk_log2 in (16, 17, 18, 19, 20):
    for slots in (2, 3, 4, 6, 8):
        st = dcs.HostStager(chunk_pairs=1 << chunk_log2, n_slots=slots)
        t = wall(lambda: st.map(dcs.pair_production, Kh, qh, STANDARD_ROCK, MUON_MASS, out=outh), reps=8)
        st.close()
        print(json.dumps({"chunk_log2": chunk_log2, "slots": slots, "ms": round(t * 1e3, 3), "Gevals_s": round(n / t / 1e9, 3)}), flush=True)
