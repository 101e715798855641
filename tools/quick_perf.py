#!/usr/bin/env python
"""Developer timing of every kernel (CUDA events, current stream). Not the judged bench."""
import ctypes, json, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from noa_b200 import dcs, grids, _lib, STANDARD_ROCK, MUON_MASS, WATER

lib = _lib.require_device()
def timeit(fn, reps=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts)), float(min(ts))

def probe():
    sink = torch.zeros(8, dtype=torch.float64, device="cuda")
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    res = {}
    for threads, bps in ((256, 8), (512, 4), (1024, 2), (256, 4), (128, 8)):
        blocks = 148 * bps; iters = 20000
        ms, best = timeit(lambda: _lib.check(lib.noa_dcs_fp64_probe(iters, blocks, threads, ctypes.c_void_p(sink.data_ptr()), st)))
        res[f"{threads}x{bps}"] = blocks * threads * iters * 16 / (best * 1e-3) / 1e12
    return res

out = {"fp64_probe_Tinstr_s": probe()}
print(json.dumps(out), flush=True)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 22
for name, gen in (("B", grids.set_b), ("A", grids.set_a)):
    K, q = gen(n); Kd, qd = torch.from_numpy(K).cuda(), torch.from_numpy(q).cuda()
    r = torch.empty_like(Kd)
    for pr in dcs.PROCESSES:
        ms, best = timeit(lambda: dcs.vmap(pr)(r, Kd, qd, STANDARD_ROCK, MUON_MASS))
        print(json.dumps({"set": name, "kernel": pr.name, "n": n, "ms": ms, "best_ms": best, "Gevals_s": n / (best * 1e-3) / 1e9}), flush=True)
    if name == "B":
        lib.noa_dcs_set_pair_mode(1)
        ms, best = timeit(lambda: dcs.vmap(dcs.pair_production)(r, Kd, qd, STANDARD_ROCK, MUON_MASS))
        lib.noa_dcs_set_pair_mode(0)
        print(json.dumps({"set": name, "kernel": "pair_lanes", "n": n, "ms": ms, "best_ms": best, "Gevals_s": n / (best * 1e-3) / 1e9}), flush=True)
        r4 = torch.empty((4, n), dtype=torch.float64, device="cuda")
        ms, best = timeit(lambda: dcs.cuda.vmap_all(r4, Kd, qd, STANDARD_ROCK, MUON_MASS))
        print(json.dumps({"set": name, "kernel": "all4", "n": n, "ms": ms, "best_ms": best, "Gevals_s": 4 * n / (best * 1e-3) / 1e9}), flush=True)
        ms, best = timeit(lambda: dcs.cuda.vmap_material(r4, Kd, qd, WATER, MUON_MASS))
        print(json.dumps({"set": name, "kernel": "water_all4", "n": n, "ms": ms, "best_ms": best, "Gevals_s": 8 * n / (best * 1e-3) / 1e9}), flush=True)
Kt = torch.from_numpy(grids.table_energies(10000)).cuda()
for mp in (180, 1000):
    ms, best = timeit(lambda: dcs.cuda.tables(Kt, 0.05, STANDARD_ROCK, MUON_MASS, mp), reps=3, warm=1)
    nodes = ((mp + 5) // 6) * 6
    print(json.dumps({"kernel": "table", "min_points": mp, "ms": ms, "best_ms": best, "Gevals_s": 10000 * nodes * 4 / (best * 1e-3) / 1e9}), flush=True)
    for pr in dcs.PROCESSES:
        ms, best = timeit(lambda: dcs.cuda.tables(Kt, 0.05, STANDARD_ROCK, MUON_MASS, mp, processes=(pr,)), reps=3, warm=1)
        print(json.dumps({"kernel": "table/" + pr.name, "min_points": mp, "ms": ms, "best_ms": best}), flush=True)
