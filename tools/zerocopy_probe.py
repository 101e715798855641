#!/usr/bin/env python
"""Developer probe: run the element-wise kernels straight on pinned host memory (UVA zero-copy:
the kernel's loads and stores cross PCIe themselves) and compare with the chunked copy pipeline."""
import ctypes, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from noa_b200 import dcs, grids, _lib, STANDARD_ROCK, MUON_MASS
lib = _lib.require_device()
n = 1 << 22
K, q = grids.set_b(n)
Kh, qh = torch.from_numpy(K).pin_memory(), torch.from_numpy(q).pin_memory()
outh = torch.empty(n, dtype=torch.float64).pin_memory()
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
A, I, Z = STANDARD_ROCK
def wall(fn, reps=10):
    fn(); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(reps): fn(); torch.cuda.synchronize()
    return (time.perf_counter() - t) / reps
ref = dcs.map(dcs.pair_production)(torch.from_numpy(K).cuda(), torch.from_numpy(q).cuda(), STANDARD_ROCK, MUON_MASS).cpu()
for pr in (1, 0, 2):
    def run():
        _lib.check(lib.noa_dcs_vmap_f64(pr, ctypes.c_void_p(Kh.data_ptr()), ctypes.c_void_p(qh.data_ptr()),
                                        ctypes.c_void_p(outh.data_ptr()), n, A, I, Z, MUON_MASS, st))
    t = wall(run)
    print(json.dumps({"process": pr, "zero_copy_ms": round(t * 1e3, 3), "Gevals_s": round(n / t / 1e9, 3),
                      "GBps": round(n * 24 / t / 1e9, 1)}), flush=True)
    if pr == 1:
        print("bit-identical to device-resident result:", bool(torch.equal(outh, ref)))
for pr in (1, 0, 2, 3):
    def run():
        _lib.check(lib.noa_dcs_vmap_pinned_f64(pr, ctypes.c_void_p(Kh.data_ptr()), ctypes.c_void_p(qh.data_ptr()),
                                               ctypes.c_void_p(outh.data_ptr()), n, A, I, Z, MUON_MASS, st))
    t = wall(run)
    print(json.dumps({"process": pr, "pinned_kernel_ms": round(t * 1e3, 3), "Gevals_s": round(n / t / 1e9, 3),
                      "GBps": round(n * 24 / t / 1e9, 1)}), flush=True)
    if pr == 1:
        print("pinned path bit-identical to device-resident result:", bool(torch.equal(outh, ref)))
# ragged sizes
Kd, qd = torch.from_numpy(K).cuda(), torch.from_numpy(q).cuda()
for m in (1, 511, 512, 513, 1000003):
    for pr in dcs.PROCESSES:
        outh[:m].zero_()
        _lib.check(lib.noa_dcs_vmap_pinned_f64(pr.index, ctypes.c_void_p(Kh.data_ptr()), ctypes.c_void_p(qh.data_ptr()),
                                               ctypes.c_void_p(outh.data_ptr()), m, A, I, Z, MUON_MASS, st))
        torch.cuda.synchronize()
        want = dcs.map(pr)(Kd[:m].contiguous(), qd[:m].contiguous(), STANDARD_ROCK, MUON_MASS).cpu()
        assert torch.equal(outh[:m], want), (m, pr)
print("ragged sizes ok")
stg = dcs.HostStager(chunk_pairs=1 << 18, n_slots=3)
for pr in (dcs.pair_production, dcs.bremsstrahlung, dcs.photonuclear):
    t = wall(lambda: stg.map(pr, Kh, qh, STANDARD_ROCK, MUON_MASS, out=outh))
    print(json.dumps({"process": pr.name, "staged_ms": round(t * 1e3, 3), "Gevals_s": round(n / t / 1e9, 3)}), flush=True)
