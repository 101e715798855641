#!/usr/bin/env python
"""FP64 pipe throughput for different operand shapes (noa_dcs_fp64_probe_mode)."""
import ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from noa_b200 import _lib
lib = _lib.require_device()
sink = torch.zeros(8, dtype=torch.float64, device="cuda")
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
names = {0: "DFMA r*c+c (1 reg src)", 1: "DFMA r*r+r (3 reg src)", 2: "DFMA r*r+c (2 reg src)", 3: "DADD r+r", 4: "DMUL r*r",
         5: "DFMA + 1 IMAD", 6: "DFMA + 2 IMAD", 7: "DFMA + 3 IMAD"}
out = {}
for mode in range(8):
    best = 0
    for _ in range(4):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        _lib.check(lib.noa_dcs_fp64_probe_mode(mode, 20000, 148 * 8, 256, ctypes.c_void_p(sink.data_ptr()), st))
        b.record(); torch.cuda.synchronize()
        best = max(best, 148 * 8 * 256 * 20000 * 16 / (a.elapsed_time(b) * 1e-3) / 1e12)
    out[names[mode]] = round(best, 3)
print(json.dumps({"fp64_pipe_Tinstr_per_s": out}))
