#!/usr/bin/env python
"""FP64 dependent-issue latency and the warps x chains needed to fill the pipe
(noa_dcs_fp64_probe_mode 10-13, 0).  Prints DFMA per clock per SM partition (SMSP); peak is 0.5."""
import ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from noa_b200 import _lib
lib = _lib.require_device()
sink = torch.zeros(8, dtype=torch.float64, device="cuda")
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
clock = 1.965e9
out = {}
for mode, chains in ((10, 1), (12, 4), (20, '1+ldc_uniform'), (21, '1+ldc_vector'), (22, '4+ldc_uniform'), (23, '4+ldc_vector'),
                     (24, '1+lds'), (25, '4+lds')):
    for warps_per_smsp in (1, 2, 4, 8):
        threads = 128 * warps_per_smsp if warps_per_smsp <= 8 else 1024
        iters = 4000
        best = 1e9
        for _ in range(3):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            _lib.check(lib.noa_dcs_fp64_probe_mode(mode, iters, 148, threads, ctypes.c_void_p(sink.data_ptr()), st))
            b.record(); torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b) * 1e-3)
        warp_instr = 148 * (threads // 32) * iters * 16
        per_smsp_clk = warp_instr / (148 * 4) / (best * clock)
        out[f"chains{chains}_warps{warps_per_smsp}"] = round(per_smsp_clk, 4)
print(json.dumps(out, indent=0))
