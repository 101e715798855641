#!/usr/bin/env python
"""Developer timing of the multi-GPU table build (run under torchrun, one rank per GPU):
local-only kernel on the rank's share, NCCL all-gather form, two-step peer form and the fused
build+exchange kernel with each fence placement.  CUDA events, max over ranks."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from noa_b200 import dcs, grids, sharding, _lib, STANDARD_ROCK, MUON_MASS

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
lib = _lib.require_device()
K = torch.from_numpy(grids.table_energies(10000)).cuda()
# EMULATE_WORLD=8 on one GPU: the per-rank share of an 8-GPU build (every 8th energy)
emulate = int(os.environ.get("EMULATE_WORLD", "1"))
if emulate > 1:
    K = K[::emulate].contiguous()
mp = int(sys.argv[1]) if len(sys.argv) > 1 else 1000

def timed(fn, reps=20, warm=3):
    for _ in range(warm): fn()
    dist.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / reps], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())

out = {"world": world, "min_points": mp, "rows": K.numel()}
gather = sharding.TableBuilder(K, rank, world)
out["local_only_ms"] = timed(lambda: gather.compute(gather.K_local, 0.05, STANDARD_ROCK, MUON_MASS, mp, out=gather.compact))
out["nccl_all_gather_ms"] = timed(lambda: gather.build(0.05, STANDARD_ROCK, MUON_MASS, mp))
if world > 1:
    two = sharding.PeerTableBuilder(K, rank, world, fused_barrier=False)
    out["peer_two_step_ms"] = timed(lambda: two.build(0.05, STANDARD_ROCK, MUON_MASS, mp))
fused = sharding.PeerTableBuilder(K, rank, world)
for mode in (0, 1, 2, 3, 4):
    lib.noa_dcs_set_exchange_fence_mode(mode)
    out[f"peer_fused_fence{mode}_ms"] = timed(lambda: fused.build(0.05, STANDARD_ROCK, MUON_MASS, mp))
lib.noa_dcs_set_exchange_fence_mode(3)
out["timeouts"] = fused.timeouts()
if rank == 0: print(json.dumps(out), flush=True)
dist.barrier(); dist.destroy_process_group()
