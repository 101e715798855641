#!/usr/bin/env python
"""Developer timing of the multi-GPU table build (run under torchrun, one rank per GPU): the rank's
share alone (no exchange), the NCCL all-gather form, the two-step peer form and the fused form, 20
back-to-back builds each, CUDA events, max over ranks.  NOA_DCS_LIB selects a variant build (e.g.
-DNOA_XCHG_NO_REMOTE=1 / -DNOA_XCHG_NO_WAIT=1 to take the exchange apart)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from noa_b200 import grids, sharding, STANDARD_ROCK, MUON_MASS

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
K = torch.from_numpy(grids.table_energies(10000)).cuda()
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

out = {"lib": os.path.basename(os.environ.get("NOA_DCS_LIB", "default")), "world": world, "min_points": mp}
gather = sharding.TableBuilder(K, rank, world)
out["local_only_ms"] = timed(lambda: gather.compute(gather.K_local, 0.05, STANDARD_ROCK, MUON_MASS, mp, out=gather.compact))
out["nccl_all_gather_ms"] = timed(lambda: gather.build(0.05, STANDARD_ROCK, MUON_MASS, mp))
if world > 1:
    two = sharding.PeerTableBuilder(K, rank, world, fused_barrier=False)
    out["peer_two_step_ms"] = timed(lambda: two.build(0.05, STANDARD_ROCK, MUON_MASS, mp))
    fused = sharding.PeerTableBuilder(K, rank, world)
    out["peer_fused_ms"] = timed(lambda: fused.build(0.05, STANDARD_ROCK, MUON_MASS, mp))
    out["multicast_stores"] = fused.multicast
    uni = sharding.PeerTableBuilder(K, rank, world, multicast=False)
    out["peer_fused_unicast_ms"] = timed(lambda: uni.build(0.05, STANDARD_ROCK, MUON_MASS, mp))
if rank == 0: print(json.dumps(out), flush=True)
dist.barrier(); dist.destroy_process_group()
