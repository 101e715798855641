"""Multi-GPU tests (need >= 2 GPUs: run with `gpurun --gpus 2 -- python -m pytest tests -m gpu`;
skipped on a single-GPU box).  Two NCCL ranks build the energy-loss tables with both exchange
strategies -- NCCL all-gather and the fused peer-memory scatter kernel -- and every rank must hold
exactly the table a single GPU computes."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from conftest import ELEMENTS, MUON_MASS, ROOT

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n_rows, tmp):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    from noa_b200 import dcs, grids, sharding
    K = torch.from_numpy(grids.table_energies(n_rows)).cuda()
    gather = sharding.TableBuilder(K, rank, world)
    a = gather.build(0.05, ELEMENTS["rock"], MUON_MASS, 180).clone()
    peer = sharding.make_table_builder(K, rank, world)
    kind = type(peer).__name__
    b = peer.build(0.05, ELEMENTS["rock"], MUON_MASS, 180)
    torch.cuda.synchronize()
    b = b.clone()
    c = peer.build(0.05, ELEMENTS["rock"], MUON_MASS, 180, processes=(dcs.pair_production,))
    torch.cuda.synchronize()
    c = c.clone()
    # the older two-step exchange (scatter kernel + barrier launches) must agree as well
    two_step = sharding.PeerTableBuilder(K, rank, world, fused_barrier=False) \
        if kind == "PeerTableBuilder" else peer
    d = two_step.build(0.05, ELEMENTS["rock"], MUON_MASS, 180).clone()
    # back-to-back builds with alternating elements: the epoch flags and the two alternating
    # tables must keep every rank consistent without any host synchronisation in between
    seq = []
    for i in range(12):
        el = ELEMENTS["Pb"] if i % 3 == 1 else ELEMENTS["rock"]
        seq.append(peer.build(0.05, el, MUON_MASS, 60 + 6 * (i % 2)).clone())
    torch.cuda.synchronize()
    timeouts = peer.timeouts() if hasattr(peer, "timeouts") else 0
    np.savez(os.path.join(tmp, f"r{rank}.npz"), gather=a.cpu().numpy(), peer=b.cpu().numpy(),
             again=c.cpu().numpy(), two_step=d.cpu().numpy(), kind=kind, timeouts=timeouts,
             seq=torch.stack(seq).cpu().numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_rows", [257, 1000])
def test_two_rank_table_build(tmp_path, n_rows):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from noa_b200 import dcs, grids
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), n_rows, str(tmp_path)), nprocs=world, join=True)
    K = torch.from_numpy(grids.table_energies(n_rows)).cuda()
    d, c = dcs.cuda.tables(K, 0.05, ELEMENTS["rock"], MUON_MASS, 180)
    want = torch.stack((d, c)).cpu().numpy()
    for r in range(world):
        got = np.load(os.path.join(str(tmp_path), f"r{r}.npz"))
        assert np.array_equal(got["gather"], want), f"all-gather build differs on rank {r}"
        assert np.array_equal(got["peer"], want), f"{got['kind']} build differs on rank {r}"
        assert np.array_equal(got["again"][:, 1], want[:, 1])
        assert np.array_equal(got["two_step"], want), f"two-step peer build differs on rank {r}"
        assert int(got["timeouts"]) == 0
        for i in range(12):
            el = ELEMENTS["Pb"] if i % 3 == 1 else ELEMENTS["rock"]
            di, ci = dcs.cuda.tables(K, 0.05, el, MUON_MASS, 60 + 6 * (i % 2))
            assert np.array_equal(got["seq"][i], torch.stack((di, ci)).cpu().numpy()), (r, i)
        print("rank", r, "builder:", got["kind"])
