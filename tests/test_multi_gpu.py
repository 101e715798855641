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
    # the peer form with per-peer stores instead of NVSwitch multicast stores (what a box without
    # multicast support, or a C++ host that only has CUDA-IPC mappings, runs)
    multicast = bool(getattr(peer, "multicast", False))
    unicast = sharding.PeerTableBuilder(K, rank, world, multicast=False) \
        if kind == "PeerTableBuilder" else peer
    e = unicast.build(0.05, ELEMENTS["rock"], MUON_MASS, 180)
    torch.cuda.synchronize()
    e = e.clone()
    np.savez(os.path.join(tmp, f"r{rank}.npz"), gather=a.cpu().numpy(), peer=b.cpu().numpy(),
             again=c.cpu().numpy(), two_step=d.cpu().numpy(), kind=kind, timeouts=timeouts,
             seq=torch.stack(seq).cpu().numpy(), unicast=e.cpu().numpy(), multicast=multicast)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n_rows", [(2, 257), (2, 1000), (4, 1000), (8, 257), (8, 1000)])
def test_multi_rank_table_build(tmp_path, world, n_rows):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    from noa_b200 import dcs, grids
    mp.spawn(_worker, args=(world, _free_port(), n_rows, str(tmp_path)), nprocs=world, join=True)
    K = torch.from_numpy(grids.table_energies(n_rows)).cuda()
    d, c = dcs.cuda.tables(K, 0.05, ELEMENTS["rock"], MUON_MASS, 180)
    want = torch.stack((d, c)).cpu().numpy()
    for r in range(world):
        got = np.load(os.path.join(str(tmp_path), f"r{r}.npz"))
        assert np.array_equal(got["gather"], want), f"all-gather build differs on rank {r}"
        assert np.array_equal(got["peer"], want), f"{got['kind']} build differs on rank {r}"
        assert np.array_equal(got["again"][:, 1], want[:, 1])
        # rows of the processes a masked build did not ask for are zero on every rank
        assert not got["again"][:, [0, 2, 3]].any(), f"stale rows after a masked build, rank {r}"
        assert np.array_equal(got["two_step"], want), f"two-step peer build differs on rank {r}"
        assert np.array_equal(got["unicast"], want), f"unicast peer build differs on rank {r}"
        assert int(got["timeouts"]) == 0
        for i in range(12):
            el = ELEMENTS["Pb"] if i % 3 == 1 else ELEMENTS["rock"]
            di, ci = dcs.cuda.tables(K, 0.05, el, MUON_MASS, 60 + 6 * (i % 2))
            assert np.array_equal(got["seq"][i], torch.stack((di, ci)).cpu().numpy()), (r, i)
        print("rank", r, "builder:", got["kind"], "multicast stores:", bool(got["multicast"]))


def test_cpp_two_rank_ipc_example():
    """examples/two_rank_table_exchange.cc: a plain C++ host (fork + CUDA IPC, no torch, no NCCL)
    drives noa_dcs_table_exchange_f64 on two GPUs and compares with a single-GPU build."""
    import subprocess
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    exe = os.path.join(ROOT, "noa_b200", "two_rank_table_exchange")
    if not os.path.exists(exe):
        pytest.skip("example not built (make -C noa_b200/csrc example)")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=180)
    print(r.stdout, r.stderr)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("equals the single-GPU build") == 8 and "OK" in r.stdout


def _nccl_worker(rank, world, port, n_rows, tmp):
    """noa_dcs_allgather_f64 with a raw ncclComm_t made through the NCCL torch has loaded."""
    import ctypes
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from noa_b200 import _lib, dcs, grids
    lib = _lib.require_device()
    torch.zeros(1, device="cuda")                       # CUDA context
    nccl = None
    for name in ("libnccl.so.2", "libnccl.so"):
        try:
            nccl = ctypes.CDLL(name, mode=ctypes.RTLD_GLOBAL)
            break
        except OSError:
            continue
    if nccl is None:
        import glob
        import site
        for sp in site.getsitepackages():
            hits = glob.glob(os.path.join(sp, "nvidia", "nccl", "lib", "libnccl.so*"))
            if hits:
                nccl = ctypes.CDLL(hits[0], mode=ctypes.RTLD_GLOBAL)
                break
    assert nccl is not None, "no NCCL library found"
    uid = (ctypes.c_byte * 128)()
    if rank == 0:
        assert nccl.ncclGetUniqueId(ctypes.byref(uid)) == 0
    box = [bytes(uid)]
    dist.broadcast_object_list(box, src=0)
    uid = (ctypes.c_byte * 128).from_buffer_copy(box[0])
    comm = ctypes.c_void_p()

    class UniqueId(ctypes.Structure):
        _fields_ = [("internal", ctypes.c_byte * 128)]

    nccl.ncclCommInitRank.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, UniqueId,
                                      ctypes.c_int]
    assert nccl.ncclCommInitRank(ctypes.byref(comm), world, UniqueId(uid), rank) == 0
    K = grids.table_energies(n_rows)
    L = (n_rows + world - 1) // world
    K_local = torch.from_numpy(np.ascontiguousarray(K[rank::world])).cuda()
    gathered = torch.zeros((world, 2, 4, L), dtype=torch.float64, device="cuda")
    n_local = K_local.numel()
    d, c = dcs.cuda.tables(K_local, 0.05, ELEMENTS["rock"], MUON_MASS, 180)
    gathered[rank, 0, :, :n_local] = d
    gathered[rank, 1, :, :n_local] = c
    vp = ctypes.c_void_p
    _lib.check(lib.noa_dcs_allgather_f64(vp(gathered.data_ptr()), 8 * L, rank, comm,
                                         vp(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    # gathered[r, c, p, l] is row l * W + r of column (c, p)
    full = gathered.permute(1, 2, 3, 0).reshape(2, 4, L * world)[:, :, :n_rows].contiguous()
    np.save(os.path.join(tmp, f"nccl_r{rank}.npy"), full.cpu().numpy())
    nccl.ncclCommDestroy.argtypes = [ctypes.c_void_p]
    nccl.ncclCommDestroy(comm)
    dist.barrier()
    dist.destroy_process_group()


def test_c_abi_nccl_allgather(tmp_path):
    """noa_dcs_allgather_f64 (the NCCL form of the exchange SURVEY 8(b) sketches) with a raw
    ncclComm_t: two ranks, cyclic rows, un-permuted table equals the single-GPU build."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from noa_b200 import dcs, grids
    world, n_rows = 2, 501
    mp.spawn(_nccl_worker, args=(world, _free_port(), n_rows, str(tmp_path)), nprocs=world,
             join=True)
    K = torch.from_numpy(grids.table_energies(n_rows)).cuda()
    d, c = dcs.cuda.tables(K, 0.05, ELEMENTS["rock"], MUON_MASS, 180)
    want = torch.stack((d, c)).cpu().numpy()
    for r in range(world):
        got = np.load(os.path.join(str(tmp_path), f"nccl_r{r}.npy"))
        assert np.array_equal(got, want), f"rank {r}"
