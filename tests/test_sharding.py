"""CPU tests of the multi-GPU host logic: partitioning helpers and, with world_size = 2 under
gloo, the cyclic table partition + all-gather + un-permute (the table arithmetic is stood in by
the oracle so that no GPU is needed)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ELEMENTS, MUON_MASS, ROOT


def test_shard_range_partitions_exactly():
    from noa_b200.sharding import shard_range, cyclic_rows
    for n in (0, 1, 7, 8, 9, 10000, (1 << 22) + 3):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
            if n < 100000:
                rows = torch.cat([cyclic_rows(n, r, world) for r in range(world)])
                assert sorted(rows.tolist()) == list(range(n))


def test_sweep_segments_cover_every_pair_once():
    from noa_b200.sharding import sweep_segments
    for n_mat, weights in ((16, 4), (10, 3), (1 << 26, [2, 1, 1, 1]), (7, 1), (12, [2, 1, 3])):
        wl = [1] * weights if isinstance(weights, int) else weights
        for world in (1, 2, 3, 4, 5, 8):
            covered, cost = [], []
            for r in range(world):
                segs = sweep_segments(n_mat, weights, r, world)
                for m, lo, hi in segs:
                    assert 0 <= m < len(wl) and 0 <= lo < hi <= n_mat
                    covered.append((m * n_mat + lo, m * n_mat + hi))
                cost.append(sum(wl[m] * (hi - lo) for m, lo, hi in segs))
            covered.sort()
            assert covered[0][0] == 0 and covered[-1][1] == n_mat * len(wl)
            for a, b in zip(covered, covered[1:]):
                assert a[1] == b[0]
            assert max(cost) - min(cost) <= 2 * max(wl)       # equal cost up to rounding
    # config 5 at 8 GPUs: 5 cost units of 2^26 over 8 ranks; rank 0 starts inside water
    assert sweep_segments(1 << 26, [2, 1, 1, 1], 0, 8) == [(0, 0, 5 << 22)]
    assert sweep_segments(1 << 26, 4, 3, 8) == [(1, 1 << 25, 1 << 26)]


def test_partition_properties_hold_for_arbitrary_sizes():
    """Property form of the two tests above (hypothesis): any pair count, any world size, any
    material weights -- the shards tile the work exactly and carry equal cost up to one pair."""
    from hypothesis import given, settings, strategies as st
    from noa_b200.sharding import shard_range, sweep_segments

    @settings(max_examples=200, deadline=None)
    @given(st.integers(0, 1 << 40), st.integers(1, 64))
    def ranges(n, world):
        spans = [shard_range(n, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [hi - lo for lo, hi in spans]
        assert min(sizes) >= 0 and max(sizes) - min(sizes) <= 1

    @settings(max_examples=200, deadline=None)
    @given(st.integers(1, 1 << 30), st.lists(st.integers(1, 4), min_size=1, max_size=6),
           st.integers(1, 16))
    def sweeps(n_mat, weights, world):
        covered, cost = [], []
        for r in range(world):
            segs = sweep_segments(n_mat, weights, r, world)
            for m, lo, hi in segs:
                assert 0 <= m < len(weights) and 0 <= lo < hi <= n_mat
                covered.append((m * n_mat + lo, m * n_mat + hi))
            cost.append(sum(weights[m] * (hi - lo) for m, lo, hi in segs))
        covered.sort()
        assert covered[0][0] == 0 and covered[-1][1] == n_mat * len(weights)
        assert all(a[1] == b[0] for a, b in zip(covered, covered[1:]))
        assert max(cost) - min(cost) <= 2 * max(weights)

    ranges()
    sweeps()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n_rows, tmp):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle
    from noa_b200 import grids
    from noa_b200.sharding import TableBuilder
    checker = oracle.load_port()

    def compute(K_local, xlow, element, mass, min_points, out, processes=None):
        k = K_local.numpy()
        for p in range(4):
            out[0][p] = torch.from_numpy(checker.vmap_integral(p, 0, k, xlow, min_points, element, mass))
            out[1][p] = torch.from_numpy(checker.vmap_integral(p, 1, k, xlow, min_points, element, mass))

    K = torch.from_numpy(grids.table_energies(n_rows, -1.0, 4.0))
    builder = TableBuilder(K, rank, world, compute=compute)
    full = builder.build(0.05, ELEMENTS["rock"], MUON_MASS, 24)
    np.save(os.path.join(tmp, f"table_{rank}.npy"), full.numpy())
    dist.destroy_process_group()


@pytest.mark.parametrize("n_rows", [11, 16])
def test_table_builder_world2_gloo(tmp_path, port, n_rows):
    from noa_b200 import grids
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), n_rows, str(tmp_path)), nprocs=world, join=True)
    K = grids.table_energies(n_rows, -1.0, 4.0)
    want = np.zeros((2, 4, n_rows))
    for p in range(4):
        for ig in range(2):
            want[ig, p] = port.vmap_integral(p, ig, K, 0.05, 24, ELEMENTS["rock"], MUON_MASS)
    for r in range(world):
        got = np.load(os.path.join(str(tmp_path), f"table_{r}.npy"))
        assert got.shape == want.shape
        assert np.array_equal(got, want), f"rank {r}"
