"""One-process-per-GPU sharding of the DCS path.

Every (K, q) evaluation and every table row is independent, so:
  * element-wise work: contiguous shard [r n/W, (r+1) n/W) per rank, outputs stay sharded, no
    collective (`shard_range`);
  * energy-loss tables: row cost grows with K (more nodes inside the pair / photonuclear
    kinematic range), so energies are dealt out cyclically (row i -> rank i mod W); each rank builds
    its [2, 4, ceil(n/W)] slice with ONE table kernel launch and the finished slices are
    all-gathered (NCCL over NVLink on the B200 box, gloo in the CPU tests) and un-permuted so every
    rank holds the complete [2 (DEL, CEL), 4 (process), n_K] table (`TableBuilder`).
The per-rank message is n_K/W x 8 columns x 8 B (80 kB at n_K = 10^4, W = 8): latency-bound.
"""
import torch
import torch.distributed as dist


def shard_range(n, rank, world):
    """Contiguous [lo, hi) of `n` items for `rank` of `world` (sizes differ by at most one)."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def cyclic_rows(n, rank, world):
    """Indices rank, rank + W, rank + 2W, ... < n."""
    return torch.arange(rank, max(n, rank), world)


class TableBuilder:
    """Builds the DEL/CEL tables of one element for all energies `K` across `world` ranks.

    `compute(K_local, xlow, element, mass, min_points, out=(del, cel))` defaults to the CUDA table
    kernel (dcs.cuda.tables); the CPU tests inject a checker-backed stand-in to exercise the
    partition / gather / un-permute logic under gloo.
    """

    def __init__(self, K, rank=0, world=1, compute=None, group=None):
        self.n = K.numel()
        self.rank, self.world, self.group = rank, world, group
        self.rows_per_rank = (self.n + world - 1) // world
        rows = cyclic_rows(self.n, rank, world).to(K.device)
        self.n_local = rows.numel()
        self.K_local = K.reshape(-1)[rows].contiguous()
        dev, L = K.device, self.rows_per_rank
        # local slice padded to L rows so every rank contributes the same byte count
        self.local = torch.zeros((2, 4, L), dtype=torch.float64, device=dev)
        self.compact = (torch.zeros((4, self.n_local), dtype=torch.float64, device=dev),
                        torch.zeros((4, self.n_local), dtype=torch.float64, device=dev))
        self.gathered = torch.zeros((world, 2, 4, L), dtype=torch.float64, device=dev) \
            if world > 1 else None
        if compute is None:
            from . import dcs
            compute = dcs.cuda.tables
        self.compute = compute

    def build(self, xlow, element, mass, min_points, processes=None):
        """Returns the full table [2, 4, n_K] (identical on every rank)."""
        kw = {} if processes is None else {"processes": processes}
        if self.n_local:
            self.compute(self.K_local, xlow, element, mass, min_points, out=self.compact, **kw)
        if self.world == 1:
            return torch.stack(self.compact)
        self.local[0, :, :self.n_local] = self.compact[0]
        self.local[1, :, :self.n_local] = self.compact[1]
        dist.all_gather_into_tensor(self.gathered.view(-1), self.local.view(-1), group=self.group)
        # gathered[r, c, p, l] is row l * W + r of column (c, p)
        full = self.gathered.permute(1, 2, 3, 0).reshape(2, 4, self.rows_per_rank * self.world)
        return full[:, :, :self.n].contiguous()
