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
import os

import torch
import torch.distributed as dist


def bind_to_gpu_numa_node(device_index):
    """Pin this process to the CPUs of the NUMA node its GPU hangs off, so that the pinned host
    buffers of the host-buffer path are allocated (first touch) in the memory next to the GPU's
    PCIe root and the ranks of one box do not all pull from one socket.  Returns the node id, or
    None when the topology cannot be read (then nothing is changed)."""
    import os
    try:
        props = torch.cuda.get_device_properties(device_index)
        bdf = f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & set(os.sched_getaffinity(0))
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except Exception:
        return None


def shard_range(n, rank, world):
    """Contiguous [lo, hi) of `n` items for `rank` of `world` (sizes differ by at most one)."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def sweep_segments(n_per_material, weights, rank, world):
    """Shard of a multi-material sweep (BASELINE.json config 5).  `weights[m]` is the cost of one
    pair of material m (its number of elements: water = H + O costs two single-element pairs), or an
    int meaning that many materials of weight 1.  The (material, pair) space is laid out
    material-major on a cost axis and cut into `world` contiguous pieces of equal cost; returns this
    rank's piece as [(material index, lo, hi), ...] with [lo, hi) a pair range inside the material."""
    if isinstance(weights, int):
        weights = [1] * weights
    starts, total = [], 0
    for w in weights:
        starts.append(total)
        total += int(w) * n_per_material

    def locate(cost):                # cost-axis position -> (material, pair)
        if cost >= total:
            return len(weights), 0
        m = max(i for i, s0 in enumerate(starts) if s0 <= cost)
        return m, (cost - starts[m]) // int(weights[m])

    (m0, p0), (m1, p1) = locate(rank * total // world), locate((rank + 1) * total // world)
    out = []
    for m in range(m0, min(m1, len(weights) - 1) + 1):
        lo = p0 if m == m0 else 0
        hi = p1 if m == m1 else n_per_material
        if lo < hi:
            out.append((m, lo, hi))
    return out


def cyclic_rows(n, rank, world):
    """Indices rank, rank + W, rank + 2W, ... < n."""
    return torch.arange(rank, max(n, rank), world)


class _HostTables:
    """Host-buffer form shared by the two builders (the e2e path: host K in -> host table out)."""

    def build_host(self, K_host, out_host, xlow, element, mass, min_points, processes=None):
        """`K_host`: the FULL energy grid [n_K] on the host (pinned for an asynchronous copy);
        `out_host`: [2, 4, n_K] host tensor that receives the complete table.  Copies the grid to
        the device, takes this rank's cyclic share, builds + exchanges, copies the table back and
        synchronises the stream.  Every rank ends with the full table on its host."""
        if K_host.is_cuda or out_host.is_cuda or K_host.dtype != torch.float64 \
                or out_host.dtype != torch.float64:
            raise ValueError("build_host expects float64 CPU tensors")
        if K_host.numel() != self.n or out_host.numel() != 8 * self.n \
                or not K_host.is_contiguous() or not out_host.is_contiguous():
            raise ValueError(f"build_host expects contiguous K [{self.n}] and out [2, 4, {self.n}]")
        dev = self.K_local.device
        if getattr(self, "K_full", None) is None:
            self.K_full = torch.empty(self.n, dtype=torch.float64, device=dev)
        self.K_full.copy_(K_host.reshape(-1), non_blocking=True)
        if self.n_local:
            self.K_local.copy_(self.K_full[self.rank::self.world])
        table = self.build(xlow, element, mass, min_points, processes)
        out_host.view(2, 4, self.n).copy_(table, non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()
        return out_host


class TableBuilder(_HostTables):
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
        # [DEL | CEL] of the local rows in one buffer: a single GPU hands it out as the table
        self.compact_buf = torch.zeros((2, 4, self.n_local), dtype=torch.float64, device=dev)
        self.compact = (self.compact_buf[0], self.compact_buf[1])
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
            return self.compact_buf      # valid until the next build()
        self.local[0, :, :self.n_local] = self.compact[0]
        self.local[1, :, :self.n_local] = self.compact[1]
        dist.all_gather_into_tensor(self.gathered.view(-1), self.local.view(-1), group=self.group)
        # gathered[r, c, p, l] is row l * W + r of column (c, p)
        full = self.gathered.permute(1, 2, 3, 0).reshape(2, 4, self.rows_per_rank * self.world)
        return full[:, :, :self.n].contiguous()


class PeerTableBuilder(_HostTables):
    """Table build fused with its exchange AND the rank barrier: ONE kernel launch per rank computes
    the rank's cyclic share of the rows, stores every finished value straight into the
    [2, 4, n_K] table of EVERY rank -- its own and, over NVLink, each peer's (buffers from torch
    symmetric memory, i.e. CUDA IPC-mapped peer allocations) -- then publishes an epoch flag to
    every peer and waits for theirs inside the same kernel.  No NCCL collective, no staging copy,
    no un-permute, no barrier launch.  C ABI: noa_dcs_table_exchange_f64.

    Two destination tables alternate (epoch parity): a rank can start build e+1 while a slower
    peer still reads table e, and build e+2 cannot start anywhere before every rank has finished
    launching e+1 behind its reads of table e (stream order).  `build()` therefore returns a view
    that stays valid until the second-next `build()` on the same stream.

    `fused_barrier=False` keeps the older two-step form (noa_dcs_table_scatter_f64 followed by a
    symmetric-memory barrier kernel) for comparison.

    A peer that does not arrive within `timeout_s` of wall-clock time is fatal: the kernel bumps
    the timeout counter and traps, so the next synchronisation on this device raises a CUDA error
    instead of `build()` having handed back a partial table (`timeouts()` reads the counter where
    the context survived).  Rows of processes outside a masked build are zero in every rank's table.

    Needs an NCCL process group + peer access between the GPUs of the box; `make_table_builder`
    falls back to the all-gather `TableBuilder` when symmetric memory cannot be set up on EVERY
    rank (the ranks agree on the outcome collectively).
    """

    FLAG_WORDS = 16     # NOA_DCS_MAX_PEERS 32-bit epoch slots

    def __init__(self, K, rank, world, group=None, fused_barrier=True, timeout_s=30.0,
                 arm=True, multicast=True):
        import ctypes
        import torch.distributed._symmetric_memory as symm_mem
        from . import _lib
        self._ctypes = ctypes
        self._lib = _lib
        self.lib = _lib.require_device()
        self.n = K.numel()
        self.rank, self.world = rank, world
        self.fused_barrier = fused_barrier
        self.timeout_s = float(timeout_s)
        rows = cyclic_rows(self.n, rank, world).to(K.device)
        self.n_local = rows.numel()
        self.K_local = K.reshape(-1)[rows].contiguous()
        group = dist.group.WORLD if group is None else group
        per_table = 2 * 4 * self.n
        # [table 0 | table 1 | flag words] in one symmetric allocation
        self.buffer = symm_mem.empty((2 * per_table + self.FLAG_WORDS // 2,), dtype=torch.float64,
                                     device=K.device)
        self.buffer.zero_()
        self.tables = [self.buffer[b * per_table:(b + 1) * per_table].view(2, 4, self.n)
                       for b in range(2)]
        self.handle = symm_mem.rendezvous(self.buffer, group)
        ptrs = [int(p) for p in self.handle.buffer_ptrs]
        assert len(ptrs) == world
        vp = ctypes.c_void_p
        half = 4 * self.n * 8
        self._del = [(vp * world)(*[p + b * per_table * 8 for p in ptrs]) for b in range(2)]
        self._cel = [(vp * world)(*[p + b * per_table * 8 + half for p in ptrs]) for b in range(2)]
        self._flags = (vp * world)(*[p + 2 * per_table * 8 for p in ptrs])
        # NVSwitch multicast alias of the same allocation (0 = not available on this box / group):
        # one multimem.st per value then reaches every rank's table
        mc = 0
        if multicast and os.environ.get("NOA_DCS_NO_MULTICAST", "0") != "1":
            try:
                mc = int(self.handle.multicast_ptr or 0)
            except Exception:
                mc = 0
        self.multicast = mc != 0
        self._mc_del = [vp(mc + b * per_table * 8) if mc else None for b in range(2)]
        self._mc_cel = [vp(mc + b * per_table * 8 + half) if mc else None for b in range(2)]
        self._mc_flags = vp(mc + 2 * per_table * 8) if mc else None
        # {CTA counter, timeouts, 6 reserved} (noa_dcs_table_exchange_f64: `sync`)
        self.done = torch.zeros(8, dtype=torch.int32, device=K.device)
        # workspace of the flat form (node terms), sized at the first build
        self.scratch = None
        self.epoch = 0
        torch.cuda.synchronize(K.device)
        if arm:
            self.arm()

    def arm(self):
        """Collective: everyone's buffer is zeroed before anyone writes.  `make_table_builder`
        defers it until all ranks have agreed that construction succeeded everywhere."""
        self.handle.barrier(channel=0)
        return self

    def build(self, xlow, element, mass, min_points, processes=None):
        """Returns the full table [2, 4, n_K]; complete, in stream order, when the launch retires."""
        mask = 0xF
        if processes is not None:
            mask = 0
            for pr in processes:
                mask |= 1 << pr.index
        A, I, Z = element
        c = self._ctypes
        dev = self.K_local.device
        self.epoch += 1
        b = self.epoch & 1
        with torch.cuda.device(dev):
            stream = c.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            if self.fused_barrier:
                need = int(self.lib.noa_dcs_table_workspace_doubles(self.n_local, int(min_points)))
                if self.scratch is None or self.scratch.numel() < need:
                    self.scratch = torch.empty(need, dtype=torch.float64, device=dev)
                self._lib.check(self.lib.noa_dcs_table_exchange_f64(
                    mask, c.c_void_p(self.K_local.data_ptr()), self.n_local, float(xlow),
                    int(min_points), float(A), float(I), int(Z), float(mass), self.world, self.rank,
                    self._del[b], self._cel[b], self._flags, self._mc_del[b], self._mc_cel[b],
                    self._mc_flags, c.c_void_p(self.done.data_ptr()),
                    c.c_void_p(self.scratch.data_ptr()), self.scratch.numel(),
                    self.epoch, self.n, self.rank, self.world, self.timeout_s, stream))
            else:
                # peers must be done reading this table before anyone overwrites it
                self.handle.barrier(channel=0)
                if self.n_local:
                    self._lib.check(self.lib.noa_dcs_table_scatter_f64(
                        mask, c.c_void_p(self.K_local.data_ptr()), self.n_local, float(xlow),
                        int(min_points), float(A), float(I), int(Z), float(mass), self.world,
                        self._del[b], self._cel[b], self.n, self.rank, self.world, stream))
                # every rank's rows have landed everywhere once all ranks pass this barrier
                self.handle.barrier(channel=0)
        return self.tables[b]

    def timeouts(self):
        """Number of exchanges in which a peer failed to arrive (synchronises the device)."""
        return int(self.done[1].item())


def make_table_builder(K, rank=0, world=1, group=None, prefer_peer=True, timeout_s=30.0):
    """PeerTableBuilder when `world` > 1 and symmetric memory works on every rank, else
    TableBuilder.  The decision is collective: each rank tries to set the peer form up (nothing in
    that attempt waits for another rank once the symmetric-memory rendezvous itself has returned),
    the ranks all-reduce(MIN) a success flag, and only a unanimous success arms the peer builders --
    so the ranks never end up running different exchange protocols."""
    if world > 1 and prefer_peer and K.is_cuda:
        builder, error = None, None
        try:
            builder = PeerTableBuilder(K, rank, world, group, timeout_s=timeout_s, arm=False)
        except Exception as exc:   # no peer access / symmetric memory unavailable
            error = exc
        ok = torch.tensor([1 if builder is not None else 0], dtype=torch.int32, device=K.device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        if int(ok.item()) == 1:
            return builder.arm()
        import warnings
        warnings.warn("peer-memory table build unavailable on at least one rank "
                      f"(this rank: {error!r}); every rank uses the NCCL all-gather")
    return TableBuilder(K, rank, world, group=group)
