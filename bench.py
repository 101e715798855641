#!/usr/bin/env python
"""Benchmark of the muon DCS hot path (the measure-dcs-calc harness of this repo).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload table|pair]
                    [--no-extras] [--builds-per-step B]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One JSON line on rank 0.  Metric: DCS evaluations per second, FP64.

Headline workload = BASELINE.json configs[3], the full energy-loss table build: standard rock,
10^4 energies x 1002 recoil nodes (min_points = 1000) x 4 processes, DEL and CEL from one DCS
evaluation per node (4.008e7 evaluations per build).  It is the configuration the metric's
"1/2/4/8 B200" and north_star's ">= 7x at 8 GPUs" are stated on and the only one with an exchange:
the energies are dealt cyclically over the N ranks, every rank builds its rows and ends with the
COMPLETE [2, 4, n_K] table (fused NVLink exchange inside the build kernels) -- strong scaling.  A
step is `--builds-per-step` (default 40) back-to-back builds so that the timed region is >= 0.5 s at
every N; `value` = evaluations of all builds / device time (CUDA events on the launching stream,
barrier + synchronize on both sides, max over ranks).  --workload pair keeps round 1's headline
(configs[1], pair production on 2^22 pairs per GPU, weak scaling); it is in `extras` otherwise.

  e2e        the same builds through the host-buffer entry point (sharding builder .build_host):
             the energy grid is copied from pinned host memory and the finished table is copied
             back to the host for every build, inside the timed region.
  parity     computed OUTSIDE the timed region on what the timed code produces: every value of
             the table against the compiled reference (oracle/_ref, or the C port where it is
             absent) on rank 0; at N > 1 every rank bit-compares its exchanged table with a local
             single-GPU build.  The extras carry their own parity samples.
  roofline   the build is FP64-pipe bound.  `frac` (= `frac_executed`) = the FP64 instructions
             the kernels of one build really execute (ncu counters of this commit,
             profiles/latest_traffic.json) / build time / the DFMA rate measured live by the probe
             library, i.e. the FP64 pipe utilisation over the build; `frac_census` = the same with
             SURVEY 8(d)'s census (per DCS value x the node evaluations INSIDE each process's
             kinematic range, counted live, + 20 per node), which over-counts (libdevice prices).
  cpu_baseline  the reference's own table integrals on the host cores (bounded sample).
--impl reference times only the CPU reference arm on the same config and prints its line.
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "dcs_evals_per_sec_fp64"
UNIT = "evals/s"
# config 4
N_K = 10000
MIN_POINTS = 1000
NODES = 1002
X_LOW = 0.05
EVALS_PER_BUILD = N_K * NODES * 4
# config 2
N_PAIRS = 1 << 22
ROTATE = 4
# algorithmic FP64-pipe instructions and bytes per evaluation (SURVEY.md 8(a),(d))
ALGO_INSTR = {"bremsstrahlung": 160, "pair_production": 3500, "photonuclear": 6900,
              "ionisation": 165}
NODE_OVERHEAD_INSTR = 20
ALGO_BYTES = 24
PROC_NAMES = ("bremsstrahlung", "pair_production", "photonuclear", "ionisation")
ROCK = (22., 0.1364E-6, 11)
MUON_MASS = 0.10565839


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


# --------------------------------------------------------------------------------------------
# clocks sampling (recipe: /opt/skills/guides/B200_PROFILING.md)
# --------------------------------------------------------------------------------------------
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.QUERY}",
                 "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._pump, daemon=True)
        self.thread.start()

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self, t0, t1):
        """Samples taken strictly inside [t0, t1] (the timed region)."""
        sm, smax, reasons, power = [], [], set(), []
        for t, line in self.rows:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                clk, mx = float(f[1]), float(f[2])
            except ValueError:
                continue
            smax.append(mx)
            if t0 <= t <= t1:
                sm.append(clk)
                try:
                    power.append(float(f[3]))
                except ValueError:
                    pass
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                      "sw_power_cap"), f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None,
                "samples": len(sm), "window_s": t1 - t0, "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------------
# CPU reference arm
# --------------------------------------------------------------------------------------------
def load_cpu_checker():
    import oracle
    ref = oracle.load_reference()
    checker, kind = (ref, "reference") if ref is not None else (oracle.load_port(), "port")
    checker.use_all_cores()
    return checker, kind


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return None


def median_time(fn, repeats=5, warm=1):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(repeats):
        t = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t)
    return statistics.median(ts)


def cpu_table_once(checker, K, threads):
    """The eight integrals of config 4 on the energies K: dcs::vmap_integral(recoil_integral(f, g))
    per (process, integrand) -- serial as the reference has it (threads = 1, dcs.hh:115-130) or a
    harness-side OpenMP loop over energies around the unmodified closure (oracle/ref_shim.cc)."""
    t = time.perf_counter()
    for process in range(4):
        for integrand in (0, 1):
            checker.vmap_integral(process, integrand, K, X_LOW, MIN_POINTS, ROCK, MUON_MASS,
                                  threads=threads)
    return time.perf_counter() - t


def table_sample(grid, n):
    """Every (n_K / n)-th energy: spans the whole cost distribution of the rows."""
    n = int(max(1, min(grid.size, n)))
    idx = np.unique(np.linspace(0, grid.size - 1, n).astype(np.int64))
    return np.ascontiguousarray(grid[idx])


def cpu_table_rate(checker, grid, threads, target_seconds):
    """(node evaluations of the job / s, sample size, seconds) for one pass over a bounded sample
    sized for about `target_seconds`.  A node evaluation counts once per (energy, process, node)
    as in SURVEY 8(d), i.e. the GPU's and the CPU's `value` are the same job over time (the
    reference evaluates the DCS twice per node, once per integrand -- that is its cost)."""
    # enough energies to keep every thread busy (the shim's OpenMP loop hands out chunks of 8)
    probe = table_sample(grid, max(32 * threads, 64) if threads > 1 else 24)
    dt = cpu_table_once(checker, probe, threads)
    n = int(min(grid.size, max(probe.size, probe.size * target_seconds / max(dt, 1e-6))))
    Ks = table_sample(grid, n)
    dt = cpu_table_once(checker, Ks, threads)
    return Ks.size * NODES * 4 / dt, Ks.size, dt


def table_config(n_gpus, builds):
    return {"workload": "full energy-loss table build (recoil-energy quadrature): standard rock, muon, "
                        "10^4 energies x 1002 nodes (min_points 1000) x 4 processes, DEL + CEL "
                        "(BASELINE.json configs[3])",
            "n_energies": N_K, "nodes_per_row": NODES, "processes": 4, "element": "standard_rock",
            "evals_per_build": EVALS_PER_BUILD, "builds_per_step": builds,
            "step": f"{builds} back-to-back builds of the same table (timed region >= 0.5 s at every N)",
            "parallelism": (f"energies dealt cyclically over {n_gpus} ranks; every rank ends with the "
                            "complete [2,4,n_K] table (fused NVLink exchange inside the build "
                            "kernels)") if n_gpus > 1 else "1 GPU",
            "l2": "compute-bound: 80 kB of input and 640 kB of output per build, no re-read data; "
                  "L2 state is immaterial (the element-wise extras rotate buffers > 126 MB L2)"}


def pair_config(n_gpus):
    return {"workload": "pair_production DCS (nested 8-node Gauss-Legendre), standard rock, muon, "
                        "2^22 (K,q) pairs per GPU, synthetic set B (BASELINE.json configs[1])",
            "pairs_per_gpu": N_PAIRS, "element": "standard_rock", "process": "pair_production",
            "parallelism": f"{n_gpus} independent shard(s), no data-path collective",
            "l2": f"inputs/outputs rotate over {ROTATE} buffer sets "
                  f"({ROTATE * N_PAIRS * 24 >> 20} MiB > 126 MB L2)"}


def run_reference_arm(args):
    """The reference's own CPU implementation of the headline workload on the host cores."""
    rank = env_int("RANK", 0)
    if rank != 0:
        return 0
    from noa_b200 import grids
    checker, kind = load_cpu_checker()
    threads = checker.max_threads
    budget = min(8.0, 150.0 / max(args.steps + args.warmup, 1))
    if args.workload == "pair":
        K, q = grids.set_b(N_PAIRS)
        probe = 1 << 15
        t = time.perf_counter()
        checker.vmap(1, K[:probe], q[:probe], ROCK, MUON_MASS, threads=threads)
        rate = probe / max(time.perf_counter() - t, 1e-6)
        n = int(min(N_PAIRS, max(probe, rate * budget)))
        idx = np.linspace(0, N_PAIRS - 1, n).astype(np.int64)
        Ks, qs = np.ascontiguousarray(K[idx]), np.ascontiguousarray(q[idx])
        step = lambda: checker.vmap(1, Ks, qs, ROCK, MUON_MASS, threads=threads)  # noqa: E731
        units = n
        sample = (f"{n} of the 2^22 set-B pairs per step (evenly strided), "
                  f"dcs::pvmap(pair_production) on {threads} OpenMP threads")
        config, scaling = pair_config(args.gpus), "weak"
    else:
        grid = grids.table_energies(N_K)
        _, n, _ = cpu_table_rate(checker, grid, threads, budget)
        Ks = table_sample(grid, n)
        step = lambda: cpu_table_once(checker, Ks, threads)  # noqa: E731
        units = Ks.size * NODES * 4
        sample = (f"{Ks.size} of the 10^4 energies per step (evenly strided), the eight "
                  f"dcs::recoil_integral columns each; the reference's driver is serial "
                  f"(dcs.hh:115-130), this is a harness-side OpenMP loop over energies around the "
                  f"unmodified closure on {threads} threads")
        config, scaling = table_config(args.gpus, args.builds_per_step), "strong"
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    value = units * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": scaling,
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind,
                         "sample": sample, "cpu": cpu_model()},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    _emit(line)
    return 0


# --------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------
# The contract is ONE JSON line on stdout.  Native libraries print there too (NCCL announces its
# version on fd 1 when the box sets NCCL_DEBUG), so fd 1 is pointed at stderr for the whole run and
# the result line goes to a private duplicate of the original stdout.
_RESULT_OUT = None


def _claim_stdout():
    global _RESULT_OUT
    if _RESULT_OUT is None:
        sys.stdout.flush()
        _RESULT_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def _emit(line):
    out = _RESULT_OUT if _RESULT_OUT is not None else sys.stdout
    print(json.dumps(line), file=out, flush=True)


def compare(got, want):
    """Parity summary of two float64 arrays: bit-exact flag and max relative error."""
    got = np.asarray(got, dtype=np.float64).reshape(-1)
    want = np.asarray(want, dtype=np.float64).reshape(-1)
    same = (got == want) | (np.isnan(got) & np.isnan(want))
    nz = (want != 0) & np.isfinite(want) & np.isfinite(got)
    rel = np.abs(got[nz] - want[nz]) / np.abs(want[nz])
    zeros_ok = bool(np.all(got[want == 0] == 0))
    return {"checked": int(got.size), "bit_exact": bool(same.all()),
            "mismatches": int((~same).sum()),
            "max_rel": float(rel.max()) if rel.size else 0.0, "exact_zeros": zeros_ok}


class Ctx:
    """What every part of the GPU arm needs."""
    pass


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="noa_b200", choices=["noa_b200", "reference"])
    ap.add_argument("--workload", default="table", choices=["table", "pair"])
    ap.add_argument("--builds-per-step", type=int, default=40)
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup
    args.builds_per_step = max(1, args.builds_per_step)

    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist
    from noa_b200 import dcs, grids, _lib, physics, sharding

    c = Ctx()
    c.args, c.torch, c.dist, c.dcs, c.grids, c.physics, c.sharding = (args, torch, dist, dcs, grids,
                                                                       physics, sharding)
    c.world = env_int("WORLD_SIZE", 1)
    c.rank = env_int("RANK", 0)
    local_rank = env_int("LOCAL_RANK", 0)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: noa_b200 has no CPU path "
                         "(use --impl reference for the CPU reference arm)")
    torch.cuda.set_device(local_rank)
    c.distributed = c.world > 1
    c.numa_node = sharding.bind_to_gpu_numa_node(local_rank) if c.distributed else None
    if c.distributed:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    c.lib = _lib.require_device()
    c.probe = _lib.load_probe()
    c._lib = _lib
    c.stream = torch.cuda.current_stream()
    c.st = ctypes.c_void_p(c.stream.cuda_stream)

    def barrier():
        if c.distributed:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if not c.distributed:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def min_over_ranks(x):
        if not c.distributed:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return float(t.item())

    c.barrier, c.max_over_ranks, c.min_over_ranks = barrier, max_over_ranks, min_over_ranks

    def timed(fn, reps=10, warm=3):
        """ms per call of fn(i): CUDA events on the launching stream, max over ranks."""
        for i in range(warm):
            fn(i)
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(c.stream)
        for i in range(reps):
            fn(i)
        b.record(c.stream)
        barrier()
        return max_over_ranks(a.elapsed_time(b)) / reps

    c.timed = timed

    # ---- FP64 peak probe (roofline denominator), timed alone ------------------------------------
    sink = torch.zeros(8, dtype=torch.float64, device="cuda")
    sms = torch.cuda.get_device_properties(local_rank).multi_processor_count

    def probe_once(iters=20000, threads=256, per_sm=8):
        blocks = sms * per_sm
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(c.stream)
        _lib.check(c.probe.noa_dcs_fp64_probe(iters, blocks, threads,
                                              ctypes.c_void_p(sink.data_ptr()), c.st))
        b.record(c.stream)
        torch.cuda.synchronize()
        return blocks * threads * iters * 16 / (a.elapsed_time(b) * 1e-3)

    probe_once(2000)
    c.fp64_peak = max(probe_once() for _ in range(5))        # DFMA / s
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        c.hbm_peak, c.hbm_src = float(peaks["hbm_gbs"]), "MEASURED_PEAKS.json"
    except Exception:
        c.hbm_peak, c.hbm_src = 6650.0, "fallback (B200_PROFILING.md)"
    try:
        c.profile = json.load(open(os.path.join(ROOT, "profiles", "latest_traffic.json")))
    except Exception:
        c.profile = {}

    c.sampler = ClockSampler(local_rank if "CUDA_VISIBLE_DEVICES" not in os.environ else
                             int(os.environ["CUDA_VISIBLE_DEVICES"].split(",")[local_rank]))
    if c.rank == 0:
        c.sampler.start()
        time.sleep(0.2)

    head = headline_table(c) if args.workload == "table" else headline_pair(c)

    if c.rank == 0:
        c.sampler.stop()

    extras = None
    if not args.no_extras:
        extras = run_extras(c)

    if c.rank == 0:
        line = {"metric": METRIC, "unit": UNIT, "n_gpus": c.world, "steps": args.steps,
                "warmup": args.warmup, "higher_is_better": True, "vs_baseline": None,
                "dtype": "f64", "data": "synthetic"}
        line.update(head)
        line["extras"] = extras
        _emit(line)
    if c.distributed:
        dist.destroy_process_group()
    return 0


# --------------------------------------------------------------------------------------------
# headline: config 4, the table build
# --------------------------------------------------------------------------------------------
def inrange_nodes(c, Kt_host):
    """Node evaluations of one build that fall inside each process's kinematic range (non-zero
    DCS), counted on the GPU with the element-wise kernels on the very nodes of the quadrature:
    the algorithmic work of the build (rows below a threshold are exact zeros by early exit)."""
    torch, dcs = c.torch, c.dcs
    x6 = np.array([0.03376524, 0.16939531, 0.38069041, 0.61930959, 0.83060469, 0.96623476])
    cells = (MIN_POINTS + 5) // 6
    counts = {}
    chunk = 1000
    for name in PROC_NAMES:
        counts[name] = 0
    for lo in range(0, Kt_host.size, chunk):
        K = Kt_host[lo:lo + chunk]
        lb, ub = np.log(K * X_LOW), np.log(K)
        h = (ub - lb) / cells
        t = (np.arange(cells)[:, None] + x6[None, :]).reshape(-1)          # [1002]
        q = np.exp(lb[:, None] + h[:, None] * t[None, :]).reshape(-1)
        Kd = torch.from_numpy(np.repeat(K, t.size)).cuda()
        qd = torch.from_numpy(q).cuda()
        for pr in dcs.PROCESSES:
            v = dcs.map(pr)(Kd, qd, c.physics.STANDARD_ROCK, c.physics.MUON_MASS)
            counts[pr.name] += int((v != 0).sum().item())
    return counts


def headline_table(c):
    torch, dcs, grids, physics, sharding, args = (c.torch, c.dcs, c.grids, c.physics, c.sharding,
                                                  c.args)
    builds = args.builds_per_step
    grid = grids.table_energies(N_K)
    Kt = torch.from_numpy(grid).cuda()
    builder = sharding.make_table_builder(Kt, c.rank, c.world)
    el, mass = physics.STANDARD_ROCK, physics.MUON_MASS

    def step(_):
        for _b in range(builds):
            builder.build(X_LOW, el, mass, MIN_POINTS)

    for i in range(args.warmup):
        step(i)
    c.barrier()
    launches0 = c.lib.noa_dcs_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.perf_counter()
    e0.record(c.stream)
    for i in range(args.steps):
        step(i)
    e1.record(c.stream)
    c.barrier()
    t_wall1 = time.perf_counter()
    launches = c.lib.noa_dcs_launch_count() - launches0
    ms_total = c.max_over_ranks(e0.elapsed_time(e1))
    ms_per_step = ms_total / args.steps
    ms_per_build = ms_per_step / builds
    value = EVALS_PER_BUILD * builds * args.steps / (ms_total * 1e-3)
    clocks = c.sampler.summary(t_wall0, t_wall1) if c.rank == 0 else None

    # ---- parity of what was just timed (outside the timed region) -------------------------------
    table = builder.build(X_LOW, el, mass, MIN_POINTS)
    torch.cuda.synchronize()
    got = table.detach().cpu().numpy().copy()
    rank_exact = 1.0
    if c.distributed:
        d1, c1 = dcs.cuda.tables(Kt, X_LOW, el, mass, MIN_POINTS)
        torch.cuda.synchronize()
        local = torch.stack((d1, c1)).cpu().numpy()
        rank_exact = 1.0 if np.array_equal(got, local) else 0.0
    all_ranks_exact = c.min_over_ranks(rank_exact) == 1.0
    parity = None
    if c.rank == 0:
        checker, kind = load_cpu_checker()
        want = np.zeros_like(got)
        t = time.perf_counter()
        for pr in range(4):
            for ig in (0, 1):
                want[ig, pr] = checker.vmap_integral(pr, ig, grid, X_LOW, MIN_POINTS, ROCK,
                                                     MUON_MASS, threads=checker.max_threads)
        parity = compare(got, want)
        parity["against"] = (f"oracle/_ref (compiled reference headers), all 8 x 10^4 values"
                             if kind == "reference" else "oracle C port, all 8 x 10^4 values")
        parity["oracle_seconds"] = time.perf_counter() - t
        if c.distributed:
            parity["every_rank_equals_single_gpu_build"] = bool(all_ranks_exact)
            parity["bit_exact"] = bool(parity["bit_exact"] and all_ranks_exact)
        parity["tolerance"] = "bit-exact asserted; north_star bar is 1e-12 relative"

    # ---- e2e: host grid in, host table out, every build -----------------------------------------
    Kh = torch.from_numpy(grid.copy()).pin_memory()
    outh = torch.empty((2, 4, N_K), dtype=torch.float64).pin_memory()
    e2e_steps = args.steps

    def e2e_step():
        for _b in range(builds):
            builder.build_host(Kh, outh, X_LOW, el, mass, MIN_POINTS)

    e2e_step()
    c.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_dt = c.max_over_ranks(time.perf_counter() - t0)
    c.barrier()
    e2e_value = EVALS_PER_BUILD * builds * e2e_steps / e2e_dt
    e2e_exact = bool(np.array_equal(outh.numpy(), got))

    # ---- roofline ------------------------------------------------------------------------------
    roofline = None
    cpu_baseline = None
    if c.rank == 0:
        inr = inrange_nodes(c, grid)
        algo = sum(inr[n] * ALGO_INSTR[n] for n in PROC_NAMES) + EVALS_PER_BUILD * NODE_OVERHEAD_INSTR
        per_gpu_s = ms_per_build * 1e-3 * c.world          # GPU-seconds of one build
        achieved = algo * 2 / per_gpu_s / 1e12
        peak = c.fp64_peak * 2 / 1e12
        executed = c.profile.get("table_build_fp64_instr_executed")
        frac_executed = (executed / per_gpu_s / c.fp64_peak) if executed else None
        # `frac` is the executed-instruction fraction (FP64 pipe utilisation over the build):
        # the census of SURVEY 8(d) prices exp / log / division at libdevice cost and credits the
        # build with more FP64 work than its kernels need (frac_census can exceed 1)
        roofline = {
            "bound": "fp64", "kernel": "table_terms_kernel<photonuclear | pair_production | "
                                       "bremsstrahlung+ionisation> + table_sum_kernel (one chained "
                                       "build, flat form)",
            "achieved": (executed * 2 / per_gpu_s / 1e12) if executed else achieved,
            "peak": peak, "unit": "TFLOP/s",
            "frac": frac_executed if frac_executed is not None else achieved / peak,
            "frac_executed": frac_executed,
            "frac_census": achieved / peak, "achieved_census": achieved,
            "traffic": c.profile.get("table_build_dram_bytes"),
            "traffic_note": "the node terms make one round trip through a workspace (16 B per node "
                            "and process written and read back: 1.28 GB per build, ~5 % of HBM "
                            "bandwidth over the build) so that no CTA ever waits on a row's "
                            "serial-order sum; input + output are 0.72 MB",
            "peak_source": "measured live: noa_dcs_fp64_probe (16 independent DFMA chains/thread), "
                           f"{c.fp64_peak / 1e12:.2f} T DFMA/s; MEASURED_PEAKS.json has no FP64 entry",
            "algorithmic": {"fp64_pipe_instr_per_build": algo,
                            "in_range_node_evals": inr, "node_evals": EVALS_PER_BUILD,
                            "per_eval": ALGO_INSTR, "per_node_overhead": NODE_OVERHEAD_INSTR,
                            "bytes_per_build": 8 * N_K + 64 * N_K,
                            "workspace_bytes_per_build": 2 * 16 * EVALS_PER_BUILD},
            "executed": {"fp64_instr_per_build": executed,
                         "source": c.profile.get("source"), "commit": c.profile.get("commit")},
            "kernel_ms": ms_per_build, "gpus": c.world,
        }
        if c.world == 1:
            checker, kind = load_cpu_checker()
            threads = checker.max_threads
            rate, n_s, dt = cpu_table_rate(checker, grid, threads, args.cpu_seconds)
            s_rate, s_n, s_dt = cpu_table_rate(checker, grid, 1, min(args.cpu_seconds, 6.0))
            cpu_baseline = {
                "value": rate, "unit": UNIT, "cores": threads, "kind": kind, "cpu": cpu_model(),
                "sample": f"{n_s} of the 10^4 energies (evenly strided), eight recoil_integral "
                          f"columns each, {dt:.1f} s: harness-side OpenMP loop over energies around "
                          f"the unmodified closure ({threads} threads)",
                "serial": {"value": s_rate, "cores": 1,
                           "sample": f"{s_n} energies, {s_dt:.1f} s: the reference's own serial "
                                     "dcs::vmap_integral (dcs.hh:115-130)"}}
    return {
        "value": value, "ms_per_step": ms_per_step, "scaling": "strong",
        "config": table_config(c.world, builds),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": builds * N_K * 8,
                "d2h_bytes_per_step": builds * 8 * N_K * 8, "steps": e2e_steps,
                "api": f"{type(builder).__name__}.build_host: pinned host grid -> device, build + "
                       "exchange, complete table -> pinned host, stream synchronised, per build",
                "host_table_equals_device_table": e2e_exact},
        "gpu_launches": int(launches), "ms_per_build": ms_per_build,
        "exchange": type(builder).__name__ if c.world > 1 else None,
        "clocks": clocks, "parity": parity, "roofline": roofline, "cpu_baseline": cpu_baseline,
    }


# --------------------------------------------------------------------------------------------
# config 2 (round 1's headline): pair production, 2^22 pairs per GPU
# --------------------------------------------------------------------------------------------
def pair_block(c, steps, warmup, with_cpu):
    torch, dcs, grids, physics = c.torch, c.dcs, c.grids, c.physics
    total = N_PAIRS * c.world
    sets = []
    for r in range(ROTATE):
        K, q = grids.set_b(total * ROTATE, (r * c.world + c.rank) * N_PAIRS, N_PAIRS)
        sets.append((torch.from_numpy(K).cuda(), torch.from_numpy(q).cuda(),
                     torch.empty(N_PAIRS, dtype=torch.float64, device="cuda")))
    K0, q0 = grids.set_b(total, c.rank * N_PAIRS, N_PAIRS)

    def step(i):
        Kd, qd, out = sets[i % ROTATE]
        dcs.vmap(dcs.pair_production)(out, Kd, qd, physics.STANDARD_ROCK, physics.MUON_MASS)

    for i in range(warmup):
        step(i)
    c.barrier()
    launches0 = c.lib.noa_dcs_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.perf_counter()
    e0.record(c.stream)
    for i in range(steps):
        step(i)
    e1.record(c.stream)
    c.barrier()
    t_wall1 = time.perf_counter()
    launches = c.lib.noa_dcs_launch_count() - launches0
    ms_total = c.max_over_ranks(e0.elapsed_time(e1))
    ms_per_step = ms_total / steps
    value = total * steps / (ms_total * 1e-3)

    # e2e: pinned (in-place kernel) and pageable (copy pipeline) host buffers
    stager = dcs.HostStager()
    Kh, qh = torch.from_numpy(K0).pin_memory(), torch.from_numpy(q0).pin_memory()
    outh = torch.empty(N_PAIRS, dtype=torch.float64).pin_memory()
    Kp, qp = torch.from_numpy(K0.copy()), torch.from_numpy(q0.copy())
    outp = torch.empty(N_PAIRS, dtype=torch.float64)

    def e2e(Ka, qa, oa, reps):
        for _ in range(2):
            stager.map(dcs.pair_production, Ka, qa, physics.STANDARD_ROCK, physics.MUON_MASS, out=oa)
        c.barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            stager.map(dcs.pair_production, Ka, qa, physics.STANDARD_ROCK, physics.MUON_MASS, out=oa)
        torch.cuda.synchronize()
        dt = c.max_over_ranks(time.perf_counter() - t0)
        c.barrier()
        return total * reps / dt

    reps = max(3, min(steps, 20))
    e2e_pinned = e2e(Kh, qh, outh, reps)
    e2e_pageable = e2e(Kp, qp, outp, reps)
    stager.close()

    # parity: a strided sample of what the timed kernel wrote, and of the host path's output
    parity = None
    rate = N_PAIRS / (ms_per_step * 1e-3)
    if c.rank == 0:
        checker, kind = load_cpu_checker()
        dcs.vmap(dcs.pair_production)(sets[0][2], *(torch.from_numpy(a).cuda() for a in (K0, q0)),
                                      physics.STANDARD_ROCK, physics.MUON_MASS)
        torch.cuda.synchronize()
        idx = np.arange(0, N_PAIRS, 61)                       # 68 760 values
        want = checker.vmap(1, K0[idx], q0[idx], ROCK, MUON_MASS, threads=checker.max_threads)
        parity = compare(sets[0][2].cpu().numpy()[idx], want)
        host = compare(outh.numpy()[idx], want)
        page = compare(outp.numpy()[idx], want)
        parity["host_pinned_bit_exact"] = host["bit_exact"]
        parity["host_pageable_bit_exact"] = page["bit_exact"]
        parity["against"] = f"oracle ({kind}), every 61st of the 2^22 pairs"
    peak = c.fp64_peak * 2 / 1e12
    achieved = rate * ALGO_INSTR["pair_production"] * 2 / 1e12
    ex = c.profile.get("pair_production_fp64_instr_per_eval_executed")
    roofline = {
        "bound": "fp64", "kernel": "vmap_kernel<pair_production>", "achieved": achieved,
        "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
        "frac_executed": rate * ex / c.fp64_peak if ex else None,
        "traffic": c.profile.get("pair_production_dram_bytes_per_launch"),
        "algorithmic": {"fp64_pipe_instr_per_eval": ALGO_INSTR["pair_production"],
                        "bytes_per_eval": ALGO_BYTES, "evals_per_launch": N_PAIRS},
        "executed": {"fp64_instr_per_eval": ex, "source": c.profile.get("source"),
                     "commit": c.profile.get("commit")},
        "kernel_ms": ms_per_step,
        "hbm_view": {"achieved": rate * ALGO_BYTES / 1e9, "peak": c.hbm_peak, "unit": "GB/s",
                     "frac": rate * ALGO_BYTES / 1e9 / c.hbm_peak, "peak_source": c.hbm_src},
    }
    cpu_baseline = None
    if with_cpu and c.rank == 0 and c.world == 1:
        checker, kind = load_cpu_checker()
        threads = checker.max_threads
        n = 1 << 20
        idx = np.linspace(0, N_PAIRS - 1, n).astype(np.int64)
        Ks, qs = np.ascontiguousarray(K0[idx]), np.ascontiguousarray(q0[idx])
        dt = median_time(lambda: checker.vmap(1, Ks, qs, ROCK, MUON_MASS, threads=threads), 5)
        cpu_baseline = {"value": n / dt, "unit": UNIT, "cores": threads, "kind": kind,
                        "sample": f"2^20 of the 2^22 set-B pairs (evenly strided), "
                                  f"dcs::pvmap(pair_production), median of 5"}
    return {
        "value": value, "ms_per_step": ms_per_step, "scaling": "weak",
        "config": pair_config(c.world),
        "e2e": {"value": e2e_pinned, "unit": UNIT, "h2d_bytes_per_step": 2 * N_PAIRS * 8,
                "d2h_bytes_per_step": N_PAIRS * 8, "steps": reps,
                "api": "dcs.HostStager.map -> noa_dcs_vmap_pinned_f64: pinned host tensors read and "
                       "written in place by the kernel over PCIe, result synchronised",
                "pageable": {"value": e2e_pageable,
                             "api": "dcs.HostStager.map -> noa_dcs_vmap_host_f64: pageable host "
                                    "tensors, chunked H2D / kernel / D2H pipeline"}},
        "gpu_launches": int(launches), "parity": parity, "roofline": roofline,
        "cpu_baseline": cpu_baseline, "_window": (t_wall0, t_wall1),
    }


def headline_pair(c):
    # a step = one pass over the 2^22 pairs; steps are multiplied up so the region is >= 0.5 s
    inner = max(1, int(600 / max(c.args.steps, 1)))
    block = pair_block(c, c.args.steps * inner, c.args.warmup, True)
    t0, t1 = block.pop("_window")
    block["ms_per_step"] *= inner
    block["config"]["step"] = f"{inner} passes over the rank's 2^22 pairs (timed region >= 0.5 s)"
    block["clocks"] = c.sampler.summary(t0, t1) if c.rank == 0 else None
    return block


# --------------------------------------------------------------------------------------------
# extras: the other kernels / BASELINE configs, each timed alone after the headline
# --------------------------------------------------------------------------------------------
def run_extras(c):
    torch, dcs, grids, physics, sharding = c.torch, c.dcs, c.grids, c.physics, c.sharding
    out = {}
    el, mass = physics.STANDARD_ROCK, physics.MUON_MASS
    rank0 = c.rank == 0
    solo = rank0 and c.world == 1
    checker, kind = load_cpu_checker() if rank0 else (None, None)

    # ---- config 2: pair production, 2^22 per GPU (round 1's headline), unless it IS the headline
    if c.args.workload == "table":
        block = pair_block(c, 50, 5, True)
        block.pop("_window")
        out["pair_production_2^22"] = block

    # ---- every element-wise kernel at its streaming size, with roofline, parity and CPU figure ---
    n_big = 1 << 24
    bufs = []
    for r in range(2):
        K, q = grids.set_b(n_big * c.world * 2, (r * c.world + c.rank) * n_big, n_big)
        bufs.append((torch.from_numpy(K).cuda(), torch.from_numpy(q).cuda(), K, q))
    res = torch.empty(n_big, dtype=torch.float64, device="cuda")
    for pr in dcs.PROCESSES:
        n = n_big if pr.index in (0, 3) else n_big // 4
        ms = c.timed(lambda i: dcs.vmap(pr)(res[:n], bufs[i % 2][0][:n], bufs[i % 2][1][:n], el,
                                            mass))
        rate = n / (ms * 1e-3)
        ex = c.profile.get(f"{pr.name}_fp64_instr_per_eval_executed")
        entry = {"pairs_per_gpu": n, "ms": ms, "evals_per_s": rate * c.world,
                 "fp64_frac": rate * ALGO_INSTR[pr.name] / c.fp64_peak,
                 "fp64_frac_executed": rate * ex / c.fp64_peak if ex else None,
                 "hbm_gbs": rate * ALGO_BYTES / 1e9,
                 "hbm_frac": rate * ALGO_BYTES / 1e9 / c.hbm_peak}
        if rank0:
            dcs.vmap(pr)(res[:n], bufs[0][0][:n], bufs[0][1][:n], el, mass)
            torch.cuda.synchronize()
            idx = np.arange(0, n, max(1, n // 70000))
            want = checker.vmap(pr.index, bufs[0][2][idx], bufs[0][3][idx], ROCK, MUON_MASS,
                                threads=checker.max_threads)
            entry["parity"] = compare(res[:n].cpu().numpy()[idx], want)
        if solo:
            m = 1 << (20 if pr.index in (0, 3) else 18)
            idx = np.linspace(0, n - 1, m).astype(np.int64)
            Ks, qs = np.ascontiguousarray(bufs[0][2][idx]), np.ascontiguousarray(bufs[0][3][idx])
            dt = median_time(lambda: checker.vmap(pr.index, Ks, qs, ROCK, MUON_MASS,
                                                  threads=checker.max_threads), 5)
            entry["cpu_baseline"] = {"value": m / dt, "unit": UNIT, "cores": checker.max_threads,
                                     "kind": kind, "sample": f"{m} pairs, dcs::pvmap, median of 5"}
        out[pr.name] = entry

    # ---- config 1: bremsstrahlung, 2^20 pairs: the reference's CPU cases beside the GPU ----------
    if solo:
        n1 = 1 << 20
        K1, q1 = grids.set_b(n1)
        K1d, q1d = torch.from_numpy(K1).cuda(), torch.from_numpy(q1).cuda()
        r1 = torch.empty_like(K1d)
        ms = c.timed(lambda i: dcs.cuda.vmap_bremsstrahlung(r1, K1d, q1d, el, mass), reps=20)
        t1 = median_time(lambda: checker.vmap(0, K1, q1, ROCK, MUON_MASS, threads=1), 5)
        tp = median_time(lambda: checker.vmap(0, K1, q1, ROCK, MUON_MASS,
                                              threads=checker.max_threads), 5)
        want = checker.vmap(0, K1, q1, ROCK, MUON_MASS, threads=checker.max_threads)
        out["config1_bremsstrahlung_2^20"] = {
            "gpu_ms": ms, "gpu_evals_per_s": n1 / (ms * 1e-3),
            "cpu_vmap_1_thread": {"evals_per_s": n1 / t1, "ms": t1 * 1e3,
                                  "case": "BremsstrahlungVectorisedLarge (measure-dcs-calc.cc:20-23), "
                                          "median of 5"},
            "cpu_pvmap": {"evals_per_s": n1 / tp, "ms": tp * 1e3, "threads": checker.max_threads,
                          "case": "BremsstrahlungVectorisedLargeOpenMP (measure-dcs-calc.cc:25-28), "
                                  "median of 5"},
            "kind": kind, "cpu": cpu_model(),
            "parity": compare(r1.cpu().numpy(), want)}
        out["bremsstrahlung"]["vs_reference_cuda"] = reference_cuda_head_to_head(c, bufs)
        del K1d, q1d, r1

    # ---- config 3: all four processes on water, 2^24 pairs --------------------------------------
    res4 = torch.empty((4, n_big), dtype=torch.float64, device="cuda")
    ms = c.timed(lambda i: dcs.cuda.vmap_material(res4, bufs[i % 2][0], bufs[i % 2][1],
                                                  physics.WATER, mass), reps=3, warm=1)
    instr = 2 * sum(ALGO_INSTR.values())
    entry = {"pairs_per_gpu": n_big, "ms": ms, "evals_per_s": 8 * n_big / (ms * 1e-3) * c.world,
             "fp64_frac": n_big / (ms * 1e-3) * instr / c.fp64_peak}
    if rank0:
        dcs.cuda.vmap_material(res4, bufs[0][0], bufs[0][1], physics.WATER, mass)
        torch.cuda.synchronize()
        idx = np.arange(0, n_big, 256)                       # 65 536 pairs x 4 processes
        sel = torch.from_numpy(idx).cuda()
        got = res4[:, sel].cpu().numpy()
        want = np.zeros_like(got)
        for p_ in range(4):
            acc = np.zeros(idx.size)
            for e_, w_ in zip(physics.WATER.elements, physics.WATER.fractions):
                acc = acc + w_ * checker.vmap(p_, bufs[0][2][idx], bufs[0][3][idx], tuple(e_),
                                              MUON_MASS, threads=checker.max_threads)
            want[p_] = acc
        entry["parity"] = compare(got, want)
        entry["parity"]["against"] = f"oracle ({kind}): sum_e w_e DCS_e on every 256th pair"
    out["water_all_four_2^24"] = entry
    del res4, res

    # ---- per-process table columns and the NCCL form of the exchange ----------------------------
    Kt = torch.from_numpy(grids.table_energies(N_K)).cuda()
    if rank0:
        d = torch.zeros((4, N_K), dtype=torch.float64, device="cuda")
        cc = torch.zeros_like(d)
        per = {}
        for pr in dcs.PROCESSES:
            per[pr.name] = c_local_timed(c, lambda i: dcs.cuda.tables(
                Kt, X_LOW, el, mass, MIN_POINTS, processes=(pr,), out=(d, cc)), reps=4, warm=1)
        per["all_four_single_gpu"] = c_local_timed(c, lambda i: dcs.cuda.tables(
            Kt, X_LOW, el, mass, MIN_POINTS, out=(d, cc)), reps=4, warm=1)
        out["table_build_per_process_ms_1gpu"] = per
    if c.world > 1:
        gather = sharding.TableBuilder(Kt, c.rank, c.world)
        ms = c.timed(lambda i: gather.build(X_LOW, el, mass, MIN_POINTS), reps=5, warm=2)
        out["table_build_nccl_all_gather"] = {"ms": ms, "evals_per_s": EVALS_PER_BUILD / (ms * 1e-3)}
        out["host_copy_bandwidth"] = host_copy_bandwidth(c)
        del gather

    # ---- the rest of the dcs.hh surface: per-energy, latency-sized launches ----
    if rank0:
        n = Kt.numel()
        z = lambda *shape: torch.zeros(shape, dtype=torch.float64, device="cuda")  # noqa: E731
        fCM, screen, fspin, invl, G, mu0, lbh, ms1 = (z(n, 2), z(n, 9), z(n), z(n), z(n, 2), z(n),
                                                      z(n), z(n))
        one = torch.ones(1, dtype=torch.float64, device="cuda")

        def coulomb_chain(i):
            dcs.coulomb_data(fCM, screen, fspin, invl, Kt, el, mass)
            dcs.coulomb_transport(G, screen, fspin, one)
            dcs.hard_scattering(mu0, lbh, G.view(1, n, 2), fCM.view(1, n, 2), screen.view(1, n, 9),
                                invl.view(1, n), fspin.view(1, n))

        out["coulomb_data+transport+hard_scattering_1e4"] = {
            "ms": c_local_timed(c, coulomb_chain), "energies": n, "launches": 3, "gpus": 1}
        ms = c_local_timed(c, lambda i: dcs.soft_scattering(ms1, Kt, el, mass))
        out["soft_scattering_1e4"] = {"ms": ms, "energies": n, "gpus": 1,
                                      "photonuclear_evals_per_s": n * 102 / (ms * 1e-3)}
        # per-material table assembly (PUMAS's steps over NOA's DCS), water, 180 nodes
        import oracle
        water = physics.WATER
        asm = {}
        ms = c_local_timed(c, lambda i: asm.update(
            dcs.cuda.material_assembly(Kt, X_LOW, water, mass, 180)), reps=8, warm=2)
        t0 = time.perf_counter()
        want = oracle.material_assembly(checker, [tuple(e) for e in water.elements],
                                        list(water.fractions), MUON_MASS,
                                        grids.table_energies(N_K), X_LOW, 180,
                                        threads=checker.max_threads)
        cpu_s = time.perf_counter() - t0
        same = all(np.array_equal(asm[k].cpu().numpy().reshape(want[k].shape), want[k],
                                  equal_nan=True)
                   for k in ("elem", "cs", "cel", "straggling", "csf", "cs_total", "xt"))
        same = same and float(asm["kt"].item()) == want["kt"] and int(asm["it"].item()) == want["it"]
        out["material_assembly_water_1e4"] = {
            "ms": ms, "energies": n, "nodes": 180, "elements": 2, "gpus": 1,
            "parity": {"bit_exact": bool(same), "against": f"oracle/material_oracle.c over the {kind} "
                       "DCS: every array (element tables, mix, cumulative fractions, threshold "
                       "row, x_t search) of all 10^4 energies"},
            "cpu_baseline": {"seconds": cpu_s, "cores": checker.max_threads, "kind": kind}}
    del bufs
    out["multi_material_sweep_2^28"] = run_sweep(c, Kt, checker, kind)
    return out


def c_local_timed(c, fn, reps=10, warm=3):
    torch = c.torch
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(c.stream)
    for i in range(reps):
        fn(i)
    b.record(c.stream)
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def reference_cuda_head_to_head(c, bufs):
    """The reference's own CUDA bremsstrahlung kernel (src/noa/pms/dcs.cuh:30-41, compiled
    unmodified into oracle/_ref/libnoa_ref_cuda.so) against noa_dcs_vmap_f64 on the same B200, same
    buffers, both synchronised: at 10^4 pairs (the published case) and 2^24."""
    torch, dcs = c.torch, c.dcs
    path = os.path.join(ROOT, "oracle", "_ref", "libnoa_ref_cuda.so")
    if not os.path.exists(path):
        return {"unavailable": "oracle/_ref/libnoa_ref_cuda.so not built (make -C oracle ref_cuda)"}
    try:
        ref = ctypes.CDLL(path)
    except OSError as exc:
        return {"unavailable": repr(exc)}
    vp, f64, i64, i32 = ctypes.c_void_p, ctypes.c_double, ctypes.c_int64, ctypes.c_int32
    ref.noa_ref_cuda_vmap_bremsstrahlung.argtypes = [vp, vp, vp, i64, f64, f64, i32, f64]
    ref.noa_ref_cuda_vmap_bremsstrahlung.restype = None
    ref.noa_ref_cuda_sync.restype = ctypes.c_int
    el, mass = c.physics.STANDARD_ROCK, c.physics.MUON_MASS
    result = {}
    for n in (10000, 1 << 24):
        Kd, qd = bufs[0][0][:n], bufs[0][1][:n]
        a = torch.empty(n, dtype=torch.float64, device="cuda")
        b = torch.empty(n, dtype=torch.float64, device="cuda")

        def theirs():
            ref.noa_ref_cuda_vmap_bremsstrahlung(vp(b.data_ptr()), vp(Kd.data_ptr()),
                                                 vp(qd.data_ptr()), n, el[0], el[1], el[2], mass)
            ref.noa_ref_cuda_sync()

        def ours():        # same level as theirs: the C entry point through ctypes, then a sync
            c.lib.noa_dcs_vmap_f64(0, vp(Kd.data_ptr()), vp(qd.data_ptr()), vp(a.data_ptr()), n,
                                   el[0], el[1], el[2], mass, None)
            ref.noa_ref_cuda_sync()

        t_ref = median_time(theirs, 31, 5)
        t_our = median_time(ours, 31, 5)
        diff = compare(a.cpu().numpy(), b.cpu().numpy())
        result[f"n={n}"] = {"reference_cuda_us": t_ref * 1e6, "noa_b200_us": t_our * 1e6,
                            "speedup": t_ref / t_our,
                            "max_rel_diff_between_the_two": diff["max_rel"],
                            "timing": "host wall clock around the C call (ctypes) + device "
                                      "synchronise, legacy default stream, median of 31"}
    result["note"] = ("reference kernel = launch_kernel<lambda> one thread per pair with libdevice "
                      "pow/log (not bit-identical to its own CPU path); ours is bit-identical to "
                      "the CPU path")
    return result


def host_copy_bandwidth(c):
    """What the host delivers to all ranks at once: every rank streams 256 MiB pinned buffers H2D
    and D2H concurrently with plain cudaMemcpyAsync (torch copy_), the ceiling of the e2e path at
    this N."""
    torch = c.torch
    n = 1 << 25
    h_in = torch.empty(n, dtype=torch.float64).pin_memory()
    h_out = torch.empty(n, dtype=torch.float64).pin_memory()
    d_in = torch.empty(n, dtype=torch.float64, device="cuda")
    d_out = torch.zeros(n, dtype=torch.float64, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    res = {}
    for label, both in (("h2d_only", False), ("h2d_and_d2h", True)):
        c.barrier()
        t0 = time.perf_counter()
        for _ in range(4):
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
            if both:
                with torch.cuda.stream(s2):
                    h_out.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize()
        dt = c.max_over_ranks(time.perf_counter() - t0)
        c.barrier()
        res[label + "_gbs_per_rank_each_direction"] = 4 * n * 8 / dt / 1e9
    res["ranks"] = c.world
    res["numa_node_of_rank0"] = c.numa_node
    return res


def run_sweep(c, Kt, checker, kind):
    """BASELINE.json configs[4]: water, standard rock, iron, lead; 2^26 (K, q) pairs per material
    (2^28 in total), all four processes per pair (per element, mass-fraction mixed for water); the
    flattened (material, pair) space is cut into contiguous shards of equal cost (a water pair
    counts twice), one per rank, outputs stay sharded; plus the DEL/CEL tables (10^4 x 1002 nodes)
    of the five distinct elements, assembled on every rank.  Strong scaling: the total work is
    fixed.  Parity: a sample of this rank's outputs of every segment and one element table."""
    torch, dcs, grids, physics, sharding = c.torch, c.dcs, c.grids, c.physics, c.sharding
    n_mat = 1 << 26
    materials = physics.SWEEP_MATERIALS
    segs = sharding.sweep_segments(n_mat, [len(m.elements) for m in materials], c.rank, c.world)
    grids_dev = {}
    for _, lo, hi in segs:
        if (lo, hi) not in grids_dev:
            K, q = grids.set_b(n_mat, lo, hi - lo)
            grids_dev[(lo, hi)] = (torch.from_numpy(K).cuda(), torch.from_numpy(q).cuda(), K, q)
    longest = max((hi - lo for _, lo, hi in segs), default=1)
    res = torch.empty(4 * longest, dtype=torch.float64, device="cuda")
    elements = []
    for m in materials:
        for e in m.elements:
            if e not in elements:
                elements.append(e)
    builder = sharding.make_table_builder(Kt, c.rank, c.world)
    tables = {}

    def sweep(i):
        for m, lo, hi in segs:
            Kd, qd = grids_dev[(lo, hi)][:2]
            dcs.cuda.vmap_material(res[:4 * (hi - lo)], Kd, qd, materials[m], physics.MUON_MASS)
        for e in elements:
            tables[e] = builder.build(X_LOW, e, physics.MUON_MASS, MIN_POINTS).clone()

    ms = c.timed(sweep, reps=2, warm=1)
    evals = sum(4 * len(m.elements) for m in materials) * n_mat + len(elements) * EVALS_PER_BUILD
    # parity on this rank's shard (every rank checks its own; the flags are reduced)
    import oracle
    port = oracle.load_port()
    ok, checked = 1.0, 0
    for m, lo, hi in segs:
        Kd, qd, K, q = grids_dev[(lo, hi)]
        n = hi - lo
        dcs.cuda.vmap_material(res[:4 * n], Kd, qd, materials[m], physics.MUON_MASS)
        torch.cuda.synchronize()
        idx = np.unique(np.linspace(0, n - 1, 4096).astype(np.int64))
        got = res[:4 * n].view(4, n)[:, torch.from_numpy(idx).cuda()].cpu().numpy()
        for p_ in range(4):
            acc = np.zeros(idx.size)
            for e_, w_ in zip(materials[m].elements, materials[m].fractions):
                acc = acc + w_ * port.vmap(p_, K[idx], q[idx], tuple(e_), MUON_MASS, threads=4)
            if not np.array_equal(got[p_], acc):
                ok = 0.0
            checked += idx.size
    e_last = elements[-1]
    idx = np.arange(0, N_K, 50)
    Ks = grids.table_energies(N_K)[idx]
    t_last = tables[e_last].cpu().numpy()
    for p_ in range(4):
        for ig in (0, 1):
            want = port.vmap_integral(p_, ig, Ks, X_LOW, MIN_POINTS, tuple(e_last), MUON_MASS,
                                      threads=4)
            if not np.array_equal(t_last[ig, p_][idx], want):
                ok = 0.0
            checked += idx.size
    all_ok = c.min_over_ranks(ok) == 1.0
    return {"ms": ms, "evals_per_s": evals / (ms * 1e-3), "scaling": "strong",
            "pairs_total": n_mat * len(materials), "materials": [m.name for m in materials],
            "table_elements": len(elements), "exchange": type(builder).__name__,
            "includes": "sharded element-wise sweep (no collective) + per-element tables on every "
                        "rank",
            "parity": {"bit_exact": bool(all_ok), "checked_on_rank0": checked,
                       "against": "oracle C port: 4096 pairs of every segment of every rank x 4 "
                                  "processes (mass-fraction mixed), and every 50th energy of the "
                                  "last element's exchanged table, all ranks reduced"}}


if __name__ == "__main__":
    sys.exit(main())
