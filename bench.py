#!/usr/bin/env python
"""Benchmark of the muon DCS hot path (the measure-dcs-calc harness of this repo).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--no-extras]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One JSON line on rank 0.  Headline workload = BASELINE.json configs[1]: pair-production DCS
(nested 8-node Gauss-Legendre) on standard rock, 2^22 (K, q) pairs per GPU ("set B" synthetic grid,
SURVEY.md 8(d)); a step is one pass over the 2^22 pairs of the rank.  Metric: DCS evaluations per
second, FP64.

  value      device-resident throughput: K back-to-back launches timed with CUDA events on the
             launching stream, barrier + synchronize on both sides, max over ranks.  Inputs rotate
             over 4 distinct buffer sets (384 MB > 126 MB L2) so no step re-reads L2-warm data.
  e2e        the same through the host-buffer entry point (pinned host tensors in, host tensor
             out): the kernel reads K, q from and writes the result to host memory over PCIe
             inside the timed region (h2d / d2h bytes are what crosses the link).
  roofline   pair production is FP64-pipe bound (about 3500 FP64-pipe instructions and 24 bytes per
             evaluation): achieved = evals/s x 3500 x 2 flop, peak = the DFMA rate
             measured live by the library's dependent-chain-free probe kernel (the driver's
             MEASURED_PEAKS.json has no FP64 entry); the HBM view (24 B/eval against
             MEASURED_PEAKS.json hbm_gbs) is given beside it.
  cpu_baseline  the reference's own CPU code (oracle/_ref = unmodified headers compiled here; else
             the C port) with all host threads on a bounded sample of the same workload.
  extras     the other kernels / BASELINE configs (streaming DCS, water, table build), measured
             after the headline region; informational.

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

N_PAIRS = 1 << 22                 # per GPU (BASELINE configs[1])
ROTATE = 4                        # distinct input/output buffer sets
METRIC = "dcs_evals_per_sec_fp64"
UNIT = "evals/s"
# algorithmic FP64-pipe instructions and bytes per evaluation (SURVEY.md 8(a),(d))
ALGO_INSTR = {"bremsstrahlung": 160, "pair_production": 3500, "photonuclear": 6900,
              "ionisation": 165}
ALGO_BYTES = 24
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
                 "--format=csv,noheader,nounits", "-lms", "100"],
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
        sm, smax, reasons, power = [], [], set(), []
        for t, line in self.rows:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                clk, mx = float(f[1]), float(f[2])
            except ValueError:
                continue
            if t0 <= t <= t1 + 0.15:
                sm.append(clk)
                try:
                    power.append(float(f[3]))
                except ValueError:
                    pass
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                      "sw_power_cap"), f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            smax.append(mx)
        if not sm:      # region shorter than the sampling period: take every sample we have
            for t, line in self.rows:
                f = [x.strip() for x in line.split(",")]
                try:
                    sm.append(float(f[1]))
                except (ValueError, IndexError):
                    pass
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------------
# CPU reference arm
# --------------------------------------------------------------------------------------------
def load_cpu_checker():
    import oracle
    ref = oracle.load_reference()
    checker, kind = (ref, "reference") if ref is not None else (oracle.load_port(), "port")
    checker.use_all_cores()
    return checker, kind


def cpu_throughput(process, K, q, target_seconds, checker):
    """evals/s of the reference CPU path (all host threads).  The whole 2^22-pair workload takes
    well under a second on a multi-core host, so the sample is the full workload repeated until
    about `target_seconds` of CPU work has been done; the best pass is reported."""
    threads = checker.max_threads
    probe = min(K.size, 1 << 16)
    t = time.perf_counter()
    checker.vmap(process, K[:probe], q[:probe], ROCK, MUON_MASS, threads=threads)
    dt = max(time.perf_counter() - t, 1e-6)
    n = int(min(K.size, max(probe, probe * target_seconds / dt)))
    if n < K.size:
        # strided sample so it spans the same (K, q) distribution as the full workload
        idx = np.linspace(0, K.size - 1, n).astype(np.int64)
        Ks, qs = np.ascontiguousarray(K[idx]), np.ascontiguousarray(q[idx])
    else:
        Ks, qs = K, q
    best, spent, passes = None, 0.0, 0
    while passes < 1 or (spent < target_seconds and passes < 50):
        t = time.perf_counter()
        checker.vmap(process, Ks, qs, ROCK, MUON_MASS, threads=threads)
        dt = time.perf_counter() - t
        best = dt if best is None else min(best, dt)
        spent += dt
        passes += 1
    return n / best, n, threads, spent, passes


def cpu_table_sample(grids, n_sample=40):
    """The reference's own table integrals (dcs::vmap_integral(recoil_integral), serial only:
    dcs.hh:115-130) timed on a bounded sample of config 4: every 250th of the 10^4 energies, the
    eight integrals each, 1000 points; plus the harness-side OpenMP loop over energies SURVEY 8(d)
    asks for (on a sample large enough to keep every thread busy).  Scaled to the full table by
    the sample fraction."""
    checker, kind = load_cpu_checker()
    grid = grids.table_energies(10000)

    def run(K, threads):
        t = time.perf_counter()
        for process in range(4):
            for integrand in (0, 1):
                checker.vmap_integral(process, integrand, K, 0.05, 1000, ROCK, MUON_MASS,
                                      threads=threads)
        return (time.perf_counter() - t) * (grid.size / K.size)

    Ks = grid[:: grid.size // n_sample].copy()
    serial = run(Ks, 1)
    threads = checker.max_threads
    Kp = grid[:: max(1, grid.size // (64 * threads))].copy()
    parallel = run(Kp, threads)
    # the reference evaluates every node once per integrand: 8 x 1002 x n_K node evaluations
    nodes = 8 * 1002 * grid.size
    return {"kind": kind, "sample": f"serial: {Ks.size} of the 10^4 energies, OpenMP: {Kp.size}; "
                                    "8 integrals each, 1000 points, scaled to the full table",
            "serial_s_full_table": serial, "serial_evals_per_s": nodes / serial,
            "omp_threads": threads, "omp_s_full_table": parallel,
            "omp_evals_per_s": nodes / parallel,
            "note": "reference is serial (dcs.hh:115-130); the OpenMP figure is a harness-side "
                    "parallel loop over energies around the unmodified closure"}


def run_reference_arm(args):
    rank = env_int("RANK", 0)
    if rank != 0:
        return 0
    from noa_b200 import grids
    checker, kind = load_cpu_checker()
    K, q = grids.set_b(N_PAIRS)
    threads = checker.max_threads
    # size the per-step sample so that (steps + warmup) finish within about two minutes
    probe = 1 << 15
    t = time.perf_counter()
    checker.vmap(1, K[:probe], q[:probe], ROCK, MUON_MASS, threads=threads)
    rate = probe / max(time.perf_counter() - t, 1e-6)
    budget = 100.0 / max(args.steps + args.warmup, 1)
    n = int(min(N_PAIRS, max(probe, rate * min(budget, 20.0))))
    idx = np.linspace(0, N_PAIRS - 1, n).astype(np.int64)
    Ks, qs = np.ascontiguousarray(K[idx]), np.ascontiguousarray(q[idx])
    for _ in range(args.warmup):
        checker.vmap(1, Ks, qs, ROCK, MUON_MASS, threads=threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        checker.vmap(1, Ks, qs, ROCK, MUON_MASS, threads=threads)
    dt = time.perf_counter() - t0
    value = n * args.steps / dt
    sample = (f"{n} of the 2^22 set-B pairs per step (evenly strided), dcs::pvmap(pair_production) "
              f"on {threads} OpenMP threads")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind,
                         "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    _emit(line)
    return 0


def workload_config(n_gpus):
    return {"workload": "pair_production DCS (nested 8-node Gauss-Legendre), standard rock, muon, "
                        "2^22 (K,q) pairs per GPU, synthetic set B (BASELINE.json configs[1])",
            "pairs_per_gpu": N_PAIRS, "element": "standard_rock", "process": "pair_production",
            "parallelism": f"{n_gpus} independent shard(s), no data-path collective; each rank bound "
                           "to the NUMA node of its GPU" if n_gpus > 1 else
                           "1 shard, no data-path collective",
            "l2": f"inputs/outputs rotate over {ROTATE} buffer sets "
                  f"({ROTATE * N_PAIRS * 24 >> 20} MiB > 126 MB L2)"}


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


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="noa_b200", choices=["noa_b200", "reference"])
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup

    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist
    from noa_b200 import dcs, grids, _lib, physics

    world = env_int("WORLD_SIZE", 1)
    rank = env_int("RANK", 0)
    local_rank = env_int("LOCAL_RANK", 0)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: noa_b200 has no CPU path "
                         "(use --impl reference for the CPU reference arm)")
    torch.cuda.set_device(local_rank)
    distributed = world > 1
    from noa_b200 import sharding as _sharding
    numa_node = _sharding.bind_to_gpu_numa_node(local_rank) if distributed else None
    if distributed:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = _lib.require_device()
    stream = torch.cuda.current_stream()
    st = ctypes.c_void_p(stream.cuda_stream)

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if not distributed:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- inputs: this rank's shard of the global set-B grid, ROTATE distinct copies ------------
    total = N_PAIRS * world
    sets = []
    for r in range(ROTATE):
        # each copy is a different slice phase of the same distribution (distinct data in HBM)
        K, q = grids.set_b(total * ROTATE, (r * world + rank) * N_PAIRS, N_PAIRS)
        sets.append((torch.from_numpy(K).cuda(), torch.from_numpy(q).cuda(),
                     torch.empty(N_PAIRS, dtype=torch.float64, device="cuda")))
    K0, q0 = grids.set_b(total, rank * N_PAIRS, N_PAIRS)

    def step(i):
        Kd, qd, out = sets[i % ROTATE]
        dcs.vmap(dcs.pair_production)(out, Kd, qd, physics.STANDARD_ROCK, physics.MUON_MASS)

    # ---- FP64 peak probe (roofline denominator), timed alone ------------------------------------
    sink = torch.zeros(8, dtype=torch.float64, device="cuda")
    sms = torch.cuda.get_device_properties(local_rank).multi_processor_count

    def probe_once(iters=20000, threads=256, per_sm=8):
        blocks = sms * per_sm
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        _lib.check(lib.noa_dcs_fp64_probe(iters, blocks, threads,
                                          ctypes.c_void_p(sink.data_ptr()), st))
        b.record(stream)
        torch.cuda.synchronize()
        return blocks * threads * iters * 16 / (a.elapsed_time(b) * 1e-3)

    probe_once(2000)
    fp64_peak_instr = max(probe_once() for _ in range(5))        # DFMA / s

    # ---- headline timed region ----------------------------------------------------------------
    sampler = ClockSampler(local_rank if "CUDA_VISIBLE_DEVICES" not in os.environ else
                           int(os.environ["CUDA_VISIBLE_DEVICES"].split(",")[local_rank]))
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    for i in range(args.warmup):
        step(i)
    barrier()
    launches0 = lib.noa_dcs_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.perf_counter()
    e0.record(stream)
    for i in range(args.steps):
        step(i)
    e1.record(stream)
    barrier()
    t_wall1 = time.perf_counter()
    launches = lib.noa_dcs_launch_count() - launches0
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    # keep the GPU busy a little longer when the region is shorter than the sampler period
    if rank == 0 and (t_wall1 - t_wall0) < 0.6:
        t_extra0 = time.perf_counter()
        while time.perf_counter() - t_extra0 < 0.8:
            step(0)
            torch.cuda.synchronize()
        t_wall1 = time.perf_counter()
    ms_per_step = ms_total / args.steps
    value = total * args.steps / (ms_total * 1e-3)

    # ---- e2e: pinned host buffers through the host entry point ----------------------------------
    stager = dcs.HostStager()
    Kh = torch.from_numpy(K0).pin_memory()
    qh = torch.from_numpy(q0).pin_memory()
    outh = torch.empty(N_PAIRS, dtype=torch.float64).pin_memory()
    e2e_steps = max(3, min(args.steps, 20))
    for _ in range(3):
        stager.map(dcs.pair_production, Kh, qh, physics.STANDARD_ROCK, physics.MUON_MASS, out=outh)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        stager.map(dcs.pair_production, Kh, qh, physics.STANDARD_ROCK, physics.MUON_MASS, out=outh)
    torch.cuda.synchronize()
    e2e_dt = max_over_ranks(time.perf_counter() - t0)
    barrier()
    e2e_value = total * e2e_steps / e2e_dt
    checksum = float(outh[::4097].sum())      # the result is read on the host
    stager.close()

    if rank == 0:
        sampler.stop()
    clocks = sampler.summary(t_wall0, t_wall1) if rank == 0 else None

    # ---- roofline of the dominant (only) kernel of the step ------------------------------------
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        hbm_peak, hbm_src = float(peaks["hbm_gbs"]), "MEASURED_PEAKS.json"
    except Exception:
        hbm_peak, hbm_src = 6650.0, "fallback (B200_PROFILING.md)"
    per_gpu_rate = N_PAIRS / (ms_per_step * 1e-3)
    achieved_tflops = per_gpu_rate * ALGO_INSTR["pair_production"] * 2 / 1e12
    peak_tflops = fp64_peak_instr * 2 / 1e12
    traffic, executed = None, None
    try:
        prof = json.load(open(os.path.join(ROOT, "profiles", "latest_traffic.json")))
        traffic = prof.get("pair_production_dram_bytes_per_launch")
        fp64_executed = prof.get("pair_production_fp64_instr_per_eval_executed")
        if fp64_executed:
            # the census numerator is SURVEY's fixed figure; this is what the kernel really issues
            executed = {"fp64_instr_per_eval": fp64_executed,
                        "instr_per_eval": prof.get("pair_production_instr_per_eval_executed"),
                        "fp64_pipe_frac": per_gpu_rate * fp64_executed / fp64_peak_instr,
                        "source": prof.get("executed_source")}
    except Exception:
        pass
    roofline = {
        "bound": "fp64", "kernel": "vmap_kernel<pair_production>",
        "achieved": achieved_tflops, "peak": peak_tflops, "unit": "TFLOP/s",
        "frac": achieved_tflops / peak_tflops, "traffic": traffic,
        "peak_source": "measured live: noa_dcs_fp64_probe (16 independent DFMA chains/thread), "
                       f"{fp64_peak_instr / 1e12:.2f} T DFMA/s of measured",
        "algorithmic": {"fp64_pipe_instr_per_eval": ALGO_INSTR["pair_production"],
                        "bytes_per_eval": ALGO_BYTES, "evals_per_launch": N_PAIRS},
        "kernel_ms": ms_per_step, "executed": executed,
        "hbm_view": {"achieved": per_gpu_rate * ALGO_BYTES / 1e9, "peak": hbm_peak,
                     "unit": "GB/s", "frac": per_gpu_rate * ALGO_BYTES / 1e9 / hbm_peak,
                     "peak_source": hbm_src + " (of measured)"},
    }

    # ---- extras: the other kernels / configs ---------------------------------------------------
    extras = None
    if not args.no_extras:
        extras = run_extras(torch, dist, dcs, grids, physics, lib, rank, world, distributed,
                            fp64_peak_instr, hbm_peak, barrier, max_over_ranks)

    # ---- CPU baseline (rank 0, N = 1 only) ------------------------------------------------------
    cpu_baseline = None
    if rank == 0 and world == 1:
        checker, kind = load_cpu_checker()
        rate, n_s, threads, spent, passes = cpu_throughput(1, K0, q0, args.cpu_seconds, checker)
        cpu_baseline = {"value": rate, "unit": UNIT, "cores": threads, "kind": kind,
                        "sample": f"{n_s} of the 2^22 set-B pairs x {passes} passes "
                                  f"({spent:.1f} s of dcs::pvmap(pair_production) on {threads} "
                                  f"OpenMP threads), best pass"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": workload_config(world),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 2 * N_PAIRS * 8,
                    "d2h_bytes_per_step": N_PAIRS * 8, "steps": e2e_steps,
                    "api": "dcs.HostStager.map -> noa_dcs_vmap_pinned_f64: pinned host tensors read and "
                           "written in place by the kernel over PCIe, result synchronised",
                    "checksum": checksum},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roofline,
            "cpu_baseline": cpu_baseline,
            "extras": extras,
        }
        _emit(line)
    if distributed:
        dist.destroy_process_group()
    return 0


def run_extras(torch, dist, dcs, grids, physics, lib, rank, world, distributed, fp64_peak, hbm_peak,
               barrier, max_over_ranks):
    """Other kernels of the path, each timed alone (CUDA events, 3 warm-up + 10 timed launches on
    rotating buffers).  Streaming kernels at 2^24 pairs per GPU; config 3 (water, 2^24) and
    config 4 (table build, strong-scaled over ranks with the all-gather inside the timed region)."""
    from noa_b200 import sharding
    stream = torch.cuda.current_stream()
    out = {}

    def timed(fn, reps=10, warm=3):
        for i in range(warm):
            fn(i)
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for i in range(reps):
            fn(i)
        b.record(stream)
        barrier()
        return max_over_ranks(a.elapsed_time(b)) / reps

    n_big = 1 << 24
    bufs = []
    for r in range(2):
        K, q = grids.set_b(n_big * world * 2, (r * world + rank) * n_big, n_big)
        bufs.append((torch.from_numpy(K).cuda(), torch.from_numpy(q).cuda()))
    res = torch.empty(n_big, dtype=torch.float64, device="cuda")
    for pr in dcs.PROCESSES:
        n = n_big if pr.index in (0, 3) else n_big // 4
        ms = timed(lambda i: dcs.vmap(pr)(res[:n], bufs[i % 2][0][:n], bufs[i % 2][1][:n],
                                          physics.STANDARD_ROCK, physics.MUON_MASS))
        rate = n / (ms * 1e-3)
        out[pr.name] = {"pairs_per_gpu": n, "ms": ms, "evals_per_s": rate * world,
                        "fp64_frac": rate * ALGO_INSTR[pr.name] / fp64_peak,
                        "hbm_gbs": rate * ALGO_BYTES / 1e9,
                        "hbm_frac": rate * ALGO_BYTES / 1e9 / hbm_peak}
    # config 3: all four processes on water, 2^24 pairs (evaluations = n x 4 processes x 2 elements)
    res4 = torch.empty((4, n_big), dtype=torch.float64, device="cuda")
    ms = timed(lambda i: dcs.cuda.vmap_material(res4, bufs[i % 2][0], bufs[i % 2][1],
                                                physics.WATER, physics.MUON_MASS), reps=3, warm=1)
    instr = 2 * sum(ALGO_INSTR.values())
    out["water_all_four_2^24"] = {"pairs_per_gpu": n_big, "ms": ms,
                                  "evals_per_s": 8 * n_big / (ms * 1e-3) * world,
                                  "fp64_frac": n_big / (ms * 1e-3) * instr / fp64_peak}
    del res4, bufs, res
    # config 4: table build, 10^4 energies x 1002 nodes x 4 processes, sharded cyclically over the
    # ranks, finished tables all-gathered (NCCL) inside the timed region -> strong scaling
    Kt = torch.from_numpy(grids.table_energies(10000)).cuda()
    nodes = 10000 * 1002 * 4
    builders = [("peer_scatter" if world > 1 else "single_gpu",
                 sharding.make_table_builder(Kt, rank, world))]
    if world > 1:
        builders.append(("nccl_all_gather", sharding.TableBuilder(Kt, rank, world)))
    for label, builder in builders:
        ms = timed(lambda i: builder.build(dcs.X_FRACTION, physics.STANDARD_ROCK,
                                           physics.MUON_MASS, 1000), reps=5, warm=2)
        key = "table_build_1e4x1002" if label != "nccl_all_gather" else \
            "table_build_1e4x1002_nccl_all_gather"
        out[key] = {
            "ms": ms, "evals_per_s": nodes / (ms * 1e-3), "scaling": "strong",
            "exchange": type(builder).__name__,
            "fp64_frac_vs_nominal_census": (10000 * 1002 * 10800 / fp64_peak) / (ms * 1e-3) / world,
            "includes": "every rank ends with the full [2,4,n_K] table" if world > 1
            else "single GPU"}
    del builders
    if rank == 0 and world == 1:
        out["table_build_1e4x1002"]["cpu_reference"] = cpu_table_sample(grids)
    # the rest of the dcs.hh surface: per-energy, latency-sized launches
    if rank == 0:
        n = Kt.numel()
        z = lambda *shape: torch.zeros(shape, dtype=torch.float64, device="cuda")
        fCM, screen, fspin, invl, G, mu0, lbh, ms1 = (z(n, 2), z(n, 9), z(n), z(n), z(n, 2), z(n),
                                                      z(n), z(n))
        one = torch.ones(1, dtype=torch.float64, device="cuda")

        def coulomb_chain(i):
            dcs.coulomb_data(fCM, screen, fspin, invl, Kt, physics.STANDARD_ROCK, physics.MUON_MASS)
            dcs.coulomb_transport(G, screen, fspin, one)
            dcs.hard_scattering(mu0, lbh, G.view(1, n, 2), fCM.view(1, n, 2), screen.view(1, n, 9),
                                invl.view(1, n), fspin.view(1, n))

        def local_timed(fn, reps=10, warm=3):
            for i in range(warm):
                fn(i)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            for i in range(reps):
                fn(i)
            b.record(stream)
            torch.cuda.synchronize()
            return a.elapsed_time(b) / reps

        ms = local_timed(coulomb_chain)
        out["coulomb_data+transport+hard_scattering_1e4"] = {"ms": ms, "energies": n,
                                                             "launches": 3, "gpus": 1}
        ms = local_timed(lambda i: dcs.soft_scattering(ms1, Kt, physics.STANDARD_ROCK,
                                                       physics.MUON_MASS))
        out["soft_scattering_1e4"] = {"ms": ms, "energies": n, "gpus": 1,
                                      "photonuclear_evals_per_s": n * 102 / (ms * 1e-3)}
    out["multi_material_sweep_2^28"] = run_sweep(torch, dcs, grids, physics, sharding, Kt, rank,
                                                 world, timed)
    return out


def run_sweep(torch, dcs, grids, physics, sharding, Kt, rank, world, timed):
    """BASELINE.json configs[4]: water, standard rock, iron, lead; 2^26 (K, q) pairs per material
    (2^28 in total), all four processes per pair (per element, mass-fraction mixed for water); the
    flattened (material, pair) space is cut into contiguous shards of equal cost (a water pair
    counts twice), one per rank, outputs stay sharded; plus the DEL/CEL tables (10^4 x 1002 nodes) of the five distinct elements, assembled
    on every rank.  Strong scaling: the total work is fixed."""
    n_mat = 1 << 26
    materials = physics.SWEEP_MATERIALS
    segs = sharding.sweep_segments(n_mat, [len(m.elements) for m in materials], rank, world)
    grids_dev = {}
    for _, lo, hi in segs:
        if (lo, hi) not in grids_dev:
            K, q = grids.set_b(n_mat, lo, hi - lo)
            grids_dev[(lo, hi)] = (torch.from_numpy(K).cuda(), torch.from_numpy(q).cuda())
    longest = max((hi - lo for _, lo, hi in segs), default=1)
    res = torch.empty(4 * longest, dtype=torch.float64, device="cuda")
    elements = []
    for m in materials:
        for e in m.elements:
            if e not in elements:
                elements.append(e)
    builder = sharding.make_table_builder(Kt, rank, world)
    tables = {}

    def sweep(i):
        for m, lo, hi in segs:
            Kd, qd = grids_dev[(lo, hi)]
            dcs.cuda.vmap_material(res[:4 * (hi - lo)], Kd, qd, materials[m], physics.MUON_MASS)
        for e in elements:
            tables[e] = builder.build(dcs.X_FRACTION, e, physics.MUON_MASS, 1000).clone()

    ms = timed(sweep, reps=2, warm=1)
    evals = sum(4 * len(m.elements) for m in materials) * n_mat + len(elements) * 10000 * 1002 * 4
    return {"ms": ms, "evals_per_s": evals / (ms * 1e-3), "scaling": "strong",
            "pairs_total": n_mat * len(materials), "materials": [m.name for m in materials],
            "table_elements": len(elements), "exchange": type(builder).__name__,
            "includes": "sharded element-wise sweep (no collective) + per-element tables on every "
                        "rank"}


if __name__ == "__main__":
    sys.exit(main())
