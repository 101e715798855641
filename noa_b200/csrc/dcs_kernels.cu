// B200 (sm_100a) kernels of the muon DCS hot path and the C ABI that exposes them
// (include/noa_dcs_b200.h).  Torch-free on purpose: this file compiles in well under a minute and
// is the whole product below the LibTorch boundary.
//
// Kernels
//   vmap_kernel<P, VEC>     element-wise DCS of one process, persistent grid-stride blocks; the two
//                           streaming processes take 4 pairs per thread and iteration (two 128-bit
//                           loads of K and of q issued before the first evaluation)
//   vmap_all_kernel         the four processes of one pair in one pass (16 B in, 32 B out)
//   vmap_mixture_element_kernel   one element's term of sum_e w_e * DCS_e (water = H + O)
//   table_rowpar / table_terms<P> / table_sum   (table_kernels.cuh) the DEL/CEL tables in the flat
//                           form: (row, node) units popped by warps, node terms through a
//                           workspace, summation in the reference's serial order by a second
//                           kernel that also delivers the rows (peers included, NVSwitch multicast
//                           stores, rank barrier) -- noa_dcs_table_ws_f64, noa_dcs_table_exchange_f64
//   table_kernel<MASK>      the same tables with one CTA per row and the terms in shared memory,
//                           for the calls that bring no workspace
//   threshold / straggling  (material_kernels.cuh) the per-material table assembly of SURVEY 8(f1)
//   coulomb_* / soft_scattering   (coulomb_kernels.cuh)
// Measurement kernels (FP64 pipe probes, the node-per-lane pair-production variant) live in
// probe_kernels.cu / libnoa_dcs_b200_probe.so, not here; this library keeps no mutable state
// beyond a launch counter and the cached SM count.
//
// Numerics: FP64 throughout, every operation IEEE (see dcs_math.cuh, glibm.cuh); this file must be
// compiled with -fmad=false.  The exp/log tables (4 KB) are staged into shared memory per CTA.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <utility>

#include "../../include/noa_dcs_b200.h"
#include "dcs_device.cuh"
#include "dcs_params.hh"

namespace noa_b200 {

// ------------------------------------------------------------------------------------------
// element-wise, one process
// ------------------------------------------------------------------------------------------
// VEC = pairs per thread and iteration: 1 (scalar loads; the quadrature-bound processes, whose
// integrand is instantiated once), 2 (one double2 of K and of q) or 4 (two of each, all four loads
// issued before the first evaluation: the streaming kernels were waiting on their own loads --
// long_scoreboard 2.5 of 12 stall cycles per issue with VEC = 2, profiles/r01_ncu_full_s4c.md).
template <int PROCESS, int VEC>
__global__ void __launch_bounds__(kThreads, MinBlocks<PROCESS>::value)
vmap_kernel(const double *__restrict__ K, const double *__restrict__ q, double *__restrict__ out,
            int64_t n, const __grid_constant__ Params p) {
    __shared__ StagedShared s_staged;
    const glibm::Tab T = stage_all(s_staged, p);
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    const int64_t tid = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    // K, q and out may also be pinned HOST buffers read and written in place over PCIe
    // (noa_dcs_vmap_pinned_f64); an explicit L2 prefetch of the next iteration's operands was
    // measured to gain nothing on either path and cost 3 % on the streaming kernels.
    if (VEC >= 2) {
        const int64_t n2 = n >> 1;
        const double2 *K2 = reinterpret_cast<const double2 *>(K);
        const double2 *q2 = reinterpret_cast<const double2 *>(q);
        double2 *o2 = reinterpret_cast<double2 *>(out);
        if (VEC == 4) {
            for (int64_t i = tid; i < n2; i += 2 * stride) {
                const int64_t i1 = i + stride;
                const bool two = i1 < n2;
                const double2 k0 = K2[i];
                const double2 r0 = q2[i];
                double2 k1 = k0, r1 = r0;
                if (two) {
                    k1 = K2[i1];
                    r1 = q2[i1];
                }
                double2 o;
                o.x = dcs_value<PROCESS, true>(k0.x, r0.x, p, T);
                o.y = dcs_value<PROCESS, true>(k0.y, r0.y, p, T);
                o2[i] = o;
                if (two) {
                    o.x = dcs_value<PROCESS, true>(k1.x, r1.x, p, T);
                    o.y = dcs_value<PROCESS, true>(k1.y, r1.y, p, T);
                    o2[i1] = o;
                }
            }
        } else {
            for (int64_t i = tid; i < n2; i += stride) {
                const double2 k = K2[i];
                const double2 r = q2[i];
                double2 o;
                o.x = dcs_value<PROCESS, true>(k.x, r.x, p, T);
                o.y = dcs_value<PROCESS, true>(k.y, r.y, p, T);
                o2[i] = o;
            }
        }
        if (tid == 0 && (n & 1)) out[n - 1] = dcs_value<PROCESS, true>(K[n - 1], q[n - 1], p, T);
    } else {
        for (int64_t i = tid; i < n; i += stride) {
            out[i] = dcs_value<PROCESS, true>(K[i], q[i], p, T);
        }
    }
}

// element-wise, all four processes of a pair: out[p * n + i]
__global__ void __launch_bounds__(kThreads, NOA_MINB_ALL)
vmap_all_kernel(const double *__restrict__ K, const double *__restrict__ q,
                double *__restrict__ out, int64_t n, const __grid_constant__ Params p) {
    __shared__ StagedShared s_staged;
    const glibm::Tab T = stage_all(s_staged, p);
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    for (int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double k = K[i], r = q[i];
        out[i] = dcs_value<0, true>(k, r, p, T);
        out[n + i] = dcs_value<1, true>(k, r, p, T);
        out[2 * n + i] = dcs_value<2, true>(k, r, p, T);
        out[3 * n + i] = dcs_value<3, true>(k, r, p, T);
    }
}

// One element of a mixture: out[slot * n + i] (+)= w * DCS_process(K[i], q[i]) for the processes of
// the mask, slot counting the processes present.  A material is evaluated element by element,
// one launch each (noa_dcs_vmap_mixture_f64): the element's Params then sit in the kernel
// parameter bank at fixed offsets and its reciprocals are staged once per CTA exactly as in the
// single-element kernels -- a kernel looping over the elements of a Mixture struct indexes the
// constant bank with a register and was 9 % slower than the sum of its parts.  FIRST: the slot is
// started as 0. + w * v, which is what `acc = 0.; acc += w * v` of the one-kernel form gives; the
// later elements add in composition order, so the result is bit-identical to that form.
template <bool FIRST>
__global__ void __launch_bounds__(kThreads, NOA_MINB_ALL)
vmap_mixture_element_kernel(const double *__restrict__ K, const double *__restrict__ q,
                            double *__restrict__ out, int64_t n, uint32_t process_mask, double w,
                            const __grid_constant__ Params p) {
    __shared__ StagedShared s_staged;
    const glibm::Tab T = stage_all(s_staged, p);
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    for (int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double k = K[i], r = q[i];
        double *o = out + i;
        if (process_mask & 1u) {
            *o = (FIRST ? 0. : *o) + w * dcs_value<0, true>(k, r, p, T);
            o += n;
        }
        if (process_mask & 2u) {
            *o = (FIRST ? 0. : *o) + w * dcs_value<1, true>(k, r, p, T);
            o += n;
        }
        if (process_mask & 4u) {
            *o = (FIRST ? 0. : *o) + w * dcs_value<2, true>(k, r, p, T);
            o += n;
        }
        if (process_mask & 8u) *o = (FIRST ? 0. : *o) + w * dcs_value<3, true>(k, r, p, T);
    }
}

}  // namespace noa_b200

#include "table_kernels.cuh"

namespace noa_b200 {

// ------------------------------------------------------------------------------------------
// host side of the C ABI
// ------------------------------------------------------------------------------------------
static std::atomic<int64_t> g_launches{0};

struct DeviceInfo {
    int sm_count = 0;
    bool ok = false;
};

static int device_info(DeviceInfo &info) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver)
                                         ? NOA_DCS_ENODEV
                                         : (int) e;
    static std::mutex mu;
    static int cached_sm[64] = {0};
    std::lock_guard<std::mutex> lock(mu);
    if (dev < 64 && cached_sm[dev]) {
        info.sm_count = cached_sm[dev];
        info.ok = true;
        return 0;
    }
    int sms = 0;
    e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return (int) e;
    if (dev < 64) cached_sm[dev] = sms;
    info.sm_count = sms;
    info.ok = true;
    return 0;
}

// Resident CTAs per SM of a kernel at kThreads threads.  The occupancy query costs a few
// microseconds -- a quarter of a 10^4-pair call -- so the answer is kept per kernel (all devices of
// a box are the same part).
static int blocks_per_sm(const void *kernel, int &per_sm, int threads = kThreads) {
    static std::mutex mu;
    static const void *keys[64];
    static int vals[64];
    static int count = 0;
    {
        std::lock_guard<std::mutex> lock(mu);
        for (int i = 0; i < count; i++)
            if (keys[i] == kernel) {
                per_sm = vals[i];
                return 0;
            }
    }
    int v = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, kernel, threads, 0);
    if (e != cudaSuccess) return (int) e;
    if (v < 1) v = 1;
    std::lock_guard<std::mutex> lock(mu);
    if (count < 64) {
        keys[count] = kernel;
        vals[count++] = v;
    }
    per_sm = v;
    return 0;
}

template <typename Kernel>
static int persistent_grid(Kernel kernel, int64_t work_items, int &blocks) {
    DeviceInfo info;
    int rc = device_info(info);
    if (rc) return rc;
    int per_sm = 0;
    rc = blocks_per_sm((const void *) kernel, per_sm);
    if (rc) return rc;
    const int64_t need = (work_items + kThreads - 1) / kThreads;
    const int64_t cap = (int64_t) info.sm_count * per_sm;
    blocks = (int) (need < cap ? need : cap);
    if (blocks < 1) blocks = 1;
    return 0;
}

static inline int after_launch() {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return (int) cudaPeekAtLastError();
}

static inline bool aligned16(const void *a, const void *b, const void *c) {
    return ((((uintptr_t) a) | ((uintptr_t) b) | ((uintptr_t) c)) & 15u) == 0;
}

template <int PROCESS>
static int launch_vmap(const double *K, const double *q, double *out, int64_t n, const Params &p,
                       cudaStream_t s) {
    int blocks = 0;
    // 128-bit loads/stores pay for the two streaming processes; the quadrature-bound ones keep
    // one pair per thread so the (large) integrand is instantiated once
    constexpr bool kStreaming = (PROCESS == 0 || PROCESS == 3);
    // latency-sized calls (less than one two-pair CTA per SM, e.g. the reference benchmark's 10^4
    // pairs) spread one pair per thread over twice as many SMs instead
    DeviceInfo info;
    int rc0 = device_info(info);
    if (rc0) return rc0;
    const bool wide = n >= (int64_t) info.sm_count * kThreads * 2;
    if (kStreaming && wide && aligned16(K, q, out)) {
        int rc = persistent_grid(vmap_kernel<PROCESS, NOA_STREAM_VEC>, n / NOA_STREAM_VEC, blocks);
        if (rc) return rc;
        vmap_kernel<PROCESS, NOA_STREAM_VEC><<<blocks, kThreads, 0, s>>>(K, q, out, n, p);
    } else {
        int rc = persistent_grid(vmap_kernel<PROCESS, 1>, n, blocks);
        if (rc) return rc;
        vmap_kernel<PROCESS, 1><<<blocks, kThreads, 0, s>>>(K, q, out, n, p);
    }
    return after_launch();
}

static int vmap_impl(int process, const double *K, const double *q, double *out, int64_t n,
                     const Params &p, cudaStream_t s) {
    switch (process) {
        case NOA_DCS_BREMSSTRAHLUNG: return launch_vmap<0>(K, q, out, n, p, s);
        case NOA_DCS_PAIR_PRODUCTION: return launch_vmap<1>(K, q, out, n, p, s);
        case NOA_DCS_PHOTONUCLEAR: return launch_vmap<2>(K, q, out, n, p, s);
        case NOA_DCS_IONISATION: return launch_vmap<3>(K, q, out, n, p, s);
    }
    return NOA_DCS_EINVAL;
}

static int vmap_all_impl(const double *K, const double *q, double *out, int64_t n, const Params &p,
                         cudaStream_t s) {
    int blocks = 0;
    int rc = persistent_grid(vmap_all_kernel, n, blocks);
    if (rc) return rc;
    vmap_all_kernel<<<blocks, kThreads, 0, s>>>(K, q, out, n, p);
    return after_launch();
}

// ---- table launches ---------------------------------------------------------------------------
using TableKernel = void (*)(const double *, int64_t, TableOut, TablePlan, Params);

static TableKernel table_kernel_for(unsigned mask) {
    switch (mask) {
        case 1u: return table_kernel<1u>;
        case 2u: return table_kernel<2u>;
        case 4u: return table_kernel<4u>;
        case 8u: return table_kernel<8u>;
    }
    return table_kernel<15u>;
}

static int rows_per_item(int process) {
    return (process == NOA_DCS_BREMSSTRAHLUNG || process == NOA_DCS_IONISATION)
                   ? TableCfg<0>::R
                   : TableCfg<1>::R;
}

// How a multi-process build of the row-per-CTA form is launched: one kernel per process, chained
// with programmatic dependent launch (each with the register budget its integrand wants), or one
// combined kernel.  Measured on config 4: 10^4 rows per process 4.31 ms split against 4.42
// combined; at 1 250 rows per process the two are within 1 % (0.592 / 0.590 ms,
// profiles/r02_flat_table_study.md) and the combined kernel is one launch.  Hence: split from
// kTableSplitRows rows on.
// NOA_DCS_TABLE_LAUNCH=split|combined (read once, not a run-time switch) forces one form so both
// stay measurable.
constexpr int64_t kTableSplitRows = 4096;
static bool table_launch_split(int64_t rows) {
    static const int forced = [] {
        const char *e = std::getenv("NOA_DCS_TABLE_LAUNCH");
        if (e && !std::strcmp(e, "combined")) return 0;
        if (e && !std::strcmp(e, "split")) return 1;
        return -1;
    }();
    if (forced >= 0) return forced == 1;
    return rows >= kTableSplitRows;
}

static int launch_table(TableKernel kernel, unsigned grid, bool dependent, cudaStream_t s,
                        const double *K, int64_t nK, const TableOut &out, const TablePlan &plan,
                        const Params &p) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kThreads);
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    if (dependent) {
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
    }
    cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, K, nK, out, plan, p);
    if (e != cudaSuccess) return (int) e;
    return after_launch();
}

constexpr int64_t kFlatRideRows = 4096;
struct TableOptions {
    double xhigh = 1.;
    int32_t second_power = 2;
    bool quadrature_only = false;
    double *workspace = nullptr;        // node terms of the flat form (table_kernels.cuh)
    int64_t workspace_doubles = 0;
};

static int64_t table_nodes(int32_t min_points) { return ((int64_t) min_points + 5) / 6 * 6; }

// {ln(K xlow), h} and {gamma, zeta} per row + 4 queue words + 4 processes x nK rows x nodes x {DEL term, CEL term}
static int64_t table_workspace_doubles(int64_t nK, int32_t min_points) {
    if (nK <= 0 || min_points < 1) return 0;
    return 4 * nK + 2 + 8 * nK * table_nodes(min_points);
}

template <typename... KArgs, typename... Args>
static int launch_chained(void (*kernel)(KArgs...), unsigned grid, unsigned block, bool dependent,
                          cudaStream_t s, Args &&...args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(block);
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    if (dependent) {
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
    }
    cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
    if (e != cudaSuccess) return (int) e;
    return after_launch();
}

template <int PROCESS>
static int launch_terms(const double *K, int64_t nK, const double2 *rowpar, const FlatQueues &fq,
                        const FlatPlan &fp, const Params &p, bool dependent, cudaStream_t s) {
    DeviceInfo info;
    int rc = device_info(info);
    if (rc) return rc;
    int per_sm = 0;
    rc = blocks_per_sm((const void *) table_terms_kernel<PROCESS>, per_sm, kFlatThreads);
    if (rc) return rc;
    // one warp per 32-node unit (FlatCfg::unit of them per pop), never more CTAs than fit at once
    const int64_t warps = (nK * (int64_t) fp.cells * 6 + 31) / 32;
    const int64_t need = (warps * 32 + kFlatThreads - 1) / kFlatThreads;
    const int64_t cap = (int64_t) info.sm_count * (per_sm < 1 ? 1 : per_sm);
    const unsigned blocks = (unsigned) (need < cap ? need : cap);
    FlatPlan plan = fp;
    plan.first_launch = dependent ? 0 : 1;
    // every terms kernel is a dependent launch: the first one of table_rowpar_kernel
    return launch_chained(table_terms_kernel<PROCESS>, blocks, (unsigned) kFlatThreads, true, s, K,
                          nK, rowpar, fq, plan, p);
}

// The flat form of a table build (table_kernels.cuh): row parameters, one terms kernel per
// process (heaviest first), one summation kernel that also delivers the rows -- to the peers too,
// rank barrier included, in the exchange form.
static int table_flat_impl(unsigned process_mask, const double *K, int64_t nK, double xlow,
                           int32_t min_points, const Params &p, TableOut out, cudaStream_t s,
                           const TableOptions &opt) {
    FlatPlan fp{};
    fp.cells = ((uint32_t) min_points + 5u) / 6u;
    fp.second_power = opt.second_power;
    fp.quadrature_only = opt.quadrature_only ? 1 : 0;
    fp.xlow = xlow;
    fp.xhigh = opt.xhigh;
    const int64_t nodes = table_nodes(min_points);
    double2 *rowpar = reinterpret_cast<double2 *>(opt.workspace);
    uint32_t *queues = reinterpret_cast<uint32_t *>(rowpar + nK);
    double2 *pairrow = rowpar + nK + 1;
    double2 *terms = pairrow + nK;
    table_rowpar_kernel<<<(unsigned) ((nK + 255) / 256), 256, 0, s>>>(K, nK, fp, rowpar, queues,
                                                                      pairrow, p);
    int rc = after_launch();
    if (rc) return rc;
    static const int heavy_first[4] = {NOA_DCS_PHOTONUCLEAR, NOA_DCS_PAIR_PRODUCTION,
                                       NOA_DCS_BREMSSTRAHLUNG, NOA_DCS_IONISATION};
    FlatSum fs{};
    fs.quadrature_only = fp.quadrature_only;
    fs.xlow = xlow;
    // Launch order: terms of photonuclear, pair production, bremsstrahlung + ionisation, then ONE
    // summation kernel over all rows.  (Adding the heavy rows up beside the last terms kernel
    // instead -- memory-bound next to FP64-bound -- was measured and lost: 4.22 against 4.16 ms, and
    // 0.603 against 0.565 ms on a 1/8 share; profiles/r02_flat_table_study.md.)
    int n_slots = 0;
    for (int i = 0; i < 4; i++) {
        const int pr = heavy_first[i];
        if (!((process_mask >> pr) & 1u)) continue;
        fs.process[n_slots] = pr;
        fs.out_row[n_slots] = pr;
        fs.terms[n_slots] = terms + (int64_t) n_slots * nK * nodes;
        n_slots++;
    }
    // bremsstrahlung and ionisation together go through one fused pass (heavy_first ends with
    // bremsstrahlung, ionisation: their slots are adjacent)
    const bool both_light = (process_mask & 9u) == 9u;
    // ... and when photonuclear is built as well they ride on its nodes (table_terms_kernel<6>) --
    // from kFlatRideRows rows on: 4.138 against 4.165 ms at 10^4 rows, but 0.575 against 0.565 ms
    // at 1 250 (longer units make a longer tail, and there are only ~11 units per warp)
    const bool ride = both_light && (process_mask & 4u) && nK >= kFlatRideRows;
    bool launched = false;
    for (int slot = 0; slot < n_slots; slot++) {
        const int pr = fs.process[slot];
        FlatQueues fq{};
        fq.terms_a = terms + (int64_t) slot * nK * nodes;
        fq.pairrow = pairrow;
        fq.queue_a = queues + slot;
        const bool dep = launched;
        if (pr == NOA_DCS_IONISATION && both_light) continue;          // rides with bremsstrahlung
        if (pr == NOA_DCS_BREMSSTRAHLUNG && ride) continue;            // rides with photonuclear
        if (pr == NOA_DCS_PHOTONUCLEAR && ride) {
            // slots in heavy-first order: photonuclear, [pair production,] bremsstrahlung, ionisation
            fq.terms_b = terms + (int64_t) (n_slots - 2) * nK * nodes;
            fq.terms_c = terms + (int64_t) (n_slots - 1) * nK * nodes;
            rc = launch_terms<6>(K, nK, rowpar, fq, fp, p, dep, s);
            if (rc) return rc;
            launched = true;
            continue;
        }
        if (pr == NOA_DCS_BREMSSTRAHLUNG && both_light) {
            fq.terms_b = fq.terms_a + nK * nodes;
            rc = launch_terms<4>(K, nK, rowpar, fq, fp, p, dep, s);
        } else {
            switch (pr) {
                case 0: rc = launch_terms<0>(K, nK, rowpar, fq, fp, p, dep, s); break;
                case 1: rc = launch_terms<1>(K, nK, rowpar, fq, fp, p, dep, s); break;
                case 2: rc = launch_terms<2>(K, nK, rowpar, fq, fp, p, dep, s); break;
                default: rc = launch_terms<3>(K, nK, rowpar, fq, fp, p, dep, s); break;
            }
        }
        if (rc) return rc;
        launched = true;
    }
    fs.first_slot = 0;
    fs.n_slots = n_slots;
    const int64_t chains = nK * n_slots;
    int64_t ctas = (chains + kSumWarps - 1) / kSumWarps;
    if (out.flags[0] != nullptr) {
        // exchange form: every CTA ends on a system-scope fence that waits for its stores into the
        // peers' tables (an NVLink round trip) while it still holds its share of the SM.  With one
        // CTA per 4 rows that is paid once per wave of CTAs -- 5.6 waves on a half share: +24 us of
        // a 2.1 ms build (tools/table_exchange_bench.py).  A grid that is resident at once, its
        // warps striding over the rows, pays it once (2.130 -> 2.111 ms at 2 GPUs; without peers the
        // capped grid is 0.4 % slower, so it is only used here).
        DeviceInfo info;
        rc = device_info(info);
        if (rc) return rc;
        int per_sm = 0;
        rc = blocks_per_sm((const void *) table_sum_kernel, per_sm, 32 * kSumWarps);
        if (rc) return rc;
        const int64_t cap = (int64_t) info.sm_count * per_sm;
        if (ctas > cap) ctas = cap;
    }
    const unsigned grid = (unsigned) ctas;
    out.total_ctas = grid;
    return launch_chained(table_sum_kernel, grid, 32u * kSumWarps, true, s, K, nK,
                          (uint32_t) nodes, fs, p, out);
}

static int table_impl(unsigned process_mask, bool single_row, const double *K, int64_t nK,
                      double xlow, int32_t min_points, double A, double I, int32_t Z, double mass,
                      TableOut out, void *stream, const TableOptions &opt = TableOptions()) {
    if (process_mask == 0 || process_mask > 15u || nK < 0 || min_points < 1)
        return NOA_DCS_EINVAL;
    cudaStream_t s = (cudaStream_t) stream;
    const bool exchange = out.flags[0] != nullptr;
    if (nK == 0 || (!out.del[0] && !out.cel[0])) {
        if (!exchange) return 0;
        out.total_ctas = 1;
        table_signal_kernel<<<1, 32, 0, s>>>(out);
        return after_launch();
    }
    if (!K) return NOA_DCS_EINVAL;
    if (nK > 0x3fffffffLL) return NOA_DCS_ERANGE;
    // the flat form counts 32-node units in 32-bit queue words
    const bool flat = !single_row && opt.workspace &&
                      opt.workspace_doubles >= table_workspace_doubles(nK, min_points) &&
                      nK * ((table_nodes(min_points) + 31) / 32) < 0xfff00000LL;
    if (exchange && !flat) return NOA_DCS_EINVAL;   // the exchange form is the flat form
    DeviceInfo info;
    int rc = device_info(info);
    if (rc) return rc;
    const Params p = make_params(A, I, Z, mass);
    if (!single_row && process_mask != 15u) {
        // rows of processes that were not asked for are zero, not whatever the tables held
        const int64_t blocks = (nK + 255) / 256;
        table_zero_rows_kernel<<<(unsigned) (blocks > 1184 ? 1184 : blocks), 256, 0, s>>>(
                nK, 15u & ~process_mask, out);
        rc = after_launch();
        if (rc) return rc;
    }

    if (flat) return table_flat_impl(process_mask, K, nK, xlow, min_points, p, out, s, opt);

    static const int heavy_first[4] = {NOA_DCS_PHOTONUCLEAR, NOA_DCS_PAIR_PRODUCTION,
                                       NOA_DCS_BREMSSTRAHLUNG, NOA_DCS_IONISATION};
    TablePlan all{};
    all.cells = ((uint32_t) min_points + 5u) / 6u;
    all.xlow = xlow;
    all.xhigh = opt.xhigh;
    all.second_power = opt.second_power;
    all.quadrature_only = opt.quadrature_only ? 1 : 0;
    for (int i = 0; i < 4; i++) {
        const int pr = heavy_first[i];
        if (!((process_mask >> pr) & 1u)) continue;
        const int R = rows_per_item(pr);
        all.process[all.n_slots] = pr;
        all.out_row[all.n_slots] = single_row ? 0 : pr;
        all.items[all.n_slots] = (uint32_t) ((nK + R - 1) / R);
        all.n_slots++;
    }

    // the launches of this build: one combined kernel, or one per process
    TablePlan plans[4];
    TableKernel kernels[4];
    unsigned grids[4];
    int n_launch = 0;
    if (all.n_slots == 1 || !table_launch_split(nK)) {
        plans[0] = all;
        const unsigned mask = all.n_slots == 1 ? (1u << all.process[0]) : 15u;
        kernels[0] = table_kernel_for(mask);
        n_launch = 1;
    } else {
        for (int i = 0; i < all.n_slots; i++) {
            TablePlan one = all;
            one.n_slots = 1;
            one.process[0] = all.process[i];
            one.out_row[0] = all.out_row[i];
            one.items[0] = all.items[i];
            plans[i] = one;
            kernels[i] = table_kernel_for(1u << all.process[i]);
        }
        n_launch = all.n_slots;
    }
    for (int i = 0; i < n_launch; i++) {
        uint64_t items = 0;
        for (int sl = 0; sl < plans[i].n_slots; sl++) items += plans[i].items[sl];
        grids[i] = (unsigned) items;
    }
    for (int i = 0; i < n_launch; i++) {
        rc = launch_table(kernels[i], grids[i], i > 0, s, K, nK, out, plans[i], p);
        if (rc) return rc;
    }
    return 0;
}

}  // namespace noa_b200

#include "coulomb_kernels.cuh"
#include "material_kernels.cuh"

using namespace noa_b200;

// staging object of noa_dcs_vmap_host_f64
struct noa_dcs_stager {
    int64_t chunk = 0;
    int32_t n_slots = 0;
    cudaStream_t *streams = nullptr;
    double **dK = nullptr, **dq = nullptr, **dout = nullptr;   // per slot; dout holds 4 * chunk
};

extern "C" {

int noa_dcs_abi_version(void) { return NOA_DCS_ABI_VERSION; }

const char *noa_dcs_strerror(int code) {
    switch (code) {
        case 0: return "success";
        case NOA_DCS_EINVAL: return "noa_dcs: invalid argument";
        case NOA_DCS_ERANGE: return "noa_dcs: size out of range";
        case NOA_DCS_ENODEV: return "noa_dcs: no CUDA device (this library has no CPU path)";
        case NOA_DCS_ENONCCL: return "noa_dcs: NCCL is not loaded in this process";
        case NOA_DCS_ELIBM: return "noa_dcs: the host libm differs from the one the kernels restate "
                                   "(results would not be bit-identical to the reference)";
    }
    if (code > 0) return cudaGetErrorString((cudaError_t) code);
    if (code <= NOA_DCS_ENCCL_BASE) return "noa_dcs: NCCL error (code = NOA_DCS_ENCCL_BASE - ncclResult_t)";
    return "noa_dcs: unknown error";
}

int noa_dcs_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        (void) cudaGetLastError();
        return 0;
    }
    return n;
}

int64_t noa_dcs_launch_count(void) { return g_launches.load(); }

int noa_dcs_div_recomputes(int64_t *count, int reset) {
    unsigned long long v = 0;
    cudaError_t e = cudaDeviceSynchronize();
    if (e == cudaSuccess) e = cudaMemcpyFromSymbol(&v, g_div_recomputes, sizeof(v));
    if (e == cudaSuccess && reset) {
        const unsigned long long zero = 0;
        e = cudaMemcpyToSymbol(g_div_recomputes, &zero, sizeof(zero));
    }
    if (e != cudaSuccess) return (int) e;
    if (count) *count = (int64_t) v;
    return 0;
}

int noa_dcs_vmap_f64(int process, const double *K, const double *q, double *result, int64_t n,
                     double A, double I, int32_t Z, double mass, void *stream) {
    if (process < 0 || process >= NOA_DCS_NPROCESS || n < 0) return NOA_DCS_EINVAL;
    if (n == 0) return 0;
    if (!K || !q || !result) return NOA_DCS_EINVAL;
    const Params p = make_params(A, I, Z, mass);
    return vmap_impl(process, K, q, result, n, p, (cudaStream_t) stream);
}

int noa_dcs_vmap_all_f64(const double *K, const double *q, double *result, int64_t n, double A,
                         double I, int32_t Z, double mass, void *stream) {
    if (n < 0) return NOA_DCS_EINVAL;
    if (n == 0) return 0;
    if (!K || !q || !result) return NOA_DCS_EINVAL;
    const Params p = make_params(A, I, Z, mass);
    return vmap_all_impl(K, q, result, n, p, (cudaStream_t) stream);
}

int noa_dcs_vmap_mixture_f64(unsigned process_mask, const double *K, const double *q,
                             double *result, int64_t n, int32_t n_elements, const double *A,
                             const double *I, const int32_t *Z, const double *w, double mass,
                             void *stream) {
    if (n < 0 || n_elements < 1 || n_elements > NOA_DCS_MAX_ELEMENTS) return NOA_DCS_EINVAL;
    if (process_mask == 0 || process_mask > 15u || !A || !I || !Z || !w) return NOA_DCS_EINVAL;
    if (n == 0) return 0;
    if (!K || !q || !result) return NOA_DCS_EINVAL;
    int blocks = 0;
    int rc = persistent_grid(vmap_mixture_element_kernel<true>, n, blocks);
    if (rc) return rc;
    for (int e = 0; e < n_elements; e++) {
        const Params p = make_params(A[e], I[e], Z[e], mass);
        if (e == 0)
            vmap_mixture_element_kernel<true><<<blocks, kThreads, 0, (cudaStream_t) stream>>>(
                    K, q, result, n, process_mask, w[e], p);
        else
            vmap_mixture_element_kernel<false><<<blocks, kThreads, 0, (cudaStream_t) stream>>>(
                    K, q, result, n, process_mask, w[e], p);
        rc = after_launch();
        if (rc) return rc;
    }
    return 0;
}

static TableOut local_out(double *del, double *cel, int64_t nK) {
    TableOut out{};
    out.n_peers = 1;
    out.n_total = nK;
    out.first_row = 0;
    out.row_stride = 1;
    out.del[0] = del;
    out.cel[0] = cel;
    return out;
}

int noa_dcs_table_f64(unsigned process_mask, const double *K, int64_t nK, double xlow,
                      int32_t min_points, double A, double I, int32_t Z, double mass, double *del,
                      double *cel, void *stream) {
    return table_impl(process_mask, false, K, nK, xlow, min_points, A, I, Z, mass,
                      local_out(del, cel, nK), stream);
}

int64_t noa_dcs_table_workspace_doubles(int64_t nK, int32_t min_points) {
    return table_workspace_doubles(nK, min_points);
}

int noa_dcs_table_ws_f64(unsigned process_mask, const double *K, int64_t nK, double xlow,
                         int32_t min_points, double A, double I, int32_t Z, double mass,
                         double *del, double *cel, double *workspace, int64_t workspace_doubles,
                         void *stream) {
    TableOptions opt;
    opt.workspace = workspace;
    opt.workspace_doubles = workspace_doubles;
    return table_impl(process_mask, false, K, nK, xlow, min_points, A, I, Z, mass,
                      local_out(del, cel, nK), stream, opt);
}

int noa_dcs_table_material_f64(unsigned process_mask, const double *K, int64_t nK, double xlow,
                               int32_t min_points, int32_t n_elements, const double *A,
                               const double *I, const int32_t *Z, const double *w, double mass,
                               double *scratch, double *table, double *workspace,
                               int64_t workspace_doubles, void *stream) {
    if (n_elements < 1 || n_elements > NOA_DCS_MAX_ELEMENTS || !A || !I || !Z || !w)
        return NOA_DCS_EINVAL;
    if (process_mask == 0 || process_mask > 15u || nK < 0 || min_points < 1) return NOA_DCS_EINVAL;
    if (nK == 0) return 0;
    if (!K || !scratch || !table) return NOA_DCS_EINVAL;
    const int64_t columns = 8 * nK;
    MixWeights m{};
    m.n_elements = n_elements;
    for (int el = 0; el < n_elements; el++) {
        m.w[el] = w[el];
        double *part = scratch + (int64_t) el * columns;
        // rows of processes outside the mask are zeroed by the build itself
        TableOptions opt;
        opt.workspace = workspace;
        opt.workspace_doubles = workspace_doubles;
        int rc = table_impl(process_mask, false, K, nK, xlow, min_points, A[el], I[el], Z[el], mass,
                            local_out(part, part + 4 * nK, nK), stream, opt);
        if (rc) return rc;
    }
    const int64_t blocks = (columns + 255) / 256;
    mix_tables_kernel<<<(unsigned) (blocks > 1184 ? 1184 : blocks), 256, 0,
                        (cudaStream_t) stream>>>(scratch, table, columns, m);
    return after_launch();
}

int noa_dcs_table_scatter_f64(unsigned process_mask, const double *K_local, int64_t n_local,
                              double xlow, int32_t min_points, double A, double I, int32_t Z,
                              double mass, int32_t n_peers, double *const *peer_del,
                              double *const *peer_cel, int64_t n_total, int64_t first_row,
                              int64_t row_stride, void *stream) {
    if (n_peers < 1 || n_peers > NOA_DCS_MAX_PEERS || !peer_del || !peer_cel)
        return NOA_DCS_EINVAL;
    if (n_total < 0 || first_row < 0 || row_stride < 1) return NOA_DCS_EINVAL;
    if (n_local > 0 && first_row + (n_local - 1) * row_stride >= n_total) return NOA_DCS_ERANGE;
    TableOut out{};
    out.n_peers = n_peers;
    out.n_total = n_total;
    out.first_row = first_row;
    out.row_stride = row_stride;
    for (int j = 0; j < n_peers; j++) {
        if (!peer_del[j] || !peer_cel[j]) return NOA_DCS_EINVAL;
        out.del[j] = peer_del[j];
        out.cel[j] = peer_cel[j];
    }
    return table_impl(process_mask, false, K_local, n_local, xlow, min_points, A, I, Z, mass, out,
                      stream);
}

int noa_dcs_table_exchange_f64(unsigned process_mask, const double *K_local, int64_t n_local,
                               double xlow, int32_t min_points, double A, double I, int32_t Z,
                               double mass, int32_t n_peers, int32_t my_peer,
                               double *const *peer_del, double *const *peer_cel,
                               uint32_t *const *peer_flags, double *multicast_del,
                               double *multicast_cel, uint32_t *multicast_flags, uint32_t *sync,
                               double *scratch,
                               int64_t scratch_doubles, uint32_t epoch, int64_t n_total,
                               int64_t first_row, int64_t row_stride, double timeout_seconds,
                               void *stream) {
    if (n_peers < 1 || n_peers > NOA_DCS_MAX_PEERS || my_peer < 0 || my_peer >= n_peers)
        return NOA_DCS_EINVAL;
    if (!peer_del || !peer_cel || !peer_flags || !sync) return NOA_DCS_EINVAL;
    if (n_total < 0 || first_row < 0 || row_stride < 1 || n_local < 0) return NOA_DCS_EINVAL;
    if (n_local > 0 && first_row + (n_local - 1) * row_stride >= n_total) return NOA_DCS_ERANGE;
    if (!(timeout_seconds > 0.)) timeout_seconds = NOA_DCS_DEFAULT_EXCHANGE_TIMEOUT_S;
    TableOut out{};
    out.n_peers = n_peers;
    out.me = my_peer;
    out.n_total = n_total;
    out.first_row = first_row;
    out.row_stride = row_stride;
    out.sync = sync;
    if (multicast_del && multicast_cel) {
        out.mc_del = multicast_del;
        out.mc_cel = multicast_cel;
        out.mc_flags = multicast_flags;
    }
    out.epoch = epoch;
    out.timeout_ns = (uint64_t) (timeout_seconds * 1e9);
    for (int j = 0; j < n_peers; j++) {
        if (!peer_del[j] || !peer_cel[j] || !peer_flags[j]) return NOA_DCS_EINVAL;
        out.del[j] = peer_del[j];
        out.cel[j] = peer_cel[j];
        out.flags[j] = peer_flags[j];
    }
    TableOptions opt;
    opt.workspace = scratch;
    opt.workspace_doubles = scratch_doubles;
    return table_impl(process_mask, false, K_local, n_local, xlow, min_points, A, I, Z, mass, out,
                      stream, opt);
}

int noa_dcs_vmap_integral_f64(int process, int integrand, const double *K, double *result,
                              int64_t n, double xlow, int32_t min_points, double A, double I,
                              int32_t Z, double mass, void *stream) {
    if (process < 0 || process >= NOA_DCS_NPROCESS || (integrand != 0 && integrand != 1))
        return NOA_DCS_EINVAL;
    if (n > 0 && !result) return NOA_DCS_EINVAL;
    return table_impl(1u << process, true, K, n, xlow, min_points, A, I, Z, mass,
                      local_out(integrand == 0 ? result : nullptr,
                                integrand == 1 ? result : nullptr, n),
                      stream);
}

int noa_dcs_vmap_integral_mode_f64(int process, int mode, const double *K, double *result,
                                   int64_t n, double xlow, double xhigh, int32_t min_points,
                                   double A, double I, int32_t Z, double mass, void *stream) {
    if (process < 0 || process >= NOA_DCS_NPROCESS || mode < 0 || mode > 2) return NOA_DCS_EINVAL;
    if (!(xlow > 0.) || !(xhigh > xlow)) return NOA_DCS_EINVAL;
    if (n > 0 && !result) return NOA_DCS_EINVAL;
    if (mode <= 1 && xhigh == 1.)      // dcs::recoil_integral proper, closed forms included
        return noa_dcs_vmap_integral_f64(process, mode, K, result, n, xlow, min_points, A, I, Z,
                                         mass, stream);
    TableOptions opt;
    opt.xhigh = xhigh;
    opt.second_power = (mode == 2) ? 3 : 2;
    opt.quadrature_only = true;
    return table_impl(1u << process, true, K, n, xlow, min_points, A, I, Z, mass,
                      local_out(mode == 0 ? result : nullptr, mode == 0 ? nullptr : result, n),
                      stream, opt);
}

int noa_dcs_material_assembly_f64(const double *K, int64_t nK, double cutoff, int32_t min_points,
                                  int32_t n_elements, const double *A, const double *I,
                                  const int32_t *Z, const double *w, double mass, double *elem,
                                  double *cs, double *cel, double *straggling, double *csf,
                                  double *cs_total, double *kt, int32_t *it, double *xt,
                                  double *workspace, int64_t workspace_doubles, void *stream) {
    if (n_elements < 1 || n_elements > NOA_DCS_MAX_ELEMENTS || !A || !I || !Z || !w)
        return NOA_DCS_EINVAL;
    if (nK < 0 || min_points < 1 || !(cutoff > 1E-06) || !(cutoff < 1.)) return NOA_DCS_EINVAL;
    if (nK == 0) return 0;
    if (!K || !elem || !cs || !cel || !straggling || !csf || !cs_total || !kt || !it || !xt)
        return NOA_DCS_EINVAL;
    cudaStream_t s = (cudaStream_t) stream;
    const int64_t n4 = 4 * nK;
    MixWeights mw{};
    MaterialParams mp{};
    mw.n_elements = mp.n_elements = n_elements;
    // step 1 (pumas.c:10786-10805): CSn / cel from the fused table kernels, the ionisation
    // straggling integral from their generalised form; the other stg rows are zero
    for (int el = 0; el < n_elements; el++) {
        mw.w[el] = mp.w[el] = w[el];
        mp.p[el] = make_params(A[el], I[el], Z[el], mass);
        double *e = elem + (int64_t) el * 3 * n4;
        TableOptions opt;
        opt.workspace = workspace;
        opt.workspace_doubles = workspace_doubles;
        int rc = table_impl(15u, false, K, nK, cutoff, min_points, A[el], I[el], Z[el], mass,
                            local_out(e, e + n4, nK), s, opt);
        if (rc) return rc;
        cudaError_t ce = cudaMemsetAsync(e + 2 * n4, 0, (size_t) n4 * sizeof(double), s);
        if (ce != cudaSuccess) return (int) ce;
        rc = noa_dcs_vmap_integral_mode_f64(NOA_DCS_IONISATION, 2, K, e + 2 * n4 + 3 * nK, nK,
                                            1E-06, cutoff, min_points, A[el], I[el], Z[el], mass,
                                            stream);
        if (rc) return rc;
    }
    const int64_t blocks = (nK + 127) / 128;
    material_mix_kernel<<<(unsigned) (blocks > 1184 ? 1184 : blocks), 128, 0, s>>>(
            elem, nK, n_elements, mw, cs, cel, straggling, csf, cs_total);
    int rc = after_launch();
    if (rc) return rc;
    material_threshold_kernel<<<1, 256, 0, s>>>(K, nK, cs_total, kt, it);
    rc = after_launch();
    if (rc) return rc;
    const int64_t xt_blocks = ((int64_t) n_elements * n4 + 127) / 128;
    material_xt_kernel<<<(unsigned) (xt_blocks > 4736 ? 4736 : xt_blocks), 128, 0, s>>>(
            K, nK, cutoff, it, mp, xt);
    return after_launch();
}

// ncclAllGather through the NCCL the process already has (torch's bundled one, or the system's):
// resolved at first use so this library carries no link-time NCCL dependency.
int noa_dcs_allgather_f64(double *table, int64_t count_per_rank, int32_t rank, void *nccl_comm,
                          void *stream) {
    if (!table || count_per_rank < 0 || rank < 0 || !nccl_comm) return NOA_DCS_EINVAL;
    if (count_per_rank == 0) return 0;
    typedef int (*AllGatherFn)(const void *, void *, size_t, int, void *, cudaStream_t);
    static AllGatherFn fn = [] {
        void *sym = dlsym(RTLD_DEFAULT, "ncclAllGather");
        if (!sym) {
            void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
            if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
            if (h) sym = dlsym(h, "ncclAllGather");
        }
        return (AllGatherFn) sym;
    }();
    if (!fn) return NOA_DCS_ENONCCL;
    const int nccl_float64 = 8;   // ncclDouble (nccl.h: ncclFloat64 = 8)
    // in place: rank r's slice sits at table + r * count_per_rank
    const int rc = fn(table + (int64_t) rank * count_per_rank, table, (size_t) count_per_rank,
                      nccl_float64, nccl_comm, (cudaStream_t) stream);
    if (rc != 0) return NOA_DCS_ENCCL_BASE - rc;
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return 0;
}

int noa_dcs_stager_create(noa_dcs_stager **out, int64_t chunk_pairs, int32_t n_slots) {
    if (!out || chunk_pairs < 1 || n_slots < 1 || n_slots > 16) return NOA_DCS_EINVAL;
    DeviceInfo info;
    int rc = device_info(info);
    if (rc) return rc;
    noa_dcs_stager *st = new (std::nothrow) noa_dcs_stager();
    if (!st) return (int) cudaErrorMemoryAllocation;
    st->chunk = chunk_pairs;
    st->n_slots = n_slots;
    st->streams = new cudaStream_t[n_slots]();
    st->dK = new double *[n_slots]();
    st->dq = new double *[n_slots]();
    st->dout = new double *[n_slots]();
    cudaError_t e = cudaSuccess;
    for (int s = 0; s < n_slots && e == cudaSuccess; s++) {
        e = cudaStreamCreateWithFlags(&st->streams[s], cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaMalloc(&st->dK[s], chunk_pairs * sizeof(double));
        if (e == cudaSuccess) e = cudaMalloc(&st->dq[s], chunk_pairs * sizeof(double));
        if (e == cudaSuccess) e = cudaMalloc(&st->dout[s], 4 * chunk_pairs * sizeof(double));
    }
    if (e != cudaSuccess) {
        noa_dcs_stager_destroy(st);
        return (int) e;
    }
    *out = st;
    return 0;
}

int noa_dcs_stager_destroy(noa_dcs_stager *st) {
    if (!st) return 0;
    for (int s = 0; s < st->n_slots; s++) {
        if (st->streams && st->streams[s]) {
            cudaStreamSynchronize(st->streams[s]);
            cudaStreamDestroy(st->streams[s]);
        }
        if (st->dK && st->dK[s]) cudaFree(st->dK[s]);
        if (st->dq && st->dq[s]) cudaFree(st->dq[s]);
        if (st->dout && st->dout[s]) cudaFree(st->dout[s]);
    }
    delete[] st->streams;
    delete[] st->dK;
    delete[] st->dq;
    delete[] st->dout;
    delete st;
    return 0;
}

int noa_dcs_vmap_host_f64(noa_dcs_stager *st, int process, const double *h_K, const double *h_q,
                          double *h_result, int64_t n, double A, double I, int32_t Z, double mass) {
    if (!st || process < 0 || process > NOA_DCS_NPROCESS || n < 0) return NOA_DCS_EINVAL;
    if (n == 0) return 0;
    if (!h_K || !h_q || !h_result) return NOA_DCS_EINVAL;
    const Params p = make_params(A, I, Z, mass);
    int rc = 0;
    int64_t c = 0;
    for (int64_t off = 0; off < n && rc == 0; off += st->chunk, c++) {
        const int s = (int) (c % st->n_slots);
        const int64_t m = (n - off < st->chunk) ? (n - off) : st->chunk;
        cudaStream_t stream = st->streams[s];
        cudaError_t e = cudaMemcpyAsync(st->dK[s], h_K + off, m * sizeof(double),
                                        cudaMemcpyHostToDevice, stream);
        if (e == cudaSuccess)
            e = cudaMemcpyAsync(st->dq[s], h_q + off, m * sizeof(double), cudaMemcpyHostToDevice,
                                stream);
        if (e != cudaSuccess) {
            rc = (int) e;
            break;
        }
        if (process == NOA_DCS_NPROCESS) {
            rc = vmap_all_impl(st->dK[s], st->dq[s], st->dout[s], m, p, stream);
            for (int pr = 0; pr < NOA_DCS_NPROCESS && rc == 0; pr++)
                rc = (int) cudaMemcpyAsync(h_result + (int64_t) pr * n + off,
                                           st->dout[s] + (int64_t) pr * m, m * sizeof(double),
                                           cudaMemcpyDeviceToHost, stream);
        } else {
            rc = vmap_impl(process, st->dK[s], st->dq[s], st->dout[s], m, p, stream);
            if (rc == 0)
                rc = (int) cudaMemcpyAsync(h_result + off, st->dout[s], m * sizeof(double),
                                           cudaMemcpyDeviceToHost, stream);
        }
    }
    for (int s = 0; s < st->n_slots; s++) {
        cudaError_t e = cudaStreamSynchronize(st->streams[s]);
        if (rc == 0 && e != cudaSuccess) rc = (int) e;
    }
    return rc;
}

// Device-side address of page-locked host memory the device can reach (cudaHostAlloc /
// cudaHostRegister); nullptr for anything else.  For registered memory the device address need not
// equal the host address, so the kernel is always given this one.
static void *device_address_of_host(const void *ptr) {
    cudaPointerAttributes attr{};
    if (cudaPointerGetAttributes(&attr, ptr) != cudaSuccess) {
        (void) cudaGetLastError();
        return nullptr;
    }
    if (attr.type != cudaMemoryTypeHost) return nullptr;
    return attr.devicePointer;
}

int noa_dcs_vmap_pinned_f64(int process, const double *h_K, const double *h_q, double *h_result,
                            int64_t n, double A, double I, int32_t Z, double mass, void *stream) {
    if (process < 0 || process >= NOA_DCS_NPROCESS || n < 0) return NOA_DCS_EINVAL;
    if (n == 0) return 0;
    if (!h_K || !h_q || !h_result) return NOA_DCS_EINVAL;
    const double *d_K = (const double *) device_address_of_host(h_K);
    const double *d_q = (const double *) device_address_of_host(h_q);
    double *d_result = (double *) device_address_of_host(h_result);
    if (!d_K || !d_q || !d_result) return NOA_DCS_EINVAL;
    // The element-wise kernels run unchanged on the mapped host addresses: their coalesced loads
    // and stores cross PCIe themselves and overlap with the arithmetic of the other resident
    // warps.  Measured alternatives (TMA bulk-copy ring with mbarriers; per-thread cp.async
    // prefetch into shared memory) were slower or equal -- profiles/r01_host_path_variants.md.
    const Params p = make_params(A, I, Z, mass);
    return vmap_impl(process, d_K, d_q, d_result, n, p, (cudaStream_t) stream);
}


// ---- host libm self-check ---------------------------------------------------------------------
// Bit-exactness against the reference's CPU path rests on the host libm being the one the device
// routines restate (glibc >= 2.28, FMA variant): make_params() evaluates pow / log / exp on the
// host exactly as the reference does, and the kernels' exp / log / log10 reproduce that libm.  On a
// different libm the results would silently be "a few ulp" off, which the pair-production
// integrand amplifies to ~1e-10.  This check compares the host's exp / log / log10 with the host
// build of glibm.cuh (same tables as the device) on 3 x 1024 arguments and pow with known answers.
static const glibm::Tables h_selfcheck_tables = {GLIBM_EXP_TABLE_INIT, GLIBM_LOG_TABLE_INIT};

int noa_dcs_selfcheck(int64_t *mismatches) {
    const glibm::Tab T = glibm::make_host_tab(&h_selfcheck_tables);
    int64_t bad = 0;
    uint64_t state = 0x9E3779B97F4A7C15ULL;
    auto next = [&state]() {          // splitmix64
        uint64_t z = (state += 0x9E3779B97F4A7C15ULL);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
        return z ^ (z >> 31);
    };
    auto same = [](double a, double b) { return glibm::to_bits(a) == glibm::to_bits(b); };
    for (int i = 0; i < 1024; i++) {
        const double u = (double) (next() >> 11) * 0x1p-53;            // [0, 1)
        const double v = (double) (next() >> 11) * 0x1p-53;
        const double xe = (u - 0.5) * 80.;                             // exp: |x| < 40
        const double xl = std::ldexp(0.5 + 0.5 * v, (int) (u * 80.) - 40);   // log: 2^-41 .. 2^40
        const double xn = 0.93 + 0.14 * v;                             // log near one
        if (!same(std::exp(xe), glibm::exp_any(xe, T))) bad++;
        if (!same(std::log(xl), glibm::log_any(xl, T))) bad++;
        if (!same(std::log(xn), glibm::log_any(xn, T))) bad++;
        if (!same(std::log10(xl), glibm::log10(xl, T))) bad++;
    }
    static const struct {
        double a, b;
        uint64_t bits;
    } pow_kat[] = {
        {11.0, 1. / 3., 0x4001cab612df9a45ULL},     {11.0, -1. / 3., 0x3fdcc6f8f0d0ed76ULL},
        {11.0, -2. / 3., 0x3fc9e108d5a254c3ULL},    {22.0, 0.27, 0x40026e48b2ce2eadULL},
        {82.0, 1. / 3., 0x401160bfc12dd090ULL},     {207.2, 0.27, 0x4010e266ff9d65c5ULL},
        {8.0, -2. / 3., 0x3fd0000000000000ULL},     {15.999, 0.27, 0x4000e9790b7a4a0aULL},
        {26.0, -1. / 3., 0x3fd59a78b28f888bULL},    {55.845, 0.27, 0x4007b39532d0b74eULL},
    };
    for (const auto &c : pow_kat) {
        volatile double a = c.a, b = c.b;       // keep the call (no constant folding)
        if (glibm::to_bits(std::pow(a, b)) != c.bits) bad++;
    }
    if (mismatches) *mismatches = bad;
    return bad == 0 ? 0 : NOA_DCS_ELIBM;
}

int noa_dcs_launch_info(int process, int32_t *blocks, int32_t *threads, int32_t *sm_count) {
    DeviceInfo info;
    int rc = device_info(info);
    if (rc) return rc;
    int b = 0;
    switch (process) {
        case 0: rc = persistent_grid(vmap_kernel<0, NOA_STREAM_VEC>, INT64_MAX / 2, b); break;
        case 1: rc = persistent_grid(vmap_kernel<1, 1>, INT64_MAX / 2, b); break;
        case 2: rc = persistent_grid(vmap_kernel<2, 1>, INT64_MAX / 2, b); break;
        case 3: rc = persistent_grid(vmap_kernel<3, NOA_STREAM_VEC>, INT64_MAX / 2, b); break;
        case 4: rc = persistent_grid(vmap_all_kernel, INT64_MAX / 2, b); break;
        default: return NOA_DCS_EINVAL;
    }
    if (rc) return rc;
    if (blocks) *blocks = b;
    if (threads) *threads = kThreads;
    if (sm_count) *sm_count = info.sm_count;
    return 0;
}

}  // extern "C"