// B200 (sm_100a) kernels of the muon DCS hot path and the C ABI that exposes them
// (include/noa_dcs_b200.h).  Torch-free on purpose: this file compiles in seconds and is the
// whole product below the LibTorch boundary.
//
// Kernels
//   vmap_kernel<P, VEC>     element-wise DCS of one process, one (K, q) pair per thread and
//                           iteration, persistent grid-stride blocks, 128-bit loads/stores
//   vmap_pair_lanes_kernel  pair production with one Gauss-Legendre node per lane (8 lanes per
//                           pair, shuffle gather, serial-order sum) -- kept for the measured
//                           comparison in DESIGN.md
//   vmap_all_kernel         the four processes of one pair in one pass (16 B in, 32 B out)
//   vmap_mixture_element_kernel   one element's term of sum_e w_e * DCS_e (water = H + O)
//   table_kernel            one CTA per (process, energy) row of the DEL/CEL tables: nodes of the
//                           composite 6-point rule across threads, node terms staged in shared
//                           memory and accumulated in the reference's serial order
//   fp64_probe_kernel       dependent-chain-free DFMA loop (roofline denominator)
//
// Numerics: FP64 throughout, every operation IEEE (see dcs_math.cuh, glibm.cuh); this file must be
// compiled with -fmad=false.  The exp/log tables (4 KB) are staged into shared memory per CTA.
#include <cuda_runtime.h>

#include <atomic>
#include <cstdio>
#include <mutex>
#include <new>

#include "../../include/noa_dcs_b200.h"
#include "dcs_math.cuh"
#include "dcs_params.hh"

namespace noa_b200 {

__device__ const glibm::Tables g_tables = {GLIBM_EXP_TABLE_INIT, GLIBM_LOG_TABLE_INIT};

#ifndef NOA_THREADS
#define NOA_THREADS 256
#endif
constexpr int kThreads = NOA_THREADS;

// Minimum resident CTAs per SM requested from ptxas (register cap = 65536 / (256 * N)).
// The kernels are bound by issue slots and fixed-latency dependencies, not by the FP64 pipe alone
// (profiles/), so occupancy matters; values chosen by measurement (tools/bounds_sweep.py,
// profiles/r01_launch_bounds_sweep_s4.txt; all within ~1.5 % of each other except where noted):
// pair 5 (48 registers), photonuclear 3 (80; 4 loses 4 %), streaming 4, fused four-process
// kernels 2 (+2.7 % over 3), table kernel 4 (3 and 5 lose 4-6 %).
#ifndef NOA_MINB_PAIR
#define NOA_MINB_PAIR 5
#endif
#ifndef NOA_MINB_PHOTO
#define NOA_MINB_PHOTO 3
#endif
#ifndef NOA_MINB_STREAM
#define NOA_MINB_STREAM 4
#endif
#ifndef NOA_MINB_ALL
#define NOA_MINB_ALL 2
#endif
#ifndef NOA_MINB_TABLE
#define NOA_MINB_TABLE 4
#endif
template <int PROCESS>
struct MinBlocks {
    static constexpr int value = (PROCESS == 1) ? NOA_MINB_PAIR
                                 : (PROCESS == 2) ? NOA_MINB_PHOTO : NOA_MINB_STREAM;
};

// 4 KB global -> shared, coalesced 128-bit copies; returns the shared-window addresses the
// lookups use
__device__ __forceinline__ glibm::Tab stage_tables(glibm::Tables &dst) {
    const uint4 *src = reinterpret_cast<const uint4 *>(&g_tables);
    uint4 *d = reinterpret_cast<uint4 *>(&dst);
    for (int i = threadIdx.x; i < (int) (sizeof(glibm::Tables) / sizeof(uint4)); i += blockDim.x)
        d[i] = src[i];
    __syncthreads();
    return glibm::make_smem_tab(dst);
}

// Evaluation of one DCS value with the folded special-case tests of folded_ops.cuh: the FoldedOps
// pass, and -- only if one of its divisions left nvcc's fast-path domain (zero / subnormal-range
// numerator, non-finite or out-of-range quotient) or an exp / log argument left the common case --
// the same value again with the plain operations, out of line.  g_div_recomputes counts those
// second passes (diagnostics: noa_dcs_div_recomputes).  NOA_FOLDED_OPS=0 builds the kernels with
// the plain operations only (measurement).
#ifndef NOA_FOLDED_OPS
#define NOA_FOLDED_OPS 1
#endif
__device__ unsigned long long g_div_recomputes = 0;

// Tables plus the refined reciprocals of the launch-invariant denominators (folded_ops.cuh: DenSlot).
struct StagedShared {
    glibm::Tables tables;
    double dens[kDenSlots];
};

__device__ __forceinline__ double den_slot_value(int slot, const Params &p) {
    switch (slot) {
        case kDenLambda2: return 0.06527;
        case kDenQ004: return 0.04;
        case kDenLogQ0L: return p.n_logq0l;
        case kDenR2: return p.p_r2;
        case kDenA: return p.A;
        case kDenMass: return p.mass;
        case kDenMe: return kElectronMass;
        default: return p.i_m2;
    }
}

__device__ __forceinline__ glibm::Tab stage_all(StagedShared &dst, const Params &p) {
    if (threadIdx.x < kDenSlots)
        dst.dens[threadIdx.x] =
                FoldedOps<true>::staged_reciprocal(den_slot_value(threadIdx.x, p));
    glibm::Tab T = stage_tables(dst.tables);
    T.aux_smem = T.exp_smem + (uint32_t) offsetof(StagedShared, dens);
    return T;
}

template <int PROCESS>
__device__ __noinline__ double dcs_eval_plain(double K, double q, const Params &p,
                                             const glibm::Tab &T) {
    atomicAdd(&g_div_recomputes, 1ULL);
    return dcs_eval<PROCESS>(K, q, p, T);
}

// STAGED = T comes from stage_all() for this very `p`
template <int PROCESS, bool STAGED>
__device__ __forceinline__ double dcs_value(double K, double q, const Params &p,
                                            const glibm::Tab &T) {
#if NOA_FOLDED_OPS
    FoldedOps<STAGED> dv;
    dv.dens = T.aux_smem;
    double v = dcs_eval<PROCESS>(K, q, p, T, dv);
    if (!dv.ok()) v = dcs_eval_plain<PROCESS>(K, q, p, T);
    return v;
#else
    return dcs_eval<PROCESS>(K, q, p, T);
#endif
}

// ------------------------------------------------------------------------------------------
// element-wise, one process
// ------------------------------------------------------------------------------------------
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
    if (VEC == 2) {
        const int64_t n2 = n >> 1;
        const double2 *K2 = reinterpret_cast<const double2 *>(K);
        const double2 *q2 = reinterpret_cast<const double2 *>(q);
        double2 *o2 = reinterpret_cast<double2 *>(out);
        for (int64_t i = tid; i < n2; i += stride) {
            const double2 k = K2[i];
            const double2 r = q2[i];
            double2 o;
            o.x = dcs_value<PROCESS, true>(k.x, r.x, p, T);
            o.y = dcs_value<PROCESS, true>(k.y, r.y, p, T);
            o2[i] = o;
        }
        if (tid == 0 && (n & 1)) out[n - 1] = dcs_value<PROCESS, true>(K[n - 1], q[n - 1], p, T);
    } else {
        for (int64_t i = tid; i < n; i += stride) {
            out[i] = dcs_value<PROCESS, true>(K[i], q[i], p, T);
        }
    }
}

// pair production, one quadrature node per lane: lanes 8g..8g+7 of a warp share pair g.
__global__ void __launch_bounds__(kThreads)
vmap_pair_lanes_kernel(const double *__restrict__ K, const double *__restrict__ q,
                       double *__restrict__ out, int64_t n, const __grid_constant__ Params p) {
    __shared__ glibm::Tables s_tables;
    const glibm::Tab T = stage_tables(s_tables);
    const int lane = threadIdx.x & 31;
    const int node = lane & 7;
    const int64_t groups = ((int64_t) gridDim.x * blockDim.x) >> 3;
    const int64_t g0 = ((int64_t) blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    const int64_t rounds = (n + groups - 1) / groups;     // uniform trip count: shuffles are warp-wide
    for (int64_t it = 0; it < rounds; it++) {
        const int64_t i = g0 + it * groups;
        const bool live = i < n;
        const double k = live ? K[i] : 1.0;
        const double r = live ? q[i] : 1.0;
        PairKinematics kin;
        const bool inside = live && pair_setup(k, r, p, T, kin);
        double term = 0.;
        if (inside) term = pair_node(c_gl8_x[node], r, kin, p, T) * c_gl8_w[node];
        // gather the 8 node terms of the group and add them in node order (numerics.hh:84-87)
        double acc = 0.;
#pragma unroll
        for (int j = 0; j < 8; j++) acc += __shfl_sync(0xffffffffu, term, (lane & 24) | j);
        if (live && node == 0) out[i] = inside ? pair_finish(k, r, acc, kin, p, T) : 0.;
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

template <bool STAGED>
__device__ __forceinline__ double dcs_dispatch(int process, double k, double r, const Params &p,
                                               const glibm::Tab &T) {
    switch (process) {
        case 0: return dcs_value<0, STAGED>(k, r, p, T);
        case 1: return dcs_value<1, STAGED>(k, r, p, T);
        case 2: return dcs_value<2, STAGED>(k, r, p, T);
        default: return dcs_value<3, STAGED>(k, r, p, T);
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

// ------------------------------------------------------------------------------------------
// energy-loss tables: dcs::vmap_integral(dcs::recoil_integral(f, del|cel_integrand))
// (src/noa/pms/dcs.hh:89-130, 955-1001; src/noa/utils/numerics.hh:72-108)
// ------------------------------------------------------------------------------------------
constexpr int kTableChunk = 1536;   // node terms staged per pass: 2 x 12 KB of shared memory

// Where the finished rows go: `n_peers` destination tables (this GPU's own and, in the multi-GPU
// build, every peer's, mapped over NVLink), each [4][n_total]; local row r is global row
// first_row + r * row_stride.
struct TableOut {
    int32_t n_peers;
    int32_t me;               // index of this GPU among the peers (exchange form only)
    int64_t n_total;
    int64_t first_row;
    int64_t row_stride;
    double *del[NOA_DCS_MAX_PEERS];
    double *cel[NOA_DCS_MAX_PEERS];
    // exchange form (noa_dcs_table_exchange_f64): flags[j] = peer j's array of n_peers epoch words,
    // done = this GPU's CTA counter {count, timeouts}; flags[0] == nullptr otherwise
    uint32_t *flags[NOA_DCS_MAX_PEERS];
    uint32_t *done;
    uint32_t epoch;
    int32_t fence_mode;       // see g_exchange_fence_mode
};

// Tail of the exchange form.  Every CTA has stored its values (local + peers) and fenced them at
// system scope; the last CTA of the grid to get here publishes this GPU's epoch into every peer's
// flag array (release, system scope) and then waits until every peer's epoch has arrived in its own
// -- so when the kernel retires, the complete table is in this GPU's memory.  No host round trip,
// no separate barrier kernel.  A peer that never shows up is reported, not waited for forever.
__device__ __forceinline__ void table_exchange_tail(const TableOut &out) {
    __syncthreads();
    if (threadIdx.x != 0) return;
    if (out.fence_mode == 1 || out.fence_mode == 3) __threadfence_system(); else __threadfence();
    const uint32_t arrived = atomicAdd(out.done, 1u);
    if (arrived != gridDim.x - 1) return;
    out.done[0] = 0;                       // re-armed for the next launch on this stream
    out.done[2] = 0;                       // (row queue of the persistent form)
    __threadfence_system();
    for (int j = 0; j < out.n_peers; j++)
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(out.flags[j] + out.me),
                     "r"(out.epoch)
                     : "memory");
    const long long t0 = clock64();
    for (int j = 0; j < out.n_peers; j++) {
        const uint32_t *slot = out.flags[out.me] + j;
        for (;;) {
            uint32_t seen;
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(slot) : "memory");
            if ((int32_t) (seen - out.epoch) >= 0) break;
            if (clock64() - t0 > (1LL << 33)) {      // ~4 s at 1.965 GHz: a peer is missing
                atomicAdd(out.done + 1, 1u);
                return;
            }
            __nanosleep(100);
        }
    }
}

struct TablePlan {
    int32_t n_slots;          // processes to build
    int32_t process[4];       // heaviest first, so the tail of the grid is cheap rows
    int32_t out_row[4];       // output row of process p (p for full tables, 0 for a single column)
    uint32_t cells;           // ceil(min_points / 6)
    double xlow;
};

// One (process, energy) row: item b of the heavy-first ordering.  Called by every thread of the
// CTA; s_del / s_cel are free to overwrite on entry.  Stores are NOT fenced here.
// Out of line on purpose: as its own function the row body is register-allocated on its own
// (216 B of spills instead of 440 B when inlined into the persistent loop) -- measured 7 % faster
// on the 180-point build and 3 % on the persistent exchange kernel (profiles/).
#ifndef NOA_TABLE_ROW_INLINE
#define NOA_TABLE_ROW_INLINE __noinline__
#endif
__device__ NOA_TABLE_ROW_INLINE void table_row(int64_t b, const double *__restrict__ K, int64_t nK,
                                          const TableOut &out, const TablePlan &plan,
                                          const Params &p, const glibm::Tab &T, double *s_del,
                                          double *s_cel, uint32_t *queue = nullptr,
                                          uint32_t *s_next = nullptr) {
    // Persistent form only (queue != nullptr): lane 64, idle while lanes 0 / 32 add the node terms
    // up, pops the CTA's next row at that point -- late enough that a heavy row in flight never
    // sits on a row another CTA could have started (a pop at row start cost 17 % at 8 GPUs),
    // early enough that the atomic's round trip is hidden behind the summation.
    const int process = plan.process[b / nK];
    const int64_t row = nK - 1 - (b % nK);
    const double k = K[row];
    const int tid = threadIdx.x;

    // destination index and the lane that owns each integrand (lane 0: DEL, lane 32: CEL)
    const int64_t at = (int64_t) plan.out_row[process] * out.n_total + out.first_row +
                       row * out.row_stride;
    double *const *dst = (tid == 0) ? out.del : out.cel;
    const bool writer = (tid == 0 || tid == 32) && dst[0] != nullptr;

    if (process == 3 && k <= p.i_kthr) {          // dcs.hh:963-966, 987-990
        if (writer) {
            const double v = ionisation_closed_form(k, plan.xlow, tid == 0 ? 0 : 1, p, T);
            for (int j = 0; j < out.n_peers; j++) dst[j][at] = v;
        }
        if (queue != nullptr && tid == 64) *s_next = atomicAdd(queue, 1u);
        return;
    }

    const double lb = glibm::log(k * plan.xlow, T);
    const double ub = glibm::log(k, T);
    const double h = (ub - lb) / plan.cells;
    const uint32_t total = plan.cells * 6u;
    double acc = 0.;
    for (uint32_t base = 0; base < total; base += kTableChunk) {
        const uint32_t count = min((uint32_t) kTableChunk, total - base);
        for (uint32_t i = base + tid; i < base + count; i += kThreads) {
            const uint32_t j = i % 6u;
            const double x = lb + h * ((i / 6u) + c_gl6_x[j]);
            const double r = glibm::exp(x, T);
            const double f = dcs_dispatch<true>(process, k, r, p, T);
            const double w = c_gl6_w[j];
            const double fr = f * r;
            s_del[i - base] = fr * h * w;           // del_integrand, dcs.hh:107-109
            s_cel[i - base] = fr * r * h * w;       // cel_integrand, dcs.hh:111-113
        }
        __syncthreads();
        // res += term, strictly in node order (numerics.hh:84-87): one lane per integrand
        if (tid == 0) {
            for (uint32_t i = 0; i < count; i++) acc += s_del[i];
        } else if (tid == 32) {
            for (uint32_t i = 0; i < count; i++) acc += s_cel[i];
        } else if (tid == 64 && queue != nullptr && base + kTableChunk >= total) {
            *s_next = atomicAdd(queue, 1u);
        }
        __syncthreads();
    }
    if (writer) {
        const double v = acc / (k + p.mass);
        // one store per destination: the local table and, over NVLink, each peer's copy
        for (int j = 0; j < out.n_peers; j++) dst[j][at] = v;
    }
}

// One CTA per row (local tables, and the scatter / per-row exchange forms).
__global__ void __launch_bounds__(kThreads, NOA_MINB_TABLE)
table_kernel(const double *__restrict__ K, int64_t nK, const __grid_constant__ TableOut out,
             const __grid_constant__ TablePlan plan, const __grid_constant__ Params p) {
    __shared__ StagedShared s_staged;
    __shared__ double s_del[kTableChunk];
    __shared__ double s_cel[kTableChunk];
    const glibm::Tab T = stage_all(s_staged, p);
    table_row(blockIdx.x, K, nK, out, plan, p, T, s_del, s_cel);
    // the writer lanes fence their own remote stores (fence_mode 0)
    if (out.n_peers > 1 && out.fence_mode == 0 && (threadIdx.x == 0 || threadIdx.x == 32))
        __threadfence_system();
    if (out.flags[0] != nullptr && out.fence_mode != 4) table_exchange_tail(out);
}

// Persistent form of the exchange: a grid that fills the GPU once; every CTA pulls rows from a
// device-side queue (out.done[2]) in the same heavy-first order, streams the finished values to
// all peers as it goes, and pays for ONE system-scope fence at the very end -- a fence per row
// costs ~7 us of CTA residency each (measured, profiles/), i.e. 8 % of the whole build.
__global__ void __launch_bounds__(kThreads, NOA_MINB_TABLE)
table_exchange_kernel(const double *__restrict__ K, int64_t nK, const __grid_constant__ TableOut out,
                      const __grid_constant__ TablePlan plan, const __grid_constant__ Params p) {
    __shared__ StagedShared s_staged;
    __shared__ double s_del[kTableChunk];
    __shared__ double s_cel[kTableChunk];
    __shared__ uint32_t s_item[2];
    const glibm::Tab T = stage_all(s_staged, p);
    const uint32_t total = (uint32_t) (nK * plan.n_slots);
    if (threadIdx.x == 64) s_item[0] = atomicAdd(out.done + 2, 1u);
    for (int cur = 0;; cur ^= 1) {
        __syncthreads();                       // s_item[cur] written; node buffers free again
        const uint32_t b = s_item[cur];
        if (b >= total) break;
        table_row(b, K, nK, out, plan, p, T, s_del, s_cel, out.done + 2, &s_item[cur ^ 1]);
    }
    table_exchange_tail(out);
}

// Material tables: out[c] = sum_e parts[e][c] * w[e], e in composition order, starting from 0
// (the per-element mixing of src/noa/3rdparty/_pumas/pumas.c:8054-8078).  `columns` = 8 n_K.
struct MixWeights {
    int32_t n_elements;
    double w[NOA_DCS_MAX_ELEMENTS];
};

__global__ void mix_tables_kernel(const double *__restrict__ parts, double *__restrict__ out,
                                  int64_t columns, const __grid_constant__ MixWeights m) {
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    for (int64_t c = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; c < columns; c += stride) {
        double acc = 0.;
        for (int e = 0; e < m.n_elements; e++) acc += parts[(int64_t) e * columns + c] * m.w[e];
        out[c] = acc;
    }
}

// A rank with no rows of its own still has to take part in the exchange.
__global__ void table_signal_kernel(const __grid_constant__ TableOut out) {
    table_exchange_tail(out);
}

// ------------------------------------------------------------------------------------------
// FP64 peak probe: 16 independent chains per thread.
//   mode 0  DFMA a = a * const + const     (1 register-pair source)  -> the roofline denominator
//   mode 1  DFMA a = a * b + c             (3 distinct register-pair sources)
//   mode 2  DFMA a = a * b + const         (2 register-pair sources)
//   mode 3  DADD a = a + b,  mode 4  DMUL a = a * b
// Modes 1-4 exist to measure how register-file bandwidth limits real instruction mixes.
// ------------------------------------------------------------------------------------------
template <int MODE>
__global__ void fp64_probe_kernel(int64_t iters, double *sink) {
    double a[16], b[4], c[4];
#pragma unroll
    for (int j = 0; j < 16; j++) a[j] = 1.0 + 1e-3 * (threadIdx.x + j);
#pragma unroll
    for (int j = 0; j < 4; j++) {
        b[j] = 0.999999 + 1e-9 * (threadIdx.x + j);
        c[j] = 1e-6 + 1e-12 * (threadIdx.x + j);
    }
    const double kb = 0.999999, kc = 1e-6;
    for (int64_t it = 0; it < iters; it++) {
#pragma unroll
        for (int j = 0; j < 16; j++) {
            if (MODE == 0) a[j] = fma(a[j], kb, kc);
            if (MODE == 1) a[j] = fma(a[j], b[j & 3], c[(j >> 2) & 3]);
            if (MODE == 2) a[j] = fma(a[j], b[j & 3], kc);
            if (MODE == 3) a[j] = __dadd_rn(a[j], c[j & 3]);
            if (MODE == 4) a[j] = __dmul_rn(a[j], b[j & 3]);
        }
    }
    double s = 0.;
#pragma unroll
    for (int j = 0; j < 16; j++) s += a[j];
    if (s == 123456.789) sink[0] = s;   // never true; keeps the chains alive
}

// Latency probe: CHAINS independent dependent-DFMA chains per thread (mode 0 is CHAINS = 16).
template <int CHAINS>
__global__ void fp64_chain_probe_kernel(int64_t iters, double *sink) {
    double a[CHAINS];
#pragma unroll
    for (int j = 0; j < CHAINS; j++) a[j] = 1.0 + 1e-3 * (threadIdx.x + j);
    const double kb = 0.999999, kc = 1e-6;
    for (int64_t it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 16 / CHAINS; r++)
#pragma unroll
            for (int j = 0; j < CHAINS; j++) a[j] = fma(a[j], kb, kc);
    }
    double s = 0.;
#pragma unroll
    for (int j = 0; j < CHAINS; j++) s += a[j];
    if (s == 123456.789) sink[0] = s;
}

// Constant-load probe: dependent DFMA chains (CHAINS per thread) whose multiplier is re-read from
// the constant bank before every DFMA (ld.const through the LDC / IDC path, as the polynomial
// coefficients of glibm are), to see what a constant load in the dependency chain costs.
__constant__ double c_probe_consts[64] = {0.999999, 0.999998, 0.999997, 0.999996};
template <int CHAINS, int UNIFORM>
__global__ void fp64_ldc_probe_kernel(int64_t iters, double *sink) {
    double a[CHAINS];
#pragma unroll
    for (int j = 0; j < CHAINS; j++) a[j] = 1.0 + 1e-3 * (threadIdx.x + j);
    const double kc = 1e-6;
    // UNIFORM = 1: the index follows the loop counter (LDCU, uniform datapath);
    // UNIFORM = 0: it comes from the chain's own value, like a table lookup (LDC with a per-thread
    // address)
    for (int64_t it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 16 / CHAINS; r++)
#pragma unroll
            for (int j = 0; j < CHAINS; j++) {
                const uint32_t idx = UNIFORM ? (uint32_t) (it + j + r)
                                             : (uint32_t) __double2loint(a[j]);
                const double kb = c_probe_consts[idx & 3u];
                a[j] = fma(a[j], kb, kc);
            }
    }
    double s = 0.;
#pragma unroll
    for (int j = 0; j < CHAINS; j++) s += a[j];
    if (s == 123456.789) sink[0] = s;
}

// Same with the multiplier read from shared memory (LDS, warp-uniform address = broadcast).
template <int CHAINS>
__global__ void fp64_lds_probe_kernel(int64_t iters, double *sink) {
    __shared__ double s_consts[64];
    if (threadIdx.x < 64) s_consts[threadIdx.x] = 0.999999 - 1e-6 * (threadIdx.x & 3);
    __syncthreads();
    double a[CHAINS];
#pragma unroll
    for (int j = 0; j < CHAINS; j++) a[j] = 1.0 + 1e-3 * (threadIdx.x + j);
    const double kc = 1e-6;
    for (int64_t it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 16 / CHAINS; r++)
#pragma unroll
            for (int j = 0; j < CHAINS; j++) {
                const double kb = s_consts[(uint32_t) (it + j + r) & 3u];
                a[j] = fma(a[j], kb, kc);
            }
    }
    double s = 0.;
#pragma unroll
    for (int j = 0; j < CHAINS; j++) s += a[j];
    if (s == 123456.789) sink[0] = s;
}

// Issue-slot probe: 16 independent DFMA chains interleaved with NINT independent 32-bit integer
// multiply-adds per DFMA.  If the time per DFMA does not grow with NINT <= 1, non-FP64 instructions
// issue in the shadow of the half-rate FP64 dispatch; if it grows, they compete for issue cycles.
template <int NINT>
__global__ void fp64_mix_probe_kernel(int64_t iters, double *sink) {
    double a[16];
    uint32_t x[16];
#pragma unroll
    for (int j = 0; j < 16; j++) {
        a[j] = 1.0 + 1e-3 * (threadIdx.x + j);
        x[j] = threadIdx.x * 2654435761u + j;
    }
    const double kb = 0.999999, kc = 1e-6;
    for (int64_t it = 0; it < iters; it++) {
#pragma unroll
        for (int j = 0; j < 16; j++) {
            a[j] = fma(a[j], kb, kc);
#pragma unroll
            for (int t = 0; t < NINT; t++)
                asm volatile("mad.lo.u32 %0, %0, 1664525, 1013904223;" : "+r"(x[j]));
        }
    }
    double s = 0.;
    uint32_t y = 0;
#pragma unroll
    for (int j = 0; j < 16; j++) {
        s += a[j];
        y ^= x[j];
    }
    if (s == 123456.789 || y == 0x12345678u) sink[0] = s + y;
}

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

static int g_max_blocks_per_sm = 0;   // measurement hook (0 = whatever fits)

template <typename Kernel>
static int persistent_grid(Kernel kernel, int64_t work_items, int &blocks) {
    DeviceInfo info;
    int rc = device_info(info);
    if (rc) return rc;
    int per_sm = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kThreads, 0);
    if (e != cudaSuccess) return (int) e;
    if (per_sm < 1) per_sm = 1;
    if (g_max_blocks_per_sm > 0 && per_sm > g_max_blocks_per_sm) per_sm = g_max_blocks_per_sm;
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
    if (kStreaming && aligned16(K, q, out) && n >= 2) {
        int rc = persistent_grid(vmap_kernel<PROCESS, 2>, n >> 1, blocks);
        if (rc) return rc;
        vmap_kernel<PROCESS, 2><<<blocks, kThreads, 0, s>>>(K, q, out, n, p);
    } else {
        int rc = persistent_grid(vmap_kernel<PROCESS, 1>, n, blocks);
        if (rc) return rc;
        vmap_kernel<PROCESS, 1><<<blocks, kThreads, 0, s>>>(K, q, out, n, p);
    }
    return after_launch();
}

// How the exchange form runs and where it fences its remote stores:
//   3 = persistent CTAs pulling rows from a queue, one fence per CTA at the end (default)
//   0 = one CTA per row, each writer lane fences right after its stores
//   1 = one CTA per row, one lane fences after the CTA barrier (NCCL's simple-protocol pattern)
//   2 = one CTA per row, only the last CTA of the grid fences (measurement only: not a sufficient
//       ordering on its own)
//   4 = one CTA per row with unfenced remote stores, then a second one-warp kernel that exchanges
//       the flags (the kernel boundary orders the stores)
static int g_exchange_fence_mode = 3;
static int g_pair_mode = 0;   // 0: one pair per thread (default), 1: one node per lane

static int vmap_impl(int process, const double *K, const double *q, double *out, int64_t n,
                     const Params &p, cudaStream_t s) {
    switch (process) {
        case NOA_DCS_BREMSSTRAHLUNG: return launch_vmap<0>(K, q, out, n, p, s);
        case NOA_DCS_PAIR_PRODUCTION:
            if (g_pair_mode == 1) {
                int blocks = 0;
                int rc = persistent_grid(vmap_pair_lanes_kernel, n * 8, blocks);
                if (rc) return rc;
                vmap_pair_lanes_kernel<<<blocks, kThreads, 0, s>>>(K, q, out, n, p);
                return after_launch();
            }
            return launch_vmap<1>(K, q, out, n, p, s);
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

}  // namespace noa_b200

#include "coulomb_kernels.cuh"

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
    }
    if (code > 0) return cudaGetErrorString((cudaError_t) code);
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

// test/bench hook: 0 = pair per thread, 1 = node per lane.  Not part of the reference surface.
int noa_dcs_set_pair_mode(int mode) {
    if (mode != 0 && mode != 1) return NOA_DCS_EINVAL;
    g_pair_mode = mode;
    return 0;
}

int noa_dcs_set_max_blocks_per_sm(int blocks) {
    if (blocks < 0) return NOA_DCS_EINVAL;
    g_max_blocks_per_sm = blocks;
    return 0;
}

int noa_dcs_set_exchange_fence_mode(int mode) {
    if (mode < 0 || mode > 4) return NOA_DCS_EINVAL;
    g_exchange_fence_mode = mode;
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

static int table_impl(unsigned process_mask, bool single_row, const double *K, int64_t nK,
                      double xlow, int32_t min_points, double A, double I, int32_t Z, double mass,
                      const TableOut &out, void *stream) {
    if (process_mask == 0 || process_mask > 15u || nK < 0 || min_points < 1)
        return NOA_DCS_EINVAL;
    if (nK == 0 || (!out.del[0] && !out.cel[0])) {
        if (!out.flags[0]) return 0;
        table_signal_kernel<<<1, 32, 0, (cudaStream_t) stream>>>(out);
        return after_launch();
    }
    if (!K) return NOA_DCS_EINVAL;
    TablePlan plan{};
    for (int i = 0; i < 4; i++) plan.out_row[i] = single_row ? 0 : i;
    static const int heavy_first[4] = {NOA_DCS_PHOTONUCLEAR, NOA_DCS_PAIR_PRODUCTION,
                                       NOA_DCS_BREMSSTRAHLUNG, NOA_DCS_IONISATION};
    for (int i = 0; i < 4; i++)
        if ((process_mask >> heavy_first[i]) & 1u) plan.process[plan.n_slots++] = heavy_first[i];
    plan.cells = ((uint32_t) min_points + 5u) / 6u;
    plan.xlow = xlow;
    const int64_t blocks = nK * plan.n_slots;
    if (blocks > 0x7fffffffLL) return NOA_DCS_ERANGE;
    DeviceInfo info;
    int rc = device_info(info);
    if (rc) return rc;
    const Params p = make_params(A, I, Z, mass);
    if (out.flags[0] != nullptr && out.fence_mode == 3) {
        int grid = 0;
        rc = persistent_grid(table_exchange_kernel, blocks * kThreads, grid);
        if (rc) return rc;
        table_exchange_kernel<<<grid, kThreads, 0, (cudaStream_t) stream>>>(K, nK, out, plan, p);
        return after_launch();
    }
    table_kernel<<<(unsigned) blocks, kThreads, 0, (cudaStream_t) stream>>>(K, nK, out, plan, p);
    if (out.flags[0] != nullptr && out.fence_mode == 4) {
        // the kernel boundary makes the rows visible system-wide; a one-warp kernel then does
        // the flag exchange
        rc = after_launch();
        if (rc) return rc;
        table_signal_kernel<<<1, 32, 0, (cudaStream_t) stream>>>(out);
    }
    return after_launch();
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

int noa_dcs_table_material_f64(unsigned process_mask, const double *K, int64_t nK, double xlow,
                               int32_t min_points, int32_t n_elements, const double *A,
                               const double *I, const int32_t *Z, const double *w, double mass,
                               double *scratch, double *table, void *stream) {
    if (n_elements < 1 || n_elements > NOA_DCS_MAX_ELEMENTS || !A || !I || !Z || !w)
        return NOA_DCS_EINVAL;
    if (process_mask == 0 || process_mask > 15u || nK < 0 || min_points < 1) return NOA_DCS_EINVAL;
    if (nK == 0) return 0;
    if (!K || !scratch || !table) return NOA_DCS_EINVAL;
    const int64_t columns = 8 * nK;
    // rows of processes outside the mask must mix to 0, not to whatever the scratch held
    cudaError_t e = cudaMemsetAsync(scratch, 0, (size_t) n_elements * columns * sizeof(double),
                                    (cudaStream_t) stream);
    if (e != cudaSuccess) return (int) e;
    MixWeights m{};
    m.n_elements = n_elements;
    for (int el = 0; el < n_elements; el++) {
        m.w[el] = w[el];
        double *part = scratch + (int64_t) el * columns;
        int rc = table_impl(process_mask, false, K, nK, xlow, min_points, A[el], I[el], Z[el], mass,
                            local_out(part, part + 4 * nK, nK), stream);
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
                               uint32_t *const *peer_flags, uint32_t *done, uint32_t epoch,
                               int64_t n_total, int64_t first_row, int64_t row_stride,
                               void *stream) {
    if (n_peers < 1 || n_peers > NOA_DCS_MAX_PEERS || my_peer < 0 || my_peer >= n_peers)
        return NOA_DCS_EINVAL;
    if (!peer_del || !peer_cel || !peer_flags || !done) return NOA_DCS_EINVAL;
    if (n_total < 0 || first_row < 0 || row_stride < 1 || n_local < 0) return NOA_DCS_EINVAL;
    if (n_local > 0 && first_row + (n_local - 1) * row_stride >= n_total) return NOA_DCS_ERANGE;
    TableOut out{};
    out.n_peers = n_peers;
    out.me = my_peer;
    out.n_total = n_total;
    out.first_row = first_row;
    out.row_stride = row_stride;
    out.done = done;
    out.epoch = epoch;
    out.fence_mode = g_exchange_fence_mode;
    for (int j = 0; j < n_peers; j++) {
        if (!peer_del[j] || !peer_cel[j] || !peer_flags[j]) return NOA_DCS_EINVAL;
        out.del[j] = peer_del[j];
        out.cel[j] = peer_cel[j];
        out.flags[j] = peer_flags[j];
    }
    return table_impl(process_mask, false, K_local, n_local, xlow, min_points, A, I, Z, mass, out,
                      stream);
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

// 1 if `ptr` is page-locked host memory the device can address (cudaHostAlloc / cudaHostRegister)
static bool device_addressable_host(const void *ptr) {
    cudaPointerAttributes attr{};
    if (cudaPointerGetAttributes(&attr, ptr) != cudaSuccess) {
        (void) cudaGetLastError();
        return false;
    }
    return attr.type == cudaMemoryTypeHost && attr.devicePointer != nullptr;
}

int noa_dcs_vmap_pinned_f64(int process, const double *h_K, const double *h_q, double *h_result,
                            int64_t n, double A, double I, int32_t Z, double mass, void *stream) {
    if (process < 0 || process >= NOA_DCS_NPROCESS || n < 0) return NOA_DCS_EINVAL;
    if (n == 0) return 0;
    if (!h_K || !h_q || !h_result) return NOA_DCS_EINVAL;
    if (!device_addressable_host(h_K) || !device_addressable_host(h_q) ||
        !device_addressable_host(h_result))
        return NOA_DCS_EINVAL;
    // The element-wise kernels run unchanged on the mapped host addresses: their coalesced loads
    // and stores cross PCIe themselves and overlap with the arithmetic of the other resident
    // warps.  Measured alternatives (TMA bulk-copy ring with mbarriers; per-thread cp.async
    // prefetch into shared memory) were slower or equal -- profiles/r01_host_path_variants.md.
    const Params p = make_params(A, I, Z, mass);
    return vmap_impl(process, h_K, h_q, h_result, n, p, (cudaStream_t) stream);
}

int noa_dcs_fp64_probe(int64_t iters, int32_t blocks, int32_t threads, double *sink,
                       void *stream) {
    if (iters < 1 || blocks < 1 || threads < 1 || threads > 1024 || !sink) return NOA_DCS_EINVAL;
    fp64_probe_kernel<0><<<blocks, threads, 0, (cudaStream_t) stream>>>(iters, sink);
    return after_launch();
}

int noa_dcs_fp64_probe_mode(int32_t mode, int64_t iters, int32_t blocks, int32_t threads,
                            double *sink, void *stream) {
    if (iters < 1 || blocks < 1 || threads < 1 || threads > 1024 || !sink) return NOA_DCS_EINVAL;
    cudaStream_t s = (cudaStream_t) stream;
    switch (mode) {
        case 0: fp64_probe_kernel<0><<<blocks, threads, 0, s>>>(iters, sink); break;
        case 1: fp64_probe_kernel<1><<<blocks, threads, 0, s>>>(iters, sink); break;
        case 2: fp64_probe_kernel<2><<<blocks, threads, 0, s>>>(iters, sink); break;
        case 3: fp64_probe_kernel<3><<<blocks, threads, 0, s>>>(iters, sink); break;
        case 4: fp64_probe_kernel<4><<<blocks, threads, 0, s>>>(iters, sink); break;
        case 5: fp64_mix_probe_kernel<1><<<blocks, threads, 0, s>>>(iters, sink); break;
        case 6: fp64_mix_probe_kernel<2><<<blocks, threads, 0, s>>>(iters, sink); break;
        case 7: fp64_mix_probe_kernel<3><<<blocks, threads, 0, s>>>(iters, sink); break;
        case 20: fp64_ldc_probe_kernel<1, 1><<<blocks, threads, 0, s>>>(iters, sink); break;
        case 21: fp64_ldc_probe_kernel<1, 0><<<blocks, threads, 0, s>>>(iters, sink); break;
        case 22: fp64_ldc_probe_kernel<4, 1><<<blocks, threads, 0, s>>>(iters, sink); break;
        case 23: fp64_ldc_probe_kernel<4, 0><<<blocks, threads, 0, s>>>(iters, sink); break;
        case 24: fp64_lds_probe_kernel<1><<<blocks, threads, 0, s>>>(iters, sink); break;
        case 25: fp64_lds_probe_kernel<4><<<blocks, threads, 0, s>>>(iters, sink); break;
        case 10: fp64_chain_probe_kernel<1><<<blocks, threads, 0, s>>>(iters, sink); break;
        case 11: fp64_chain_probe_kernel<2><<<blocks, threads, 0, s>>>(iters, sink); break;
        case 12: fp64_chain_probe_kernel<4><<<blocks, threads, 0, s>>>(iters, sink); break;
        case 13: fp64_chain_probe_kernel<8><<<blocks, threads, 0, s>>>(iters, sink); break;
        default: return NOA_DCS_EINVAL;
    }
    return after_launch();
}

int noa_dcs_launch_info(int process, int32_t *blocks, int32_t *threads, int32_t *sm_count) {
    DeviceInfo info;
    int rc = device_info(info);
    if (rc) return rc;
    int b = 0;
    switch (process) {
        case 0: rc = persistent_grid(vmap_kernel<0, 2>, INT64_MAX / 2, b); break;
        case 1: rc = persistent_grid(vmap_kernel<1, 1>, INT64_MAX / 2, b); break;
        case 2: rc = persistent_grid(vmap_kernel<2, 1>, INT64_MAX / 2, b); break;
        case 3: rc = persistent_grid(vmap_kernel<3, 2>, INT64_MAX / 2, b); break;
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
