// Device-side helpers shared by the kernels of the product library (dcs_kernels.cu) and of the
// measurement library (probe_kernels.cu): the exp/log tables and per-launch reciprocals staged in
// shared memory, and the evaluation of one DCS value with the folded special-case tests.
#pragma once

#include <cuda_runtime.h>

#include "../../include/noa_dcs_b200.h"
#include "dcs_math.cuh"

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
// kernels 2 (+2.7 % over 3).  Table kernels (table_kernels.cuh; profiles/r02_table_variants.jsonl):
// pair 4 (64 registers: 1.677 ms for the pair rows of config 4 against 1.733 at 5), photonuclear
// 3, the two cheap processes 4 (5 changes nothing); the combined four-process kernel takes the
// budget of its largest member (3 CTAs = 80 registers).
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
#ifndef NOA_MINB_TABLE_PAIR
#define NOA_MINB_TABLE_PAIR 4
#endif
#ifndef NOA_MINB_TABLE_PHOTO
#define NOA_MINB_TABLE_PHOTO 3
#endif
#ifndef NOA_MINB_TABLE_LIGHT
#define NOA_MINB_TABLE_LIGHT 4
#endif
#ifndef NOA_MINB_TABLE_ALL
#define NOA_MINB_TABLE_ALL 3
#endif
// pairs per thread and iteration of the two streaming kernels (2 or 4); 4 (all four 128-bit loads
// issued before the first evaluation) measured 2 % SLOWER on bremsstrahlung and equal on
// ionisation (profiles/r02_table_variants.jsonl), so 2 stays
#ifndef NOA_STREAM_VEC
#define NOA_STREAM_VEC 2
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

// pair production with the kinetic-energy-only part (PairRow) handed in; an evaluation whose
// flag drops is redone with the plain operations from scratch, as everywhere else
__device__ __forceinline__ double dcs_value_pair_row(double K, double q, const PairRow &row,
                                                     const Params &p, const glibm::Tab &T) {
#if NOA_FOLDED_OPS
    FoldedOps<true> dv;
    dv.dens = T.aux_smem;
    double v = pair_production(K, q, p, T, dv, &row);
    if (!dv.ok()) v = dcs_eval_plain<1>(K, q, p, T);
    return v;
#else
    return dcs_eval<1>(K, q, p, T);
#endif
}

}  // namespace noa_b200
