// Per-material table assembly: what PUMAS does with per-element DCS
// integrals when it builds the tables of one material (src/noa/3rdparty/_pumas/pumas.c, v1.2.1),
// on the GPU and over NOA's own DCS.  Included by dcs_kernels.cu; oracle: oracle/material_oracle.c.
//
//   step 1  per (element, process, energy): CSn, cel over x in [cutoff, 1] -- the existing fused
//           table kernels -- and the ionisation straggling integral (mode 2, dcs q^3) over
//           x in [1e-6, cutoff] -- the same kernels in their generalised form
//                                                              pumas.c:10768-10808, 10901-10955
//   step 2  material_mix_kernel: one thread per energy; mass-fraction mixing in composition
//           order, straggling, total cross-section and the normalised cumulative fractions CSf
//                                                              pumas.c:8054-8111, 8130-8131
//   step 3  material_threshold_kernel (one CTA): first row >= 1 with a non-zero total
//           cross-section -> Kt; rows below it take that row's value       pumas.c:10820-10833
//   step 4  material_xt_kernel: one thread per (element, process, energy); doubling from the
//           cutoff until the DCS is positive, then bisection down to 1 % of the cutoff -- every
//           probe is one evaluation of the element-wise DCS, and because those are bit-identical
//           to the reference's, every branch of the search goes the same way
//                                                              pumas.c:10839-10880
#pragma once

#include "dcs_device.cuh"

namespace noa_b200 {

struct MaterialParams {
    int32_t n_elements;
    double w[NOA_DCS_MAX_ELEMENTS];
    Params p[NOA_DCS_MAX_ELEMENTS];
};

// elem: [ne][3][4][nK] (CSn, cel, stg).  One thread per energy.
__global__ void material_mix_kernel(const double *__restrict__ elem, int64_t nK, int32_t ne,
                                    const __grid_constant__ MixWeights m, double *__restrict__ cs,
                                    double *__restrict__ cel, double *__restrict__ straggling,
                                    double *__restrict__ csf, double *__restrict__ cs_total) {
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    const int64_t n4 = 4 * nK;
    for (int64_t row = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; row < nK; row += stride) {
        double frct_cel[4] = {0., 0., 0., 0.};
        double frct_cs[4] = {0., 0., 0., 0.};
        double strag = 0.;
        for (int ic = 0; ic < ne; ic++) {
            const double *e = elem + (int64_t) ic * 3 * n4;
            const double w = m.w[ic];
#pragma unroll
            for (int ip = 0; ip < 4; ip++) {
                const double f = e[ip * nK + row] * w;
                csf[(int64_t) ic * n4 + ip * nK + row] = f;
                frct_cs[ip] += f;
                frct_cel[ip] += e[n4 + ip * nK + row] * w;
                strag += e[2 * n4 + ip * nK + row] * w;
            }
        }
        double frct_cs_del = 0.;
#pragma unroll
        for (int ip = 0; ip < 4; ip++) {
            frct_cs_del += frct_cs[ip];
            cs[ip * nK + row] = frct_cs[ip];
            cel[ip * nK + row] = frct_cel[ip];
        }
        if (frct_cs_del <= 0.) {
            for (int ic = 0; ic < ne; ic++)
                for (int ip = 0; ip < 4; ip++) csf[(int64_t) ic * n4 + ip * nK + row] = 0.;
        } else {
            double sum_tot = 0.;
            for (int ic = 0; ic < ne; ic++)
                for (int ip = 0; ip < 4; ip++) sum_tot += csf[(int64_t) ic * n4 + ip * nK + row];
            double sum = 0.;
            for (int ic = 0; ic < ne; ic++)
                for (int ip = 0; ip < 4; ip++) {
                    double *slot = csf + (int64_t) ic * n4 + ip * nK + row;
                    sum += *slot;
                    *slot = sum / sum_tot;
                }
            csf[(int64_t) (ne - 1) * n4 + 3 * nK + row] = 1.;   // protect against rounding
        }
        straggling[row] = strag;
        cs_total[row] = frct_cs_del;
    }
}

// One CTA: it = first row >= 1 with cs_total != 0 (nK if none), kt = K[it], rows 1 .. it-1 <- cs0.
__global__ void material_threshold_kernel(const double *__restrict__ K, int64_t nK,
                                          double *__restrict__ cs_total, double *__restrict__ kt,
                                          int32_t *__restrict__ it_out) {
    __shared__ int s_it;
    if (threadIdx.x == 0) s_it = (int) nK;
    __syncthreads();
    int first = (int) nK;
    for (int64_t row = 1 + threadIdx.x; row < nK; row += blockDim.x)
        if (cs_total[row] != 0.) {
            first = (int) row;
            break;                      // rows ascend per thread: the first hit is its smallest
        }
    first = __reduce_min_sync(0xffffffffu, first);
    if ((threadIdx.x & 31) == 0) atomicMin(&s_it, first);
    __syncthreads();
    const int it = s_it;
    if (threadIdx.x == 0) {
        *it_out = it;
        *kt = (it < nK) ? K[it] : 0.;
    }
    if (it >= nK) return;
    const double cs0 = cs_total[it];
    __syncthreads();
    for (int64_t row = 1 + threadIdx.x; row < it; row += blockDim.x) cs_total[row] = cs0;
}

template <int PROCESS>
__device__ __noinline__ double material_xt_search(double k, double cutoff, const Params &p,
                                                  const glibm::Tab &T) {
    double x = cutoff;
    while ((x < 1.) && (dcs_eval<PROCESS>(k, k * x, p, T) <= 0.)) x *= 2;
    if (x >= 1.) return 1.;
    if (x > cutoff) {
        const double eps = 1E-02 * cutoff;
        double x0 = 0.5 * x;
        double dcs = 0.;
        for (;;) {
            if (dcs == 0.)
                x0 += 0.5 * (x - x0);
            else {
                const double dx = x - x0;
                x = x0;
                x0 -= 0.5 * dx;
            }
            if ((x - x0) <= eps) break;
            dcs = dcs_eval<PROCESS>(k, k * x0, p, T);
        }
    }
    return x;
}

// xt: [ne][4][nK]; thread index = (element, process) major, energy minor, so warps are uniform in
// the integrand they run
__global__ void __launch_bounds__(128)
material_xt_kernel(const double *__restrict__ K, int64_t nK, double cutoff,
                   const int32_t *__restrict__ it_ptr, const __grid_constant__ MaterialParams m,
                   double *__restrict__ xt) {
    __shared__ glibm::Tables s_tables;
    const glibm::Tab T = stage_tables(s_tables);
    const int64_t it = *it_ptr;
    const int64_t total = (int64_t) m.n_elements * 4 * nK;
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    for (int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int64_t row = i % nK;
        const int ip = (int) ((i / nK) & 3);
        const int iel = (int) (i / (4 * nK));
        double x = 1.;
        if (row >= it) {
            const double k = K[row];
            const Params &p = m.p[iel];
            switch (ip) {
                case 0: x = material_xt_search<0>(k, cutoff, p, T); break;
                case 1: x = material_xt_search<1>(k, cutoff, p, T); break;
                case 2: x = material_xt_search<2>(k, cutoff, p, T); break;
                default: x = material_xt_search<3>(k, cutoff, p, T); break;
            }
        }
        xt[i] = x;
    }
}

}  // namespace noa_b200
