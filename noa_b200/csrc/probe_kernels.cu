// Measurement kernels (libnoa_dcs_b200_probe.so; C ABI in include/noa_dcs_b200_probe.h).  Nothing
// here is part of the product: bench.py uses the FP64 probe for the roofline denominator, the
// tools/ scripts use the rest for the studies under profiles/.
//
//   fp64_probe_kernel<MODE>        dependent-chain-free FP64 loops (pipe peak, operand shapes)
//   fp64_chain / ldc / lds / mix   latency, constant-load and issue-slot probes
//   vmap_pair_lanes_kernel         pair production with one Gauss-Legendre node per lane (8 lanes
//                                  per pair, shuffle gather, serial-order sum): the mapping
//                                  north_star names, measured against the product's one pair per
//                                  thread (2.43 vs 4.56 G evals/s) and bit-identical to it
#include <cuda_runtime.h>

#include "../../include/noa_dcs_b200_probe.h"
#include "dcs_device.cuh"
#include "dcs_params.hh"

namespace noa_b200 {

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
}  // namespace noa_b200

using namespace noa_b200;

extern "C" {

int noa_dcs_probe_pair_lanes_f64(const double *K, const double *q, double *result, int64_t n,
                                 double A, double I, int32_t Z, double mass, void *stream) {
    if (n < 0) return NOA_DCS_EINVAL;
    if (n == 0) return 0;
    if (!K || !q || !result) return NOA_DCS_EINVAL;
    int dev = 0, sms = 0, per_sm = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e == cudaSuccess)
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, vmap_pair_lanes_kernel, kThreads,
                                                          0);
    if (e != cudaSuccess) return (int) e;
    const int64_t need = (n * 8 + kThreads - 1) / kThreads;
    const int64_t cap = (int64_t) sms * (per_sm < 1 ? 1 : per_sm);
    const Params p = make_params(A, I, Z, mass);
    vmap_pair_lanes_kernel<<<(unsigned) (need < cap ? need : cap), kThreads, 0,
                             (cudaStream_t) stream>>>(K, q, result, n, p);
    return (int) cudaPeekAtLastError();
}

int noa_dcs_fp64_probe(int64_t iters, int32_t blocks, int32_t threads, double *sink,
                       void *stream) {
    return noa_dcs_fp64_probe_mode(0, iters, blocks, threads, sink, stream);
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
    return (int) cudaPeekAtLastError();
}

}  // extern "C"
