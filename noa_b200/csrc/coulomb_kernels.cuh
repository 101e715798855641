// Kernels and C ABI of the Coulomb / soft-scattering part of the reference's dcs.hh
// (src/noa/pms/dcs.hh:499-952); included by dcs_kernels.cu (one translation unit, one library).
//
//   coulomb_data_kernel        one energy per thread: frame factors, spin factor, 9 screening
//                              factors, inverse Wentzel path                (dcs.hh:600-622)
//   coulomb_transport_kernel   one energy per thread                        (dcs.hh:674-693)
//   hard_scattering_kernel     one energy per thread, Ridders root inside   (dcs.hh:843-872)
//   soft_scattering_kernel     one CTA (128 threads) per energy: the 102 nodes of the photonuclear
//                              transport quadrature across threads, each running the 9-node
//                              photonuclear DCS; terms added in node order  (dcs.hh:901-952)
// Per-energy work with a handful of outputs: these are latency-sized launches (10^4 energies), not
// roofline kernels; what matters is that they are bit-identical to the reference's CPU results.
#pragma once

#include "coulomb_math.cuh"

namespace noa_b200 {

constexpr int kCoulombThreads = 128;

__global__ void __launch_bounds__(kCoulombThreads)
coulomb_data_kernel(const double *__restrict__ K, int64_t n, double *__restrict__ fcm,
                    double *__restrict__ screening, double *__restrict__ fspin,
                    double *__restrict__ invlambda, const __grid_constant__ CoulombParams c) {
    __shared__ glibm::Tables s_tables;
    const glibm::Tab T = stage_tables(s_tables);
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    for (int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        double f0, f1, ps[kScreenFactors];
        const double kinetic0 = coulomb_frame(K[i], c, f0, f1);
        fcm[2 * i] = f0;
        fcm[2 * i + 1] = f1;
        fspin[i] = coulomb_spin(kinetic0, c.mass);
        invlambda[i] = coulomb_screening(kinetic0, c, T, ps);
#pragma unroll
        for (int j = 0; j < kScreenFactors; j++) screening[kScreenFactors * i + j] = ps[j];
    }
}

__global__ void __launch_bounds__(kCoulombThreads)
coulomb_transport_kernel(const double *__restrict__ screening, const double *__restrict__ fspin,
                         const double *__restrict__ mu, int64_t n_mu, int64_t n,
                         double *__restrict__ coefficients) {
    __shared__ glibm::Tables s_tables;
    const glibm::Tab T = stage_tables(s_tables);
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    for (int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        double ps[kScreenFactors];
#pragma unroll
        for (int j = 0; j < kScreenFactors; j++) ps[j] = screening[kScreenFactors * i + j];
        double g0, g1;
        coulomb_transport_coefficients(ps, fspin[i], mu[n_mu == 1 ? 0 : i], T, g0, g1);
        coefficients[2 * i] = g0;
        coefficients[2 * i + 1] = g1;
    }
}

__global__ void __launch_bounds__(kCoulombThreads)
hard_scattering_kernel(const double *__restrict__ G, const double *__restrict__ fcm,
                       const double *__restrict__ screening, const double *__restrict__ invlambda,
                       const double *__restrict__ fspin, int32_t nel, int64_t nkin,
                       double max_mu0, double *__restrict__ mu0, double *__restrict__ lb_h) {
    __shared__ glibm::Tables s_tables;
    const glibm::Tab T = stage_tables(s_tables);
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    for (int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; i < nkin; i += stride) {
        HardView v;
        v.G = G + 2 * i;
        v.fcm = fcm + 2 * i;
        v.screen = screening + kScreenFactors * i;
        v.invlambda = invlambda + i;
        v.fspin = fspin + i;
        v.nel = nel;
        v.nkin = nkin;
        double m, l;
        coulomb_hard_scattering(v, max_mu0, T, m, l);
        mu0[i] = m;
        lb_h[i] = l;
    }
}

__global__ void __launch_bounds__(kCoulombThreads)
soft_scattering_kernel(const double *__restrict__ K, int64_t n, double *__restrict__ ms1,
                       const __grid_constant__ Params p, const __grid_constant__ CoulombParams c) {
    __shared__ StagedShared s_staged;
    __shared__ double s_term[kSoftNodes];
    const glibm::Tab T = stage_all(s_staged, p);
    // the 102 photonuclear DCS values of a row with the folded special-case tests (folded_ops.cuh)
    const auto photonuclear_dcs = [&](double kk, double qq) { return dcs_value<2, true>(kk, qq, p, T); };
    for (int64_t row = blockIdx.x; row < n; row += gridDim.x) {
        const double k = K[row];
        if (threadIdx.x < kSoftNodes)
            s_term[threadIdx.x] = soft_photonuclear_term(threadIdx.x, k, p, c, T, photonuclear_dcs);
        __syncthreads();
        if (threadIdx.x == 0) {
            double acc = 0.;
            for (int i = 0; i < kSoftNodes; i++) acc += s_term[i];     // numerics.hh:84-87
            ms1[row] = transverse_transport_ionisation(k, p, c, T) + 2. * acc;   // dcs.hh:947-949
        }
        __syncthreads();
    }
}

static int per_thread_grid(int64_t n) {
    const int64_t blocks = (n + kCoulombThreads - 1) / kCoulombThreads;
    return (int) (blocks < 1 ? 1 : (blocks > 148 * 16 ? 148 * 16 : blocks));
}

}  // namespace noa_b200

extern "C" {

int noa_dcs_coulomb_data_f64(const double *K, int64_t n, double A, double I, int32_t Z, double mass,
                             double *fcm, double *screening, double *fspin, double *invlambda,
                             void *stream) {
    using namespace noa_b200;
    if (n < 0) return NOA_DCS_EINVAL;
    if (n == 0) return 0;
    if (!K || !fcm || !screening || !fspin || !invlambda) return NOA_DCS_EINVAL;
    DeviceInfo info;
    int rc = device_info(info);
    if (rc) return rc;
    const CoulombParams c = make_coulomb_params(A, I, Z, mass);
    coulomb_data_kernel<<<per_thread_grid(n), kCoulombThreads, 0, (cudaStream_t) stream>>>(
            K, n, fcm, screening, fspin, invlambda, c);
    return after_launch();
}

int noa_dcs_coulomb_transport_f64(const double *screening, const double *fspin, const double *mu,
                                  int64_t n_mu, int64_t n, double *coefficients, void *stream) {
    using namespace noa_b200;
    if (n < 0 || (n_mu != 1 && n_mu != n)) return NOA_DCS_EINVAL;
    if (n == 0) return 0;
    if (!screening || !fspin || !mu || !coefficients) return NOA_DCS_EINVAL;
    DeviceInfo info;
    int rc = device_info(info);
    if (rc) return rc;
    coulomb_transport_kernel<<<per_thread_grid(n), kCoulombThreads, 0, (cudaStream_t) stream>>>(
            screening, fspin, mu, n_mu, n, coefficients);
    return after_launch();
}

int noa_dcs_hard_scattering_f64(const double *coefficients, const double *fcm,
                                const double *screening, const double *invlambda,
                                const double *fspin, int32_t nel, int64_t nkin, double *mu0,
                                double *lb_h, void *stream) {
    using namespace noa_b200;
    if (nel < 1 || nkin < 0) return NOA_DCS_EINVAL;
    if (nkin == 0) return 0;
    if (!coefficients || !fcm || !screening || !invlambda || !fspin || !mu0 || !lb_h)
        return NOA_DCS_EINVAL;
    DeviceInfo info;
    int rc = device_info(info);
    if (rc) return rc;
    // MAX_MU0 (physics.hh:82-83) does not depend on the element
    const double max_mu0 = make_coulomb_params(1., 1., 1, 1.).h_max_mu0;
    hard_scattering_kernel<<<per_thread_grid(nkin), kCoulombThreads, 0, (cudaStream_t) stream>>>(
            coefficients, fcm, screening, invlambda, fspin, nel, nkin, max_mu0, mu0, lb_h);
    return after_launch();
}

int noa_dcs_soft_scattering_f64(const double *K, int64_t n, double A, double I, int32_t Z,
                                double mass, double *ms1, void *stream) {
    using namespace noa_b200;
    if (n < 0) return NOA_DCS_EINVAL;
    if (n == 0) return 0;
    if (!K || !ms1) return NOA_DCS_EINVAL;
    DeviceInfo info;
    int rc = device_info(info);
    if (rc) return rc;
    const Params p = make_params(A, I, Z, mass);
    const CoulombParams c = make_coulomb_params(A, I, Z, mass);
    const int64_t cap = (int64_t) info.sm_count * 8;
    const int grid = (int) (n < cap ? n : cap);
    soft_scattering_kernel<<<grid, kCoulombThreads, 0, (cudaStream_t) stream>>>(K, n, ms1, p, c);
    return after_launch();
}

}  // extern "C"
