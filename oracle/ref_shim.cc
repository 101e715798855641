// TEST INFRASTRUCTURE ONLY -- never linked or imported by the product path (noa_b200/).
//
// C-ABI shim around the UNMODIFIED reference headers, compiled where they lie under
// /root/reference/src (recipe: oracle/Makefile, output: oracle/_ref/libnoa_ref.so, git-ignored).
// It lets the tests / bench.py cpu_baseline call the reference's own CPU implementation of the
// DCS hot path on raw double buffers:
//   noa::pms::dcs::vmap / pvmap           (src/noa/pms/dcs.hh:35-75)
//   noa::pms::dcs::vmap_integral          (src/noa/pms/dcs.hh:115-130)
//   noa::pms::dcs::recoil_integral        (src/noa/pms/dcs.hh:89-105, 955-1001)
//   noa::pms::dcs::coulomb_data / coulomb_transport / hard_scattering / soft_scattering
//                                         (src/noa/pms/dcs.hh:600-622, 674-693, 843-872, 940-952)
// No reference source is copied into this repository; this file only *calls* it.
#include <noa/pms/dcs.hh>

#include <omp.h>
#include <cstdint>

using namespace noa::pms;

namespace {
    inline torch::Tensor wrap(const double *p, int64_t n) {
        return torch::from_blob(const_cast<double *>(p), {n}, torch::kFloat64);
    }

    template<typename F>
    inline void run_vmap(const F &f, int parallel, double *out, const double *K, const double *q,
                         int64_t n, const AtomicElement &el, double mass) {
        auto r = wrap(out, n);
        auto k = wrap(K, n);
        auto qq = wrap(q, n);
        if (parallel) dcs::pvmap(f)(r, k, qq, el, mass);
        else dcs::vmap(f)(r, k, qq, el, mass);
    }

    template<typename F, typename G>
    inline void run_integral(const F &f, const G &g, int parallel, double *out, const double *K,
                             int64_t n, double xlow, const AtomicElement &el, double mass,
                             int32_t min_points) {
        if (!parallel) {
            // the reference's own (serial-only) driver
            auto r = wrap(out, n);
            auto k = wrap(K, n);
            dcs::vmap_integral(dcs::recoil_integral(f, g))(r, k, xlow, el, mass, min_points);
        } else {
            // harness-side OpenMP loop over energies around the unmodified closure
            // ("harness-parallelised reference arithmetic")
#pragma omp parallel for schedule(dynamic, 8)
            for (int64_t i = 0; i < n; i++)
                out[i] = dcs::recoil_integral(f, g)(K[i], xlow, el, mass, min_points);
        }
    }
}

extern "C" {

int noa_ref_threads() { return omp_get_max_threads(); }

// torchrun exports OMP_NUM_THREADS=1; the bench's reference arm restores the host's core count
void noa_ref_set_threads(int n) { if (n > 0) omp_set_num_threads(n); }

// process: 0 brems, 1 pair, 2 photonuclear, 3 ionisation
int noa_ref_vmap(int process, int parallel, const double *K, const double *q, double *out,
                 int64_t n, double A, double I, int32_t Z, double mass) {
    const AtomicElement el{A, I, Z};
    switch (process) {
        case 0: run_vmap(dcs::bremsstrahlung, parallel, out, K, q, n, el, mass); return 0;
        case 1: run_vmap(dcs::pair_production, parallel, out, K, q, n, el, mass); return 0;
        case 2: run_vmap(dcs::photonuclear, parallel, out, K, q, n, el, mass); return 0;
        case 3: run_vmap(dcs::ionisation, parallel, out, K, q, n, el, mass); return 0;
    }
    return 1;
}

// integrand: 0 = del_integrand (dcs*q), 1 = cel_integrand (dcs*q*q)
int noa_ref_vmap_integral(int process, int integrand, int parallel, const double *K, double *out,
                          int64_t n, double xlow, int32_t min_points, double A, double I,
                          int32_t Z, double mass) {
    const AtomicElement el{A, I, Z};
#define NOA_REF_CASE(P, F)                                                                       \
    case P:                                                                                      \
        if (integrand == 0)                                                                      \
            run_integral(F, dcs::del_integrand, parallel, out, K, n, xlow, el, mass, min_points); \
        else                                                                                     \
            run_integral(F, dcs::cel_integrand, parallel, out, K, n, xlow, el, mass, min_points); \
        return 0;
    switch (process) {
        NOA_REF_CASE(0, dcs::bremsstrahlung)
        NOA_REF_CASE(1, dcs::pair_production)
        NOA_REF_CASE(2, dcs::photonuclear)
        NOA_REF_CASE(3, dcs::ionisation)
    }
#undef NOA_REF_CASE
    return 1;
}

// ---- scalar entry points for oracle/material_oracle.c ---------------------
// The reference's own scalar DCS lambdas and its own quadrature; the mode-2 (straggling) integrand
// and the upper bound x_high follow PUMAS's compute_dcs_integral (pumas.c:10901-10955) in the shape
// of dcs::recoil_integral (dcs.hh:89-105).  Modes 0 / 1 with x_high = 1 ARE dcs::recoil_integral.
double noa_ref_dcs_scalar(int process, double K, double q, double A, double I, int32_t Z,
                          double mass) {
    const AtomicElement el{A, I, Z};
    switch (process) {
        case 0: return dcs::bremsstrahlung(K, q, el, mass);
        case 1: return dcs::pair_production(K, q, el, mass);
        case 2: return dcs::photonuclear(K, q, el, mass);
        default: return dcs::ionisation(K, q, el, mass);
    }
}

}  // extern "C"

namespace {
    template<typename F>
    double integral_mode(const F &f, int mode, double K, double xlow, double xhigh,
                         const AtomicElement &el, double mass, int32_t min_points) {
        if (mode <= 1 && xhigh == 1.)
            return (mode == 0)
                   ? dcs::recoil_integral(f, dcs::del_integrand)(K, xlow, el, mass, min_points)
                   : dcs::recoil_integral(f, dcs::cel_integrand)(K, xlow, el, mass, min_points);
        return noa::utils::numerics::quadrature6<Scalar>(
                log(K * xlow), log(K * xhigh),
                [&](const Scalar &t) {
                    const Scalar q = exp(t);
                    Scalar y = f(K, q, el, mass) * q;
                    if (mode > 0) y *= q;
                    if (mode > 1) y *= q;
                    return y;
                },
                min_points) / (K + mass);
    }
}

extern "C" {

double noa_ref_integral_scalar(int process, int mode, double K, double xlow, double xhigh,
                               double A, double I, int32_t Z, double mass, int32_t min_points) {
    const AtomicElement el{A, I, Z};
    switch (process) {
        case 0: return integral_mode(dcs::bremsstrahlung, mode, K, xlow, xhigh, el, mass, min_points);
        case 1: return integral_mode(dcs::pair_production, mode, K, xlow, xhigh, el, mass, min_points);
        case 2: return integral_mode(dcs::photonuclear, mode, K, xlow, xhigh, el, mass, min_points);
        default: return integral_mode(dcs::ionisation, mode, K, xlow, xhigh, el, mass, min_points);
    }
}

// ---- Coulomb and soft scattering ------------------------------------
int noa_ref_coulomb_data(double *fcm, double *screening, double *fspin, double *invlambda,
                         const double *K, int64_t n, double A, double I, int32_t Z, double mass) {
    const AtomicElement el{A, I, Z};
    auto opt = torch::kFloat64;
    dcs::coulomb_data(torch::from_blob(fcm, {n, 2}, opt), torch::from_blob(screening, {n, 9}, opt),
                      wrap(fspin, n), wrap(invlambda, n), wrap(K, n), el, mass);
    return 0;
}

int noa_ref_coulomb_transport(double *coef, const double *screening, const double *fspin,
                              const double *mu, int64_t n_mu, int64_t n) {
    auto opt = torch::kFloat64;
    dcs::coulomb_transport(torch::from_blob(coef, {n, 2}, opt),
                           torch::from_blob(const_cast<double *>(screening), {n, 9}, opt),
                           wrap(fspin, n), wrap(mu, n_mu));
    return 0;
}

int noa_ref_hard_scattering(double *mu0, double *lb_h, const double *G, const double *fcm,
                            const double *screening, const double *invlambda, const double *fspin,
                            int32_t nel, int32_t nkin) {
    auto opt = torch::kFloat64;
    auto c = [](const double *p) { return const_cast<double *>(p); };
    dcs::hard_scattering(wrap(mu0, nkin), wrap(lb_h, nkin),
                         torch::from_blob(c(G), {nel, nkin, 2}, opt),
                         torch::from_blob(c(fcm), {nel, nkin, 2}, opt),
                         torch::from_blob(c(screening), {nel, nkin, 9}, opt),
                         torch::from_blob(c(invlambda), {nel, nkin}, opt),
                         torch::from_blob(c(fspin), {nel, nkin}, opt));
    return 0;
}

int noa_ref_soft_scattering(double *ms1, const double *K, int64_t n, double A, double I, int32_t Z,
                            double mass) {
    const AtomicElement el{A, I, Z};
    dcs::soft_scattering(wrap(ms1, n), wrap(K, n), el, mass);
    return 0;
}

}  // extern "C"
