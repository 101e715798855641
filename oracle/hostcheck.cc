// TEST INFRASTRUCTURE ONLY (built by tests/, never by or for the product).
//
// Compiles the kernels' scalar arithmetic (noa_b200/csrc/dcs_math.cuh, glibm.cuh) for the HOST so
// that, in a container without a GPU, tests can check
//   (1) glibm::exp/log/log10 == the system libm, bit for bit, on ~10^8 arguments, and
//   (2) the hoisted-invariant DCS formulas == the oracle, bit for bit.
// The GPU runs the same IEEE operation sequence (-fmad=false, IEEE div/sqrt), so this is the
// cheap pre-flight for the -m gpu parity tests.  Build: see tests/conftest.py (g++ -O2
// -ffp-contract=off -mfma).
#include <cmath>
#include <cstdint>
#include <cstring>
#include <random>

#include "../noa_b200/csrc/coulomb_math.cuh"
#include "../noa_b200/csrc/dcs_math.cuh"
#include "../noa_b200/csrc/dcs_params.hh"

using namespace noa_b200;

static const glibm::Tables kTables = {GLIBM_EXP_TABLE_INIT, GLIBM_LOG_TABLE_INIT};
static const glibm::Tab kT = glibm::make_host_tab(&kTables);

static inline bool same(double a, double b) {
    return std::memcmp(&a, &b, 8) == 0 || (std::isnan(a) && std::isnan(b));
}

extern "C" {

// returns the number of mismatches against libm over n pseudo-random arguments per function
int64_t hostcheck_glibm(int64_t n, uint64_t seed) {
    std::mt19937_64 rng(seed);
    std::uniform_real_distribution<double> U(0, 1);
    int64_t bad = 0;
    for (int64_t i = 0; i < n; i++) {
        const double u = U(rng), v = U(rng);
        double xe, xl;
        switch (i & 7) {
            case 0: xe = -750 * u; break;
            case 1: xe = 710 * u; break;
            case 2: xe = -40 * u; break;
            case 3: xe = (u - 0.5) * 2; break;
            case 4: xe = -1e-3 * u; break;
            case 5: xe = (u - 0.5) * 1e-10; break;
            case 6: xe = -100 * u; break;
            default: xe = (u - 0.5) * 1500;
        }
        switch (i & 7) {
            case 0: xl = std::exp((v - 0.5) * 1400); break;
            case 1: xl = 1 + (v - 0.5) * 0.2; break;
            case 2: xl = 1 + (v - 0.5) * 1e-6; break;
            case 3: xl = v * 10; break;
            case 4: xl = 1 + v * 1e-3; break;
            case 5: xl = v * 1e-310; break;
            case 6: xl = 3 + v * 1000; break;
            default: xl = 1 + 1 / (v * 1e3 + 1e-3);
        }
        bad += !same(glibm::exp(xe, kT), std::exp(xe));
        bad += !same(glibm::log(xl, kT), std::log(xl));
        bad += !same(glibm::log10(xl, kT), std::log10(xl));
    }
    const double sp[] = {0.0, -0.0, 1.0, INFINITY, -INFINITY, NAN, -1.0, 5e-324, 1e-310, 1.7e308,
                         709.9, -745.2, -746, 800, -800, 0x1p-54, 0x1p-55, 512, -512, 1024, -1024,
                         0.9375, 1.0647, 1.06469, 0.93749};
    for (double x : sp) {
        bad += !same(glibm::exp(x, kT), std::exp(x));
        bad += !same(glibm::log(x, kT), std::log(x));
        bad += !same(glibm::log10(x, kT), std::log10(x));
    }
    return bad;
}

// the kernels' scalar DCS on the host (serial loop), for comparison with the oracle
int hostcheck_dcs(int process, const double *K, const double *q, double *out, int64_t n, double A,
                  double I, int32_t Z, double mass) {
    const Params p = make_params(A, I, Z, mass);
    for (int64_t i = 0; i < n; i++) {
        switch (process) {
            case 0: out[i] = dcs_eval<0>(K[i], q[i], p, kT); break;
            case 1: out[i] = dcs_eval<1>(K[i], q[i], p, kT); break;
            case 2: out[i] = dcs_eval<2>(K[i], q[i], p, kT); break;
            case 3: out[i] = dcs_eval<3>(K[i], q[i], p, kT); break;
            default: return 1;
        }
    }
    return 0;
}

// pair production with the kinetic-energy-only part (PairRow: Lorentz factor, zeta) computed
// separately and handed in, the way the table kernels hoist it out of the node loop
int hostcheck_pair_with_row_part(const double *K, const double *q, double *out, int64_t n, double A,
                                 double I, int32_t Z, double mass) {
    const Params p = make_params(A, I, Z, mass);
    for (int64_t i = 0; i < n; i++) {
        PlainOps dv;
        PairRow row;
        row.gamma = pair_gamma(K[i], p, dv);
        row.zeta = pair_zeta(row.gamma, p, kT, dv);
        out[i] = pair_production(K[i], q[i], p, kT, dv, &row);
    }
    return 0;
}

int hostcheck_ionisation_closed_form(int integrand, const double *K, double *out, int64_t n,
                                     double xlow, double A, double I, int32_t Z, double mass) {
    const Params p = make_params(A, I, Z, mass);
    for (int64_t i = 0; i < n; i++) out[i] = ionisation_closed_form(K[i], xlow, integrand, p, kT);
    return 0;
}

// ---- Coulomb / soft scattering: the kernels' per-energy arithmetic, serial on the host -----------
int hostcheck_coulomb_data(const double *K, int64_t n, double A, double I, int32_t Z, double mass,
                           double *fcm, double *screening, double *fspin, double *invlambda) {
    const CoulombParams c = make_coulomb_params(A, I, Z, mass);
    for (int64_t i = 0; i < n; i++) {
        const double k0 = coulomb_frame(K[i], c, fcm[2 * i], fcm[2 * i + 1]);
        fspin[i] = coulomb_spin(k0, c.mass);
        invlambda[i] = coulomb_screening(k0, c, kT, screening + kScreenFactors * i);
    }
    return 0;
}

int hostcheck_coulomb_transport(const double *screening, const double *fspin, const double *mu,
                                int64_t n_mu, int64_t n, double *coef) {
    for (int64_t i = 0; i < n; i++)
        coulomb_transport_coefficients(screening + kScreenFactors * i, fspin[i],
                                       mu[n_mu == 1 ? 0 : i], kT, coef[2 * i], coef[2 * i + 1]);
    return 0;
}

int hostcheck_hard_scattering(const double *G, const double *fcm, const double *screening,
                              const double *invlambda, const double *fspin, int32_t nel,
                              int64_t nkin, double *mu0, double *lb_h) {
    const double max_mu0 = make_coulomb_params(1., 1., 1, 1.).h_max_mu0;
    for (int64_t i = 0; i < nkin; i++) {
        HardView v{G + 2 * i, fcm + 2 * i, screening + kScreenFactors * i, invlambda + i,
                   fspin + i, nel, nkin};
        coulomb_hard_scattering(v, max_mu0, kT, mu0[i], lb_h[i]);
    }
    return 0;
}

int hostcheck_soft_scattering(const double *K, int64_t n, double A, double I, int32_t Z,
                              double mass, double *ms1) {
    const Params p = make_params(A, I, Z, mass);
    const CoulombParams c = make_coulomb_params(A, I, Z, mass);
    for (int64_t i = 0; i < n; i++) {
        double acc = 0.;
        for (int j = 0; j < kSoftNodes; j++) acc += soft_photonuclear_term(j, K[i], p, c, kT);
        ms1[i] = transverse_transport_ionisation(K[i], p, c, kT) + 2. * acc;
    }
    return 0;
}

}  // extern "C"
