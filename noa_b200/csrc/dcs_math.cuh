// Scalar muon DCS arithmetic shared by every kernel in dcs_kernels.cu.
//
// Each function evaluates the same sequence of IEEE-754 double operations as the reference's CPU
// code (cited per function, paths relative to the reference tree), with two changes that do not
// alter a single bit of the result:
//   * sub-expressions that depend only on (element, projectile mass) are evaluated once on the
//     host (dcs_params.hh) with the same operand order and handed over in `Params`;
//   * exp/log/log10 are glibm:: (glibm.cuh), the table-driven routines glibc itself runs.
// Division, exp, log and log10 go through a policy object (folded_ops.cuh): PlainOps is the plain
// `/` and glibm::exp / log / log10; FoldedOps (device only) produces the same values with the
// special-case tests of all operations of one DCS value folded into one flag.  sqrt is CUDA's
// IEEE-correct one.  The translation unit is compiled with
// -fmad=false so that no multiply-add is contracted; the reference's benchmark/test builds
// (-O3, x86-64 baseline) contain no FMA either.
//
// The header also compiles for the host (g++ -ffp-contract=off -mfma) so oracle/hostcheck.cc can
// compare it with the oracle without a GPU.  That host build is a test fixture only; the product
// has no CPU path.
#pragma once

#include "glibm.cuh"
#include "folded_ops.cuh"

namespace noa_b200 {

constexpr double kElectronMass = 0.510998910E-03;   // src/noa/pms/physics.hh:57
constexpr double kAvogadro = 6.02214076E+23;        // src/noa/pms/physics.hh:54
constexpr double kXFraction = 5E-02;                // src/noa/pms/physics.hh:76

// Everything that depends only on the atomic element and the projectile mass.
// Filled by make_params() in dcs_params.hh (host, same libm as the reference's CPU path).
struct Params {
    double A, I, mass, Zd;
    int32_t Z;
    int32_t z_is_one;
    // bremsstrahlung (src/noa/pms/physics.hh:119-133)
    double b_bzn, b_bze, b_dn, b_phie, b_pref, b_hm2, b_c1, b_bzem, b_hm2me;
    // pair production (src/noa/pms/dcs.hh:153-166, 235-241, 255)
    double p_z13, p_thr, p_r, p_r2, p_hr2, p_az13, p_cl, p_cle, p_raz13, p_z15, p_g1, p_g2, p_cz;
    // photonuclear (src/noa/pms/dcs.hh:290, 313-315, 349, 374)
    double n_logA, n_logq0l, n_alow, n_halfA, n_2m2, n_m2, n_qpi;
    // ionisation (src/noa/pms/dcs.hh:422, 431, 435-436)
    double i_wmin, i_kthr, i_m2;
};

// Unroll factor of the in-thread node loops (1 = rolled).  The rolled form keeps the integrand
// instantiated once; see DESIGN.md for the measured choice.
#ifndef NOA_NODE_UNROLL
#define NOA_NODE_UNROLL 1
#endif
#define NOA_PRAGMA_(x) _Pragma(#x)
#define NOA_UNROLL_(n) NOA_PRAGMA_(unroll n)
#define NOA_NODE_LOOP NOA_UNROLL_(NOA_NODE_UNROLL)

// ---- quadrature rules (src/noa/utils/numerics.hh:97-100, 116-121, 137-144) -------------------
#define NOA_GL6_X {0.03376524, 0.16939531, 0.38069041, 0.61930959, 0.83060469, 0.96623476}
#define NOA_GL6_W {0.08566225, 0.18038079, 0.23395697, 0.23395697, 0.18038079, 0.08566225}
#define NOA_GL8_X {0.01985507, 0.10166676, 0.2372338, 0.40828268, 0.59171732, 0.7627662, \
                   0.89833324, 0.98014493}
#define NOA_GL8_W {0.05061427, 0.11119052, 0.15685332, 0.18134189, 0.18134189, 0.15685332, \
                   0.11119052, 0.05061427}
#define NOA_GL9_X {0.0000000000000000, -0.8360311073266358, 0.8360311073266358,               \
                   -0.9681602395076261, 0.9681602395076261, -0.3242534234038089,              \
                   0.3242534234038089, -0.6133714327005904, 0.6133714327005904}
#define NOA_GL9_W {0.3302393550012598, 0.1806481606948574, 0.1806481606948574,                \
                   0.0812743883615744, 0.0812743883615744, 0.3123470770400029,                \
                   0.3123470770400029, 0.2606106964029354, 0.2606106964029354}

// Node/weight lookup: __constant__ memory on the device (uniform or near-uniform index), plain
// static arrays in the host test build.
#if defined(__CUDACC__)
__constant__ double c_gl6_x[6] = NOA_GL6_X;
__constant__ double c_gl6_w[6] = NOA_GL6_W;
__constant__ double c_gl8_x[8] = NOA_GL8_X;
__constant__ double c_gl8_w[8] = NOA_GL8_W;
__constant__ double c_gl9_x[9] = NOA_GL9_X;
__constant__ double c_gl9_w[9] = NOA_GL9_W;
#endif
static const double h_gl6_x[6] = NOA_GL6_X;
static const double h_gl6_w[6] = NOA_GL6_W;
static const double h_gl8_x[8] = NOA_GL8_X;
static const double h_gl8_w[8] = NOA_GL8_W;
static const double h_gl9_x[9] = NOA_GL9_X;
static const double h_gl9_w[9] = NOA_GL9_W;
#if defined(__CUDA_ARCH__)
#define NOA_GL(rule, what, j) c_gl##rule##_##what[j]
#else
#define NOA_GL(rule, what, j) h_gl##rule##_##what[j]
#endif

// Literal constants of bremsstrahlung and ionisation that do not fit an instruction's 32-bit
// immediate: kept in __constant__ memory on the device (one LDCU.64 per use instead of two UMOVs,
// see PhotoConsts below).  Products of literals are folded as the reference's compiler folds them.
struct StreamConsts {
    double me, two_me, sqrte_brems, four_thirds, avogadro, x_fraction, ion_pref, ion_rad;
};
#define NOA_STREAM_CONSTS_INIT                                                                   \
    {0.510998910E-03, 2. * 0.510998910E-03, 1.648721271, 4. / 3., 6.02214076E+23, 5E-02,        \
     1.535336E-05, 1.16141E-03}
#if defined(__CUDACC__)
__constant__ StreamConsts c_stream = NOA_STREAM_CONSTS_INIT;
#endif
static const StreamConsts h_stream = NOA_STREAM_CONSTS_INIT;
#if defined(__CUDA_ARCH__)
#define SK(name) (::noa_b200::c_stream.name)
#else
#define SK(name) (::noa_b200::h_stream.name)
#endif

// ------------------------------------------------------------------------------------------
// Bremsstrahlung -- src/noa/pms/physics.hh:114-153
// ------------------------------------------------------------------------------------------
template <class DV>
NOA_HD double bremsstrahlung(double K, double q, const Params &p, const glibm::Tab &T, DV &dv) {
    const double me = SK(me);
    const double sqrte = SK(sqrte_brems);
    const double E = K + p.mass;
    const typename DV::Den by_E = dv.den(E);
    const double delta_factor = dv.div(p.b_hm2, by_E);
    const double nu = dv.div(q, by_E);
    const double delta = dv.div(delta_factor * nu, 1. - nu);
    double phi_n = dv.log(dv.div(p.b_bzn * (p.mass + delta * p.b_c1),
                                 p.b_dn * (me + delta * sqrte * p.b_bzn)), T);
    if (phi_n < 0.) phi_n = 0.;
    double phi_e = 0.;
    // q < qe_max = E / (1 + hm2 / (me E)) (physics.hh:137, 146): qe_max is needed for nothing else,
    // so the policy may decide the comparison without forming the two quotients
    if (dv.below_ratio(q, E, by_E, p.b_hm2, me, p.b_hm2me)) {
        phi_e = dv.log(dv.div(p.b_bzem,
                              (1. + delta * p.b_phie) * (me + delta * sqrte * p.b_bze)), T);
        if (phi_e < 0.) phi_e = 0.;
    }
    const double s = p.b_pref * (p.Zd * phi_n + phi_e) *
                     (SK(four_thirds) * (dv.rcp(nu) - 1.) + nu);
    return (s < 0.) ? 0. : dv.div_slot(s * 1E+03 * SK(avogadro), p.A, kDenA);
}

// ------------------------------------------------------------------------------------------
// e+e- pair production -- src/noa/pms/dcs.hh:144-258
// ------------------------------------------------------------------------------------------
struct PairKinematics {   // per-(K,q) quantities of src/noa/pms/dcs.hh:159-176
    double tmin, beta, xi_factor, gamma;
};

// Kinematic window and integration bound; false = the DCS is exactly 0 (dcs.hh:151-156,174-175)
// The part of a pair-production value that depends on the kinetic energy alone: the Lorentz factor
// (dcs.hh:171) and the atomic-electron correction zeta (dcs.hh:229-241).  The table build evaluates
// ~10^3 recoil energies per kinetic energy, so it computes these once per row
// (table_rowpar_kernel) and hands them in; same operations on the same operands, same bits.
struct PairRow {
    double gamma, zeta;
};

template <class DV>
NOA_HD double pair_gamma(double K, const Params &p, DV &dv) {
    return 1. + dv.div_slot(K, p.mass, kDenMass);
}

template <class DV>
NOA_HD double pair_zeta(double gamma, const Params &p, const glibm::Tab &T, DV &dv) {
    double zeta;
    if (gamma <= 35.)
        zeta = 0.;
    else {
        zeta = 0.073 * dv.log(dv.div(gamma, 1. + p.p_g1 * gamma * p.p_z13 * p.p_z13), T) -
               0.26;
        if (zeta <= 0.)
            zeta = 0.;
        else
            zeta = dv.div(zeta, 0.058 * dv.log(dv.div(gamma, 1. + p.p_g2 * gamma * p.p_z13),
                                                   T) - 0.14);
    }
    return zeta;
}

template <class DV>
NOA_HD bool pair_setup(double K, double q, const Params &p, const glibm::Tab &T,
                       PairKinematics &k, DV &dv, const PairRow *row = nullptr) {
    if (q <= 4. * kElectronMass) return false;
    if (q >= K + p.p_thr) return false;
    const double nu = dv.div(q, K + p.mass);
    k.beta = dv.div(0.5 * nu * nu, 1. - nu);
    k.xi_factor = p.p_hr2 * k.beta;
    k.gamma = row ? row->gamma : pair_gamma(K, p, dv);
    const double x0 = dv.div(4. * kElectronMass, q);
    const double x1 = dv.div(6., k.gamma * (k.gamma - dv.div_slot(q, p.mass, kDenMass)));
    const double argmin = dv.div(x0 + 2. * (1. - x0) * x1, 1. + (1. - x1) * sqrt(1. - x0));
    if ((argmin >= 1.) || (argmin <= 0.)) return false;
    k.tmin = dv.log(argmin, T);
    return true;
}

// Integrand of the t = ln(1-rho) integral at node t (dcs.hh:179-227)
template <class DV>
NOA_HD double pair_node(double t, double q, const PairKinematics &k, const Params &p,
                        const glibm::Tab &T, DV &dv) {
    const double beta = k.beta;
    const double eps = dv.exp(t * k.tmin, T);
    const double rho = 1. - eps;
    const double rho2 = rho * rho;
    const double rho21 = eps * (2. - eps);
    const double xi = k.xi_factor * rho21;
    const double xi_i = dv.rcp(xi);
    const typename DV::Den by_1xi = dv.den(1. + xi);     // shared by Be and Bmu

    double Be;
    if (xi >= 1E+03)
        Be = 0.5 * xi_i * ((3 - rho2) + 2. * beta * (1. + rho2));
    else
        Be = ((2. + rho2) * (1. + beta) + xi * (3. + rho2)) * dv.log(1. + xi_i, T) +
             dv.div(rho21 - beta, by_1xi) - 3. - rho2;
    const double Ye = dv.div(5. - rho2 + 4. * beta * (1. + rho2),
                             2. * (1. + 3. * beta) * dv.log(3. + xi_i, T) - rho2 -
                                     2. * beta * (2. - rho2));
    const double xe = (1. + xi) * (1. + Ye);
    const double cLi = dv.div(p.p_cl, rho21);
    const double Le = dv.log(dv.div(p.p_az13 * sqrt(xe) * q, q + cLi * xe), T) -
                      0.5 * dv.log(1. + p.p_cle * xe, T);
    double phi_e = Be * Le;
    if (phi_e < 0.) phi_e = 0.;

    double Bmu;
    if (xi <= 1E-03)
        Bmu = 0.5 * xi * (5. - rho2 + beta * (3. + rho2));
    else
        Bmu = ((1. + rho2) * (1. + 1.5 * beta) - xi_i * (1. + 2. * beta) * rho21) *
                      dv.log(1. + xi, T) +
              dv.div(xi * (rho21 - beta), by_1xi) + (1. + 2. * beta) * rho21;
    const double Ymu = dv.div(4. + rho2 + 3. * beta * (1. + rho2),
                              (1. + rho2) * (1.5 + 2. * beta) * dv.log(3. + xi, T) + 1. -
                                      1.5 * rho2);
    const double xmu = (1. + xi) * (1. + Ymu);
    const double Lmu = dv.log(dv.div(p.p_raz13 * q, p.p_z15 * (q + cLi * xmu)), T);
    double phi_mu = Bmu * Lmu;
    if (phi_mu < 0.) phi_mu = 0.;
    return -(phi_e + dv.div_slot(phi_mu, p.p_r2, kDenR2)) * (1. - rho) * k.tmin;
}

// Atomic-electron form factor and normalisation (dcs.hh:229-257); `integral` is the 8-node sum
template <class DV>
NOA_HD double pair_finish(double K, double q, double integral, const PairKinematics &k,
                          const Params &p, const glibm::Tab &T, DV &dv,
                          const PairRow *row = nullptr) {
    const double zeta = row ? row->zeta : pair_zeta(k.gamma, p, T, dv);
    const double E = K + p.mass;
    const double s = dv.div(p.p_cz * (p.Zd + zeta) * (E - q) * integral, q * E);
    return (s < 0.) ? 0. : dv.div_slot(s * 1E+03 * kAvogadro * (p.mass + K), p.A, kDenA);
}

// One thread does all 8 nodes; accumulation order of numerics.hh:84-87 (h = 1, lb = 0)
template <class DV>
NOA_HD double pair_production(double K, double q, const Params &p, const glibm::Tab &T, DV &dv,
                              const PairRow *row = nullptr) {
    PairKinematics k;
    if (!pair_setup(K, q, p, T, k, dv, row)) return 0.;
    double acc = 0.;
    NOA_NODE_LOOP
    for (int j = 0; j < 8; j++)
        acc += pair_node(NOA_GL(8, x, j), q, k, p, T, dv) * NOA_GL(8, w, j);
    return pair_finish(K, q, acc, k, p, T, dv, row);
}

// ------------------------------------------------------------------------------------------
// Photonuclear -- src/noa/pms/dcs.hh:261-405
// ------------------------------------------------------------------------------------------
// Literal constants of the photonuclear model.  On the device they sit in __constant__ memory: an
// FP64 instruction takes an arbitrary 64-bit literal only through a uniform register, which costs
// two UMOVs per use against one LDCU.64 from the constant bank (these kernels are bound by issue
// slots; ~700 of the 9 300 instructions of a photonuclear value were such UMOVs).  Differences of
// literals are folded here exactly as the reference's compiler folds them (IEEE double).
struct PhotoConsts {
    // ALLM97 (dcs.hh:262-287)
    double m02, mP2, mR2, Q02, Lambda2, M2;
    double cP1, cP1m2, cP3, aP1, aP1m2, aP3, bP1, bP2, bP3;
    double cR1, cR2, cR3, aR1, aR2, aR3, bR1, bR2, bR3;
    // DRSS shadowing (dcs.hh:312-318)
    double x_lo, x_hi, s_a, s_b, s_c1, s_c2, s_c3;
    // Whitlow R (dcs.hh:323-331)
    double q2_min, w_c, w_q0, w_a, w_b, w_c2, w_d;
    // d2sigma (dcs.hh:338)
    double cf;
};
#define NOA_PHOTO_CONSTS_INIT                                                                     \
    {0.31985, 49.457, 0.15052, 0.52544, 0.06527, 0.8803505929,                                    \
     0.28067, 0.28067 - 0.22291, 2.1979, -0.0808, -0.0808 - -0.44812, 1.1709, 0.36292, 1.8917,    \
     1.8439, 0.80107, 0.97307, 3.4942, 0.58400, 0.37888, 2.6063, 0.01147, 3.7582, 0.49338,        \
     0.0014, 0.04, 0.069, 0.097, -1.85, 2.45, -2.35,                                              \
     0.3, 0.015625, 0.04, 0.635, 0.5747, 0.09, 0.3534,                                            \
     2.603096E-35}
#if defined(__CUDACC__)
__constant__ PhotoConsts c_photo = NOA_PHOTO_CONSTS_INIT;
#endif
static const PhotoConsts h_photo = NOA_PHOTO_CONSTS_INIT;
#if defined(__CUDA_ARCH__)
#define PK(name) (::noa_b200::c_photo.name)
#else
#define PK(name) (::noa_b200::h_photo.name)
#endif

// ALLM97 F2 (dcs.hh:261-307)
template <class DV>
NOA_HD double f2_allm(double x, double Q2, const Params &p, const glibm::Tab &T, DV &dv) {
    const double W2 = PK(M2) + Q2 * (dv.rcp(x) - 1.0);
    const double t = dv.log(dv.div_slot(dv.log(dv.div_slot(Q2 + PK(Q02), PK(Lambda2), kDenLambda2),
                                               T),
                                        p.n_logq0l, kDenLogQ0L), T);
    const double xP = dv.div(Q2 + PK(mP2), Q2 + PK(mP2) + W2 - PK(M2));
    const double xR = dv.div(Q2 + PK(mR2), Q2 + PK(mR2) + W2 - PK(M2));
    const double lnt = dv.log(t, T);
    const double cP = PK(cP1) + PK(cP1m2) * (dv.rcp(1.0 + dv.exp(PK(cP3) * lnt, T)) - 1.0);
    const double aP = PK(aP1) + PK(aP1m2) * (dv.rcp(1.0 + dv.exp(PK(aP3) * lnt, T)) - 1.0);
    const double bP = PK(bP1) + PK(bP2) * dv.exp(PK(bP3) * lnt, T);
    const double cR = PK(cR1) + PK(cR2) * dv.exp(PK(cR3) * lnt, T);
    const double aR = PK(aR1) + PK(aR2) * dv.exp(PK(aR3) * lnt, T);
    const double bR = PK(bR1) + PK(bR2) * dv.exp(PK(bR3) * lnt, T);

    const double l1x = dv.log(1 - x, T);
    const double F2P = cP * dv.exp(aP * dv.log(xP, T) + bP * l1x, T);
    const double F2R = cR * dv.exp(aR * dv.log(xR, T) + bR * l1x, T);
    return dv.div(Q2, Q2 + PK(m02)) * (F2P + F2R);
}

// DRSS shadowing (dcs.hh:310-319)
template <class DV>
NOA_HD double f2a_drss(double x, double F2p, const Params &p, const glibm::Tab &T, DV &dv) {
    double a = 1.0;
    if (x < PK(x_lo))
        a = p.n_alow;
    else if (x < PK(x_hi))
        a = dv.exp((PK(s_a) * dv.log10(x, T) + PK(s_b)) * p.n_logA, T);
    return (p.n_halfA * a * (2.0 + x * (PK(s_c1) + x * (PK(s_c2) + x * (PK(s_c3) + x)))) * F2p);
}

// Whitlow R (dcs.hh:322-332)
template <class DV>
NOA_HD double r_whitlow(double x, double Q2, const glibm::Tab &T, DV &dv) {
    double q2 = Q2;
    if (Q2 < PK(q2_min)) q2 = PK(q2_min);
    const double theta = 1 + dv.div(dv.div(12.0 * q2, 1.0 + q2) * PK(w_c), PK(w_c) + x * x);
    return (dv.div(PK(w_a), dv.log(dv.div_slot(q2, PK(w_q0), kDenQ004), T)) * theta +
            dv.div(PK(w_b), q2) - dv.div(PK(w_d), PK(w_c2) + q2 * q2));
}

template <class DV>
struct PhotoKinematics {   // per-(K,q) quantities of dcs.hh:372-387 and 340-342
    double centre, width, y;
    typename DV::Den by_Mq, by_E2, by_q;   // denominators shared by the 9 nodes
};

template <class DV>
NOA_HD bool photonuclear_setup(double K, double q, const Params &p, const glibm::Tab &T,
                               PhotoKinematics<DV> &k, DV &dv) {
    if ((q < 1.) || (q < 2E-03 * K)) return false;            // dcs.hh:357-359
    const double M = 0.931494;
    const double mpi = 0.134977;
    const double E = K + p.mass;
    if ((q >= (E - p.mass)) || (q <= p.n_qpi)) return false;  // dcs.hh:374
    const double y = dv.div(q, E);
    const double Q2min = dv.div(p.n_m2 * y * y, 1 - y);
    const double Q2max = 2.0 * M * (q - mpi) - mpi * mpi;
    if ((Q2max < Q2min) | (Q2min < 0)) return false;
    const double lo = dv.log(Q2min, T);
    const double hi = dv.log(Q2max, T);
    k.width = hi - lo;
    k.centre = 0.5 * (hi + lo);
    k.y = y;
    k.by_Mq = dv.den(M * q);
    k.by_E2 = dv.den(E * E);
    k.by_q = dv.den(q);
    return true;
}

// d2sigma/dq dQ2 * Q2 at node t in [-1,1] (dcs.hh:335-355, 397-402)
template <class DV>
NOA_HD double photonuclear_node(double t, const PhotoKinematics<DV> &k, const Params &p,
                                const glibm::Tab &T, DV &dv) {
    const double Q2 = dv.exp(k.centre + 0.5 * k.width * t, T);
    const double y = k.y;
    const double x = dv.div(0.5 * Q2, k.by_Mq);
    const double F2p = f2_allm(x, Q2, p, T, dv);
    const double F2A = f2a_drss(x, F2p, p, T, dv);
    const double R = r_whitlow(x, Q2, T, dv);
    const double dds =
            dv.div(1 - y + dv.div(0.5 * (1 - dv.div(p.n_2m2, Q2)) * (y * y + dv.div(Q2, k.by_E2)),
                                  1 + R),
                   Q2 * Q2) -
            dv.div(0.25, k.by_E2.b * Q2);
    return dv.div(PK(cf) * F2A * dds, k.by_q) * Q2;
}

template <class DV>
NOA_HD double photonuclear(double K, double q, const Params &p, const glibm::Tab &T, DV &dv) {
    PhotoKinematics<DV> k;
    if (!photonuclear_setup(K, q, p, T, k, dv)) return 0.;
    double acc = 0.;
    NOA_NODE_LOOP
    for (int j = 0; j < 9; j++)
        acc += photonuclear_node(NOA_GL(9, x, j), k, p, T, dv) * NOA_GL(9, w, j);
    return (acc < 0.) ? 0.
                      : dv.div_slot(0.5 * acc * k.width * 1E+03 * kAvogadro * (p.mass + K), p.A,
                                    kDenA);
}

// ------------------------------------------------------------------------------------------
// Ionisation -- src/noa/pms/dcs.hh:408-443
// ------------------------------------------------------------------------------------------
template <class DV>
NOA_HD double ionisation(double K, double q, const Params &p, const glibm::Tab &T, DV &dv) {
    const double me = SK(me);
    const double P2 = K * (K + 2. * p.mass);
    const double E = K + p.mass;
    const double Wmax = dv.div(SK(two_me) * P2, p.i_m2 + me * (me + 2. * E));
    if ((Wmax < SK(x_fraction) * K) || (q > Wmax)) return 0.;
    if (q <= p.i_wmin) return 0.;
    const typename DV::Den by_P2 = dv.den(P2);
    const double a0 = dv.div(0.5, by_P2);
    const double a1 = dv.div(-1., Wmax);
    const double a2 = dv.div(E * E, by_P2);
    const typename DV::Den by_q = dv.den(q);
    const double cs = dv.div_slot(SK(ion_pref) * E * p.Zd, p.A, kDenA) *
                      (a0 + dv.div(1., by_q) * (a1 + dv.div(a2, by_q)));
    double Delta = 0.;
    if (K >= p.i_kthr) {
        const double L1 = dv.log(1. + dv.div_slot(2. * q, me, kDenMe), T);
        Delta = SK(ion_rad) * L1 *
                (dv.log(dv.div_slot(4. * E * (E - q), p.i_m2, kDenIm2), T) - L1);
    }
    return cs * (1. + Delta);
}

// Closed-form ionisation integrals (dcs.hh:446-496); integrand 0 = DEL, 1 = CEL
NOA_HD double ionisation_closed_form(double K, double xlow, int integrand, const Params &p,
                                     const glibm::Tab &T) {
    const double me = kElectronMass;
    const double P2 = K * (K + 2. * p.mass);
    const double E = K + p.mass;
    const double Wmax = 2. * me * P2 / (p.i_m2 + me * (me + 2. * E));
    if (Wmax < kXFraction * K) return 0.;
    double Wmin = p.i_wmin;
    const double qlow = K * xlow;
    if (qlow >= Wmin) Wmin = qlow;
    if (Wmax <= Wmin) return 0.;
    const double a0 = 0.5 / P2, a1 = -1. / Wmax, a2 = E * E / P2;
    double term;
    if (integrand == 0)
        term = a0 * (Wmax - Wmin) + a1 * glibm::log(Wmax / Wmin, T) +
               a2 * (1. / Wmin - 1. / Wmax);
    else
        term = 0.5 * a0 * (Wmax * Wmax - Wmin * Wmin) + a1 * (Wmax - Wmin) +
               a2 * glibm::log(Wmax / Wmin, T);
    return 1.535336E-05 * p.Zd / p.A * term;
}

// ---- plain-division forms (host build, table and Coulomb kernels, recompute path) -------------
NOA_HD double bremsstrahlung(double K, double q, const Params &p, const glibm::Tab &T) {
    PlainOps dv;
    return bremsstrahlung(K, q, p, T, dv);
}
NOA_HD bool pair_setup(double K, double q, const Params &p, const glibm::Tab &T,
                       PairKinematics &k) {
    PlainOps dv;
    return pair_setup(K, q, p, T, k, dv);
}
NOA_HD double pair_node(double t, double q, const PairKinematics &k, const Params &p,
                        const glibm::Tab &T) {
    PlainOps dv;
    return pair_node(t, q, k, p, T, dv);
}
NOA_HD double pair_finish(double K, double q, double integral, const PairKinematics &k,
                          const Params &p, const glibm::Tab &T) {
    PlainOps dv;
    return pair_finish(K, q, integral, k, p, T, dv);
}
NOA_HD double pair_production(double K, double q, const Params &p, const glibm::Tab &T) {
    PlainOps dv;
    return pair_production(K, q, p, T, dv);
}
NOA_HD double photonuclear(double K, double q, const Params &p, const glibm::Tab &T) {
    PlainOps dv;
    return photonuclear(K, q, p, T, dv);
}
NOA_HD double ionisation(double K, double q, const Params &p, const glibm::Tab &T) {
    PlainOps dv;
    return ionisation(K, q, p, T, dv);
}

template <int PROCESS, class DV>
NOA_HD double dcs_eval(double K, double q, const Params &p, const glibm::Tab &T, DV &dv) {
    if (PROCESS == 0) return bremsstrahlung(K, q, p, T, dv);
    if (PROCESS == 1) return pair_production(K, q, p, T, dv);
    if (PROCESS == 2) return photonuclear(K, q, p, T, dv);
    return ionisation(K, q, p, T, dv);
}

template <int PROCESS>
NOA_HD double dcs_eval(double K, double q, const Params &p, const glibm::Tab &T) {
    if (PROCESS == 0) return bremsstrahlung(K, q, p, T);
    if (PROCESS == 1) return pair_production(K, q, p, T);
    if (PROCESS == 2) return photonuclear(K, q, p, T);
    return ionisation(K, q, p, T);
}

}  // namespace noa_b200
