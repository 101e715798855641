// Scalar arithmetic of the Coulomb-scattering and soft-scattering functions of the reference's
// dcs.hh (src/noa/pms/dcs.hh:499-952), shared by the kernels in coulomb_kernels.cuh.
//
// Same contract as dcs_math.cuh: every function evaluates the reference's sequence of IEEE-754
// double operations (operand order kept, no contraction), exp/log are glibm::, and only
// sub-expressions that depend on (element, projectile mass) alone are hoisted to the host
// (make_coulomb_params, same operand order, host libm pow/cos/log as on the reference's CPU path).
// Also compiles for the host so oracle/hostcheck.cc can compare it with the oracle without a GPU.
#pragma once

#include "dcs_math.cuh"

namespace noa_b200 {

constexpr int kScreenFactors = 9;          // NSF, src/noa/pms/physics.hh:86
constexpr double kKinCutoff = 1E-9;        // KIN_CUTOFF, physics.hh:78
constexpr double kEhsPathMax = 1E+9;       // EHS_PATH_MAX, physics.hh:79
constexpr double kEhsOverMsc = 1E-4;       // EHS_OVER_MSC, physics.hh:80

struct CoulombParams {
    double mass, Zd;
    int32_t series_terms;      // 10 + Z (dcs.hh:572)
    int32_t pad_;
    // frame (dcs.hh:505-520)
    double f_Ma, f_M2, f_2Ma, f_mMa, f_rM2;
    // screening (dcs.hh:548-585), Wentzel path (dcs.hh:538)
    double s_R1sq, s_R2sq, s_ps0, s_wentzel;
    // hard scattering (physics.hh:82-83)
    double h_max_mu0;
    // soft scattering: ionisation (dcs.hh:880-898) and the photonuclear quadrature (dcs.hh:907-937)
    double t_m2, t_wmin062, t_pref, t_lb, t_h;
};

// std::min(a, b): b if b < a, else a
NOA_HD double std_min(double a, double b) { return (b < a) ? b : a; }

// ---- dcs::coulomb_data --------------------------------------------------------------------------
// coulomb_frame_parameters, dcs.hh:499-523
NOA_HD double coulomb_frame(double K, const CoulombParams &c, double &f0, double &f1) {
    const double sCM12i = 1. / sqrt(c.f_M2 + c.f_2Ma * K);
    f0 = (K + c.mass + c.f_Ma) * sCM12i;
    double kinetic0 = (K * c.f_Ma + c.f_mMa) * sCM12i - c.mass;
    if (kinetic0 < kKinCutoff) kinetic0 = kKinCutoff;
    const double etot = K + c.mass + c.f_Ma;
    const double betaCM2 = K * (K + 2. * c.mass) / (etot * etot);
    f1 = sqrt(c.f_rM2 * (1. - betaCM2) + betaCM2);
    return kinetic0;
}

// coulomb_spin_factor, dcs.hh:526-529
NOA_HD double coulomb_spin(double K, double mass) {
    const double e = K + mass;
    return K * (e + mass) / (e * e);
}

// coulomb_screening_parameters + coulomb_wentzel_path, dcs.hh:532-598; ps[9]
NOA_HD double coulomb_screening(double K, const CoulombParams &c, const glibm::Tab &T,
                                double *ps) {
    const double p2 = K * (K + 2. * c.mass);
    const double d = 5.8406E-02 / p2;
    ps[1] = d / c.s_R1sq;
    ps[2] = d / c.s_R2sq;
    const double etot = K + c.mass;
    const double ZE = c.Zd * etot;
    const double zeta2 = 5.3251346E-05 * (ZE * ZE) / p2;
    double cK;
    if (zeta2 > 1.) {
        double f = 0.;
        for (int32_t i = 1; i <= c.series_terms; i++)
            f += zeta2 / ((double) i * ((double) (i * i) + zeta2));
        cK = glibm::exp(f, T);
    } else {
        cK = glibm::exp(1. - 1. / (1. + zeta2) +
                        zeta2 * (0.2021 + zeta2 * (0.0083 * zeta2 - 0.0369)), T);
    }
    const double cM = 1. + 3.34 * zeta2;
    double r = K / etot;
    r *= r;
    const double cc = r * cK + (1. - r) * cM;
    ps[0] = c.s_ps0 * cc / p2;
    const double d01 = 1. / (ps[0] - ps[1]);
    const double d02 = 1. / (ps[0] - ps[2]);
    const double d12 = 1. / (ps[1] - ps[2]);
    ps[6] = d01 * d01 * d02 * d02;
    ps[7] = d01 * d01 * d12 * d12;
    ps[8] = d12 * d12 * d02 * d02;
    ps[3] = 2. * ps[6] * (d01 + d02);
    ps[4] = 2. * ps[7] * (d12 - d01);
    ps[5] = -2. * ps[8] * (d12 + d02);
    const double dw = K * (K + 2. * c.mass) / (c.Zd * (K + c.mass));
    return 1. / (c.s_wentzel * ps[0] * (1. + ps[0]) * dw * dw);
}

// ---- dcs::coulomb_transport ---------------------------------------------------------------------
// coulomb_transport_coefficients, dcs.hh:624-672
NOA_HD void coulomb_transport_coefficients(const double *ps, double fspin, double mu,
                                           const glibm::Tab &T, double &g0, double &g1) {
    const double nuclear_screening = (ps[1] < ps[2]) ? ps[1] : ps[2];
    if (mu < 1E-08 * nuclear_screening) {
        const double L = glibm::log(1. + mu / ps[0], T);
        const double r = mu / (mu + ps[0]);
        const double k = ps[0] * (1. + ps[0]);
        g0 = k * (r / ps[0] - fspin * (L - r));
        const double I2 = mu - ps[0] * (r - 2. * L);
        g1 = 2. * k * (L - r - fspin * I2);
        return;
    }
    double I0[3], I1[3], I2[3], J0[3], J1[3], J2[3];
    const double mu2 = 0.5 * mu * mu;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        double r = mu / (mu + ps[i]);
        double L = glibm::log(1. + mu / ps[i], T);
        double mu1 = mu;
        I0[i] = r / ps[i];
        J0[i] = L;
        I1[i] = L - r;
        r *= ps[i];
        L *= ps[i];
        J1[i] = mu1 - L;
        I2[i] = mu1 - 2. * L + r;
        L *= ps[i];
        mu1 *= ps[i];
        J2[i] = mu2 + L - mu1;
    }
    const double k = ps[0] * (1. + ps[0]) * ps[1] * ps[1] * ps[2] * ps[2];
    double a = 0., b = 0.;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        a += ps[3 + i] * (J0[i] - fspin * J1[i]) + ps[6 + i] * (I0[i] - fspin * I1[i]);
        b += ps[3 + i] * (J1[i] - fspin * J2[i]) + ps[6 + i] * (I1[i] - fspin * I2[i]);
    }
    g0 = a * k;
    g1 = b * (2. * k);
}

// ---- dcs::hard_scattering -----------------------------------------------------------------------
// coulomb_restricted_cs, dcs.hh:696-737
NOA_HD double coulomb_restricted_cs(double mu, double fspin, const double *ps,
                                    const glibm::Tab &T) {
    if (mu >= 1.) return 0.;
    const double nuclear_screening = (ps[1] < ps[2]) ? ps[1] : ps[2];
    if (mu < 1E-08 * nuclear_screening) {
        const double L = glibm::log((ps[0] + 1.) / (ps[0] + mu), T);
        const double r = (1. - mu) / ((ps[0] + mu) * (ps[0] + 1.));
        const double k = ps[0] * (1. + ps[0]);
        return k * (r - fspin * (L - ps[0] * r));
    }
    double I0[3], I1[3], J0[3], J1[3];
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const double L = glibm::log((ps[i] + 1.) / (ps[i] + mu), T);
        const double r = (1. - mu) / ((ps[i] + mu) * (ps[i] + 1.));
        I0[i] = r;
        J0[i] = L;
        I1[i] = L - ps[i] * r;
        J1[i] = mu - ps[i] * L;
    }
    const double k = ps[0] * (1. + ps[0]) * ps[1] * ps[1] * ps[2] * ps[2];
    double cs = 0.;
#pragma unroll
    for (int i = 0; i < 3; i++)
        cs += ps[3 + i] * (J0[i] - fspin * J1[i]) + ps[6 + i] * (I0[i] - fspin * I1[i]);
    return k * cs;
}

// Views of one kinetic energy's column in the [nel][nkin][...] arrays of dcs::hard_scattering
struct HardView {
    const double *G, *fcm, *screen, *invlambda, *fspin;
    int32_t nel;
    int64_t nkin;
};

// cutoff_objective, dcs.hh:739-753
NOA_HD double cutoff_objective(double cs_h, double mu, const HardView &v, const glibm::Tab &T) {
    double cs_tot = 0.;
    for (int32_t iel = 0; iel < v.nel; iel++) {
        const int64_t off = (int64_t) iel * v.nkin;
        cs_tot += v.invlambda[off] *
                  coulomb_restricted_cs(mu, v.fspin[off], v.screen + kScreenFactors * off, T);
    }
    return cs_tot - cs_h;
}

// utils::numerics::ridders_root (src/noa/utils/numerics.hh:155-218) on cutoff_objective
NOA_HD bool ridders_cutoff(double xa, double xb, double fa, double fb, double xtol, double rtol,
                           uint32_t max_iter, double cs_h, const HardView &v, const glibm::Tab &T,
                           double &root) {
    if (fa * fb > 0) return false;
    if (fa == 0) {
        root = xa;
        return true;
    }
    if (fb == 0) {
        root = xb;
        return true;
    }
    const double tol = xtol + rtol * std_min(fabs(xa), fabs(xb));
    for (uint32_t i = 0; i < max_iter; i++) {
        double dm = 0.5 * (xb - xa);
        const double xm = xa + dm;
        const double fm = cutoff_objective(cs_h, xm, v, T);
        double sgn = (fb > fa) ? 1. : -1.;
        double dn = sgn * dm * fm / sqrt(fm * fm - fa * fb);
        sgn = (dn > 0.) ? 1. : -1.;
        dn = fabs(dn);
        dm = fabs(dm) - 0.5 * tol;
        if (dn < dm) dm = dn;
        const double xn = xm - sgn * dm;
        const double fn = cutoff_objective(cs_h, xn, v, T);
        if (fn * fm < 0.0) {
            xa = xn;
            fa = fn;
            xb = xm;
            fb = fm;
        } else if (fn * fa < 0.0) {
            xb = xn;
            fb = fn;
        } else {
            xa = xn;
            fa = fn;
        }
        if (fn == 0.0 || fabs(xb - xa) < tol) {
            root = xn;
            return true;
        }
    }
    return false;
}

// coulomb_hard_scattering, dcs.hh:755-840
NOA_HD void coulomb_hard_scattering(const HardView &v, double max_mu0, const glibm::Tab &T,
                                    double &mu0, double &lb_h) {
    double invlb_m = 0., invlb1_m = 0.;
    double s_m_l = 0., s_m_h = 0.;
    for (int32_t iel = 0; iel < v.nel; iel++) {
        const int64_t off = (int64_t) iel * v.nkin;
        const double invlb = v.invlambda[off];
        const double scr = v.screen[kScreenFactors * off];
        invlb_m += invlb * v.G[2 * off];
        s_m_h += scr * invlb;
        s_m_l += invlb / scr;
        const double d = 1. / (v.fcm[2 * off] * (1. + v.fcm[1 + 2 * off]));
        invlb1_m += invlb * v.G[1 + 2 * off] * d * d;
    }
    const double lb_m = 1. / invlb_m;
    lb_h = std_min(kEhsOverMsc / invlb1_m, kEhsPathMax);
    if (!(lb_m < lb_h)) {
        lb_h = lb_m;
        mu0 = 0;
        return;
    }
    const double s_m = (lb_h > 2. * lb_m) ? s_m_h * lb_m : 1. / (s_m_l * lb_m);
    mu0 = s_m * (lb_h - lb_m) / (s_m * lb_h + lb_m);
    const double cs_h = 1. / lb_h;
    double mu_max = std_min(4. * mu0, 1.);
    double mu_min = 0.25 * mu0;
    double fmax = cutoff_objective(cs_h, mu_max, v, T);
    if (fmax > 0.) return;      // dcs.hh:797-802 only re-brackets; mu0 and lb_h keep these values
    double fmin = cutoff_objective(cs_h, mu_min, v, T);
    if (fmin < 0.) {
        mu_max = mu_min;
        fmax = fmin;
        mu_min = 0.;
        fmin = cutoff_objective(cs_h, mu_min, v, T);
    }
    if (mu_min < max_mu0) {
        mu_max = std_min(mu_max, max_mu0);
        double best;
        if (ridders_cutoff(mu_min, mu_max, fmin, fmax, 1E-6 * mu0, 1E-6, 100, cs_h, v, T, best))
            mu0 = best;
    }
    mu0 = std_min(mu0, max_mu0);
    lb_h = cutoff_objective(cs_h, mu0, v, T) + cs_h;
    lb_h = (lb_h <= 1. / kEhsPathMax) ? kEhsPathMax : 1. / lb_h;
}

// ---- dcs::soft_scattering -----------------------------------------------------------------------
// transverse_transport_ionisation, dcs.hh:874-899
NOA_HD double transverse_transport_ionisation(double K, const Params &p, const CoulombParams &c,
                                              const glibm::Tab &T) {
    const double me = kElectronMass;
    const double momentum2 = K * (K + 2. * c.mass);
    const double E = K + c.mass;
    const double Wmax = 2. * me * momentum2 / (c.t_m2 + me * (me + 2. * E));
    const double W0 = 2. * momentum2 / me;
    const double mu_max = Wmax / W0;
    double mu3 = K * kXFraction / W0;
    if (mu3 > mu_max) mu3 = mu_max;
    const double mu2 = c.t_wmin062 / W0;
    if (mu2 >= mu3) return 0.;
    const double a0 = 0.5 * W0 / momentum2;
    const double a1 = -1. / Wmax;
    const double a2 = E * E / (W0 * momentum2);
    return c.t_pref * (0.5 * a0 * (mu3 * mu3 - mu2 * mu2) + a1 * (mu3 - mu2) +
                       a2 * glibm::log(mu3 / mu2, T));
}

// integrand of transverse_transport_photonuclear at node t, dcs.hh:909-936; `dcs` evaluates
// dcs::photonuclear(K, q) for this element (the kernel passes its folded-operations form)
template <class PhotonuclearFn>
NOA_HD double transverse_transport_photonuclear_node(double t, double K, const Params &p,
                                                     const glibm::Tab &T, const PhotonuclearFn &dcs) {
    const double E = K + p.mass;
    const double nu = kXFraction * glibm::exp(t, T);
    const double q = nu * K;
    const double m02 = 0.4;
    const double q2 = q * q;
    const double tmax = 1.876544 * q;
    const double tmin = q2 * p.mass * p.mass / (E * (E - q));
    const double b1 = 1. / (1. - q2 / m02);
    const double c1 = 1. / (1. - m02 / q2);
    double L1 = b1 * glibm::log((q2 + tmax) / (q2 + tmin), T);
    double L2 = c1 * glibm::log((m02 + tmax) / (m02 + tmin), T);
    const double I0 = glibm::log(tmax / tmin, T) - L1 - L2;
    L1 *= q2;
    L2 *= m02;
    const double I1 = L1 + L2;
    L1 *= q2;
    L2 *= m02;
    const double I2 = (tmax - tmin) * (b1 * q2 + c1 * m02) - L1 - L2;
    const double ratio = (I1 * tmax - I2) / ((I0 * tmax - I1) * K * (K + 2. * p.mass));
    return dcs(K, q) * ratio * nu;
}

constexpr int kSoftCells = (100 + 6 - 1) / 6;     // quadrature6(..., 100), numerics.hh:79
constexpr int kSoftNodes = kSoftCells * 6;        // 102

// term i of the composite rule: f(lb + h ((i / 6) + x_j)) h w_j   (numerics.hh:84-87)
template <class PhotonuclearFn>
NOA_HD double soft_photonuclear_term(uint32_t i, double K, const Params &p, const CoulombParams &c,
                                     const glibm::Tab &T, const PhotonuclearFn &dcs) {
    const uint32_t j = i % 6u;
    const double x = c.t_lb + c.t_h * ((i / 6u) + NOA_GL(6, x, j));
    return transverse_transport_photonuclear_node(x, K, p, T, dcs) * c.t_h * NOA_GL(6, w, j);
}

struct PlainPhotonuclear {      // dcs::photonuclear with the plain operations
    const Params &p;
    const glibm::Tab &T;
    NOA_HD double operator()(double K, double q) const { return photonuclear(K, q, p, T); }
};

NOA_HD double soft_photonuclear_term(uint32_t i, double K, const Params &p,
                                     const CoulombParams &c, const glibm::Tab &T) {
    const PlainPhotonuclear dcs{p, T};
    return soft_photonuclear_term(i, K, p, c, T, dcs);
}

}  // namespace noa_b200
