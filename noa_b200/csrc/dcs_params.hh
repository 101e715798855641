// Host-side evaluation of the (element, projectile-mass)-only sub-expressions of the four DCS.
// Same operand order as the reference so every value is bit-identical to what its scalar
// functions recompute per call; pow/log/exp are the host libm's, exactly as on the reference's
// CPU path (src/noa/pms/physics.hh:119-130, src/noa/pms/dcs.hh:153-166, 235-241, 290, 313, 374,
// 422, 435-436).
#pragma once

#include <cmath>

#include "coulomb_math.cuh"
#include "dcs_math.cuh"

namespace noa_b200 {

inline Params make_params(double A, double I, int32_t Z, double mass) {
    Params p{};
    const double me = kElectronMass;
    p.A = A;
    p.I = I;
    p.mass = mass;
    p.Z = Z;
    p.Zd = (double) Z;
    p.z_is_one = (Z == 1);

    {   // bremsstrahlung
        const double sqrte = 1.648721271;
        const double rem = 5.63588E-13 * me / mass;
        p.b_phie = mass / (me * me * sqrte);
        p.b_bzn = (Z == 1) ? 202.4 : 182.7 * std::pow((double) Z, -1. / 3.);
        p.b_bze = (Z == 1) ? 446. : 1429. * std::pow((double) Z, -2. / 3.);
        p.b_dn = 1.54 * std::pow(A, 0.27);
        p.b_pref = 7.297182E-07 * rem * rem * Z;
        p.b_hm2 = 0.5 * mass * mass;
        p.b_hm2me = p.b_hm2 / me;      // only used for an approximate pre-test (folded_ops.cuh)
        p.b_c1 = p.b_dn * sqrte - 2.;
        p.b_bzem = p.b_bze * mass;
    }
    {   // pair production
        const double sqrte = 1.6487212707;
        const double Z13 = std::pow((double) Z, 1. / 3.);
        const double r = mass / me;
        const double A_ = (Z == 1) ? 202.4 : 183.;
        p.p_z13 = Z13;
        p.p_thr = mass * (1. - 0.75 * sqrte * Z13);
        p.p_r = r;
        p.p_r2 = r * r;
        p.p_hr2 = 0.5 * r * r;
        p.p_az13 = A_ / Z13;
        p.p_cl = 2. * sqrte * me * p.p_az13;
        p.p_cle = 2.25 * Z13 * Z13 / (r * r);
        p.p_raz13 = r * p.p_az13;
        p.p_z15 = 1.5 * Z13;
        p.p_g1 = (Z == 1) ? 4.4E-05 : 1.95E-05;
        p.p_g2 = (Z == 1) ? 4.8E-05 : 5.30E-05;
        p.p_cz = 1.794664E-34 * Z;
    }
    {   // photonuclear
        const double M = 0.931494;
        const double mpi = 0.134977;
        const double Q02 = 0.52544, Lambda2 = 0.06527;
        p.n_logA = std::log(A);
        p.n_logq0l = std::log(Q02 / Lambda2);
        p.n_alow = std::exp(-0.1 * std::log(A));
        p.n_halfA = 0.5 * A;
        p.n_2m2 = 2 * mass * mass;
        p.n_m2 = mass * mass;
        p.n_qpi = mpi * (1.0 + 0.5 * mpi / M);
    }
    {   // ionisation
        const double m1 = mass - me;
        p.i_wmin = 0.62 * I;
        p.i_kthr = 0.5 * m1 * m1 / me;
        p.i_m2 = mass * mass;
    }
    return p;
}

// (element, mass)-only sub-expressions of the Coulomb / soft-scattering functions
// (src/noa/pms/dcs.hh:505-520, 548-585, 538, 880-898, 907; src/noa/pms/physics.hh:82-83).
inline CoulombParams make_coulomb_params(double A, double I, int32_t Z, double mass) {
    CoulombParams c{};
    c.mass = mass;
    c.Zd = (double) Z;
    c.series_terms = 10 + Z;
    {   // centre-of-mass frame
        const double Ma = A * 0.931494;
        double M2 = mass + Ma;
        M2 *= M2;
        double rM2 = mass / Ma;
        rM2 *= rM2;
        c.f_Ma = Ma;
        c.f_M2 = M2;
        c.f_2Ma = 2. * Ma;
        c.f_mMa = mass * (mass + Ma);
        c.f_rM2 = rM2;
    }
    {   // screening
        const double third = 1. / 3;
        const double A13 = std::pow(A, third);
        const double R1 = 1.02934 * A13 + 0.435;
        const double R2 = 2.;
        c.s_R1sq = R1 * R1;
        c.s_R2sq = R2 * R2;
        c.s_ps0 = 5.179587126E-12 * std::pow((double) Z, 2. / 3.);
        c.s_wentzel = A * 2.54910918E+08;
    }
    c.h_max_mu0 = 0.5 * (1. - std::cos(1E+00 * M_PI / 180.));
    {   // soft scattering
        const double cs0 = 1.535336E-05 / A;
        c.t_m2 = mass * mass;
        c.t_wmin062 = 0.62 * I;
        c.t_pref = 2. * cs0 * Z;
        c.t_lb = std::log(1E-06);
        c.t_h = (0. - c.t_lb) / kSoftCells;
    }
    return c;
}

}  // namespace noa_b200
