// Host-side evaluation of the (element, projectile-mass)-only sub-expressions of the four DCS.
// Same operand order as the reference so every value is bit-identical to what its scalar
// functions recompute per call; pow/log/exp are the host libm's, exactly as on the reference's
// CPU path (src/noa/pms/physics.hh:119-130, src/noa/pms/dcs.hh:153-166, 235-241, 290, 313, 374,
// 422, 435-436).
#pragma once

#include <cmath>

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

}  // namespace noa_b200
