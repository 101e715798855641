/*
 * TEST INFRASTRUCTURE ONLY.
 *
 * CPU oracle for the muon DCS hot path: a plain-C restatement of the reference's algorithm.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load it; the product (noa_b200/) never does and fails loudly without its CUDA library.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this file
 *   (1) bit-for-bit against the compiled reference itself (oracle/_ref/libnoa_ref.so, built from
 *       the unmodified headers in /root/reference/src by oracle/Makefile) whenever that is present,
 *   (2) against tests/golden/ fixtures generated from that compiled reference
 *       (tests/golden/make_golden.py), and
 *   (3) against the values the reference prints in docs/pms/muon_dcs_calc.ipynb:325,535,670,741.
 *
 * Build: gcc -O2 -ffp-contract=off -fopenmp -fPIC -shared  (see oracle/Makefile).  Contraction is
 * disabled so the floating-point operation order is exactly the reference's x86-64 baseline build.
 *
 * Each function cites the reference lines it follows (paths relative to /root/reference).
 */
#include <math.h>
#include <stdint.h>

/* src/noa/pms/physics.hh:54-58,76 */
#define N_AVOGADRO 6.02214076E+23
#define M_ELECTRON 0.510998910E-03
#define CUT_FRACTION 5E-02

typedef struct {
    double A; /* g/mol */
    double I; /* GeV   */
    int32_t Z;
} oracle_element; /* src/noa/pms/physics.hh:39-43 */

typedef double (*dcs_fn)(double K, double q, const oracle_element *el, double mass);

/* ------------------------------------------------------------------------------------------
 * Composite Gauss-Legendre rule, src/noa/utils/numerics.hh:72-89.  Strictly serial
 * accumulation  res += f(x) * h * w[j]  with x = lb + h * ((i / order) + x[j]).
 * ------------------------------------------------------------------------------------------ */
typedef double (*integrand_fn)(double t, const void *ctx);

static double composite_gl(double lb, double ub, integrand_fn f, const void *ctx,
                           uint32_t min_points, uint32_t order, const double *node,
                           const double *weight) {
    const uint32_t cells = (min_points + order - 1) / order;
    const double h = (ub - lb) / cells;
    const uint32_t total = cells * order;
    double acc = 0;
    for (uint32_t i = 0; i < total; i++) {
        const uint32_t j = i % order;
        acc += f(lb + h * ((i / order) + node[j]), ctx) * h * weight[j];
    }
    return acc;
}

/* src/noa/utils/numerics.hh:97-100 (8-decimal nodes on [0,1]) */
static const double GL6_X[6] = {0.03376524, 0.16939531, 0.38069041,
                                0.61930959, 0.83060469, 0.96623476};
static const double GL6_W[6] = {0.08566225, 0.18038079, 0.23395697,
                                0.23395697, 0.18038079, 0.08566225};
/* src/noa/utils/numerics.hh:116-121 */
static const double GL8_X[8] = {0.01985507, 0.10166676, 0.2372338,  0.40828268,
                                0.59171732, 0.7627662,  0.89833324, 0.98014493};
static const double GL8_W[8] = {0.05061427, 0.11119052, 0.15685332, 0.18134189,
                                0.18134189, 0.15685332, 0.11119052, 0.05061427};
/* src/noa/utils/numerics.hh:137-144 (nodes on [-1,1], used with bounds (0,1)) */
static const double GL9_X[9] = {0.0000000000000000,  -0.8360311073266358, 0.8360311073266358,
                                -0.9681602395076261, 0.9681602395076261,  -0.3242534234038089,
                                0.3242534234038089,  -0.6133714327005904, 0.6133714327005904};
static const double GL9_W[9] = {0.3302393550012598, 0.1806481606948574, 0.1806481606948574,
                                0.0812743883615744, 0.0812743883615744, 0.3123470770400029,
                                0.3123470770400029, 0.2606106964029354, 0.2606106964029354};

/* ------------------------------------------------------------------------------------------
 * Bremsstrahlung, src/noa/pms/physics.hh:114-153
 * ------------------------------------------------------------------------------------------ */
double oracle_bremsstrahlung(double K, double q, const oracle_element *el, double mass) {
    const int32_t Z = el->Z;
    const double A = el->A;
    const double me = M_ELECTRON;
    const double sqrte = 1.648721271;
    const double phie_factor = mass / (me * me * sqrte);
    const double rem = 5.63588E-13 * me / mass;

    const double BZ_n = (Z == 1) ? 202.4 : 182.7 * pow(Z, -1. / 3.);
    const double BZ_e = (Z == 1) ? 446. : 1429. * pow(Z, -2. / 3.);
    const double D_n = 1.54 * pow(A, 0.27);
    const double E = K + mass;
    const double pref = 7.297182E-07 * rem * rem * Z;

    const double delta_factor = 0.5 * mass * mass / E;
    const double qe_max = E / (1. + 0.5 * mass * mass / (me * E));

    const double nu = q / E;
    const double delta = delta_factor * nu / (1. - nu);
    double phi_n = log(BZ_n * (mass + delta * (D_n * sqrte - 2.)) /
                       (D_n * (me + delta * sqrte * BZ_n)));
    if (phi_n < 0.) phi_n = 0.;
    double phi_e = 0.;
    if (q < qe_max) {
        phi_e = log(BZ_e * mass / ((1. + delta * phie_factor) * (me + delta * sqrte * BZ_e)));
        if (phi_e < 0.) phi_e = 0.;
    }
    const double s = pref * (Z * phi_n + phi_e) * (4. / 3. * (1. / nu - 1.) + nu);
    return (s < 0.) ? 0. : s * 1E+03 * N_AVOGADRO / A;
}

/* ------------------------------------------------------------------------------------------
 * e+e- pair production, src/noa/pms/dcs.hh:144-258
 * ------------------------------------------------------------------------------------------ */
typedef struct {
    double tmin, beta, xi_factor, cL, cLe, AZ13, Z13, r, q;
} pair_ctx;

/* integrand of the t = ln(1-rho) integral, src/noa/pms/dcs.hh:179-227 */
static double pair_node(double t, const void *vctx) {
    const pair_ctx *c = (const pair_ctx *) vctx;
    const double tmin = c->tmin, beta = c->beta, q = c->q, r = c->r;
    const double eps = exp(t * tmin);
    const double rho = 1. - eps;
    const double rho2 = rho * rho;
    const double rho21 = eps * (2. - eps);
    const double xi = c->xi_factor * rho21;
    const double xi_i = 1. / xi;

    double Be;
    if (xi >= 1E+03)
        Be = 0.5 * xi_i * ((3 - rho2) + 2. * beta * (1. + rho2));
    else
        Be = ((2. + rho2) * (1. + beta) + xi * (3. + rho2)) * log(1. + xi_i) +
             (rho21 - beta) / (1. + xi) - 3. - rho2;
    const double Ye = (5. - rho2 + 4. * beta * (1. + rho2)) /
                      (2. * (1. + 3. * beta) * log(3. + xi_i) - rho2 - 2. * beta * (2. - rho2));
    const double xe = (1. + xi) * (1. + Ye);
    const double cLi = c->cL / rho21;
    const double Le = log(c->AZ13 * sqrt(xe) * q / (q + cLi * xe)) - 0.5 * log(1. + c->cLe * xe);
    double phi_e = Be * Le;
    if (phi_e < 0.) phi_e = 0.;

    double Bmu;
    if (xi <= 1E-03)
        Bmu = 0.5 * xi * (5. - rho2 + beta * (3. + rho2));
    else
        Bmu = ((1. + rho2) * (1. + 1.5 * beta) - xi_i * (1. + 2. * beta) * rho21) * log(1. + xi) +
              xi * (rho21 - beta) / (1. + xi) + (1. + 2. * beta) * rho21;
    const double Ymu = (4. + rho2 + 3. * beta * (1. + rho2)) /
                       ((1. + rho2) * (1.5 + 2. * beta) * log(3. + xi) + 1. - 1.5 * rho2);
    const double xmu = (1. + xi) * (1. + Ymu);
    const double Lmu = log(r * c->AZ13 * q / (1.5 * c->Z13 * (q + cLi * xmu)));
    double phi_mu = Bmu * Lmu;
    if (phi_mu < 0.) phi_mu = 0.;
    return -(phi_e + phi_mu / (r * r)) * (1. - rho) * tmin;
}

double oracle_pair_production(double K, double q, const oracle_element *el, double mass) {
    const int32_t Z = el->Z;
    const double A = el->A;
    if (q <= 4. * M_ELECTRON) return 0.;
    const double sqrte = 1.6487212707;
    const double Z13 = pow(Z, 1. / 3.);
    if (q >= K + mass * (1. - 0.75 * sqrte * Z13)) return 0.;

    pair_ctx c;
    const double nu = q / (K + mass);
    c.r = mass / M_ELECTRON;
    c.beta = 0.5 * nu * nu / (1. - nu);
    c.xi_factor = 0.5 * c.r * c.r * c.beta;
    const double A_ = (Z == 1) ? 202.4 : 183.;
    c.AZ13 = A_ / Z13;
    c.Z13 = Z13;
    c.cL = 2. * sqrte * M_ELECTRON * c.AZ13;
    c.cLe = 2.25 * Z13 * Z13 / (c.r * c.r);
    c.q = q;

    const double gamma = 1. + K / mass;
    const double x0 = 4. * M_ELECTRON / q;
    const double x1 = 6. / (gamma * (gamma - q / mass));
    const double argmin = (x0 + 2. * (1. - x0) * x1) / (1. + (1. - x1) * sqrt(1. - x0));
    if ((argmin >= 1.) || (argmin <= 0.)) return 0.;
    c.tmin = log(argmin);

    /* quadrature8<Scalar>(0.f, 1.f, ...), src/noa/pms/dcs.hh:179 */
    const double I = composite_gl(0., 1., pair_node, &c, 1, 8, GL8_X, GL8_W);

    /* atomic-electron form factor, src/noa/pms/dcs.hh:230-251 */
    double zeta;
    if (gamma <= 35.)
        zeta = 0.;
    else {
        double g1, g2;
        if (Z == 1.) {
            g1 = 4.4E-05;
            g2 = 4.8E-05;
        } else {
            g1 = 1.95E-05;
            g2 = 5.30E-05;
        }
        zeta = 0.073 * log(gamma / (1. + g1 * gamma * Z13 * Z13)) - 0.26;
        if (zeta <= 0.)
            zeta = 0.;
        else
            zeta /= 0.058 * log(gamma / (1. + g2 * gamma * Z13)) - 0.14;
    }

    const double E = K + mass;
    const double s = 1.794664E-34 * Z * (Z + zeta) * (E - q) * I / (q * E);
    return (s < 0.) ? 0. : s * 1E+03 * N_AVOGADRO * (mass + K) / A;
}

/* ------------------------------------------------------------------------------------------
 * Photonuclear, src/noa/pms/dcs.hh:261-405
 * ------------------------------------------------------------------------------------------ */
/* ALLM97 proton structure function, src/noa/pms/dcs.hh:261-307 */
static double f2_allm(double x, double Q2) {
    const double m02 = 0.31985, mP2 = 49.457, mR2 = 0.15052, Q02 = 0.52544, Lambda2 = 0.06527;
    const double cP1 = 0.28067, cP2 = 0.22291, cP3 = 2.1979;
    const double aP1 = -0.0808, aP2 = -0.44812, aP3 = 1.1709;
    const double bP1 = 0.36292, bP2 = 1.8917, bP3 = 1.8439;
    const double cR1 = 0.80107, cR2 = 0.97307, cR3 = 3.4942;
    const double aR1 = 0.58400, aR2 = 0.37888, aR3 = 2.6063;
    const double bR1 = 0.01147, bR2 = 3.7582, bR3 = 0.49338;
    const double M2 = 0.8803505929;

    const double W2 = M2 + Q2 * (1.0 / x - 1.0);
    const double t = log(log((Q2 + Q02) / Lambda2) / log(Q02 / Lambda2));
    const double xP = (Q2 + mP2) / (Q2 + mP2 + W2 - M2);
    const double xR = (Q2 + mR2) / (Q2 + mR2 + W2 - M2);
    const double lnt = log(t);
    const double cP = cP1 + (cP1 - cP2) * (1.0 / (1.0 + exp(cP3 * lnt)) - 1.0);
    const double aP = aP1 + (aP1 - aP2) * (1.0 / (1.0 + exp(aP3 * lnt)) - 1.0);
    const double bP = bP1 + bP2 * exp(bP3 * lnt);
    const double cR = cR1 + cR2 * exp(cR3 * lnt);
    const double aR = aR1 + aR2 * exp(aR3 * lnt);
    const double bR = bR1 + bR2 * exp(bR3 * lnt);

    const double F2P = cP * exp(aP * log(xP) + bP * log(1 - x));
    const double F2R = cR * exp(aR * log(xR) + bR * log(1 - x));
    return Q2 / (Q2 + m02) * (F2P + F2R);
}

/* DRSS nuclear shadowing, src/noa/pms/dcs.hh:310-319 */
static double f2a_drss(double x, double F2p, double A) {
    double a = 1.0;
    if (x < 0.0014)
        a = exp(-0.1 * log(A));
    else if (x < 0.04)
        a = exp((0.069 * log10(x) + 0.097) * log(A));
    return (0.5 * A * a * (2.0 + x * (-1.85 + x * (2.45 + x * (-2.35 + x)))) * F2p);
}

/* Whitlow R = sigma_L / sigma_T, src/noa/pms/dcs.hh:322-332 */
static double r_whitlow(double x, double Q2) {
    double q2 = Q2;
    if (Q2 < 0.3) q2 = 0.3;
    const double theta = 1 + 12.0 * q2 / (1.0 + q2) * 0.015625 / (0.015625 + x * x);
    return (0.635 / log(q2 / 0.04) * theta + 0.5747 / q2 - 0.3534 / (0.09 + q2 * q2));
}

/* doubly differential cross-section, src/noa/pms/dcs.hh:335-355 */
static double photonuclear_d2(double A, double mass, double K, double q, double Q2) {
    const double cf = 2.603096E-35;
    const double M = 0.931494;
    const double E = K + mass;
    const double y = q / E;
    const double x = 0.5 * Q2 / (M * q);
    const double F2p = f2_allm(x, Q2);
    const double F2A = f2a_drss(x, F2p, A);
    const double R = r_whitlow(x, Q2);
    const double dds =
            (1 - y + 0.5 * (1 - 2 * mass * mass / Q2) * (y * y + Q2 / (E * E)) / (1 + R)) /
            (Q2 * Q2) -
            0.25 / (E * E * Q2);
    return cf * F2A * dds / q;
}

typedef struct {
    double A, mass, K, q, centre, width;
} photo_ctx;

/* src/noa/pms/dcs.hh:397-402 */
static double photo_node(double t, const void *vctx) {
    const photo_ctx *c = (const photo_ctx *) vctx;
    const double Q2 = exp(c->centre + 0.5 * c->width * t);
    return photonuclear_d2(c->A, c->mass, c->K, c->q, Q2) * Q2;
}

double oracle_photonuclear(double K, double q, const oracle_element *el, double mass) {
    /* dcs_photonuclear_check, src/noa/pms/dcs.hh:357-359 */
    if ((q < 1.) || (q < 2E-03 * K)) return 0.;
    const double A = el->A;
    const double M = 0.931494;
    const double mpi = 0.134977;
    const double E = K + mass;
    if ((q >= (E - mass)) || (q <= (mpi * (1.0 + 0.5 * mpi / M)))) return 0.;

    const double y = q / E;
    const double Q2min = mass * mass * y * y / (1 - y);
    const double Q2max = 2.0 * M * (q - mpi) - mpi * mpi;
    if ((Q2max < Q2min) | (Q2min < 0)) return 0.;

    const double lo = log(Q2min);
    const double hi = log(Q2max);
    photo_ctx c;
    c.A = A;
    c.mass = mass;
    c.K = K;
    c.q = q;
    c.width = hi - lo;
    c.centre = 0.5 * (hi + lo);
    /* quadrature9<Scalar>(0.f, 1.f, ...), src/noa/pms/dcs.hh:394-402 */
    const double ds = composite_gl(0., 1., photo_node, &c, 1, 9, GL9_X, GL9_W);
    return (ds < 0.) ? 0. : 0.5 * ds * c.width * 1E+03 * N_AVOGADRO * (mass + K) / A;
}

/* ------------------------------------------------------------------------------------------
 * Ionisation, src/noa/pms/dcs.hh:408-443
 * ------------------------------------------------------------------------------------------ */
double oracle_ionisation(double K, double q, const oracle_element *el, double mass) {
    const double A = el->A;
    const int32_t Z = el->Z;
    const double P2 = K * (K + 2. * mass);
    const double E = K + mass;
    const double Wmax =
            2. * M_ELECTRON * P2 / (mass * mass + M_ELECTRON * (M_ELECTRON + 2. * E));
    if ((Wmax < CUT_FRACTION * K) || (q > Wmax)) return 0.;
    const double Wmin = 0.62 * el->I;
    if (q <= Wmin) return 0.;

    const double a0 = 0.5 / P2;
    const double a1 = -1. / Wmax;
    const double a2 = E * E / P2;
    const double cs = 1.535336E-05 * E * Z / A * (a0 + 1. / q * (a1 + a2 / q));

    double Delta = 0.;
    const double m1 = mass - M_ELECTRON;
    if (K >= 0.5 * m1 * m1 / M_ELECTRON) {
        const double L1 = log(1. + 2. * q / M_ELECTRON);
        Delta = 1.16141E-03 * L1 * (log(4. * E * (E - q) / (mass * mass)) - L1);
    }
    return cs * (1. + Delta);
}

/* closed-form ionisation integrals, src/noa/pms/dcs.hh:446-496 */
static double ionisation_closed_form(double K, double xlow, const oracle_element *el, double mass,
                                     int integrand) {
    const double P2 = K * (K + 2. * mass);
    const double E = K + mass;
    const double Wmax =
            2. * M_ELECTRON * P2 / (mass * mass + M_ELECTRON * (M_ELECTRON + 2. * E));
    if (Wmax < CUT_FRACTION * K) return 0.;
    double Wmin = 0.62 * el->I;
    const double qlow = K * xlow;
    if (qlow >= Wmin) Wmin = qlow;
    if (Wmax <= Wmin) return 0.;
    const double a0 = 0.5 / P2, a1 = -1. / Wmax, a2 = E * E / P2;
    double term;
    if (integrand == 0) /* src/noa/pms/dcs.hh:446-455 */
        term = a0 * (Wmax - Wmin) + a1 * log(Wmax / Wmin) + a2 * (1. / Wmin - 1. / Wmax);
    else /* src/noa/pms/dcs.hh:457-466 */
        term = 0.5 * a0 * (Wmax * Wmax - Wmin * Wmin) + a1 * (Wmax - Wmin) +
               a2 * log(Wmax / Wmin);
    return 1.535336E-05 * el->Z / el->A * term;
}

/* ------------------------------------------------------------------------------------------
 * Recoil-energy integrals, src/noa/pms/dcs.hh:89-113 and 955-1001
 * ------------------------------------------------------------------------------------------ */
typedef struct {
    dcs_fn f;
    double K, mass;
    const oracle_element *el;
    int integrand; /* 0: dcs*q (del_integrand), 1: dcs*q*q (cel_integrand) */
} recoil_ctx;

static double recoil_node(double t, const void *vctx) {
    const recoil_ctx *c = (const recoil_ctx *) vctx;
    const double q = exp(t);
    const double s = c->f(c->K, q, c->el, c->mass);
    return c->integrand == 0 ? s * q : s * q * q;
}

static dcs_fn pick(int process) {
    switch (process) {
        case 0: return oracle_bremsstrahlung;
        case 1: return oracle_pair_production;
        case 2: return oracle_photonuclear;
        case 3: return oracle_ionisation;
    }
    return 0;
}

double oracle_recoil_integral(int process, int integrand, double K, double xlow,
                              const oracle_element *el, double mass, int32_t min_points) {
    if (process == 3) { /* src/noa/pms/dcs.hh:963-966, 987-990 */
        const double m1 = mass - M_ELECTRON;
        if (K <= 0.5 * m1 * m1 / M_ELECTRON)
            return ionisation_closed_form(K, xlow, el, mass, integrand);
    }
    recoil_ctx c = {pick(process), K, mass, el, integrand};
    return composite_gl(log(K * xlow), log(K), recoil_node, &c, (uint32_t) min_points, 6, GL6_X,
                        GL6_W) /
           (K + mass);
}

/* ------------------------------------------------------------------------------------------
 * Scalar entry points for material_oracle.c (per-material table assembly).
 * oracle_integral_scalar generalises the recoil integral the way PUMAS's compute_dcs_integral is
 * shaped (src/noa/3rdparty/_pumas/pumas.c:10901-10955): an upper bound x_high and a third
 * integrand, mode 2 = dcs*q*q*q (the straggling integrand, pumas.c:10945-10949), evaluated with
 * NOA's own composite 6-point rule and Jacobian (src/noa/pms/dcs.hh:89-105).  Modes 0 / 1 with
 * x_high = 1 are exactly dcs::recoil_integral(f, del|cel_integrand), closed forms included.
 * ------------------------------------------------------------------------------------------ */
double oracle_dcs_scalar(int process, double K, double q, double A, double I, int32_t Z,
                         double mass) {
    const oracle_element el = {A, I, Z};
    return pick(process)(K, q, &el, mass);
}

typedef struct {
    dcs_fn f;
    double K, mass;
    const oracle_element *el;
    int mode;
} recoil_mode_ctx;

static double recoil_mode_node(double t, const void *vctx) {
    const recoil_mode_ctx *c = (const recoil_mode_ctx *) vctx;
    const double q = exp(t);
    double y = c->f(c->K, q, c->el, c->mass) * q;
    if (c->mode > 0) y *= q;
    if (c->mode > 1) y *= q;
    return y;
}

double oracle_integral_scalar(int process, int mode, double K, double xlow, double xhigh, double A,
                              double I, int32_t Z, double mass, int32_t min_points) {
    const oracle_element el = {A, I, Z};
    if (mode <= 1 && xhigh == 1.)
        return oracle_recoil_integral(process, mode, K, xlow, &el, mass, min_points);
    recoil_mode_ctx c = {pick(process), K, mass, &el, mode};
    return composite_gl(log(K * xlow), log(K * xhigh), recoil_mode_node, &c, (uint32_t) min_points,
                        6, GL6_X, GL6_W) /
           (K + mass);
}

/* ------------------------------------------------------------------------------------------
 * Array drivers: dcs::vmap / pvmap (src/noa/pms/dcs.hh:35-75, src/noa/utils/common.hh:146-183)
 * and dcs::vmap_integral (src/noa/pms/dcs.hh:115-130).  `threads` <= 1 is the reference's serial
 * loop; > 1 is its `omp parallel for` (pvmap) or, for integrals, a harness-side loop over energies.
 * ------------------------------------------------------------------------------------------ */
#ifdef _OPENMP
#include <omp.h>
int oracle_max_threads(void) { return omp_get_max_threads(); }
#else
int oracle_max_threads(void) { return 1; }
#endif

int oracle_vmap(int process, int threads, const double *K, const double *q, double *out,
                int64_t n, double A, double I, int32_t Z, double mass) {
    const dcs_fn f = pick(process);
    if (!f) return 1;
    const oracle_element el = {A, I, Z};
    if (threads <= 1) {
        for (int64_t i = 0; i < n; i++) out[i] = f(K[i], q[i], &el, mass);
    } else {
#pragma omp parallel for num_threads(threads)
        for (int64_t i = 0; i < n; i++) out[i] = f(K[i], q[i], &el, mass);
    }
    return 0;
}

int oracle_vmap_integral(int process, int integrand, int threads, const double *K, double *out,
                         int64_t n, double xlow, int32_t min_points, double A, double I, int32_t Z,
                         double mass) {
    if (!pick(process)) return 1;
    const oracle_element el = {A, I, Z};
    if (threads <= 1) {
        for (int64_t i = 0; i < n; i++)
            out[i] = oracle_recoil_integral(process, integrand, K[i], xlow, &el, mass, min_points);
    } else {
#pragma omp parallel for num_threads(threads) schedule(dynamic, 8)
        for (int64_t i = 0; i < n; i++)
            out[i] = oracle_recoil_integral(process, integrand, K[i], xlow, &el, mass, min_points);
    }
    return 0;
}
