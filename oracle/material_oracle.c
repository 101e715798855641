/*
 * TEST INFRASTRUCTURE ONLY -- never linked or imported by the product path (noa_b200/).
 *
 * Per-material table assembly: a plain-C restatement of what PUMAS does
 * with per-element DCS integrals when it builds the tables of one material, with the DCS and the
 * recoil integrals left as CALLBACKS so that the same algorithm can be driven by
 *   - the reference's own CPU DCS (oracle/ref_shim.cc: noa_ref_dcs_scalar / noa_ref_integral_scalar,
 *     i.e. the unmodified src/noa/pms/dcs.hh lambdas and src/noa/utils/numerics.hh quadrature), or
 *   - the C port (dcs_oracle.c: oracle_dcs_scalar / oracle_integral_scalar).
 * NOA itself has no such driver (its API is one element at a time); the arithmetic below follows
 * src/noa/3rdparty/_pumas/pumas.c (v1.2.1) line by line where cited.  PUMAS's own DCS functions
 * differ from NOA's, so only the assembly ALGORITHM is taken from PUMAS; every
 * number it consumes comes from NOA's DCS.
 *
 *   step 1  per (element, process, energy): CSn (mode 0), cel (mode 1) over x in [cutoff, 1] and,
 *           for ionisation, the straggling integral (mode 2) over x in [1e-6, cutoff]
 *                                                       pumas.c:10768-10808, 10901-10955
 *   step 2  mass-fraction mixing per energy: material cs / cel per process, straggling, and the
 *           normalised cumulative fractions CSf used to pick (element, process)
 *                                                       pumas.c:8054-8111, 8130-8131
 *   step 3  kinetic threshold Kt of the material (first tabulated energy with a non-zero total
 *           cross-section) and regularisation of the cross-section below it
 *                                                       pumas.c:10816-10833
 *   step 4  per (process, element, energy) fractional threshold Xt: doubling from the cutoff until
 *           the DCS is positive, then bisection to 1 % of the cutoff
 *                                                       pumas.c:10839-10880
 */
#include <stdint.h>
#include <stdlib.h>

#define N_DEL_PROCESSES 4 /* pumas.c: bremsstrahlung, pair production, photonuclear, ionisation */

typedef double (*scalar_dcs_fn)(int process, double K, double q, double A, double I, int32_t Z,
                                double mass);
/* mode 0: dcs*q, 1: dcs*q*q, 2: dcs*q*q*q, integrated in ln q over [ln(K xlow), ln(K xhigh)] and
 * divided by (K + mass).  Modes 0 and 1 with xhigh = 1 are dcs::recoil_integral(f, del|cel). */
typedef double (*scalar_integral_fn)(int process, int mode, double K, double xlow, double xhigh,
                                     double A, double I, int32_t Z, double mass,
                                     int32_t min_points);

/* pumas.c:10839-10880 for one (process, element, energy) */
static double threshold_fraction(scalar_dcs_fn dcs_func, int ip, double k, double cutoff, double A,
                                 double I, int32_t Z, double mass) {
    double x = cutoff;
    while ((x < 1.) && (dcs_func(ip, k, k * x, A, I, Z, mass) <= 0.)) x *= 2;
    if (x >= 1.)
        x = 1.;
    else if (x > cutoff) {
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
            dcs = dcs_func(ip, k, k * x0, A, I, Z, mass);
        }
    }
    return x;
}

/*
 * elem        [n_elements][3][4][nK]   CSn, cel, stg per element and process (step 1)
 * cs, cel     [4][nK]                  material restricted cross-section / energy loss per process
 * straggling  [nK]
 * csf         [n_elements][4][nK]      normalised cumulative fractions
 * cs_total    [nK]                     sum over processes, regularised below Kt (rows 1 .. it-1)
 * kt, it      the threshold energy and its row
 * xt          [n_elements][4][nK]
 * Returns 0; 1 if no row has a non-zero cross-section (then kt / xt are not defined: PUMAS would
 * read past the table).
 */
int oracle_material_assembly(int32_t n_elements, const double *A, const double *I,
                             const int32_t *Z, const double *w, double mass, const double *K,
                             int64_t nK, double cutoff, int32_t min_points, scalar_dcs_fn dcs,
                             scalar_integral_fn integral, double *elem, double *cs, double *cel,
                             double *straggling, double *csf, double *cs_total, double *kt,
                             int32_t *it_out, double *xt, int threads) {
    const int64_t n4 = 4 * nK;
    (void) threads;
    /* step 1: pumas.c:10786-10805 */
#pragma omp parallel for num_threads(threads > 0 ? threads : 1) schedule(dynamic, 4)
    for (int64_t row = 0; row < nK; row++) {
        const double kinetic = K[row];
        for (int iel = 0; iel < n_elements; iel++)
            for (int ip = 0; ip < N_DEL_PROCESSES; ip++) {
                double *e = elem + (int64_t) iel * 3 * n4;
                e[0 * n4 + ip * nK + row] = integral(ip, 0, kinetic, cutoff, 1., A[iel], I[iel],
                                                     Z[iel], mass, min_points);
                e[1 * n4 + ip * nK + row] = integral(ip, 1, kinetic, cutoff, 1., A[iel], I[iel],
                                                     Z[iel], mass, min_points);
                /* compute_dcs_integral(physics, 2, element, kinetic, dcs, 0, cutoff, 180) with
                 * "if (xlow <= 0) xlow = 1E-06" (pumas.c:10797-10801, 10905-10907) */
                e[2 * n4 + ip * nK + row] =
                        (ip == 3) ? integral(ip, 2, kinetic, 1E-06, cutoff, A[iel], I[iel], Z[iel],
                                             mass, min_points)
                                  : 0.;
            }
    }

    /* step 2: pumas.c:8054-8111, 8130-8131 */
    for (int64_t row = 0; row < nK; row++) {
        double frct_cel[] = {0., 0., 0., 0.};
        double frct_cs[] = {0., 0., 0., 0.};
        double strag = 0.;
        for (int ic = 0; ic < n_elements; ic++) {
            const double *e = elem + (int64_t) ic * 3 * n4;
            for (int ip = 0; ip < N_DEL_PROCESSES; ip++) {
                const double f = e[0 * n4 + ip * nK + row] * w[ic];
                csf[(int64_t) ic * n4 + ip * nK + row] = f;
                frct_cs[ip] += f;
                frct_cel[ip] += e[1 * n4 + ip * nK + row] * w[ic];
                strag += e[2 * n4 + ip * nK + row] * w[ic];
            }
        }
        double frct_cs_del = 0.;
        for (int ip = 0; ip < N_DEL_PROCESSES; ip++) {
            frct_cs_del += frct_cs[ip];
            cs[ip * nK + row] = frct_cs[ip];
            cel[ip * nK + row] = frct_cel[ip];
        }
        if (frct_cs_del <= 0.) {
            for (int ic = 0; ic < n_elements; ic++)
                for (int ip = 0; ip < N_DEL_PROCESSES; ip++)
                    csf[(int64_t) ic * n4 + ip * nK + row] = 0.;
        } else {
            double sum_tot = 0.;
            for (int ic = 0; ic < n_elements; ic++)
                for (int ip = 0; ip < N_DEL_PROCESSES; ip++)
                    sum_tot += csf[(int64_t) ic * n4 + ip * nK + row];
            double sum = 0.;
            for (int ic = 0; ic < n_elements; ic++)
                for (int ip = 0; ip < N_DEL_PROCESSES; ip++) {
                    sum += csf[(int64_t) ic * n4 + ip * nK + row];
                    csf[(int64_t) ic * n4 + ip * nK + row] = sum / sum_tot;
                }
            /* protect against rounding errors */
            csf[(int64_t) (n_elements - 1) * n4 + (N_DEL_PROCESSES - 1) * nK + row] = 1.;
        }
        straggling[row] = strag;
        cs_total[row] = frct_cs_del;
    }

    /* step 3: pumas.c:10820-10833 (the NI_in update needs PUMAS's dE tables and is not part of
     * this path) */
    int64_t it;
    double cs0 = 0.;
    for (it = 1; it < nK; it++)
        if ((cs0 = cs_total[it]) != 0) break;
    if (it >= nK) {
        *it_out = (int32_t) nK;
        *kt = 0.;
        return 1;
    }
    *kt = K[it];
    *it_out = (int32_t) it;
    for (int64_t row = 1; row < it; row++) cs_total[row] = cs0;

    /* step 4: pumas.c:10839-10880 */
#pragma omp parallel for num_threads(threads > 0 ? threads : 1) schedule(dynamic, 4)
    for (int64_t row = 0; row < nK; row++)
        for (int iel = 0; iel < n_elements; iel++)
            for (int ip = 0; ip < N_DEL_PROCESSES; ip++)
                xt[(int64_t) iel * n4 + ip * nK + row] =
                        (row < it) ? 1.
                                   : threshold_fraction(dcs, ip, K[row], cutoff, A[iel], I[iel],
                                                        Z[iel], mass);
    return 0;
}
