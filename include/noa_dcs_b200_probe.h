/*
 * noa_dcs_b200_probe -- measurement kernels (libnoa_dcs_b200_probe.so).  NOT part of the product
 * ABI (include/noa_dcs_b200.h): bench.py takes the FP64-pipe peak its rooflines are quoted
 * against from noa_dcs_fp64_probe, tools/ use the rest for the studies kept under profiles/.
 * Same conventions: device pointers, `void *stream` = cudaStream_t, 0 on success.
 */
#ifndef NOA_DCS_B200_PROBE_H
#define NOA_DCS_B200_PROBE_H

#include <stdint.h>

#include "noa_dcs_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* A dependent-chain-free DFMA loop: executes blocks * threads * iters * 16 DFMA. */
int noa_dcs_fp64_probe(int64_t iters, int32_t blocks, int32_t threads, double *sink, void *stream);

/* Same loop with other operand shapes, to measure what register-file bandwidth allows:
 * mode 0 = the probe above (DFMA, one register-pair source), 1 = DFMA with three distinct
 * register-pair sources, 2 = DFMA with two, 3 = DADD, 4 = DMUL; 5 / 6 / 7 = mode 0 with one / two /
 * three independent 32-bit integer multiply-adds issued per DFMA (do non-FP64 instructions issue in
 * the shadow of the half-rate FP64 dispatch, or do they take issue cycles of their own?);
 * 10 / 11 / 12 / 13 = 1 / 2 / 4 / 8 dependent DFMA chains per thread (still 16 DFMA per thread and
 * iteration): with one warp per scheduler the rate gives the dependent-issue latency;
 * 20-25 = chains whose multiplier is re-read from the constant bank (uniform / per-thread index)
 * or from shared memory before every DFMA. */
int noa_dcs_fp64_probe_mode(int32_t mode, int64_t iters, int32_t blocks, int32_t threads,
                            double *sink, void *stream);

/* Pair-production DCS with one Gauss-Legendre node per lane (8 lanes per pair, shuffle gather,
 * node terms added in node order): the lane mapping north_star names, kept as the measured
 * alternative to the product's one pair per thread.  Same results as
 * noa_dcs_vmap_f64(NOA_DCS_PAIR_PRODUCTION, ...), bit for bit. */
int noa_dcs_probe_pair_lanes_f64(const double *K, const double *q, double *result, int64_t n,
                                 double A, double I, int32_t Z, double mass, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* NOA_DCS_B200_PROBE_H */
