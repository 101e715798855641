/*
 * noa_dcs_b200 -- C ABI of the B200-native muon DCS hot path (libnoa_dcs_b200.so).
 *
 * Plain pointers and sizes, no torch types.  Every device entry point is asynchronous on the
 * CUDA stream it is given (a `cudaStream_t` passed as `void *`; NULL = the legacy default stream,
 * which is what the reference launches on, src/noa/utils/common.cuh:55), never allocates or frees
 * caller memory, performs no hidden synchronisation and is re-entrant.  All arrays are contiguous
 * FP64.  Return value: 0 on success, a positive `cudaError_t`, or a negative NOA_DCS_E* code;
 * noa_dcs_strerror() explains either.
 *
 * Reference interfaces these entry points replace (paths relative to the reference tree):
 *   noa_dcs_vmap_f64           dcs::cuda::vmap_bremsstrahlung       src/noa/pms/dcs.hh:1006-1011,
 *                                                                   src/noa/pms/dcs.cuh:30-41
 *                              and, on the GPU, dcs::vmap(f) / dcs::pvmap(f) for
 *                              f = bremsstrahlung | pair_production | photonuclear | ionisation
 *                                                                   src/noa/pms/dcs.hh:35-75
 *   noa_dcs_vmap_all_f64       the four dcs::vmap(f) calls of docs/pms/muon_dcs.cc:8-27 fused
 *   noa_dcs_vmap_mixture_f64   per-element DCS mixed by mass fraction (a "material"; the
 *                              reference has single elements only, src/noa/pms/physics.hh:39-43)
 *   noa_dcs_vmap_integral_f64  dcs::vmap_integral(dcs::recoil_integral(f, del|cel_integrand))
 *                                                                   src/noa/pms/dcs.hh:89-130,
 *                                                                   955-1001
 *   noa_dcs_table_f64          the eight such columns (4 processes x DEL, CEL) of one element in
 *                              one launch, one DCS evaluation per node feeding both integrands
 *   noa_dcs_table_material_f64 element tables mixed by mass fraction (PUMAS-style material tables)
 *   noa_dcs_table_scatter_f64  the same for a cyclic shard of the energies, finished rows written
 *                              straight into every GPU's table over NVLink (no reference
 *                              counterpart: the reference is single-process)
 *   noa_dcs_table_exchange_f64 the same plus the rank barrier, inside the same launches
 *   noa_dcs_allgather_f64      the NCCL all-gather of table slices SURVEY.md 8(b) proposes
 *   noa_dcs_vmap_host_f64      dcs::map(f) on CPU tensors           src/noa/pms/dcs.hh:50-60
 *                              (host buffers in, host buffers out; copies pipelined with compute)
 *   noa_dcs_vmap_pinned_f64    the same for pinned CPU tensors: one kernel streams the host arrays
 *                              over PCIe itself, no staging copies at all
 */
#ifndef NOA_DCS_B200_H
#define NOA_DCS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NOA_DCS_ABI_VERSION 3

/* process ids: PUMAS / NOA order (NPR = 4, src/noa/pms/physics.hh:86) */
#define NOA_DCS_BREMSSTRAHLUNG 0
#define NOA_DCS_PAIR_PRODUCTION 1
#define NOA_DCS_PHOTONUCLEAR 2
#define NOA_DCS_IONISATION 3
#define NOA_DCS_NPROCESS 4

#define NOA_DCS_MAX_ELEMENTS 8 /* elements per material in noa_dcs_vmap_mixture_f64 */
#define NOA_DCS_MAX_PEERS 16   /* destination tables in noa_dcs_table_scatter_f64 */

/* negative error codes (positive ones are cudaError_t) */
#define NOA_DCS_EINVAL (-1)  /* bad process id / mask / count / null pointer */
#define NOA_DCS_ERANGE (-2)  /* size out of the supported range */
#define NOA_DCS_ENODEV (-3)  /* no CUDA device: there is no CPU fallback */
#define NOA_DCS_ENONCCL (-4) /* noa_dcs_allgather_f64: no NCCL in this process */
#define NOA_DCS_ELIBM (-5)   /* noa_dcs_selfcheck: host libm is not the one the kernels restate */
#define NOA_DCS_ENCCL_BASE (-1000) /* NCCL failure: code = NOA_DCS_ENCCL_BASE - ncclResult_t */

#define NOA_DCS_DEFAULT_EXCHANGE_TIMEOUT_S 30.0

int noa_dcs_abi_version(void);
const char *noa_dcs_strerror(int code);

/* Number of CUDA devices visible; <= 0 means the library cannot run (no CPU path exists). */
int noa_dcs_device_count(void);

/*
 * result[i] = f_process(K[i], q[i], element, mass), i in [0, n).   K, q, result: device pointers.
 * Same semantics as the reference's scalar functions: values outside the kinematic range give
 * exactly 0, NaNs propagate, bremsstrahlung has no range guard (src/noa/pms/physics.hh:135-152).
 */
int noa_dcs_vmap_f64(int process, const double *K, const double *q, double *result, int64_t n,
                     double A, double I, int32_t Z, double mass, void *stream);

/* result[p * n + i] for the four processes p (one pass over K, q). */
int noa_dcs_vmap_all_f64(const double *K, const double *q, double *result, int64_t n, double A,
                         double I, int32_t Z, double mass, void *stream);

/*
 * Material DCS: for every process p in process_mask (bit p set), in increasing p, slot s = rank of
 * p within the mask:
 *   result[s * n + i] = sum_e w[e] * f_p(K[i], q[i], element e, mass)      (e = 0 .. n_elements-1,
 * accumulated in that order starting from 0).   A, I, Z, w are HOST arrays of n_elements entries.
 */
int noa_dcs_vmap_mixture_f64(unsigned process_mask, const double *K, const double *q,
                             double *result, int64_t n, int32_t n_elements, const double *A,
                             const double *I, const int32_t *Z, const double *w, double mass,
                             void *stream);

/*
 * Energy-loss tables.  For every process p in process_mask and every energy K[i]:
 *   del[p * nK + i] = recoil_integral(f_p, del_integrand)(K[i], xlow, element, mass, min_points)
 *   cel[p * nK + i] = recoil_integral(f_p, cel_integrand)(K[i], xlow, element, mass, min_points)
 * i.e. composite 6-point Gauss-Legendre in ln q over [ln(K xlow), ln K] with
 * ceil(min_points / 6) cells, nodes accumulated in the reference's serial order, divided by
 * (K + mass); ionisation uses the closed form for K <= 0.5 (m - me)^2 / me.
 * del / cel are device arrays of 4 * nK doubles; rows of processes not in the mask are set to
 * zero.  Either may be NULL to skip that integrand.  A full build is four kernels (one per process,
 * each with the register budget its integrand wants, chained with programmatic dependent launch
 * so they overlap at their boundaries; work enqueued after the call sees all of them complete).
 */
int noa_dcs_table_f64(unsigned process_mask, const double *K, int64_t nK, double xlow,
                      int32_t min_points, double A, double I, int32_t Z, double mass, double *del,
                      double *cel, void *stream);

/*
 * noa_dcs_table_f64 with a caller-provided workspace for the node terms ("flat" form): every
 * (row, node) of a process is evaluated as one flat index space by a persistent grid (no barrier,
 * no per-row summation phase, balanced to 256 nodes whatever the number of rows), the two terms of
 * a node make one round trip through `workspace`, and one summation kernel adds each row's terms
 * up in node order (numerics.hh:84-87) -- the same additions in the same order, so the same bits
 * as noa_dcs_table_f64.  `workspace` = device array of at least
 * noa_dcs_table_workspace_doubles(nK, min_points) doubles (4 nK + 2 + 8 nK * 6 ceil(min_points / 6):
 * 16 B per node and process), contents irrelevant, free again when the launches have run on
 * `stream`.  A NULL or too small workspace falls back to noa_dcs_table_f64's launches.
 */
int64_t noa_dcs_table_workspace_doubles(int64_t nK, int32_t min_points);
int noa_dcs_table_ws_f64(unsigned process_mask, const double *K, int64_t nK, double xlow,
                         int32_t min_points, double A, double I, int32_t Z, double mass,
                         double *del, double *cel, double *workspace, int64_t workspace_doubles,
                         void *stream);

/*
 * Tables of a material (mass-fraction mix of elements): table[c][p][i] = sum_e t_e[c][p][i] * w[e]
 * with t_e the element tables of noa_dcs_table_f64 (c = 0 DEL, 1 CEL), accumulated in composition
 * order from 0 -- the per-element mixing PUMAS applies to its cross-section and energy-loss
 * tables (src/noa/3rdparty/_pumas/pumas.c:8054-8078); the reference's own API has single elements
 * only.  table: [2][4][nK] device doubles; scratch: n_elements x 8 x nK device doubles (the
 * element tables, left there for the caller).  A, I, Z, w: HOST arrays.  n_elements + 1 launches.
 */
int noa_dcs_table_material_f64(unsigned process_mask, const double *K, int64_t nK, double xlow,
                               int32_t min_points, int32_t n_elements, const double *A,
                               const double *I, const int32_t *Z, const double *w, double mass,
                               double *scratch, double *table, double *workspace,
                               int64_t workspace_doubles, void *stream);

/*
 * Multi-GPU form of noa_dcs_table_f64: builds the rows of the energies K_local[0 .. n_local) and
 * stores each finished value into n_peers destination tables at once -- this GPU's own and every
 * peer's, the latter mapped into this process (CUDA IPC / symmetric memory) and written over
 * NVLink from inside the table kernel, so the build and the "all-gather" are one launch.
 * Local row r is row first_row + r * row_stride of the [4][n_total] destination tables
 * (cyclic partition: first_row = rank, row_stride = world).  peer_del / peer_cel are HOST arrays of
 * n_peers device pointers.  The caller synchronises the ranks afterwards (a barrier over the same
 * stream); remote stores are followed by a system-scope fence.
 */
int noa_dcs_table_scatter_f64(unsigned process_mask, const double *K_local, int64_t n_local,
                              double xlow, int32_t min_points, double A, double I, int32_t Z,
                              double mass, int32_t n_peers, double *const *peer_del,
                              double *const *peer_cel, int64_t n_total, int64_t first_row,
                              int64_t row_stride, void *stream);

/*
 * noa_dcs_table_scatter_f64 with the rank synchronisation inside the same launches: after the last
 * row of the build is stored and fenced, the build writes `epoch` into slot `my_peer` of every
 * peer's flag array (peer_flags[j] = peer j's array of n_peers 32-bit words, mapped here like the
 * tables) and returns only when all n_peers slots of its own array have reached `epoch`, i.e. when
 * every peer's rows have landed in this GPU's table; work enqueued after the call on `stream` sees
 * the complete table.  `epoch` must increase by one per
 * call on all ranks (flags start at 0, first epoch 1); callers alternate between two destination
 * tables so a fast rank never overwrites rows a slow rank is still reading.
 * `sync` = EIGHT zero-initialised 32-bit words on this device {CTA counter, timeout count, 6
 * reserved}.
 * `scratch` = workspace of noa_dcs_table_ws_f64 for the LOCAL rows, REQUIRED
 * (scratch_doubles >= noa_dcs_table_workspace_doubles(n_local, min_points), else
 * NOA_DCS_EINVAL): the exchange build is the flat form -- terms kernels over the local (row, node)
 * space, then the summation kernel, which stores every finished row into all n_peers tables and
 * whose last CTA runs the flag exchange.
 * `multicast_del` / `multicast_cel` (optional, both or neither; NULL otherwise) = NVSwitch
 * multicast addresses that alias the DEL / CEL halves of ALL n_peers destination tables (a CUDA
 * multicast object bound to the same allocations, e.g. torch symmetric memory's `multicast_ptr`):
 * each finished value is then written with ONE `multimem.st` that the switch replicates into every
 * GPU's table, this one's included, instead of n_peers stores.  `multicast_flags` (optional, only
 * with the two above) = the same kind of alias of the flag arrays: the epoch is then published to
 * all peers with one store as well.  A peer that does not arrive within `timeout_seconds` of wall-clock time
 * (<= 0: NOA_DCS_DEFAULT_EXCHANGE_TIMEOUT_S) is FATAL: the timeout count is bumped and the kernel
 * traps, so the stream reports a launch failure instead of handing back a partial table.
 */
int noa_dcs_table_exchange_f64(unsigned process_mask, const double *K_local, int64_t n_local,
                               double xlow, int32_t min_points, double A, double I, int32_t Z,
                               double mass, int32_t n_peers, int32_t my_peer,
                               double *const *peer_del, double *const *peer_cel,
                               uint32_t *const *peer_flags, double *multicast_del,
                               double *multicast_cel, uint32_t *multicast_flags, uint32_t *sync,
                               double *scratch,
                               int64_t scratch_doubles, uint32_t epoch, int64_t n_total,
                               int64_t first_row, int64_t row_stride, double timeout_seconds,
                               void *stream);

/*
 * The exchange SURVEY.md 8(b) sketches for hosts that hold an NCCL communicator instead of
 * peer-mapped tables: in-place ncclAllGather (FP64) of `count_per_rank` doubles per rank, rank r's
 * slice at table + r * count_per_rank, on `stream`.  `nccl_comm` is the caller's ncclComm_t.  NCCL
 * is resolved at first use from what the process has loaded (no link-time dependency);
 * NOA_DCS_ENONCCL if there is none, NOA_DCS_ENCCL_BASE - ncclResult_t on an NCCL error.
 * With the cyclic row partition of the table builders the gathered layout is [rank][2][4][rows per
 * rank]; noa_b200/sharding.py and examples/ show the un-permute.
 */
int noa_dcs_allgather_f64(double *table, int64_t count_per_rank, int32_t rank, void *nccl_comm,
                          void *stream);

/*
 * One column of the above, with the reference's own call shape
 *   dcs::vmap_integral(dcs::recoil_integral(f_process, integrand))(result, K, xlow, element, mass,
 *                                                                  min_points)
 * (src/noa/pms/dcs.hh:115-130).  integrand: 0 = del_integrand (dcs * q), 1 = cel_integrand
 * (dcs * q * q), src/noa/pms/dcs.hh:107-113.  result: n doubles on the device.
 */
int noa_dcs_vmap_integral_f64(int process, int integrand, const double *K, double *result,
                              int64_t n, double xlow, int32_t min_points, double A, double I,
                              int32_t Z, double mass, void *stream);

/*
 * The recoil integral in the shape of PUMAS's compute_dcs_integral
 * (src/noa/3rdparty/_pumas/pumas.c:10901-10955) over NOA's DCS and NOA's composite 6-point rule:
 *   result[i] = 1 / (K[i] + mass) * integral over ln q in [ln(K[i] xlow), ln(K[i] xhigh)] of
 *               dcs(K[i], q) q^(1 + mode)          mode 0 cross-section, 1 energy loss, 2 straggling
 * Modes 0 / 1 with xhigh = 1 are noa_dcs_vmap_integral_f64 (closed forms included); everything
 * else is evaluated by quadrature.  xlow > 0, xhigh > xlow.
 */
int noa_dcs_vmap_integral_mode_f64(int process, int mode, const double *K, double *result,
                                   int64_t n, double xlow, double xhigh, int32_t min_points,
                                   double A, double I, int32_t Z, double mass, void *stream);

/*
 * Per-material table assembly: the steps PUMAS runs on per-element DCS
 * integrals when it tabulates a material, over NOA's DCS (the reference has no such driver; the
 * oracle is oracle/material_oracle.c driving the reference's scalar DCS).  All outputs are device
 * arrays; A, I, Z, w are HOST arrays of n_elements entries.
 *   elem       [n_elements][3][4][nK]  per element and process: CSn (cross-section, x in
 *              [cutoff, 1]), cel (energy loss, same range), stg (straggling, ionisation only, x in
 *              [1e-6, cutoff])                                  pumas.c:10768-10808
 *   cs, cel    [4][nK]  mass-fraction mix per process; straggling [nK]; cs_total [nK] the sum over
 *              processes, regularised below the threshold; csf [n_elements][4][nK] the normalised
 *              cumulative fractions used to pick (element, process)
 *                                                               pumas.c:8054-8111, 8130-8131
 *   kt, it     (one double / one int32 on the device) the first tabulated energy >= row 1 with a
 *              non-zero total cross-section and its row           pumas.c:10820-10833
 *   xt         [n_elements][4][nK]  fractional threshold per process, element and energy: 1 below
 *              row `it`, else doubling from `cutoff` until the DCS is positive followed by a
 *              bisection to 1 % of the cutoff                     pumas.c:10839-10880
 * 4 n_elements + 3 launches (+ the chained table launches), all on `stream`.  `workspace`
 * (optional, NULL allowed): as in noa_dcs_table_ws_f64, sized for nK rows; the element tables are
 * then built in the flat form.
 */
int noa_dcs_material_assembly_f64(const double *K, int64_t nK, double cutoff, int32_t min_points,
                                  int32_t n_elements, const double *A, const double *I,
                                  const int32_t *Z, const double *w, double mass, double *elem,
                                  double *cs, double *cel, double *straggling, double *csf,
                                  double *cs_total, double *kt, int32_t *it, double *xt,
                                  double *workspace, int64_t workspace_doubles, void *stream);

/*
 * Coulomb scattering and soft scattering -- the rest of the reference's dcs.hh surface
 * (SURVEY.md 8(f)); same argument meaning and array layouts as the reference's functors, device
 * pointers, FP64, bit-identical results.
 *
 * noa_dcs_coulomb_data_f64       dcs::coulomb_data         src/noa/pms/dcs.hh:600-622
 *     per energy K[i]: fcm[2 i .. 2 i + 1] (CM Lorentz factors), screening[9 i .. 9 i + 8]
 *     (3 screening factors + 6 pole-reduction factors, NSF = 9), fspin[i], invlambda[i].
 * noa_dcs_coulomb_transport_f64  dcs::coulomb_transport    src/noa/pms/dcs.hh:674-693
 *     coefficients[2 i .. 2 i + 1] from screening, fspin and the angular cutoff mu
 *     (n_mu = 1: one cutoff for all energies, else n_mu = n).
 * noa_dcs_hard_scattering_f64    dcs::hard_scattering      src/noa/pms/dcs.hh:843-872
 *     coefficients, fcm: [nel][nkin][2]; screening: [nel][nkin][9]; invlambda, fspin: [nel][nkin];
 *     writes the cutoff angle mu0[nkin] (Ridders root, src/noa/utils/numerics.hh:155-218) and the
 *     hard-scattering mean free path lb_h[nkin].
 * noa_dcs_soft_scattering_f64    dcs::soft_scattering      src/noa/pms/dcs.hh:940-952
 *     ms1[i] = transverse transport of the soft ionisation + photonuclear interactions at K[i]
 *     (the latter a 102-node quadrature of the photonuclear DCS, dcs.hh:901-938).
 */
int noa_dcs_coulomb_data_f64(const double *K, int64_t n, double A, double I, int32_t Z, double mass,
                             double *fcm, double *screening, double *fspin, double *invlambda,
                             void *stream);
int noa_dcs_coulomb_transport_f64(const double *screening, const double *fspin, const double *mu,
                                  int64_t n_mu, int64_t n, double *coefficients, void *stream);
int noa_dcs_hard_scattering_f64(const double *coefficients, const double *fcm,
                                const double *screening, const double *invlambda,
                                const double *fspin, int32_t nel, int64_t nkin, double *mu0,
                                double *lb_h, void *stream);
int noa_dcs_soft_scattering_f64(const double *K, int64_t n, double A, double I, int32_t Z,
                                double mass, double *ms1, void *stream);

/*
 * Host-buffer form of noa_dcs_vmap_f64 (process 0..3) and noa_dcs_vmap_all_f64 (process = 4,
 * h_result holds 4 * n doubles): copies h_K, h_q to the device in chunks, evaluates, copies the
 * result back, the three stages overlapped on separate streams.  Blocks until h_result is
 * complete.  Pinned (page-locked) host buffers give full PCIe speed; pageable ones still work.
 * The stager owns the device scratch and streams so repeated calls do not allocate.
 */
typedef struct noa_dcs_stager noa_dcs_stager;
int noa_dcs_stager_create(noa_dcs_stager **out, int64_t chunk_pairs, int32_t n_slots);
int noa_dcs_stager_destroy(noa_dcs_stager *stager);
int noa_dcs_vmap_host_f64(noa_dcs_stager *stager, int process, const double *h_K,
                          const double *h_q, double *h_result, int64_t n, double A, double I,
                          int32_t Z, double mass);

/*
 * The same on PINNED host buffers without any copy calls: h_K, h_q, h_result must be page-locked
 * and device-addressable (cudaHostAlloc / cudaHostRegister; torch pin_memory() is).  One persistent
 * kernel reads K / q from host memory over PCIe itself (coalesced loads on the mapped addresses) and
 * stores the results straight into h_result; reads, arithmetic and writes of the resident warps
 * overlap inside the kernel.  Asynchronous on `stream`: h_result is complete when the stream has
 * reached this point (synchronise before reading it).  Returns NOA_DCS_EINVAL for pageable
 * buffers -- use noa_dcs_vmap_host_f64 for those.
 */
int noa_dcs_vmap_pinned_f64(int process, const double *h_K, const double *h_q, double *h_result,
                            int64_t n, double A, double I, int32_t Z, double mass, void *stream);

/*
 * Host libm self-check.  Results are bit-identical to the reference's CPU path only when the
 * host's exp / log / log10 / pow are the ones the device routines restate (glibc >= 2.28, FMA
 * variant): this compares them on 3 x 1024 arguments plus pow known answers, entirely on the
 * host.  Returns 0 when identical, NOA_DCS_ELIBM otherwise (`mismatches`, if not NULL, gets the
 * count).  noa_b200/_lib.py runs it at load time and refuses to continue on a mismatch unless
 * NOA_DCS_ALLOW_LIBM_MISMATCH=1.
 */
int noa_dcs_selfcheck(int64_t *mismatches);

/* Launch geometry the element-wise kernels use on the current device (for reporting). */
int noa_dcs_launch_info(int process, int32_t *blocks, int32_t *threads, int32_t *sm_count);

/* Measurement kernels (FP64 pipe probes, the node-per-lane pair-production variant) are in a
 * separate library: include/noa_dcs_b200_probe.h. */

/* Number of kernels this library has launched since it was loaded (bench.py's gpu_launches). */
int64_t noa_dcs_launch_count(void);

/* Diagnostics of the folded special-case tests (csrc/folded_ops.cuh): how many DCS values on the
 * current device had to be evaluated a second time with the plain operations because a division
 * left the fast-path domain (zero or subnormal-range numerator, non-finite operand) or an exp / log
 * argument left the common case.
 * Synchronises the device; `reset` != 0 clears the counter afterwards. */
int noa_dcs_div_recomputes(int64_t *count, int reset);

#ifdef __cplusplus
}
#endif
#endif /* NOA_DCS_B200_H */
