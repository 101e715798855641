// Energy-loss tables: dcs::vmap_integral(dcs::recoil_integral(f, del|cel_integrand))
// (src/noa/pms/dcs.hh:89-130, 955-1001; src/noa/utils/numerics.hh:72-108), every (process, energy)
// row of the DEL / CEL tables from one DCS evaluation per node.
//
// Work item = R rows of one process, handled by one CTA:
//   * the nodes of the composite 6-point rule are spread over the 256 threads; every thread writes
//     its two node terms (f q h w, f q q h w) into shared memory;
//   * the terms are then added strictly in node order (numerics.hh:84-87: `res += f(x) h w[j]`), one
//     lane per (row, integrand) chain, the 2R chains in adjacent lanes of warp 0 (one DADD per step
//     for all of them).  A chain is 1002 dependent additions = 4.4 us whatever else happens, so
//     - the two cheap processes (bremsstrahlung, ionisation: a row evaluates in less time than it
//       takes to add up) put R = 4 rows into a pass and quarter that cost per row (config 4,
//       bremsstrahlung rows: 0.257 / 0.204 / 0.181 ms at R = 1 / 2 / 4);
//     - terms that are exactly zero are skipped: x + (+-0) = x for every x a chain can hold (it
//       starts at +0 and +0 + -0 = +0), so only the index range that holds non-zero terms is
//       walked -- rows below a process's kinematic threshold cost no summation at all.
//   * lanes 0 .. 2R-1 finish with acc / (K + mass) and store the value into the local table and,
//     in the multi-GPU forms, into every peer's table over NVLink.
// Items are ordered heavy first (photonuclear, pair production, bremsstrahlung, ionisation; energy
// descending inside a process) so the end of the schedule is made of cheap rows.
//
// The row body is instantiated per process (`table_item<PROCESS>`, out of line): every integrand
// gets its own register allocation instead of one 64-register body that holds all four (636 B of
// spills, 145 MB of local-memory traffic per build in round 1).  table_kernel<MASK>: one CTA per
// item; MASK = 15 is the combined kernel (all processes, register budget of the largest), single-bit
// masks are per-process kernels with their own launch bounds, chained with programmatic dependent
// launch so the tail of one process overlaps the head of the next.
//
// This row-per-CTA form serves the calls that bring no workspace (noa_dcs_table_f64,
// noa_dcs_vmap_integral*_f64, noa_dcs_table_scatter_f64).  Builds with a workspace -- the Python /
// C++ `tables`, the multi-GPU exchange -- run the flat form further down.
#pragma once

#include <climits>

#include "dcs_device.cuh"

namespace noa_b200 {

// measurement-only switches (tools/table_exchange_bench.py builds variants with them; the product
// build leaves both at 0): skip the stores into the peers' tables / do not wait for the peers' flags
#ifndef NOA_XCHG_NO_REMOTE
#define NOA_XCHG_NO_REMOTE 0
#endif
#ifndef NOA_XCHG_NO_WAIT
#define NOA_XCHG_NO_WAIT 0
#endif

constexpr int kTableTerms = 1536;   // node terms staged per pass and integrand: 2 x 12 KB
#ifndef NOA_TABLE_LIGHT_ROWS
#define NOA_TABLE_LIGHT_ROWS 4
#endif
constexpr int kTableMaxRows = 4;

template <int PROCESS>
struct TableCfg {
    static constexpr int R = (PROCESS == 0 || PROCESS == 3) ? NOA_TABLE_LIGHT_ROWS : 1;
};

// Where the finished rows go: `n_peers` destination tables (this GPU's own and, in the multi-GPU
// build, every peer's, mapped over NVLink), each [4][n_total]; local row r is global row
// first_row + r * row_stride.
struct TableOut {
    int32_t n_peers;
    int32_t me;               // index of this GPU among the peers (exchange form only)
    int64_t n_total;
    int64_t first_row;
    int64_t row_stride;
    double *del[NOA_DCS_MAX_PEERS];
    double *cel[NOA_DCS_MAX_PEERS];
    // exchange form (noa_dcs_table_exchange_f64): flags[j] = peer j's array of n_peers epoch words;
    // sync = this GPU's words {CTA counter, timeouts, 6 reserved}; flags[0] == nullptr otherwise
    uint32_t *flags[NOA_DCS_MAX_PEERS];
    uint32_t *sync;
    // NVSwitch multicast addresses of the destination table (exchange form, optional): one
    // multimem.st reaches every GPU's copy, this one's included, instead of n_peers stores
    double *mc_del, *mc_cel;
    uint32_t *mc_flags;       // multicast alias of the flag arrays (all peers' at once), or nullptr
    uint32_t epoch;
    uint32_t total_ctas;      // CTAs of all launches of this build (the last one to finish signals)
    uint64_t timeout_ns;      // how long to wait for a peer before giving up (trap)
};

struct TablePlan {
    int32_t n_slots;          // processes in this launch
    int32_t process[4];       // heaviest first
    int32_t out_row[4];       // output row of process p (p for full tables, 0 for a single column)
    uint32_t items[4];        // items of slot s = ceil(nK / R(process))
    uint32_t cells;           // ceil(min_points / 6)
    double xlow;
    // generalised form (noa_dcs_vmap_integral_mode_f64; PUMAS's compute_dcs_integral shape,
    // pumas.c:10901-10955): upper bound ln(K xhigh) instead of ln K, the second chain's integrand
    // dcs q^second_power (2 = cel_integrand, 3 = straggling), no ionisation closed form
    double xhigh;
    int32_t second_power;
    int32_t quadrature_only;
};

struct TableShared {
    StagedShared staged;
    double2 gl6[6];                         // {node, weight} of the 6-point rule
    double terms[2 * kTableTerms];          // [node][chain], chain = 2 row + integrand
    double row_k[kTableMaxRows], row_lb[kTableMaxRows], row_h[kTableMaxRows];
    double row_gamma, row_zeta;             // pair production (one row per item): PairRow
    int64_t row_at[kTableMaxRows];          // destination index of the row, -1 = no such row
    int32_t row_quad[kTableMaxRows];        // 1 = quadrature, 0 = closed form / no row
    int32_t lo, hi;                         // index range of the non-zero terms of the pass
};

// ---- programmatic dependent launch ------------------------------------------------------------
__device__ __forceinline__ void pdl_release_dependents() {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
// no-op unless the launch carried the programmatic-serialisation attribute; then: the previous
// kernel of the stream has completed and its writes are visible
__device__ __forceinline__ void pdl_wait_prerequisites() {
    asm volatile("griddepcontrol.wait;" ::: "memory");
}

__device__ __forceinline__ uint64_t global_timer_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// Tail of the exchange form.  Every CTA has stored its values (local + peers) and fenced them at
// system scope; the last CTA of the build to get here publishes this GPU's epoch into every peer's
// flag array (release, system scope) and then waits until every peer's epoch has arrived in its own
// -- so when the last kernel of the build retires, the complete table is in this GPU's memory.  No
// host round trip, no separate barrier kernel.  A peer that does not show up within
// `timeout_ns` (wall clock) is fatal: the timeout counter is bumped and the kernel traps, so the
// launch fails instead of handing back a partial table.
__device__ __forceinline__ void table_exchange_tail(const TableOut &out) {
    __syncthreads();
    if (threadIdx.x != 0) return;
    __threadfence_system();
    const uint32_t arrived = atomicAdd(out.sync, 1u);
    if (arrived != out.total_ctas - 1) return;
    // re-armed for the next build on this stream
    out.sync[0] = 0;
    __threadfence_system();
    if (out.mc_flags != nullptr) {
        // one store, replicated by the switch into slot `me` of every peer's flag array
        asm volatile("multimem.st.release.sys.global.b32 [%0], %1;" ::"l"(out.mc_flags + out.me),
                     "r"(out.epoch)
                     : "memory");
    } else {
        for (int j = 0; j < out.n_peers; j++)
            asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(out.flags[j] + out.me),
                         "r"(out.epoch)
                         : "memory");
    }
    const uint64_t t0 = global_timer_ns();
    for (int j = 0; j < (NOA_XCHG_NO_WAIT ? 0 : out.n_peers); j++) {
        const uint32_t *slot = out.flags[out.me] + j;
        for (;;) {
            uint32_t seen;
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(slot) : "memory");
            if ((int32_t) (seen - out.epoch) >= 0) break;
            if (global_timer_ns() - t0 > out.timeout_ns) {
                atomicAdd(out.sync + 1, 1u);
                __threadfence_system();
                __trap();
            }
            __nanosleep(100);
        }
    }
}

// Measured alternative (NOA_TABLE_EVAL_CALL=1): the two quadrature-bound integrands called, not
// inlined, from the row body so that they get the register budget to themselves.  It loses 2.7 %
// (config 4: 4.494 against 4.380 ms, profiles/r02_table_variants.jsonl) -- the caller-saved state
// around the call costs more than the spills it avoids -- so the integrand stays inline.
#ifndef NOA_TABLE_EVAL_CALL
#define NOA_TABLE_EVAL_CALL 0
#endif
template <int PROCESS>
__device__ __noinline__ double dcs_value_call(double K, double q, const Params &p,
                                              const glibm::Tab &T) {
    return dcs_value<PROCESS, true>(K, q, p, T);
}

// The two terms of node i of a row (dcs.hh:107-113; numerics.hh:84-87): f q h w and f q q h w.
template <int PROCESS>
__device__ __forceinline__ void table_node_terms(uint32_t i, double k, double lb, double h,
                                                 const TablePlan &plan, const Params &p,
                                                 const glibm::Tab &T, const double2 *gl6,
                                                 double &td, double &tc) {
    const uint32_t cell = i / 6u;
    const uint32_t j = i - cell * 6u;
    const double2 xw = gl6[j];
    const double x = lb + h * (cell + xw.x);
    const double q = glibm::exp(x, T);
    const double f = (NOA_TABLE_EVAL_CALL && (PROCESS == 1 || PROCESS == 2))
                             ? dcs_value_call<PROCESS>(k, q, p, T)
                             : dcs_value<PROCESS, true>(k, q, p, T);
    const double w = xw.y;
    const double fq = f * q;
    td = fq * h * w;                        // del_integrand, dcs.hh:107-109
    double y = fq * q;                      // cel_integrand, dcs.hh:111-113
    if (plan.second_power == 3) y *= q;     // straggling, pumas.c:10945-10949
    tc = y * h * w;
}

// pair production with the row's {gamma, zeta} handed in
__device__ __forceinline__ void table_node_terms_pair(uint32_t i, double k, double lb, double h,
                                                      const PairRow &pre, const TablePlan &plan,
                                                      const Params &p, const glibm::Tab &T,
                                                      const double2 *gl6, double &td, double &tc) {
    const uint32_t cell = i / 6u;
    const double2 xw = gl6[i - cell * 6u];
    const double q = glibm::exp(lb + h * (cell + xw.x), T);
    const double fq = dcs_value_pair_row(k, q, pre, p, T) * q;
    td = fq * h * xw.y;
    double y = fq * q;
    if (plan.second_power == 3) y *= q;
    tc = y * h * xw.y;
}

// One item: rows nK-1 - (item R + r), r = 0 .. R-1, of `PROCESS` (descending energy).  Called by
// every thread of the CTA with the shared buffers free to overwrite.
template <int PROCESS>
__device__ __noinline__ void table_item(uint32_t item, int slot, const double *__restrict__ K,
                                        int64_t nK, const TableOut &out, const TablePlan &plan,
                                        const Params &p, const glibm::Tab &T, TableShared &s) {
    constexpr int R = TableCfg<PROCESS>::R;
    constexpr int CH = 2 * R;
    static_assert(R <= kTableMaxRows && kTableTerms % R == 0, "rows per pass");
    const int tid = threadIdx.x;

    if (tid < R) {
        const int64_t row = nK - 1 - ((int64_t) item * R + tid);
        int quad = 0;
        int64_t at = -1;
        if (row >= 0) {
            const double k = K[row];
            at = (int64_t) plan.out_row[slot] * out.n_total + out.first_row + row * out.row_stride;
            s.row_k[tid] = k;
            // dcs.hh:963-966, 987-990
            quad = !(PROCESS == 3 && k <= p.i_kthr && !plan.quadrature_only);
            if (quad) {
                const double lb = glibm::log(k * plan.xlow, T);
                const double ub = (plan.xhigh == 1.) ? glibm::log(k, T)
                                                     : glibm::log(k * plan.xhigh, T);
                s.row_lb[tid] = lb;
                s.row_h[tid] = (ub - lb) / plan.cells;
                if (PROCESS == 1) {     // R = 1: what the row's nodes share (dcs_math.cuh: PairRow)
                    PlainOps dv;
                    const double gamma = pair_gamma(k, p, dv);
                    s.row_gamma = gamma;
                    s.row_zeta = pair_zeta(gamma, p, T, dv);
                }
            }
        }
        s.row_at[tid] = at;
        s.row_quad[tid] = quad;
    }
    if (tid == 0) {
        s.lo = INT_MAX;
        s.hi = -1;
    }
    __syncthreads();

    bool any_quad = false;
#pragma unroll
    for (int r = 0; r < R; r++) any_quad |= (s.row_quad[r] != 0);

    // chain owned by this lane (tid < CH): row tid / 2, integrand tid % 2 (0 DEL, 1 CEL)
    const int my_row = (tid < CH) ? (tid >> 1) : 0;
    const bool my_quad = (tid < CH) && s.row_quad[my_row] != 0;
    double acc = 0.;

    if (any_quad) {
        const uint32_t total = plan.cells * 6u;
        constexpr uint32_t per_pass = kTableTerms / R;
        for (uint32_t base = 0; base < total; base += per_pass) {
            const uint32_t count = min(per_pass, total - base);
            int lo = INT_MAX, hi = -1;
            for (uint32_t e = tid; e < R * count; e += kThreads) {
                uint32_t r = 0, il = e;
                if (R > 1) {
#pragma unroll
                    for (int t = 1; t < R; t++)
                        if (il >= count) {
                            il -= count;
                            r++;
                        }
                }
                if (!s.row_quad[r]) continue;
                double td, tc;
                if (PROCESS == 1) {
                    PairRow pre;
                    pre.gamma = s.row_gamma;
                    pre.zeta = s.row_zeta;
                    table_node_terms_pair(base + il, s.row_k[0], s.row_lb[0], s.row_h[0], pre, plan,
                                          p, T, s.gl6, td, tc);
                } else {
                    table_node_terms<PROCESS>(base + il, s.row_k[r], s.row_lb[r], s.row_h[r], plan,
                                              p, T, s.gl6, td, tc);
                }
                s.terms[il * CH + 2 * r] = td;
                s.terms[il * CH + 2 * r + 1] = tc;
                if (td != 0. || tc != 0.) {             // NaN counts as non-zero
                    lo = min(lo, (int) il);
                    hi = max(hi, (int) il);
                }
            }
            lo = __reduce_min_sync(0xffffffffu, lo);
            hi = __reduce_max_sync(0xffffffffu, hi);
            if ((tid & 31) == 0 && hi >= 0) {
                atomicMin(&s.lo, lo);
                atomicMax(&s.hi, hi);
            }
            __syncthreads();
            // res += term, strictly in node order (numerics.hh:84-87), one lane per chain
            if (tid < 32) {
                const int first = s.lo, last = s.hi;
                if (my_quad)
                    for (int il = first; il <= last; il++) acc += s.terms[il * CH + tid];
                __syncwarp();
                if (tid == 0) {
                    s.lo = INT_MAX;
                    s.hi = -1;
                }
            }
            __syncthreads();
        }
    }

    if (tid < CH && s.row_at[my_row] >= 0) {
        const int integrand = tid & 1;
        double *const *dst = integrand ? out.cel : out.del;
        if (dst[0] != nullptr) {
            const double k = s.row_k[my_row];
            const double v = my_quad ? acc / (k + p.mass)
                                     : ionisation_closed_form(k, plan.xlow, integrand, p, T);
            const int64_t at = s.row_at[my_row];
            // one store per destination: the local table and, over NVLink, each peer's copy
            for (int j = 0; j < out.n_peers; j++) dst[j][at] = v;
        }
    }
}

template <unsigned MASK>
struct TableMinBlocks {
    static constexpr int value = (MASK == 9u)   ? NOA_MINB_TABLE_LIGHT
                                 : (MASK == 2u) ? NOA_MINB_TABLE_PAIR
                                 : (MASK == 4u) ? NOA_MINB_TABLE_PHOTO
                                 : (MASK == 15u) ? NOA_MINB_TABLE_ALL
                                                 : NOA_MINB_TABLE_LIGHT;
};

template <unsigned MASK>
__device__ __forceinline__ void table_dispatch(uint32_t b, const double *__restrict__ K, int64_t nK,
                                               const TableOut &out, const TablePlan &plan,
                                               const Params &p, const glibm::Tab &T, TableShared &s) {
    int slot = 0;
    while (slot < plan.n_slots - 1 && b >= plan.items[slot]) b -= plan.items[slot++];
    switch (plan.process[slot]) {      // CTA-uniform
        case 0:
            if (MASK & 1u) table_item<0>(b, slot, K, nK, out, plan, p, T, s);
            break;
        case 1:
            if (MASK & 2u) {
                table_item<1>(b, slot, K, nK, out, plan, p, T, s);
            }
            break;
        case 2:
            if (MASK & 4u) {
                table_item<2>(b, slot, K, nK, out, plan, p, T, s);
            }
            break;
        default:
            if (MASK & 8u) table_item<3>(b, slot, K, nK, out, plan, p, T, s);
            break;
    }
}

template <unsigned MASK>
__global__ void __launch_bounds__(kThreads, TableMinBlocks<MASK>::value)
table_kernel(const double *__restrict__ K, int64_t nK, const __grid_constant__ TableOut out,
             const __grid_constant__ TablePlan plan, const __grid_constant__ Params p) {
    __shared__ TableShared s;
    if (threadIdx.x < 6) {
        s.gl6[threadIdx.x] = make_double2(c_gl6_x[threadIdx.x], c_gl6_w[threadIdx.x]);
    }
    const glibm::Tab T = stage_all(s.staged, p);
    // the next launch of the build (another process: disjoint rows) may start filling SMs as
    // soon as every CTA of this one is resident or done
    pdl_release_dependents();
    table_dispatch<MASK>(blockIdx.x, K, nK, out, plan, p, T, s);
    // scatter form: the writer lanes fence their own remote stores
    if (out.n_peers > 1 && threadIdx.x < 32) __threadfence_system();
    // completion order along the chain: this kernel does not retire before its predecessor has
    if (threadIdx.x == 0) pdl_wait_prerequisites();
}

// the two terms of a node whose recoil energy q = exp(x) is already there
template <int PROCESS>
__device__ __forceinline__ void table_terms_at(double k, double q, double h, double w,
                                               const TablePlan &plan, const Params &p,
                                               const glibm::Tab &T, double &td, double &tc) {
    const double fq = dcs_value<PROCESS, true>(k, q, p, T) * q;
    td = fq * h * w;
    double y = fq * q;
    if (plan.second_power == 3) y *= q;
    tc = y * h * w;
}

// bremsstrahlung (td, tc) and, if `ion`, ionisation (ud, uc) terms of one node
__device__ __forceinline__ void table_node_terms_light(uint32_t i, double k, double lb, double h,
                                                       bool ion, const TablePlan &plan,
                                                       const Params &p, const glibm::Tab &T,
                                                       const double2 *gl6, double &td, double &tc,
                                                       double &ud, double &uc) {
    const uint32_t cell = i / 6u;
    const uint32_t j = i - cell * 6u;
    const double2 xw = gl6[j];
    const double x = lb + h * (cell + xw.x);
    const double q = glibm::exp(x, T);
    const double w = xw.y;
    {
        const double fq = dcs_value<0, true>(k, q, p, T) * q;
        td = fq * h * w;
        double y = fq * q;
        if (plan.second_power == 3) y *= q;
        tc = y * h * w;
    }
    if (ion) {
        const double fq = dcs_value<3, true>(k, q, p, T) * q;
        ud = fq * h * w;
        double y = fq * q;
        if (plan.second_power == 3) y *= q;
        uc = y * h * w;
    }
}

// ---- flat form (builds that come with a workspace for the node terms) ----------------------------
// With 16 B of workspace per node the rows need no CTA-per-row structure at all:
//   table_rowpar_kernel      {ln(K xlow), h} of every row; zeroes the queue words
//   table_terms_kernel<P>    every (row, node) of process P as one flat index space cut into
//                            units of 32 consecutive nodes of one row (8 x 32 for the two cheap
//                            processes) that WARPS pop from a device-side queue, rows in
//                            descending energy: no barrier, no summation phase, warps never wait
//                            for each other, and the schedule is balanced to a unit whatever the
//                            number of rows.  (Static deals of 256-node chunks to a persistent
//                            grid were built first and lost: a chunk of out-of-range nodes is
//                            free, and 4 chunks per row lock onto a grid of 4 x 111 CTAs.)
//   table_sum_kernel         one warp per (process, row): the row's terms through a cp.async
//                            ring in shared memory, res += term in node order by lanes 0 / 1,
//                            / (K + mass) -- or the closed form of an ionisation row -- stored to
//                            the local table and every peer's; in the exchange form its last CTA
//                            runs the rank barrier.
// The launches of a build are chained with programmatic dependent launch.  The terms make one
// round trip through L2 / HBM (160 MB each way per process on config 4, hidden under the
// FP64-bound evaluation); what is bought is the 4.4 us of two-lane summation per row during which
// the other 254 threads of the row's CTA idled (barrier stalls: 14 % of the pair kernel's warp
// cycles, 8 % of photonuclear's, profiles/r02_ncu_full_s3.md).  Every variant tried on the way is
// in profiles/r02_flat_table_study.md.
struct FlatPlan {
    uint32_t cells;
    int32_t second_power;
    int32_t quadrature_only;
    int32_t first_launch;           // terms kernel: 1 = the build's first (waits for rowpar)
    double xlow, xhigh;
};

// queue words of the flat form live in the workspace right behind rowpar; table_rowpar_kernel
// zeroes them
__global__ void table_rowpar_kernel(const double *__restrict__ K, int64_t nK,
                                    const __grid_constant__ FlatPlan plan,
                                    double2 *__restrict__ rowpar, uint32_t *__restrict__ queues,
                                    double2 *__restrict__ pairrow,
                                    const __grid_constant__ Params p) {
    __shared__ glibm::Tables s_tables;
    pdl_release_dependents();       // the first terms kernel may get resident and stage its tables
    const glibm::Tab T = stage_tables(s_tables);
    const int64_t row = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (row < 4) queues[row] = 0;
    if (row >= nK) return;
    const double k = K[row];
    const double lb = glibm::log(k * plan.xlow, T);
    const double ub = (plan.xhigh == 1.) ? glibm::log(k, T) : glibm::log(k * plan.xhigh, T);
    rowpar[row] = make_double2(lb, (ub - lb) / plan.cells);
    // what pair production needs of the kinetic energy alone (dcs_math.cuh: PairRow)
    PlainOps dv;
    const double gamma = pair_gamma(k, p, dv);
    pairrow[row] = make_double2(gamma, pair_zeta(gamma, p, T, dv));
}

// Work unit = kFlatUnit x 32 consecutive nodes of one row, popped by a WARP from a device-side
// queue (lane 0's atomic, one unit ahead so its latency is hidden): heaviest rows first, no
// barrier anywhere, no static assignment whose period could lock onto the rows'.
// PROCESS = 4: bremsstrahlung and ionisation of a node in one pass (one exp, one launch less);
// `terms` then holds the bremsstrahlung terms and `terms_b` the ionisation ones.
// PROCESS = 6: photonuclear with bremsstrahlung (`terms_b`) and ionisation (`terms_c`) riding on
// its nodes: the two cheap processes are 4.5 % of the work but, launched on their own, one partial
// wave of pure latency on a rank with few rows (45 us of a 565 us build on a 1/8 share); inside
// the units of the heaviest kernel they cost their arithmetic and nothing else.
template <int PROCESS>
struct FlatCfg {
    static constexpr uint32_t unit = (PROCESS == 1 || PROCESS == 2 || PROCESS == 6) ? 1u : 8u;
    static constexpr unsigned mask = (PROCESS == 4) ? 9u : (PROCESS == 6) ? 4u : (1u << PROCESS);
};

#ifndef NOA_FLAT_THREADS
#define NOA_FLAT_THREADS 256
#endif
constexpr int kFlatThreads = NOA_FLAT_THREADS;

// the units of one queue: PROCESS 0..3, or 4 = bremsstrahlung + ionisation fused
template <int PROCESS>
__device__ __forceinline__ void flat_run_units(const double *__restrict__ K, int64_t nK,
                                               const double2 *__restrict__ rowpar,
                                               double2 *__restrict__ terms,
                                               double2 *__restrict__ terms_b,
                                               double2 *__restrict__ terms_c,
                                               const double2 *__restrict__ pairrow,
                                               uint32_t *__restrict__ queue, const FlatPlan &fp,
                                               const Params &p, const glibm::Tab &T,
                                               const double2 *gl6) {
    TablePlan plan{};
    plan.second_power = fp.second_power;
    const uint32_t nodes = fp.cells * 6u;
    constexpr uint32_t span = 32u * FlatCfg<PROCESS>::unit;
    const uint32_t per_row = (nodes + span - 1) / span;
    const uint64_t units = (uint64_t) nK * per_row;
    const uint32_t lane = threadIdx.x & 31u;
    uint32_t next = 0;
    if (lane == 0) next = atomicAdd(queue, 1u);
    for (;;) {
        const uint32_t u = __shfl_sync(0xffffffffu, next, 0);
        if (u >= units) break;
        if (lane == 0) next = atomicAdd(queue, 1u);
        const uint32_t rr = u / per_row;
        const int64_t row = nK - 1 - (int64_t) rr;      // descending energy: the schedule ends cheap
        const double k = K[row];
        // dcs.hh:963-966, 987-990: closed form, nothing to integrate
        const bool ion_closed = k <= p.i_kthr && !fp.quadrature_only;
        if (PROCESS == 3 && ion_closed) continue;
        const double2 lbh = __ldcg(rowpar + row);
        double2 *row_terms = terms + row * nodes;
        const uint32_t first = (u - rr * per_row) * span + lane;
#pragma unroll 1
        for (uint32_t i = first; i < min(nodes, first + span); i += 32u) {
            double td, tc;
            if (PROCESS == 6) {
                // photonuclear, bremsstrahlung and ionisation of the node from one exp(x)
                const uint32_t cell = i / 6u;
                const double2 xw = gl6[i - cell * 6u];
                const double q = glibm::exp(lbh.x + lbh.y * (cell + xw.x), T);
                double ud, uc;
                table_terms_at<0>(k, q, lbh.y, xw.y, plan, p, T, ud, uc);
                terms_b[row * nodes + i] = make_double2(ud, uc);
                if (!ion_closed) {
                    table_terms_at<3>(k, q, lbh.y, xw.y, plan, p, T, ud, uc);
                    terms_c[row * nodes + i] = make_double2(ud, uc);
                }
                table_terms_at<2>(k, q, lbh.y, xw.y, plan, p, T, td, tc);
            } else if (PROCESS == 1) {
                // gamma and zeta of the row come from table_rowpar_kernel
                const double2 gz = __ldcg(pairrow + row);
                PairRow pre;
                pre.gamma = gz.x;
                pre.zeta = gz.y;
                table_node_terms_pair(i, k, lbh.x, lbh.y, pre, plan, p, T, gl6, td, tc);
            } else if (PROCESS == 4) {
                double ud = 0., uc = 0.;
                table_node_terms_light(i, k, lbh.x, lbh.y, !ion_closed, plan, p, T, gl6, td, tc, ud,
                                       uc);
                if (!ion_closed) terms_b[row * nodes + i] = make_double2(ud, uc);
            } else {
                table_node_terms<(PROCESS >= 4 ? 0 : PROCESS)>(i, k, lbh.x, lbh.y, plan, p, T, gl6,
                                                                td, tc);
            }
            row_terms[i] = make_double2(td, tc);
        }
    }
}

// One launch = the units of one queue.  (Pair production and then, from a second queue, the
// bremsstrahlung + ionisation units in ONE launch -- same register budget, the end of the pair
// kernel filled with short units -- was measured and lost 0.4 % / 1.3 % on a full / a 1/8 share:
// the combined kernel spills 148 B against 68; profiles/r02_flat_table_study.md.)
struct FlatQueues {
    double2 *terms_a, *terms_b, *terms_c;   // terms of the process; of the processes riding with it
    const double2 *pairrow;                 // {gamma, zeta} per row (pair production)
    uint32_t *queue_a;
};

template <int PROCESS>
__global__ void __launch_bounds__(kFlatThreads, TableMinBlocks<FlatCfg<PROCESS>::mask>::value *
                                                        (kThreads / kFlatThreads))
table_terms_kernel(const double *__restrict__ K, int64_t nK, const double2 *__restrict__ rowpar,
                   const __grid_constant__ FlatQueues fq, const __grid_constant__ FlatPlan fp,
                   const __grid_constant__ Params p) {
    __shared__ StagedShared s_staged;
    __shared__ double2 s_gl6[6];
    if (threadIdx.x < 6)
        s_gl6[threadIdx.x] = make_double2(c_gl6_x[threadIdx.x], c_gl6_w[threadIdx.x]);
    const glibm::Tab T = stage_all(s_staged, p);
    // The first terms kernel of a build is itself a dependent launch of table_rowpar_kernel (its
    // CTAs stage their tables while that one runs): it waits here for the row parameters and
    // the zeroed queues, and only then lets the next launch in -- so every later terms kernel
    // starts after them too.  The later ones release at once: the next launch of the build
    // (another process: other terms) may fill SMs as they free up.
    if (fp.first_launch) pdl_wait_prerequisites();
    pdl_release_dependents();
    flat_run_units<PROCESS>(K, nK, rowpar, fq.terms_a, fq.terms_b, fq.terms_c, fq.pairrow,
                            fq.queue_a, fp, p, T, s_gl6);
    // completion order along the chain: this kernel does not retire before its predecessor has,
    // so the summation kernel only has to wait for the last one
    __syncthreads();
    if (threadIdx.x == 0) pdl_wait_prerequisites();
}

struct FlatSum {
    int32_t first_slot, n_slots;    // the slots this launch adds up
    int32_t process[4];
    int32_t out_row[4];             // output row of slot s
    const double2 *terms[4];        // [nK][nodes]
    int32_t quadrature_only;
    double xlow;
};

#ifndef NOA_SUM_WARPS
#define NOA_SUM_WARPS 8
#endif
#ifndef NOA_SUM_STAGE
#define NOA_SUM_STAGE 128
#endif
// One WARP per (process, row).  The row's terms stream through a private shared-memory ring
// (cp.async, kSumStage nodes = 4 KB per stage, the next stage in flight while this one is added
// up: enough bytes in flight per SM to keep HBM busy with a few thousand rows, and a row costs
// its chain of additions -- 4.4 us at 1002 nodes -- plus one stage of latency however few rows
// there are).  Lanes 0 and 1 own the two chains: res += f(x) h w[j] in node order
// (numerics.hh:84-87).
constexpr int kSumWarps = NOA_SUM_WARPS;
constexpr uint32_t kSumStage = NOA_SUM_STAGE;

__device__ __forceinline__ void cp_async_16(uint32_t smem_addr, const void *gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(gmem)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

__global__ void __launch_bounds__(32 * kSumWarps)
table_sum_kernel(const double *__restrict__ K, int64_t nK, uint32_t nodes,
                 const __grid_constant__ FlatSum fs, const __grid_constant__ Params p,
                 const __grid_constant__ TableOut out) {
    __shared__ glibm::Tables s_tables;
    __shared__ double2 s_ring[kSumWarps][2][kSumStage];
    const glibm::Tab T = stage_tables(s_tables);     // for the closed-form rows
    pdl_wait_prerequisites();       // the terms kernels of the build have completed
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t w = threadIdx.x >> 5;
    const int64_t chains = nK * fs.n_slots;
    for (int64_t chain = (int64_t) blockIdx.x * kSumWarps + w; chain < chains;
         chain += (int64_t) gridDim.x * kSumWarps) {
        const int slot = fs.first_slot + (int) (chain / nK);
        const int64_t row = nK - 1 - chain % nK;
        const double k = K[row];
        const bool closed = fs.process[slot] == 3 && k <= p.i_kthr && !fs.quadrature_only;
        double acc = 0.;
        if (!closed) {
            const double2 *t = fs.terms[slot] + row * nodes;
            const uint32_t ring = (uint32_t) __cvta_generic_to_shared(&s_ring[w][0][0]);
            const uint32_t n_stages = (nodes + kSumStage - 1) / kSumStage;
            auto fetch = [&](uint32_t st) {
                const uint32_t base = st * kSumStage;
                const uint32_t dst = ring + (st & 1u) * kSumStage * 16u;
#pragma unroll
                for (uint32_t e = 0; e < kSumStage; e += 32u)
                    if (base + e + lane < nodes) cp_async_16(dst + (e + lane) * 16u, t + base + e + lane);
                cp_async_commit();
            };
            fetch(0);
            for (uint32_t st = 0; st < n_stages; st++) {
                if (st + 1 < n_stages) {
                    fetch(st + 1);
                    cp_async_wait<1>();
                } else {
                    cp_async_wait<0>();
                }
                __syncwarp();
                if (lane < 2) {
                    const double *src = reinterpret_cast<const double *>(&s_ring[w][st & 1u][0]) + lane;
                    const uint32_t count = min(kSumStage, nodes - st * kSumStage);
#pragma unroll 8
                    for (uint32_t il = 0; il < count; il++) acc += src[2 * il];
                }
                __syncwarp();       // the stage may be overwritten by the fetch after next
            }
        }
        if (lane < 2) {
            double *const *dst = lane ? out.cel : out.del;
            if (dst[0] != nullptr) {
                const double v = closed ? ionisation_closed_form(k, fs.xlow, (int) lane, p, T)
                                        : acc / (k + p.mass);
                const int64_t at = (int64_t) fs.out_row[slot] * out.n_total + out.first_row +
                                   row * out.row_stride;
                double *mc = lane ? out.mc_cel : out.mc_del;
                if (mc != nullptr && !NOA_XCHG_NO_REMOTE) {
                    // the switch replicates the store into all n_peers tables
                    asm volatile("multimem.st.relaxed.sys.global.f64 [%0], %1;" ::"l"(mc + at), "d"(v)
                                 : "memory");
                } else {
                    for (int j = 0; j < out.n_peers; j++)
                        if (!NOA_XCHG_NO_REMOTE || j == out.me) dst[j][at] = v;
                }
                // scatter form without flags: the writer lanes fence their own remote stores
                if (out.n_peers > 1 && out.flags[0] == nullptr) __threadfence_system();
            }
        }
    }
    if (out.flags[0] != nullptr) table_exchange_tail(out);
}

// Rows of processes outside the mask are defined to be zero in every destination table.
__global__ void table_zero_rows_kernel(int64_t n_local, unsigned zero_mask,
                                       const __grid_constant__ TableOut out) {
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    for (int64_t r = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; r < n_local; r += stride)
        for (int pr = 0; pr < 4; pr++) {
            if (!((zero_mask >> pr) & 1u)) continue;
            const int64_t at = (int64_t) pr * out.n_total + out.first_row + r * out.row_stride;
            for (int j = 0; j < out.n_peers; j++) {
                if (out.del[j]) out.del[j][at] = 0.;
                if (out.cel[j]) out.cel[j][at] = 0.;
            }
        }
}

// Material tables: out[c] = sum_e parts[e][c] * w[e], e in composition order, starting from 0
// (the per-element mixing of src/noa/3rdparty/_pumas/pumas.c:8054-8078).
struct MixWeights {
    int32_t n_elements;
    double w[NOA_DCS_MAX_ELEMENTS];
};

__global__ void mix_tables_kernel(const double *__restrict__ parts, double *__restrict__ out,
                                  int64_t columns, const __grid_constant__ MixWeights m) {
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    for (int64_t c = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; c < columns; c += stride) {
        double acc = 0.;
        for (int e = 0; e < m.n_elements; e++) acc += parts[(int64_t) e * columns + c] * m.w[e];
        out[c] = acc;
    }
}

// A rank with no rows of its own still has to take part in the exchange.
__global__ void table_signal_kernel(const __grid_constant__ TableOut out) {
    if (threadIdx.x == 0) pdl_wait_prerequisites();
    table_exchange_tail(out);
}

}  // namespace noa_b200
