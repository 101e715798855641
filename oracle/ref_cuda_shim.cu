// TEST / MEASUREMENT INFRASTRUCTURE ONLY -- never linked or imported by the product path.
//
// C-ABI shim around the reference's OWN CUDA path, compiled where it lies under
// /root/reference/src (recipe: oracle/Makefile `make ref_cuda`, output
// oracle/_ref/libnoa_ref_cuda.so, git-ignored): noa::pms::dcs::cuda::vmap_bremsstrahlung
// (src/noa/pms/dcs.cuh:30-41 -> utils::cuda::vmapi -> launch_kernel, src/noa/utils/common.cuh:39-76),
// the only GPU kernel the reference has and the one SURVEY 2a / BASELINE.md name as "the GPU
// baseline to beat".  bench.py times it on the same B200 beside noa_dcs_vmap_f64 and the parity
// tests compare the two.  The reference launches on the legacy default stream without a sync, so
// the shim exposes an explicit synchronise.  No reference source is copied; this file only calls it.
#include <noa/kernels.cuh>

#include <cuda_runtime.h>
#include <cstdint>

using namespace noa::pms;

namespace {
    inline torch::Tensor wrap(const double *p, int64_t n) {
        return torch::from_blob(const_cast<double *>(p), {n},
                                torch::TensorOptions().dtype(torch::kFloat64).device(torch::kCUDA));
    }
}

extern "C" {

// dcs::cuda::vmap_bremsstrahlung(result, K, q, element, mass) on raw device pointers
void noa_ref_cuda_vmap_bremsstrahlung(double *out, const double *K, const double *q, int64_t n,
                                      double A, double I, int32_t Z, double mass) {
    const AtomicElement el{A, I, Z};
    dcs::cuda::vmap_bremsstrahlung(wrap(out, n), wrap(K, n), wrap(q, n), el, mass);
}

// dcs::cuda::map_bremsstrahlung(K, q, element, mass): allocates (zeros_like) and launches, as the
// reference's benchmark case BremsstrahlungVectorisedCUDA's callers see it; the result is dropped
void noa_ref_cuda_map_bremsstrahlung(const double *K, const double *q, int64_t n, double A, double I,
                                     int32_t Z, double mass) {
    const AtomicElement el{A, I, Z};
    (void) dcs::cuda::map_bremsstrahlung(wrap(K, n), wrap(q, n), el, mass);
}

int noa_ref_cuda_sync() { return (int) cudaDeviceSynchronize(); }

}
