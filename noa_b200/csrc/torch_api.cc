// LibTorch boundary: defines noa::pms::dcs::cuda::* (include/noa_b200/pms_dcs_cuda.hh) on top of
// the C ABI (include/noa_dcs_b200.h).  CUDA tensors are passed straight through (asynchronous on
// the current stream); CPU tensors -- the reference's CPU call sites, dcs::vmap(f)(result, K, q,
// ...) on host memory (src/noa/pms/dcs.hh:35-60, test/unit/test-dcs-calc.cc:44-48) -- run on the
// GPU through the host-buffer entry points and are complete on return, like the CPU path they
// replace: pinned tensors are read and written in place by the kernel
// (noa_dcs_vmap_pinned_f64), pageable ones go through the chunked copy pipeline
// (noa_dcs_vmap_host_f64).  There is still no CPU implementation.  Compiled by the host compiler only -- no nvcc, no kernels
// here -- so it builds in about a minute despite <torch/...> (the reference's dcs.cuh TU needs
// ~5 min under nvcc).
#include "../../include/noa_b200/pms_dcs_cuda.hh"
#include "../../include/noa_dcs_b200.h"

#include <c10/cuda/CUDAFunctions.h>
#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>
#include <torch/torch.h>

#include <mutex>

namespace noa::pms::dcs::cuda {

    namespace {
        void check_layout(const torch::Tensor &t, const char *name) {
            TORCH_CHECK(t.defined(), name, " is undefined");
            TORCH_CHECK(t.is_cuda() || t.is_cpu(), name, " must be a CUDA or a CPU tensor");
            TORCH_CHECK(t.scalar_type() == torch::kFloat64, name, " must be float64, got ",
                        t.scalar_type());
            TORCH_CHECK(t.is_contiguous(), name, " must be contiguous");
        }

        void check_tensor(const torch::Tensor &t, const char *name) {
            check_layout(t, name);
            TORCH_CHECK(t.is_cuda(), name, " must be a CUDA tensor");
        }

        bool pinned(const torch::Tensor &t) {
            try {
                return t.is_pinned();
            } catch (const c10::Error &) {      // no CUDA hooks in this process
                return false;
            }
        }

        // device scratch + streams of the pageable host path, created on first use per device
        struct HostStager {
            std::mutex mu;
            noa_dcs_stager *stager = nullptr;
            int device = -1;
            ~HostStager() {
                if (stager) noa_dcs_stager_destroy(stager);
            }
        };
        HostStager &host_stager() {
            static HostStager s;
            return s;
        }

        // dcs::vmap(f)(result, K, q, element, mass) on CPU tensors; process 4 = all four
        void host_vmap(int process, const torch::Tensor &result, const torch::Tensor &K,
                       const torch::Tensor &q, const AtomicElement &el, const ParticleMass &mass) {
            const int64_t n = K.numel();
            if (n == 0) return;
            if (process < NOA_DCS_NPROCESS && pinned(K) && pinned(q) && pinned(result)) {
                const auto stream = c10::cuda::getCurrentCUDAStream();
                const int rc = noa_dcs_vmap_pinned_f64(process, K.data_ptr<double>(),
                                                       q.data_ptr<double>(),
                                                       result.data_ptr<double>(), n, el.A, el.I,
                                                       el.Z, mass, (void *) stream.stream());
                TORCH_CHECK(rc == 0, "noa_dcs_vmap_pinned_f64 failed: ", noa_dcs_strerror(rc));
                stream.synchronize();
                return;
            }
            auto &hs = host_stager();
            const std::lock_guard<std::mutex> lock(hs.mu);
            const int device = (int) c10::cuda::current_device();
            if (hs.stager == nullptr || hs.device != device) {
                if (hs.stager) noa_dcs_stager_destroy(hs.stager);
                hs.stager = nullptr;
                const int rc = noa_dcs_stager_create(&hs.stager, int64_t(1) << 18, 3);
                TORCH_CHECK(rc == 0, "noa_dcs_stager_create failed: ", noa_dcs_strerror(rc));
                hs.device = device;
            }
            const int rc = noa_dcs_vmap_host_f64(hs.stager, process, K.data_ptr<double>(),
                                                 q.data_ptr<double>(), result.data_ptr<double>(),
                                                 n, el.A, el.I, el.Z, mass);
            TORCH_CHECK(rc == 0, "noa_dcs_vmap_host_f64 failed: ", noa_dcs_strerror(rc));
        }

        void check_pair(const torch::Tensor &K, const torch::Tensor &q) {
            check_layout(K, "kinetic_energies");
            check_layout(q, "recoil_energies");
            TORCH_CHECK(K.numel() == q.numel(), "kinetic_energies and recoil_energies differ in "
                        "size: ", K.numel(), " vs ", q.numel());
            TORCH_CHECK(K.device() == q.device(), "tensors are on different devices");
        }

        void check_rc(int rc, const char *what) {
            TORCH_CHECK(rc == 0, what, " failed: ", noa_dcs_strerror(rc), " (", rc, ")");
        }

        void *current_stream(const torch::Tensor &t) {
            return (void *) c10::cuda::getCurrentCUDAStream(t.device().index()).stream();
        }

    }  // namespace

    void vmap_dcs(int process, const Calculation &result, const Energies &K, const Energies &q,
                  const AtomicElement &el, const ParticleMass &mass) {
        {
            TORCH_CHECK(process >= 0 && process < NOA_DCS_NPROCESS, "process must be 0..3");
            check_pair(K, q);
            check_layout(result, "result");
            TORCH_CHECK(result.numel() == K.numel(), "result has ", result.numel(),
                        " elements, expected ", K.numel());
            TORCH_CHECK(result.device() == K.device(), "result is on a different device");
            if (!K.is_cuda()) {
                host_vmap(process, result, K, q, el, mass);
                return;
            }
            const c10::cuda::CUDAGuard guard(K.device());
            check_rc(noa_dcs_vmap_f64(process, K.data_ptr<double>(), q.data_ptr<double>(),
                                      result.data_ptr<double>(), K.numel(), el.A, el.I, el.Z, mass,
                                      current_stream(K)),
                     "noa_dcs_vmap_f64");
        }
    }

    Calculation map_dcs(int process, const Energies &K, const Energies &q, const AtomicElement &el,
                        const ParticleMass &mass) {
        // the reference allocates zeros_like (dcs.cuh:48, dcs.hh:57); every element is overwritten
        const auto result = torch::empty_like(K);
        vmap_dcs(process, result, K, q, el, mass);
        return result;
    }

#define NOA_B200_DEFINE_PROCESS(NAME, ID)                                                        \
    void vmap_##NAME(const Calculation &result, const Energies &kinetic_energies,                \
                     const Energies &recoil_energies, const AtomicElement &element,              \
                     const ParticleMass &mass) {                                                 \
        vmap_dcs(ID, result, kinetic_energies, recoil_energies, element, mass);              \
    }                                                                                            \
    Calculation map_##NAME(const Energies &kinetic_energies, const Energies &recoil_energies,    \
                           const AtomicElement &element, const ParticleMass &mass) {             \
        return map_dcs(ID, kinetic_energies, recoil_energies, element, mass);                \
    }

    NOA_B200_DEFINE_PROCESS(bremsstrahlung, NOA_DCS_BREMSSTRAHLUNG)
    NOA_B200_DEFINE_PROCESS(pair_production, NOA_DCS_PAIR_PRODUCTION)
    NOA_B200_DEFINE_PROCESS(photonuclear, NOA_DCS_PHOTONUCLEAR)
    NOA_B200_DEFINE_PROCESS(ionisation, NOA_DCS_IONISATION)
#undef NOA_B200_DEFINE_PROCESS

    void vmap_all(const Calculation &result, const Energies &K, const Energies &q,
                  const AtomicElement &el, const ParticleMass &mass) {
        check_pair(K, q);
        check_layout(result, "result");
        TORCH_CHECK(result.numel() == 4 * K.numel(), "result must hold 4 x ", K.numel(),
                    " elements");
        TORCH_CHECK(result.device() == K.device(), "result is on a different device");
        if (!K.is_cuda()) {
            host_vmap(NOA_DCS_NPROCESS, result, K, q, el, mass);
            return;
        }
        const c10::cuda::CUDAGuard guard(K.device());
        check_rc(noa_dcs_vmap_all_f64(K.data_ptr<double>(), q.data_ptr<double>(),
                                      result.data_ptr<double>(), K.numel(), el.A, el.I, el.Z, mass,
                                      current_stream(K)),
                 "noa_dcs_vmap_all_f64");
    }

    Calculation map_all(const Energies &K, const Energies &q, const AtomicElement &el,
                        const ParticleMass &mass) {
        auto shape = K.sizes().vec();
        shape.insert(shape.begin(), 4);
        const auto result = torch::empty(shape, K.options());
        vmap_all(result, K, q, el, mass);
        return result;
    }

    Calculation map_material(const Energies &K, const Energies &q,
                             const std::vector<AtomicElement> &elements,
                             const std::vector<Scalar> &mass_fractions, const ParticleMass &mass) {
        check_pair(K, q);
        if (!K.is_cuda())       // host tensors: evaluate on the device, hand back a host tensor
            return map_material(K.to(torch::kCUDA), q.to(torch::kCUDA), elements, mass_fractions,
                                mass).to(torch::kCPU);
        TORCH_CHECK(!elements.empty() && elements.size() == mass_fractions.size() &&
                    elements.size() <= NOA_DCS_MAX_ELEMENTS,
                    "a material has 1..", NOA_DCS_MAX_ELEMENTS, " elements with one mass fraction "
                    "each");
        std::vector<double> A, I;
        std::vector<int32_t> Z;
        for (const auto &e : elements) {
            A.push_back(e.A);
            I.push_back(e.I);
            Z.push_back(e.Z);
        }
        auto shape = K.sizes().vec();
        shape.insert(shape.begin(), 4);
        const auto result = torch::empty(shape, K.options());
        const c10::cuda::CUDAGuard guard(K.device());
        check_rc(noa_dcs_vmap_mixture_f64(0xFu, K.data_ptr<double>(), q.data_ptr<double>(),
                                          result.data_ptr<double>(), K.numel(),
                                          (int32_t) elements.size(), A.data(), I.data(), Z.data(),
                                          mass_fractions.data(), mass, current_stream(K)),
                 "noa_dcs_vmap_mixture_f64");
        return result;
    }

    void vmap_integral(int process, int integrand, const Calculation &result, const Energies &K,
                       const EnergyTransfer &xlow, const AtomicElement &el,
                       const ParticleMass &mass, const Index min_points) {
        check_layout(K, "kinetic_energies");
        check_layout(result, "result");
        TORCH_CHECK(result.numel() == K.numel(), "result has ", result.numel(),
                    " elements, expected ", K.numel());
        TORCH_CHECK(result.device() == K.device(), "result is on a different device");
        if (!K.is_cuda()) {
            // dcs::vmap_integral on CPU tensors (test/unit/test-dcs-calc.cc:22-42): the energies go
            // to the device, the column comes back into `result` (blocking copy)
            if (K.numel() == 0) return;
            const auto Kd = K.to(torch::kCUDA);
            const auto rd = torch::empty_like(Kd);
            vmap_integral(process, integrand, rd, Kd, xlow, el, mass, min_points);
            result.copy_(rd.view_as(result));
            return;
        }
        const c10::cuda::CUDAGuard guard(K.device());
        check_rc(noa_dcs_vmap_integral_f64(process, integrand, K.data_ptr<double>(),
                                           result.data_ptr<double>(), K.numel(), xlow, min_points,
                                           el.A, el.I, el.Z, mass, current_stream(K)),
                 "noa_dcs_vmap_integral_f64");
    }

#define NOA_B200_DEFINE_INTEGRAL(NAME, ID)                                                        \
    void vmap_del_integral_##NAME(const Calculation &result, const Energies &K,                   \
                                  const EnergyTransfer &xlow, const AtomicElement &element,       \
                                  const ParticleMass &mass, const Index min_points) {             \
        vmap_integral(ID, 0, result, K, xlow, element, mass, min_points);                         \
    }                                                                                             \
    void vmap_cel_integral_##NAME(const Calculation &result, const Energies &K,                   \
                                  const EnergyTransfer &xlow, const AtomicElement &element,       \
                                  const ParticleMass &mass, const Index min_points) {             \
        vmap_integral(ID, 1, result, K, xlow, element, mass, min_points);                         \
    }

    NOA_B200_DEFINE_INTEGRAL(bremsstrahlung, NOA_DCS_BREMSSTRAHLUNG)
    NOA_B200_DEFINE_INTEGRAL(pair_production, NOA_DCS_PAIR_PRODUCTION)
    NOA_B200_DEFINE_INTEGRAL(photonuclear, NOA_DCS_PHOTONUCLEAR)
    NOA_B200_DEFINE_INTEGRAL(ionisation, NOA_DCS_IONISATION)
#undef NOA_B200_DEFINE_INTEGRAL

    Calculation tables(const Energies &K, const EnergyTransfer &xlow, const AtomicElement &el,
                       const ParticleMass &mass, const Index min_points) {
        check_layout(K, "kinetic_energies");
        if (!K.is_cuda()) return tables(K.to(torch::kCUDA), xlow, el, mass, min_points).to(torch::kCPU);
        const auto result = torch::zeros({2, 4, K.numel()}, K.options());
        if (K.numel() == 0) return result;
        const c10::cuda::CUDAGuard guard(K.device());
        double *base = result.data_ptr<double>();
        // flat form: the node terms go through a workspace (16 B per node and process) taken from
        // the stream-ordered caching allocator for the duration of the launches
        const int64_t need = noa_dcs_table_workspace_doubles(K.numel(), min_points);
        const auto workspace = torch::empty({need}, K.options());
        check_rc(noa_dcs_table_ws_f64(0xFu, K.data_ptr<double>(), K.numel(), xlow, min_points, el.A,
                                      el.I, el.Z, mass, base, base + 4 * K.numel(),
                                      workspace.data_ptr<double>(), need, current_stream(K)),
                 "noa_dcs_table_ws_f64");
        return result;
    }

    namespace {
        void check_numel(const torch::Tensor &t, const char *name, int64_t expected) {
            check_tensor(t, name);
            TORCH_CHECK(t.numel() == expected, name, " has ", t.numel(), " elements, expected ",
                        expected);
        }
    }  // namespace

    void coulomb_data(const torch::Tensor &fCM, const torch::Tensor &screening,
                      const torch::Tensor &fspin, const torch::Tensor &invlambda,
                      const Energies &K, const AtomicElement &el, const ParticleMass &mass) {
        check_tensor(K, "kinetic_energies");
        const int64_t n = K.numel();
        check_numel(fCM, "fCM", 2 * n);
        check_numel(screening, "screening", 9 * n);
        check_numel(fspin, "fspin", n);
        check_numel(invlambda, "invlambda", n);
        const c10::cuda::CUDAGuard guard(K.device());
        check_rc(noa_dcs_coulomb_data_f64(K.data_ptr<double>(), n, el.A, el.I, el.Z, mass,
                                          fCM.data_ptr<double>(), screening.data_ptr<double>(),
                                          fspin.data_ptr<double>(), invlambda.data_ptr<double>(),
                                          current_stream(K)),
                 "noa_dcs_coulomb_data_f64");
    }

    void coulomb_transport(const torch::Tensor &coefficients, const torch::Tensor &screening,
                           const torch::Tensor &fspin, const torch::Tensor &mu) {
        check_tensor(fspin, "fspin");
        const int64_t n = fspin.numel();
        check_numel(coefficients, "coefficients", 2 * n);
        check_numel(screening, "screening", 9 * n);
        check_tensor(mu, "mu");
        TORCH_CHECK(mu.numel() == 1 || mu.numel() == n, "mu must hold 1 or ", n, " elements");
        const c10::cuda::CUDAGuard guard(fspin.device());
        check_rc(noa_dcs_coulomb_transport_f64(screening.data_ptr<double>(),
                                               fspin.data_ptr<double>(), mu.data_ptr<double>(),
                                               mu.numel(), n, coefficients.data_ptr<double>(),
                                               current_stream(fspin)),
                 "noa_dcs_coulomb_transport_f64");
    }

    void hard_scattering(const torch::Tensor &mu0, const torch::Tensor &lb_h,
                         const torch::Tensor &coefficients, const torch::Tensor &transform,
                         const torch::Tensor &screening, const torch::Tensor &invlambdas,
                         const torch::Tensor &fspins) {
        check_tensor(invlambdas, "invlambdas");
        TORCH_CHECK(invlambdas.dim() == 2, "invlambdas must be [nel, nkin]");
        const int64_t nel = invlambdas.size(0), nkin = invlambdas.size(1);
        check_numel(fspins, "fspins", nel * nkin);
        check_numel(coefficients, "coefficients", 2 * nel * nkin);
        check_numel(transform, "transform", 2 * nel * nkin);
        check_numel(screening, "screening", 9 * nel * nkin);
        check_numel(mu0, "mu0", nkin);
        check_numel(lb_h, "lb_h", nkin);
        const c10::cuda::CUDAGuard guard(invlambdas.device());
        check_rc(noa_dcs_hard_scattering_f64(coefficients.data_ptr<double>(),
                                             transform.data_ptr<double>(),
                                             screening.data_ptr<double>(),
                                             invlambdas.data_ptr<double>(),
                                             fspins.data_ptr<double>(), (int32_t) nel, nkin,
                                             mu0.data_ptr<double>(), lb_h.data_ptr<double>(),
                                             current_stream(invlambdas)),
                 "noa_dcs_hard_scattering_f64");
    }

    void soft_scattering(const Calculation &ms1, const Energies &K, const AtomicElement &el,
                         const ParticleMass &mass) {
        check_tensor(K, "kinetic_energies");
        check_numel(ms1, "ms1", K.numel());
        const c10::cuda::CUDAGuard guard(K.device());
        check_rc(noa_dcs_soft_scattering_f64(K.data_ptr<double>(), K.numel(), el.A, el.I, el.Z,
                                             mass, ms1.data_ptr<double>(), current_stream(K)),
                 "noa_dcs_soft_scattering_f64");
    }

}  // namespace noa::pms::dcs::cuda
