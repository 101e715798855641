// LibTorch boundary of the B200 muon DCS path: `noa::pms::dcs::cuda`.
//
// Drop-in for the reference's CUDA entry points, which it declares in
// src/noa/pms/dcs.hh:1004-1019 and defines in src/noa/pms/dcs.cuh:30-51 (a TU that includes
// <noa/kernels.cuh>, e.g. test/kernels.cu).  Link noa_b200/csrc/torch_api.cc + libnoa_dcs_b200.so
// INSTEAD of that TU and every existing caller keeps compiling and running:
//   test/unit/test-dcs-calc-cuda.cc:17, benchmark/measure-dcs-calc-cuda.cc:18,29,
//   docs/pms/muon_dcs.cu:9.
// The same header adds the GPU forms the reference lacks ("Only a CPU version is available",
// docs/pms/muon_dcs_calc.ipynb:514-515): the other three processes, the fused forms and the
// recoil-integral (energy-loss table) drivers, all torch::Tensor in / torch::Tensor out.
//
// Build modes
//   -DNOA_B200_WITH_REFERENCE_HEADERS  include the reference's own <noa/pms/dcs.hh> for the types
//                                      (use this inside the NOA tree);
//   default                            a minimal mirror of the types of src/noa/pms/physics.hh
//                                      so the boundary builds without the reference present.
//
// Semantics: tensors must be float64, contiguous, equal numel, all on one device (checked; the
// reference assumes it, src/noa/utils/common.cuh:45-56).
//   * CUDA tensors: work is enqueued on the current CUDA stream of the tensors' device (the
//     reference uses the legacy default stream) and is not synchronised.
//   * CPU tensors (the reference's CPU call sites, dcs::vmap(f) on host memory,
//     src/noa/pms/dcs.hh:35-60): evaluated on the current CUDA device through the host-buffer
//     entry points of the C ABI -- pinned tensors are read and written in place by the kernel over
//     PCIe, pageable ones go through a chunked copy pipeline -- and complete on return, like the
//     CPU path they replace.  The element-wise and recoil-integral entry points and `tables` take
//     them; the Coulomb entry points are CUDA-only.
// `vmap_*` write `result` in place; `map_*` return a fresh tensor shaped like `kinetic_energies`.
// include/noa_b200/pms_dcs.hh adds the reference's own CPU call SHAPES (dcs::vmap(dcs::f)(...),
// dcs::vmap_integral(dcs::recoil_integral(f, g))(...)) on top of these.
#pragma once

#ifdef NOA_B200_WITH_REFERENCE_HEADERS
#include <noa/pms/dcs.hh>
#else
#include <torch/types.h>

namespace noa::pms {
    // src/noa/pms/physics.hh:29-71
    using Scalar = double_t;
    using Index = int32_t;
    using ParticleMass = Scalar;
    using AtomicNumber = Index;
    using AtomicMass = Scalar;
    using MeanExcitation = Scalar;
    using EnergyTransfer = Scalar;
    struct AtomicElement {
        AtomicMass A;
        MeanExcitation I;
        AtomicNumber Z;
    };
    using Energies = torch::Tensor;
    using Calculation = torch::Tensor;
    constexpr ParticleMass ELECTRON_MASS = 0.510998910E-03;
    constexpr ParticleMass MUON_MASS = 0.10565839;
    constexpr ParticleMass TAU_MASS = 1.77682;
    constexpr AtomicElement STANDARD_ROCK = AtomicElement{22., 0.1364E-6, 11};
    namespace dcs {
        constexpr EnergyTransfer X_FRACTION = 5E-02;   // physics.hh:76
        namespace cuda {
            // src/noa/pms/dcs.hh:1006-1017
            void vmap_bremsstrahlung(const Calculation &result, const Energies &kinetic_energies,
                                     const Energies &recoil_energies, const AtomicElement &element,
                                     const ParticleMass &mass);
            Calculation map_bremsstrahlung(const Energies &kinetic_energies,
                                           const Energies &recoil_energies,
                                           const AtomicElement &element, const ParticleMass &mass);
        }
    }
}
#endif

#include <vector>

namespace noa::pms::dcs::cuda {

    // ---- by process id (0 bremsstrahlung, 1 pair_production, 2 photonuclear, 3 ionisation): what
    // the named forms below and include/noa_b200/pms_dcs.hh forward to
    void vmap_dcs(int process, const Calculation &result, const Energies &kinetic_energies,
                  const Energies &recoil_energies, const AtomicElement &element,
                  const ParticleMass &mass);
    Calculation map_dcs(int process, const Energies &kinetic_energies,
                        const Energies &recoil_energies, const AtomicElement &element,
                        const ParticleMass &mass);

    // ---- same call shape as vmap_bremsstrahlung / map_bremsstrahlung for the other processes:
    // GPU forms of dcs::vmap(dcs::pair_production) etc. (src/noa/pms/dcs.hh:35-60, 144-443)
    void vmap_pair_production(const Calculation &result, const Energies &kinetic_energies,
                              const Energies &recoil_energies, const AtomicElement &element,
                              const ParticleMass &mass);
    Calculation map_pair_production(const Energies &kinetic_energies,
                                    const Energies &recoil_energies, const AtomicElement &element,
                                    const ParticleMass &mass);
    void vmap_photonuclear(const Calculation &result, const Energies &kinetic_energies,
                           const Energies &recoil_energies, const AtomicElement &element,
                           const ParticleMass &mass);
    Calculation map_photonuclear(const Energies &kinetic_energies, const Energies &recoil_energies,
                                 const AtomicElement &element, const ParticleMass &mass);
    void vmap_ionisation(const Calculation &result, const Energies &kinetic_energies,
                         const Energies &recoil_energies, const AtomicElement &element,
                         const ParticleMass &mass);
    Calculation map_ionisation(const Energies &kinetic_energies, const Energies &recoil_energies,
                               const AtomicElement &element, const ParticleMass &mass);

    // ---- the four processes in one pass; result / return value is [4, numel]
    void vmap_all(const Calculation &result, const Energies &kinetic_energies,
                  const Energies &recoil_energies, const AtomicElement &element,
                  const ParticleMass &mass);
    Calculation map_all(const Energies &kinetic_energies, const Energies &recoil_energies,
                        const AtomicElement &element, const ParticleMass &mass);

    // ---- material = mass-fraction mix of elements: result[p] = sum_e w_e DCS_p(element e)
    Calculation map_material(const Energies &kinetic_energies, const Energies &recoil_energies,
                             const std::vector<AtomicElement> &elements,
                             const std::vector<Scalar> &mass_fractions, const ParticleMass &mass);

    // ---- dcs::vmap_integral(dcs::recoil_integral(f, del|cel_integrand)) on the GPU
    // (src/noa/pms/dcs.hh:89-130, 955-1001).  process: 0 bremsstrahlung, 1 pair_production,
    // 2 photonuclear, 3 ionisation;  integrand: 0 del_integrand, 1 cel_integrand.
    void vmap_integral(int process, int integrand, const Calculation &result,
                       const Energies &kinetic_energies, const EnergyTransfer &xlow,
                       const AtomicElement &element, const ParticleMass &mass,
                       const Index min_points);
    void vmap_del_integral_bremsstrahlung(const Calculation &result, const Energies &K,
                                          const EnergyTransfer &xlow, const AtomicElement &element,
                                          const ParticleMass &mass, const Index min_points);
    void vmap_cel_integral_bremsstrahlung(const Calculation &result, const Energies &K,
                                          const EnergyTransfer &xlow, const AtomicElement &element,
                                          const ParticleMass &mass, const Index min_points);
    // ... the remaining six named forms are spelled vmap_{del,cel}_integral_{pair_production,
    // photonuclear,ionisation} and declared below
    void vmap_del_integral_pair_production(const Calculation &result, const Energies &K,
                                           const EnergyTransfer &xlow, const AtomicElement &element,
                                           const ParticleMass &mass, const Index min_points);
    void vmap_cel_integral_pair_production(const Calculation &result, const Energies &K,
                                           const EnergyTransfer &xlow, const AtomicElement &element,
                                           const ParticleMass &mass, const Index min_points);
    void vmap_del_integral_photonuclear(const Calculation &result, const Energies &K,
                                        const EnergyTransfer &xlow, const AtomicElement &element,
                                        const ParticleMass &mass, const Index min_points);
    void vmap_cel_integral_photonuclear(const Calculation &result, const Energies &K,
                                        const EnergyTransfer &xlow, const AtomicElement &element,
                                        const ParticleMass &mass, const Index min_points);
    void vmap_del_integral_ionisation(const Calculation &result, const Energies &K,
                                      const EnergyTransfer &xlow, const AtomicElement &element,
                                      const ParticleMass &mass, const Index min_points);
    void vmap_cel_integral_ionisation(const Calculation &result, const Energies &K,
                                      const EnergyTransfer &xlow, const AtomicElement &element,
                                      const ParticleMass &mass, const Index min_points);

    // ---- fused table build: returns [2 (DEL, CEL), 4 (process), numel(K)]; one DCS evaluation per
    // node feeds both integrands, one kernel launch for everything.
    Calculation tables(const Energies &kinetic_energies, const EnergyTransfer &xlow,
                       const AtomicElement &element, const ParticleMass &mass,
                       const Index min_points);

    // ---- Coulomb scattering and soft scattering on the GPU: same argument order, shapes and
    // in-place semantics as the reference's CPU functors dcs::coulomb_data (dcs.hh:600-622),
    // dcs::coulomb_transport (:674-693), dcs::hard_scattering (:843-872) and dcs::soft_scattering
    // (:940-952); callers: test/unit/test-dcs-calc.cc:134-178.
    //   fCM [n, 2], screening [n, 9], fspin [n], invlambda [n], coefficients [n, 2];
    //   mu: one cutoff (numel 1) or one per energy;
    //   hard_scattering: coefficients / transform [nel, nkin, 2], screening [nel, nkin, 9],
    //   invlambdas / fspins [nel, nkin] -> mu0 [nkin], lb_h [nkin].
    void coulomb_data(const torch::Tensor &fCM, const torch::Tensor &screening,
                      const torch::Tensor &fspin, const torch::Tensor &invlambda,
                      const Energies &kinetic_energies, const AtomicElement &element,
                      const ParticleMass &mass);
    void coulomb_transport(const torch::Tensor &coefficients, const torch::Tensor &screening,
                           const torch::Tensor &fspin, const torch::Tensor &mu);
    void hard_scattering(const torch::Tensor &mu0, const torch::Tensor &lb_h,
                         const torch::Tensor &coefficients, const torch::Tensor &transform,
                         const torch::Tensor &screening, const torch::Tensor &invlambdas,
                         const torch::Tensor &fspins);
    void soft_scattering(const Calculation &ms1, const Energies &kinetic_energies,
                         const AtomicElement &element, const ParticleMass &mass);

}  // namespace noa::pms::dcs::cuda
