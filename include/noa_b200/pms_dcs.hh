// The reference's CPU call shapes of noa::pms::dcs, executed on the B200.
//
// The reference's call sites build closures from its scalar DCS lambdas,
//   dcs::vmap(dcs::pair_production)(result, K, q, STANDARD_ROCK, MUON_MASS);        // dcs.hh:35-48
//   auto r = dcs::map(dcs::photonuclear)(K, q, STANDARD_ROCK, MUON_MASS);           // dcs.hh:50-60
//   dcs::pvmap(dcs::ionisation)(result, K, q, el, mass);  dcs::pmap(f)(K, q, ...)    // dcs.hh:62-87
//   dcs::vmap_integral(dcs::recoil_integral(dcs::bremsstrahlung, dcs::cel_integrand))(
//       result, K, dcs::X_FRACTION, STANDARD_ROCK, MUON_MASS, 180);                 // dcs.hh:89-130
// (test/unit/test-dcs-calc.cc:22-131, docs/pms/muon_dcs.cc:8-27, benchmark/measure-dcs-calc.cc).
// Here the same expressions compile against tokens instead of host lambdas and run the CUDA
// kernels: `noa::pms::dcs::b200::` holds them; without the reference's headers in the build the
// names are also placed in `noa::pms::dcs` itself, so those call sites compile UNCHANGED and only
// the link line differs (libnoa_dcs_b200_torch.so + libnoa_dcs_b200.so).  Inside the NOA tree
// (-DNOA_B200_WITH_REFERENCE_HEADERS) the reference owns those names; write `dcs::b200::vmap(...)`
// or `namespace dcs = noa::pms::dcs::b200;` at the call site.
// CPU and CUDA tensors are both accepted (see pms_dcs_cuda.hh for what happens to each).
#pragma once

#include "pms_dcs_cuda.hh"

namespace noa::pms::dcs::b200 {

    struct Process {
        int id;
    };
    inline constexpr Process bremsstrahlung{0};     // physics.hh:108-153, dcs.hh:132-142
    inline constexpr Process pair_production{1};    // dcs.hh:144-258
    inline constexpr Process photonuclear{2};       // dcs.hh:362-405
    inline constexpr Process ionisation{3};         // dcs.hh:408-443

    struct Integrand {
        int id;
    };
    inline constexpr Integrand del_integrand{0};    // dcs * q      (dcs.hh:107-109)
    inline constexpr Integrand cel_integrand{1};    // dcs * q * q  (dcs.hh:111-113)

    struct RecoilIntegral {
        Process process;
        Integrand integrand;
    };

    // dcs.hh:35-48
    inline auto vmap(const Process &dcs_func) {
        return [dcs_func](const Calculation &result, const Energies &kinetic_energies,
                          const Energies &recoil_energies, const AtomicElement &element,
                          const ParticleMass &mass) {
            cuda::vmap_dcs(dcs_func.id, result, kinetic_energies, recoil_energies, element, mass);
        };
    }

    // dcs.hh:50-60
    inline auto map(const Process &dcs_func) {
        return [dcs_func](const Energies &kinetic_energies, const Energies &recoil_energies,
                          const AtomicElement &element, const ParticleMass &mass) {
            return cuda::map_dcs(dcs_func.id, kinetic_energies, recoil_energies, element, mass);
        };
    }

    // dcs.hh:62-87: the OpenMP forms are the same GPU call
    inline auto pvmap(const Process &dcs_func) { return vmap(dcs_func); }
    inline auto pmap(const Process &dcs_func) { return map(dcs_func); }

    // dcs.hh:89-105, 955-1001
    inline RecoilIntegral recoil_integral(const Process &dcs_func, const Integrand &integrand) {
        return RecoilIntegral{dcs_func, integrand};
    }

    // dcs.hh:115-130
    inline auto vmap_integral(const RecoilIntegral &cs_integral) {
        return [cs_integral](const Calculation &result, const Energies &kinetic_energies,
                             const EnergyTransfer &xlow, const AtomicElement &element,
                             const ParticleMass &mass, const Index min_points) {
            cuda::vmap_integral(cs_integral.process.id, cs_integral.integrand.id, result,
                                kinetic_energies, xlow, element, mass, min_points);
        };
    }

}  // namespace noa::pms::dcs::b200

#ifndef NOA_B200_WITH_REFERENCE_HEADERS
namespace noa::pms::dcs {
    using b200::bremsstrahlung;
    using b200::pair_production;
    using b200::photonuclear;
    using b200::ionisation;
    using b200::del_integrand;
    using b200::cel_integrand;
    using b200::vmap;
    using b200::map;
    using b200::pvmap;
    using b200::pmap;
    using b200::recoil_integral;
    using b200::vmap_integral;
}
#endif
