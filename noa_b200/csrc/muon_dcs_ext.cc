// pybind11 torch extension with the surface of the reference's notebook extensions
// docs/pms/muon_dcs.cu:8-16 (`bremsstrahlung` on CUDA) and docs/pms/muon_dcs.cc:8-45 (the four
// processes + `serialise`), all on the GPU: standard rock, muon, GIL released during the call.
// It goes through the LibTorch boundary (torch_api.cc), i.e. exactly what a C++ user links.
#include "../../include/noa_b200/pms_dcs.hh"

#include <torch/extension.h>

using namespace noa::pms;

inline torch::Tensor bremsstrahlung(torch::Tensor kinetic_energies, torch::Tensor recoil_energies) {
    return dcs::cuda::map_bremsstrahlung(kinetic_energies, recoil_energies, STANDARD_ROCK,
                                         MUON_MASS);
}

inline torch::Tensor pair_production(torch::Tensor kinetic_energies,
                                     torch::Tensor recoil_energies) {
    return dcs::cuda::map_pair_production(kinetic_energies, recoil_energies, STANDARD_ROCK,
                                          MUON_MASS);
}

inline torch::Tensor photonuclear(torch::Tensor kinetic_energies, torch::Tensor recoil_energies) {
    return dcs::cuda::map_photonuclear(kinetic_energies, recoil_energies, STANDARD_ROCK,
                                       MUON_MASS);
}

inline torch::Tensor ionisation(torch::Tensor kinetic_energies, torch::Tensor recoil_energies) {
    return dcs::cuda::map_ionisation(kinetic_energies, recoil_energies, STANDARD_ROCK, MUON_MASS);
}

inline torch::Tensor all_processes(torch::Tensor kinetic_energies, torch::Tensor recoil_energies) {
    return dcs::cuda::map_all(kinetic_energies, recoil_energies, STANDARD_ROCK, MUON_MASS);
}

inline torch::Tensor tables(torch::Tensor kinetic_energies, double xlow, int min_points) {
    return dcs::cuda::tables(kinetic_energies, xlow, STANDARD_ROCK, MUON_MASS, min_points);
}

inline torch::Tensor recoil_integral(int process, int integrand, torch::Tensor kinetic_energies,
                                     double xlow, int min_points) {
    const auto result = torch::empty_like(kinetic_energies);
    dcs::cuda::vmap_integral(process, integrand, result, kinetic_energies, xlow, STANDARD_ROCK,
                             MUON_MASS, min_points);
    return result;
}

inline torch::Tensor water(torch::Tensor kinetic_energies, torch::Tensor recoil_energies) {
    return dcs::cuda::map_material(kinetic_energies, recoil_energies,
                                   {AtomicElement{1.0087, 19.2E-9, 1},
                                    AtomicElement{15.999, 95.0E-9, 8}},
                                   {0.111894, 0.888106}, MUON_MASS);
}

// dcs::coulomb_data -> coulomb_transport(mu) -> hard_scattering for standard rock, with the call
// shapes of test/unit/test-dcs-calc.cc:134-172; returns {fCM, screening, fspin, invlambda, G, mu0,
// lb_h}
inline std::vector<torch::Tensor> coulomb_hard_scattering(torch::Tensor kinetic_energies,
                                                          double mu) {
    const int64_t nkin = kinetic_energies.numel();
    const auto opt = kinetic_energies.options();
    auto fCM = torch::zeros({nkin, 2}, opt);
    auto screen = torch::zeros({nkin, 9}, opt);
    auto fspin = torch::zeros_like(kinetic_energies);
    auto invlambda = torch::zeros_like(kinetic_energies);
    dcs::cuda::coulomb_data(fCM, screen, fspin, invlambda, kinetic_energies, STANDARD_ROCK,
                            MUON_MASS);
    auto G = torch::zeros_like(fCM);
    dcs::cuda::coulomb_transport(G, screen, fspin, torch::full({1}, mu, opt));
    const auto lb_h = torch::zeros_like(kinetic_energies);
    const auto mu0 = torch::zeros_like(kinetic_energies);
    dcs::cuda::hard_scattering(mu0, lb_h, G.view({1, nkin, 2}), fCM.view({1, nkin, 2}),
                               screen.view({1, nkin, 9}), invlambda.view({1, nkin}),
                               fspin.view({1, nkin}));
    return {fCM, screen, fspin, invlambda, G, mu0, lb_h};
}

inline torch::Tensor soft_scattering(torch::Tensor kinetic_energies) {
    const auto result = torch::zeros_like(kinetic_energies);
    dcs::cuda::soft_scattering(result, kinetic_energies, STANDARD_ROCK, MUON_MASS);
    return result;
}

// The reference's CPU call sites, spelled exactly as in test/unit/test-dcs-calc.cc:22-131 and
// docs/pms/muon_dcs.cc:8-27 (include/noa_b200/pms_dcs.hh makes them compile unchanged); `K`, `q` are
// CPU (or CUDA) tensors.  Returns {the four dcs::vmap results, the eight vmap_integral columns at
// 180 points, dcs::map / pmap / pvmap of pair production}.
inline std::vector<torch::Tensor> reference_call_sites(torch::Tensor kinetic_energies,
                                                       torch::Tensor recoil_energies) {
    std::vector<torch::Tensor> out;
    {
        const auto result = torch::zeros_like(kinetic_energies);
        dcs::vmap(dcs::bremsstrahlung)(result, kinetic_energies, recoil_energies, STANDARD_ROCK,
                                       MUON_MASS);
        out.push_back(result);
    }
    {
        const auto result = torch::zeros_like(kinetic_energies);
        dcs::vmap(dcs::pair_production)(
                result,
                kinetic_energies,
                recoil_energies,
                STANDARD_ROCK, MUON_MASS);
        out.push_back(result);
    }
    {
        const auto result = torch::zeros_like(kinetic_energies);
        dcs::vmap(dcs::photonuclear)(result, kinetic_energies, recoil_energies, STANDARD_ROCK,
                                     MUON_MASS);
        out.push_back(result);
    }
    {
        const auto result = torch::zeros_like(kinetic_energies);
        dcs::vmap(dcs::ionisation)(result, kinetic_energies, recoil_energies, STANDARD_ROCK,
                                   MUON_MASS);
        out.push_back(result);
    }
#define NOA_B200_COLUMN(F, G)                                                                     \
    {                                                                                             \
        const auto result = torch::zeros_like(kinetic_energies);                                  \
        dcs::vmap_integral(                                                                       \
                dcs::recoil_integral(dcs::F, dcs::G))(                                            \
                result,                                                                           \
                kinetic_energies,                                                                 \
                dcs::X_FRACTION, STANDARD_ROCK, MUON_MASS, 180);                                  \
        out.push_back(result);                                                                    \
    }
    NOA_B200_COLUMN(bremsstrahlung, del_integrand)
    NOA_B200_COLUMN(bremsstrahlung, cel_integrand)
    NOA_B200_COLUMN(pair_production, del_integrand)
    NOA_B200_COLUMN(pair_production, cel_integrand)
    NOA_B200_COLUMN(photonuclear, del_integrand)
    NOA_B200_COLUMN(photonuclear, cel_integrand)
    NOA_B200_COLUMN(ionisation, del_integrand)
    NOA_B200_COLUMN(ionisation, cel_integrand)
#undef NOA_B200_COLUMN
    out.push_back(dcs::map(dcs::pair_production)(kinetic_energies, recoil_energies, STANDARD_ROCK,
                                                 MUON_MASS));
    out.push_back(dcs::pmap(dcs::pair_production)(kinetic_energies, recoil_energies, STANDARD_ROCK,
                                                  MUON_MASS));
    {
        const auto result = torch::zeros_like(kinetic_energies);
        dcs::pvmap(dcs::pair_production)(result, kinetic_energies, recoil_energies, STANDARD_ROCK,
                                         MUON_MASS);
        out.push_back(result);
    }
    return out;
}

inline void serialise(torch::Tensor tensor, std::string path) { torch::save(tensor, path); }

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
    m.def("bremsstrahlung", &bremsstrahlung, py::call_guard<py::gil_scoped_release>(),
          "Standard Rock Bremsstrahlung DCS for Muons on CUDA");
    m.def("pair_production", &pair_production, py::call_guard<py::gil_scoped_release>(),
          "Standard Rock Pair production DCS for Muons on CUDA");
    m.def("photonuclear", &photonuclear, py::call_guard<py::gil_scoped_release>(),
          "Standard Rock Photonuclear DCS for Muons on CUDA");
    m.def("ionisation", &ionisation, py::call_guard<py::gil_scoped_release>(),
          "Standard Rock Ionisation DCS for Muons on CUDA");
    m.def("all_processes", &all_processes, py::call_guard<py::gil_scoped_release>(),
          "The four DCS in one pass, [4, n]");
    m.def("tables", &tables, py::call_guard<py::gil_scoped_release>(),
          "DEL/CEL tables [2, 4, n_K] for standard rock");
    m.def("recoil_integral", &recoil_integral, py::call_guard<py::gil_scoped_release>(),
          "vmap_integral(recoil_integral(process, integrand)) for standard rock");
    m.def("water", &water, py::call_guard<py::gil_scoped_release>(),
          "The four DCS on water (H + O mass-fraction mix), [4, n]");
    m.def("coulomb_hard_scattering", &coulomb_hard_scattering,
          py::call_guard<py::gil_scoped_release>(),
          "coulomb_data + coulomb_transport(mu) + hard_scattering for standard rock");
    m.def("soft_scattering", &soft_scattering, py::call_guard<py::gil_scoped_release>(),
          "Soft-scattering transverse transport for standard rock");
    m.def("reference_call_sites", &reference_call_sites, py::call_guard<py::gil_scoped_release>(),
          "The reference's CPU call expressions (dcs::vmap(dcs::f)(...), vmap_integral(...)) on "
          "CPU or CUDA tensors");
    m.def("serialise", &serialise, py::call_guard<py::gil_scoped_release>(), "Save tensor to disk");
}
