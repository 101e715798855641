// measure_dcs_calc_cuda -- the reference's benchmark/measure-dcs-calc{,-cuda}.cc cases on the B200
// path, through the C++ LibTorch boundary (noa::pms::dcs::cuda::*), printing Google-Benchmark-style
// console lines so the output can be diffed against upstream's `measure_dcs_calc`.
//
// Case names follow the reference (benchmark/measure-dcs-calc.cc:10-139,
// benchmark/measure-dcs-calc-cuda.cc:10-30) with its CUDA suffix:
//   <Process>VectorisedCUDA / <Process>VectorisedLargeCUDA        dcs::cuda::vmap_<process>
//   DEL<Process>VectorisedCUDA / CEL<Process>VectorisedCUDA       vmap_integral(recoil_integral), 180 nodes
//   CoulombHardScatteringCUDA / CoulombSoftScatteringCUDA         test/unit/test-dcs-calc.cc:134-178
// Google Benchmark and the reference's input tensors (noa-test-data) cannot be fetched offline, so
// this file carries a small timing loop of its own and uses the notebook grid
// (docs/pms/muon_dcs_calc.ipynb:174: K = linspace(1e-3, 1e6, 10000), q = 0.0505 K); "Large" is that
// grid repeat_interleave(1000), exactly as measure-dcs-calc.hh:41-42 does with its own data.
// Every iteration is timed with CUDA events around the call (device time, what the reference's
// un-synchronised loop does NOT measure) and, in the CPU column, host wall time per call.
//
// Build: noa_b200/csrc/build_torch_ext.py (g++, links libnoa_dcs_b200_torch.so).
#include "../include/noa_b200/pms_dcs_cuda.hh"

#include <c10/cuda/CUDAStream.h>
#include <cuda_runtime_api.h>
#include <torch/torch.h>

#include <chrono>
#include <cstdio>
#include <cstring>
#include <functional>
#include <string>
#include <vector>

using namespace noa::pms;

namespace {
    struct Case {
        std::string name;
        std::function<void()> body;
    };

    void run(const Case &c, double min_seconds) {
        cudaStream_t stream = c10::cuda::getCurrentCUDAStream().stream();
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        for (int i = 0; i < 3; i++) c.body();
        cudaStreamSynchronize(stream);
        int64_t iters = 0;
        double dev_ns = 0., cpu_ns = 0.;
        const auto begin = std::chrono::steady_clock::now();
        while (true) {
            const auto t0 = std::chrono::steady_clock::now();
            cudaEventRecord(e0, stream);
            c.body();
            cudaEventRecord(e1, stream);
            const auto t1 = std::chrono::steady_clock::now();
            cudaEventSynchronize(e1);
            float ms = 0.f;
            cudaEventElapsedTime(&ms, e0, e1);
            dev_ns += ms * 1e6;
            cpu_ns += std::chrono::duration<double, std::nano>(t1 - t0).count();
            iters++;
            const double spent =
                    std::chrono::duration<double>(std::chrono::steady_clock::now() - begin).count();
            if ((spent > min_seconds && iters >= 10) || iters >= 100000) break;
        }
        std::printf("%-58s %12.0f ns %12.0f ns %12lld\n", ("DCSBenchmark/" + c.name).c_str(),
                    dev_ns / iters, cpu_ns / iters, (long long) iters);
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
    }
}  // namespace

int main(int argc, char **argv) {
    double min_seconds = 0.2;
    std::string filter;
    for (int i = 1; i < argc; i++) {
        if (!std::strncmp(argv[i], "--benchmark_min_time=", 21)) min_seconds = std::atof(argv[i] + 21);
        if (!std::strncmp(argv[i], "--benchmark_filter=", 19)) filter = argv[i] + 19;
    }
    if (!torch::cuda::is_available()) {
        std::fprintf(stderr, "measure_dcs_calc_cuda needs a CUDA device (there is no CPU path)\n");
        return 2;
    }
    const auto opt = torch::dtype(torch::kFloat64).device(torch::kCUDA);
    const auto K = torch::linspace(1e-3, 1e6, 10000, opt);
    const auto q = 0.0505 * K;
    const auto r = torch::zeros_like(K);
    const auto KL = K.repeat_interleave(1000), qL = q.repeat_interleave(1000);
    const auto rL = torch::zeros_like(KL);
    const auto element = STANDARD_ROCK;
    const auto mu = MUON_MASS;

    std::vector<Case> cases;
    struct Proc {
        const char *name;
        int id;
        void (*vmap)(const Calculation &, const Energies &, const Energies &, const AtomicElement &,
                     const ParticleMass &);
    };
    const Proc procs[] = {{"Bremsstrahlung", 0, dcs::cuda::vmap_bremsstrahlung},
                          {"PairProduction", 1, dcs::cuda::vmap_pair_production},
                          {"Photonuclear", 2, dcs::cuda::vmap_photonuclear},
                          {"Ionisation", 3, dcs::cuda::vmap_ionisation}};
    for (const auto &p : procs) {
        cases.push_back({std::string(p.name) + "VectorisedCUDA",
                         [=] { p.vmap(r, K, q, element, mu); }});
        cases.push_back({std::string(p.name) + "VectorisedLargeCUDA",
                         [=] { p.vmap(rL, KL, qL, element, mu); }});
        cases.push_back({std::string("DEL") + p.name + "VectorisedCUDA", [=] {
                             dcs::cuda::vmap_integral(p.id, 0, r, K, dcs::X_FRACTION, element, mu, 180);
                         }});
        cases.push_back({std::string("CEL") + p.name + "VectorisedCUDA", [=] {
                             dcs::cuda::vmap_integral(p.id, 1, r, K, dcs::X_FRACTION, element, mu, 180);
                         }});
    }
    cases.push_back({"AllProcessesVectorisedLargeCUDA",
                     [=] { (void) dcs::cuda::map_all(KL, qL, element, mu); }});
    cases.push_back({"TablesVectorisedCUDA",
                     [=] { (void) dcs::cuda::tables(K, dcs::X_FRACTION, element, mu, 180); }});
    {
        const int64_t n = K.numel();
        const auto fCM = torch::zeros({n, 2}, opt), screen = torch::zeros({n, 9}, opt);
        const auto fspin = torch::zeros_like(K), invlambda = torch::zeros_like(K);
        const auto G = torch::zeros({n, 2}, opt), mu0 = torch::zeros_like(K), lb_h = torch::zeros_like(K);
        const auto one = torch::ones({1}, opt);
        cases.push_back({"CoulombHardScatteringCUDA", [=] {
                             dcs::cuda::coulomb_data(fCM, screen, fspin, invlambda, K, element, mu);
                             dcs::cuda::coulomb_transport(G, screen, fspin, one);
                             dcs::cuda::hard_scattering(mu0, lb_h, G.view({1, n, 2}),
                                                        fCM.view({1, n, 2}), screen.view({1, n, 9}),
                                                        invlambda.view({1, n}), fspin.view({1, n}));
                         }});
        cases.push_back({"CoulombSoftScatteringCUDA",
                         [=] { dcs::cuda::soft_scattering(r, K, element, mu); }});
    }

    cudaDeviceProp prop{};
    cudaGetDeviceProperties(&prop, 0);
    std::printf("Run on (%s, %d SMs)\n", prop.name, prop.multiProcessorCount);
    std::printf("%s\n", std::string(100, '-').c_str());
    std::printf("%-58s %15s %15s %12s\n", "Benchmark", "Time", "CPU", "Iterations");
    std::printf("%s\n", std::string(100, '-').c_str());
    for (const auto &c : cases)
        if (filter.empty() || c.name.find(filter) != std::string::npos) run(c, min_seconds);
    return 0;
}
