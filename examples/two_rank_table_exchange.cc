// Two ranks, two GPUs, no Python, no NCCL, no torch: the fused table build + NVLink exchange of
// the C ABI (noa_dcs_table_exchange_f64, include/noa_dcs_b200.h) driven by a plain C++ host with
// CUDA IPC.  Each rank builds the DEL/CEL rows of its cyclic share of the energies and ends with
// the COMPLETE [2][4][n] table of standard rock in its own memory; the result is compared, bit for
// bit, with a single-GPU build (noa_dcs_table_f64) of all rows.
//
//   g++ -std=c++17 -O2 examples/two_rank_table_exchange.cc -Iinclude -I/usr/local/cuda/include \
//       -Lnoa_b200 -lnoa_dcs_b200 -L/usr/local/cuda/lib64 -lcudart -Wl,-rpath,'$ORIGIN' \
//       -o noa_b200/two_rank_table_exchange          (noa_b200/csrc/Makefile: `make example`)
//
// What a multi-process C++ host does: allocate [table 0 | table 1 | flag words]
// per rank, swap cudaIpcMemHandle_t with the peers, open them, and from then on every build is ONE
// call per rank with an increasing epoch; two tables alternate by epoch parity.
#include <cuda_runtime.h>
#include <sys/socket.h>
#include <sys/wait.h>
#include <unistd.h>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include "noa_dcs_b200.h"

#define CHECK(call)                                                                      \
    do {                                                                                 \
        const int rc_ = (int) (call);                                                    \
        if (rc_ != 0) {                                                                  \
            std::fprintf(stderr, "rank %d: %s -> %d (%s)\n", rank, #call, rc_,           \
                         noa_dcs_strerror(rc_));                                         \
            return 1;                                                                    \
        }                                                                                \
    } while (0)

static bool swap_bytes(int sock, const void *mine, void *theirs, size_t n) {
    return write(sock, mine, n) == (ssize_t) n && read(sock, theirs, n) == (ssize_t) n;
}

static int run_rank(int rank, int sock) {
    const int world = 2;
    const int64_t n = 1000;                       // energies 1e-2 .. 1e6 GeV
    const double A = 22., I = 0.1364E-6, mass = 0.10565839, xlow = 0.05;
    const int32_t Z = 11, min_points = 180;
    CHECK(cudaSetDevice(rank));
    std::vector<double> K(n), K_local;
    for (int64_t i = 0; i < n; i++) K[i] = std::pow(10., -2. + 8. * (double) i / (double) (n - 1));
    for (int64_t i = rank; i < n; i += world) K_local.push_back(K[i]);

    // [table 0 | table 1 | 16 flag words], one allocation per rank, shared with the peer
    const size_t table_doubles = 2 * 4 * n, bytes = (2 * table_doubles + 8) * sizeof(double);
    double *mine = nullptr, *peer = nullptr;
    CHECK(cudaMalloc(&mine, bytes));
    CHECK(cudaMemset(mine, 0, bytes));
    CHECK(cudaDeviceSynchronize());
    cudaIpcMemHandle_t my_handle, peer_handle;
    CHECK(cudaIpcGetMemHandle(&my_handle, mine));
    if (!swap_bytes(sock, &my_handle, &peer_handle, sizeof(my_handle))) return 1;
    CHECK(cudaIpcOpenMemHandle((void **) &peer, peer_handle, cudaIpcMemLazyEnablePeerAccess));
    char token = 1;                               // both buffers are zeroed before anyone writes
    if (!swap_bytes(sock, &token, &token, 1)) return 1;

    double *bases[2] = {rank == 0 ? mine : peer, rank == 0 ? peer : mine};   // by rank
    double *d_K_local = nullptr, *d_K = nullptr, *d_ref = nullptr;
    uint32_t *d_sync = nullptr;
    CHECK(cudaMalloc(&d_K_local, K_local.size() * sizeof(double)));
    CHECK(cudaMalloc(&d_K, n * sizeof(double)));
    CHECK(cudaMalloc(&d_ref, table_doubles * sizeof(double)));
    const size_t sync_words = 8;                          // see noa_dcs_table_exchange_f64
    CHECK(cudaMalloc(&d_sync, sync_words * sizeof(uint32_t)));
    CHECK(cudaMemset(d_sync, 0, sync_words * sizeof(uint32_t)));
    // workspace of the flat build: 16 B per node and process of the local rows
    const int64_t scratch_doubles =
            noa_dcs_table_workspace_doubles((int64_t) K_local.size(), min_points);
    double *d_scratch = nullptr;
    CHECK(cudaMalloc(&d_scratch, scratch_doubles * sizeof(double)));
    CHECK(cudaMemcpy(d_K_local, K_local.data(), K_local.size() * sizeof(double),
                     cudaMemcpyHostToDevice));
    CHECK(cudaMemcpy(d_K, K.data(), n * sizeof(double), cudaMemcpyHostToDevice));
    cudaStream_t stream;
    CHECK(cudaStreamCreate(&stream));
    // the single-GPU table every rank must end up with
    CHECK(noa_dcs_table_f64(0xF, d_K, n, xlow, min_points, A, I, Z, mass, d_ref, d_ref + 4 * n,
                            stream));

    std::vector<double> got(table_doubles), want(table_doubles);
    int bad = 0;
    for (uint32_t epoch = 1; epoch <= 4; epoch++) {
        const size_t off = (epoch & 1) * table_doubles;
        double *del[2], *cel[2];
        uint32_t *flags[2];
        for (int r = 0; r < world; r++) {
            del[r] = bases[r] + off;
            cel[r] = bases[r] + off + 4 * n;
            flags[r] = (uint32_t *) (bases[r] + 2 * table_doubles);
        }
        CHECK(noa_dcs_table_exchange_f64(0xF, d_K_local, (int64_t) K_local.size(), xlow,
                                         min_points, A, I, Z, mass, world, rank, del, cel, flags,
                                         /*multicast_del=*/nullptr, /*multicast_cel=*/nullptr,
                                         /*multicast_flags=*/nullptr,
                                         d_sync, d_scratch, scratch_doubles, epoch, n,
                                         /*first_row=*/rank,
                                         /*row_stride=*/world, /*timeout_seconds=*/20., stream));
        CHECK(cudaStreamSynchronize(stream));
        CHECK(cudaMemcpy(got.data(), mine + off, table_doubles * sizeof(double),
                         cudaMemcpyDeviceToHost));
        CHECK(cudaMemcpy(want.data(), d_ref, table_doubles * sizeof(double),
                         cudaMemcpyDeviceToHost));
        const bool same = std::memcmp(got.data(), want.data(), table_doubles * sizeof(double)) == 0;
        std::printf("rank %d epoch %u: exchanged table %s the single-GPU build\n", rank, epoch,
                    same ? "equals" : "DIFFERS FROM");
        bad += !same;
    }
    if (!swap_bytes(sock, &token, &token, 1)) return 1;   // peer is done with my memory too
    cudaIpcCloseMemHandle(peer);
    cudaFree(mine);
    return bad;
}

int main() {
    int socks[2];
    if (socketpair(AF_UNIX, SOCK_STREAM, 0, socks) != 0) return 2;
    pid_t pids[2];
    for (int rank = 0; rank < 2; rank++) {       // fork before any CUDA call
        pids[rank] = fork();
        if (pids[rank] == 0) {
            const int rc = run_rank(rank, socks[rank]);
            std::fflush(stdout);
            std::fflush(stderr);
            _exit(rc);
        }
    }
    int failed = 0;
    for (int rank = 0; rank < 2; rank++) {
        int status = 0;
        waitpid(pids[rank], &status, 0);
        failed += !(WIFEXITED(status) && WEXITSTATUS(status) == 0);
    }
    std::printf(failed ? "FAILED\n" : "OK\n");
    return failed;
}
