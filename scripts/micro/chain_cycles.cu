// Where the cycles of the sequential coarsest-level solve go: k_chain<Heat1D<32,33>> with one team, instrumented
// with the MGB_T probes of common.cuh / sweeps.cuh.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -DMGB_CYCLES -Ipymgrit_b200/csrc \
//        -o scripts/micro/chain_cycles.bin scripts/micro/chain_cycles.cu -Lpymgrit_b200/lib -lmgrit_b200
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../include/mgrit_b200.h"
#include "sweeps.cuh"

using namespace mgb;
using Phi = Heat1D<32, 33>;

int main(int argc, char **argv) {
    const int npts = argc > 1 ? atoi(argv[1]) : 513, n = 1023, pitch = 1024;
    const int with_g = argc > 2 ? atoi(argv[2]) : 1, nin = argc > 3 ? atoi(argv[3]) : 4;
    const int cw = mgb_step_consts_width(MGB_APP_HEAT1D, 32, 33);
    std::vector<double> sc(cw), host((size_t)npts * pitch);
    mgb_heat1d_step_consts(2.0 * 2048, n, 32, 33, sc.data());
    for (size_t k = 0; k < host.size(); ++k) host[k] = (k % pitch) < (size_t)n ? 1e-3 * ((k * 2654435761u) % 1000) : 0.0;
    LevelDev L = {};
    double *u, *g, *scd, *rx, *rt;
    cudaMalloc(&u, host.size() * 8);
    cudaMalloc(&g, host.size() * 8);
    cudaMalloc(&scd, cw * 8);
    cudaMalloc(&rx, 33 * 32 * 8);
    cudaMalloc(&rt, npts * 8);
    cudaMemcpy(u, host.data(), host.size() * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(g, host.data(), host.size() * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(scd, sc.data(), cw * 8, cudaMemcpyHostToDevice);
    cudaMemset(rx, 0, 33 * 32 * 8);
    cudaMemset(rt, 0, npts * 8);
    L.u = u;
    L.g = with_g ? g : nullptr;
    L.npts = npts;
    L.n = n;
    L.pitch = pitch;
    L.ndt = 1;
    L.cw = cw;
    L.sconst = scd;
    L.nrhs = 1;
    L.rhs_x = rx;
    L.rhs_t = rt;
    L.nsys = 1;
    L.tile = pitch;
    L.nrow = n;
    const size_t smem = kHeaderBytes + (size_t)(nin + 1) * Phi::SH::SLOT_BYTES;
    cudaFuncSetAttribute(k_chain<Phi>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    long long zero[24] = {0}, cyc[24];
    const char *names[24] = {"pop: mbarrier wait", "pop: LDS + select + sync", "pop: issue next load (thread 0)",
                             "push: wait_group.read + sync", "push: STS", "push: fence.proxy.async + sync",
                             "push: bulk store issue", "", "advance: step consts / dense rhs", "advance: Phi",
                             "advance: pop g + add"};
    for (int rep = 0; rep < 2; ++rep) {
        cudaMemcpyToSymbol(g_cyc, zero, sizeof zero);
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        cudaEventRecord(e0);
        k_chain<Phi><<<1, 32, smem>>>(L, 1, 0, nin);
        cudaEventRecord(e1);
        cudaDeviceSynchronize();
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        cudaMemcpyFromSymbol(cyc, g_cyc, sizeof cyc);
        if (rep == 1) {
            printf("%d points, g=%d, nin=%d: %.1f us, %.3f us per step (%s)\n", npts, with_g, nin, ms * 1e3,
                   ms * 1e3 / (npts - 1), cudaGetErrorString(cudaGetLastError()));
            for (int k = 0; k < 11; ++k)
                if (names[k][0]) printf("  %-36s %8.1f cycles per step\n", names[k], (double)cyc[k] / (npts - 1));
        }
    }
    return 0;
}
