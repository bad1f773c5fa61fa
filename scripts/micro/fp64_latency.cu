// Micro-benchmark: FP64 dependent-issue latency and single-warp throughput on sm_100a (one warp, one CTA).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/fp64_latency scripts/micro/fp64_latency.cu && /tmp/fp64_latency
#include <cstdio>
#include <cuda_runtime.h>

template <int CHAINS>
__global__ void k_dfma(double *out, long long *cyc, double a, double b, int iters) {
    double x[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) x[c] = threadIdx.x + c;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int c = 0; c < CHAINS; ++c) x[c] = fma(a, x[c], b);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s += x[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// same, but the multiplier and/or the addend are (loop-invariant) per-thread registers instead of uniform operands:
// REGS = 2: fma(y, x, b), REGS = 3: fma(y, x, z)
template <int CHAINS, int REGS>
__global__ void k_dfma_regs(double *out, long long *cyc, double a, double b, int iters) {
    double x[CHAINS], y[CHAINS], z[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) {
        x[c] = threadIdx.x + c;
        y[c] = a + 1e-9 * (threadIdx.x + c);
        z[c] = b + 1e-9 * (threadIdx.x * 3 + c);
    }
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int c = 0; c < CHAINS; ++c) x[c] = REGS == 3 ? fma(y[c], x[c], z[c]) : fma(y[c], x[c], b);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s += x[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int CHAINS, int REGS>
void run_regs(int warps) {
    double *out;
    long long *cyc;
    cudaMalloc(&out, sizeof(double) * 1024);
    cudaMalloc(&cyc, sizeof(long long));
    const int iters = 2000;
    k_dfma_regs<CHAINS, REGS><<<1, 32 * warps>>>(out, cyc, 0.999, 1e-3, iters);
    k_dfma_regs<CHAINS, REGS><<<1, 32 * warps>>>(out, cyc, 0.999, 1e-3, iters);
    long long h;
    cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    printf("DFMA with %d register operands, chains=%2d warps/CTA=%2d: %.2f cycles per DFMA per warp\n", REGS, CHAINS, warps,
           (double)h / (iters * 8.0 * CHAINS));
    cudaFree(out);
    cudaFree(cyc);
}

__global__ void k_shfl_dfma(double *out, long long *cyc, double a, int iters) {
    double x = threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            double t = __shfl_up_sync(0xffffffffu, x, 1);
            x = fma(a, t, x);
        }
    }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

__global__ void k_ffma(float *out, long long *cyc, float a, float b, int iters) {
    float x = threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 8; ++r) x = fmaf(a, x, b);
    }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

template <int CHAINS>
void run(int warps, int blocks = 1) {
    double *out;
    long long *cyc;
    cudaMalloc(&out, sizeof(double) * 1024 * blocks);
    cudaMalloc(&cyc, sizeof(long long) * blocks);
    const int iters = 2000;
    k_dfma<CHAINS><<<blocks, 32 * warps>>>(out, cyc, 0.999, 1e-3, iters);
    k_dfma<CHAINS><<<blocks, 32 * warps>>>(out, cyc, 0.999, 1e-3, iters);
    long long h;
    cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    printf("DFMA chains=%2d warps/CTA=%2d: %.2f cycles per DFMA per warp, %.2f cycles per dependent step\n", CHAINS, warps,
           (double)h / (iters * 8.0 * CHAINS), (double)h / (iters * 8.0));
    cudaFree(out);
    cudaFree(cyc);
}

int main() {
    run<1>(1);
    run<2>(1);
    run<3>(1);
    run<4>(1);
    run<6>(1);
    run<8>(1);
    run<11>(1);
    run<16>(1);
    run<1>(4);
    run<1>(8);
    run<1>(16);
    run<3>(4);
    run<3>(8);
    run<3>(12);
    run<3>(16);
    run<8>(4);
    run<8>(8);
    run<8>(16);
    run_regs<8, 2>(4);
    run_regs<8, 3>(4);
    run_regs<8, 2>(8);
    run_regs<8, 3>(8);
    run_regs<8, 3>(12);
    run_regs<3, 3>(12);
    double *out;
    long long *cyc;
    cudaMalloc(&out, 8 * 1024);
    cudaMalloc(&cyc, 8);
    long long h;
    k_shfl_dfma<<<1, 32>>>(out, cyc, 0.5, 2000);
    k_shfl_dfma<<<1, 32>>>(out, cyc, 0.5, 2000);
    cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    printf("SHFL.up(double) + DFMA dependent pair: %.2f cycles\n", (double)h / (2000 * 8.0));
    k_ffma<<<1, 32>>>((float *)out, cyc, 0.5f, 0.1f, 2000);
    cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    printf("FFMA dependent: %.2f cycles\n", (double)h / (2000 * 8.0));
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
