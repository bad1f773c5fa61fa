// generic.cu -- MGRIT sweeps for applications whose Phi is not one of the fused team kernels: the "batched" path.
//
// The reference's plug-in idea is Application.step (core/application.py:99): any time integrator.  A Phi that does not fit
// the register-resident team kernels of sweeps.cuh (a 2-D periodic solve, a user's own device code) still runs all
// coarse intervals of a sweep at once: the host walks the positions inside an interval, and for every position ONE
// batched Phi call advances every interval of the level (Application.step_rows), followed by the row-wise combinations
// below.  The arithmetic per row is the reference's, operation by operation (mgrit.py:312-327, 354-368, 497-547, 722-726).
//
// Also here: the batched Phi of the Allen-Cahn IMEX scheme (allen_cahn/allen_cahn.py:191-197), which needs a 2-D periodic
// solve per step.
#include "../../include/mgrit_b200.h"
#include "phi.cuh"
#include "table.h"

namespace mgb {

int heat2d_fail(const char *msg);  // api.cu: records the message, returns MGB_EINVAL
int launch_gemm(int M, int N, int K, const double *A, int lda, long sA, const double *B, int ldb, long sB, double *C,
                int ldc, long sC, int count, cudaStream_t st);  // heat2d.cu: batched FP64 product

__device__ __forceinline__ const double *row_of(const double *base, int pitch, const int *idx, int k) {
    return base + (size_t)(idx ? idx[k] : k) * pitch;
}

// out[oi[k]] = ((a X[xi[k]] + b Y[yi[k]]) + c Z[zi[k]]), products and sums rounded one by one like NumPy does; Y and Z may
// be null.  One CTA per row (grid-stride over rows).
__global__ void __launch_bounds__(256) k_rows_lincomb(int count, int n, double *__restrict__ out, int out_pitch,
                                                      const int *__restrict__ oi, double a, const double *X, int x_pitch,
                                                      const int *__restrict__ xi, double b, const double *Y, int y_pitch,
                                                      const int *__restrict__ yi, double c, const double *Z, int z_pitch,
                                                      const int *__restrict__ zi, const int *__restrict__ stop) {
    if (stop != nullptr && *stop != 0) return;
    for (int k = blockIdx.x; k < count; k += gridDim.x) {
        double *o = out + (size_t)(oi ? oi[k] : k) * out_pitch;
        const double *x = row_of(X, x_pitch, xi, k);
        const double *y = Y ? row_of(Y, y_pitch, yi, k) : nullptr;
        const double *z = Z ? row_of(Z, z_pitch, zi, k) : nullptr;
        for (int q = threadIdx.x; q < n; q += blockDim.x) {
            double v = (a == 1.0) ? x[q] : __dmul_rn(a, x[q]);
            if (y) v = __dadd_rn(v, (b == 1.0) ? y[q] : (b == -1.0 ? -y[q] : __dmul_rn(b, y[q])));
            if (z) v = __dadd_rn(v, (c == 1.0) ? z[q] : (c == -1.0 ? -z[q] : __dmul_rn(c, z[q])));
            o[q] = v;
        }
    }
}

// out_sq[k] = sum_q X[xi[k]][q]^2, fixed summation order (deterministic)
__global__ void __launch_bounds__(256) k_rows_sumsq(int count, int n, const double *__restrict__ X, int pitch,
                                                    const int *__restrict__ xi, double *__restrict__ out_sq,
                                                    const int *__restrict__ stop) {
    if (stop != nullptr && *stop != 0) return;
    __shared__ double s[8];
    for (int k = blockIdx.x; k < count; k += gridDim.x) {
        const double *x = row_of(X, pitch, xi, k);
        double acc = 0.0;
        for (int q = threadIdx.x; q < n; q += blockDim.x) acc = fma(x[q], x[q], acc);
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
        __syncthreads();
        if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            double tot = 0.0;
            for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += s[w];
            out_sq[k] = tot;
        }
    }
}

// ---- Allen-Cahn, IMEX (allen_cahn.py:191-197) --------------------------------------------------------------------
//   rhs = u + dt (1/eps^2) u (1 - u^nu);   (I - dt L) y = rhs,   L = 5-point Laplacian, periodic, on nx x nx nodes.
// L = L1 (x) I + I (x) L1 with the circulant second difference L1 = Q diag(-mu) Q^T, Q the real orthonormal Fourier
// basis (built by the host): y = Q [ (Q^T rhs Q) / (1 + dt (mu_i + mu_j)) ] Q^T -- four small dense products per step,
// batched over all the steps of a call.
__global__ void __launch_bounds__(256) k_ac_rhs(int count, int nn, const double *__restrict__ src, int src_pitch,
                                                const int *__restrict__ si, const double *__restrict__ dt, double inv_eps2,
                                                int nu, double *__restrict__ w, const int *__restrict__ stop) {
    if (stop != nullptr && *stop != 0) return;
    for (int k = blockIdx.y; k < count; k += gridDim.y) {
        const double *u = row_of(src, src_pitch, si, k);
        const double h = dt[k];
        for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < nn; q += gridDim.x * blockDim.x) {
            const double v = u[q];
            double p = v;  // v ** nu for the integer nu >= 1 of the reference (nu = 2: NumPy squares, v * v)
            for (int e = 1; e < nu; ++e) p = __dmul_rn(p, v);
            // new + dt * (1 / eps^2 * new * (1 - new^nu)), evaluated left to right as NumPy does
            const double react = __dmul_rn(__dmul_rn(inv_eps2, v), __dsub_rn(1.0, p));
            w[(size_t)k * nn + q] = __dadd_rn(v, __dmul_rn(h, react));
        }
    }
}

__global__ void __launch_bounds__(256) k_ac_scale(int count, int nx, const double *__restrict__ mu,
                                                  const double *__restrict__ dt, double *__restrict__ w,
                                                  const int *__restrict__ stop) {
    if (stop != nullptr && *stop != 0) return;
    const int nn = nx * nx;
    for (int k = blockIdx.y; k < count; k += gridDim.y) {
        const double h = dt[k];
        for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < nn; q += gridDim.x * blockDim.x) {
            const int i = q / nx, j = q - i * nx;
            w[(size_t)k * nn + q] = __ddiv_rn(w[(size_t)k * nn + q], fma(h, mu[i] + mu[j], 1.0));
        }
    }
}

__global__ void __launch_bounds__(256) k_rows_scatter(int count, int n, const double *__restrict__ w, double *__restrict__ dst,
                                                      int dst_pitch, const int *__restrict__ di,
                                                      const int *__restrict__ stop) {
    if (stop != nullptr && *stop != 0) return;
    for (int k = blockIdx.y; k < count; k += gridDim.y) {
        double *o = dst + (size_t)(di ? di[k] : k) * dst_pitch;
        for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < n; q += gridDim.x * blockDim.x) o[q] = w[(size_t)k * n + q];
    }
}

}  // namespace mgb

using namespace mgb;

extern "C" {

int mgb_rows_lincomb(int32_t count, int32_t n, double *out_dev, int32_t out_pitch, const int32_t *out_idx_dev, double a,
                     const double *x_dev, int32_t x_pitch, const int32_t *x_idx_dev, double b, const double *y_dev,
                     int32_t y_pitch, const int32_t *y_idx_dev, double c, const double *z_dev, int32_t z_pitch,
                     const int32_t *z_idx_dev, void *stream) {
    if (count < 0 || n < 1 || out_dev == nullptr || x_dev == nullptr || out_pitch < n || x_pitch < n ||
        (y_dev != nullptr && y_pitch < n) || (z_dev != nullptr && z_pitch < n))
        return heat2d_fail("rows_lincomb: bad argument");
    const DeviceInfo *di = device_info();
    if (di == nullptr) return MGB_ECUDA;
    if (count == 0) return MGB_OK;
    const int grid = count < 8 * di->sms ? count : 8 * di->sms;
    k_rows_lincomb<<<grid, 256, 0, (cudaStream_t)stream>>>(count, n, out_dev, out_pitch, out_idx_dev, a, x_dev, x_pitch,
                                                          x_idx_dev, b, y_dev, y_pitch, y_idx_dev, c, z_dev, z_pitch,
                                                          z_idx_dev, stop_flag());
    return cuda_fail(cudaGetLastError(), "rows_lincomb");
}

int mgb_rows_sumsq(int32_t count, int32_t n, const double *x_dev, int32_t pitch, const int32_t *x_idx_dev,
                   double *out_sq_dev, void *stream) {
    if (count < 0 || n < 1 || x_dev == nullptr || out_sq_dev == nullptr || pitch < n)
        return heat2d_fail("rows_sumsq: bad argument");
    const DeviceInfo *di = device_info();
    if (di == nullptr) return MGB_ECUDA;
    if (count == 0) return MGB_OK;
    const int grid = count < 8 * di->sms ? count : 8 * di->sms;
    k_rows_sumsq<<<grid, 256, 0, (cudaStream_t)stream>>>(count, n, x_dev, pitch, x_idx_dev, out_sq_dev, stop_flag());
    return cuda_fail(cudaGetLastError(), "rows_sumsq");
}

int mgb_allen_cahn_imex_rows(int32_t nx, int32_t count, const double *src_dev, int32_t src_pitch,
                             const int32_t *src_idx_dev, double *dst_dev, int32_t dst_pitch, const int32_t *dst_idx_dev,
                             const double *dt_dev, double inv_eps2, int32_t nu, const double *q_dev, const double *qt_dev,
                             const double *mu_dev, double *work1_dev, double *work2_dev, void *stream) {
    const long nn = (long)nx * nx;
    if (nx < 1 || count < 0 || nu < 1 || src_dev == nullptr || dst_dev == nullptr || dt_dev == nullptr || q_dev == nullptr ||
        qt_dev == nullptr || mu_dev == nullptr || work1_dev == nullptr || work2_dev == nullptr || src_pitch < nn ||
        dst_pitch < nn)
        return heat2d_fail("allen_cahn_imex_rows: bad argument");
    const DeviceInfo *di = device_info();
    if (di == nullptr) return MGB_ECUDA;
    if (count == 0) return MGB_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const dim3 grid((unsigned)((nn + 255) / 256 < 64 ? (nn + 255) / 256 : 64), (unsigned)(count < 4096 ? count : 4096));
    k_ac_rhs<<<grid, 256, 0, st>>>(count, (int)nn, src_dev, src_pitch, src_idx_dev, dt_dev, inv_eps2, nu, work1_dev,
                                   stop_flag());
    // work2 = Q^T work1;  work1 = work2 Q;  scale;  work2 = Q work1;  work1 = work2 Q^T;  scatter
    if (int rc = launch_gemm(nx, nx, nx, qt_dev, nx, 0, work1_dev, nx, nn, work2_dev, nx, nn, count, st)) return rc;
    if (int rc = launch_gemm(nx, nx, nx, work2_dev, nx, nn, q_dev, nx, 0, work1_dev, nx, nn, count, st)) return rc;
    k_ac_scale<<<grid, 256, 0, st>>>(count, nx, mu_dev, dt_dev, work1_dev, stop_flag());
    if (int rc = launch_gemm(nx, nx, nx, q_dev, nx, 0, work1_dev, nx, nn, work2_dev, nx, nn, count, st)) return rc;
    if (int rc = launch_gemm(nx, nx, nx, work2_dev, nx, nn, qt_dev, nx, 0, work1_dev, nx, nn, count, st)) return rc;
    k_rows_scatter<<<grid, 256, 0, st>>>(count, (int)nn, work1_dev, dst_dev, dst_pitch, dst_idx_dev, stop_flag());
    return cuda_fail(cudaGetLastError(), "allen_cahn_imex_rows");
}

}  // extern "C"
