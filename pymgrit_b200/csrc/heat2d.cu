// heat2d.cu -- where Heat2D values enter or leave the solver: the orthonormal sine transforms between the reference's
// (nx, ny) node array (heat/heat_2d.py:20-136, boundary nodes included) and the level-row layout of phi.cuh's Heat2D:
//
//     row = [ Sx U_int Sy  (nx-2)(ny-2) coefficients, index k (ny-2) + l | zero padding to a whole tile |
//             boundary values: i = 0 row, i = nx-1 row, j = 0 column (i = 1..nx-2), j = ny-1 column | zero padding ]
//
// S_n[j][k] = sqrt(2/(n+1)) sin(pi (j+1)(k+1)/(n+1)) is symmetric and its own inverse, so both directions are the same
// two matrix products (first along y, then along x) and only differ in where the node values live.  The products are
// plain FP64 tiled GEMMs: they run once per value entering/leaving (initial condition, right-hand-side factors,
// get_values()), never inside a sweep.
#include "../../include/mgrit_b200.h"
#include "phi.cuh"
#include "table.h"

namespace mgb {

int heat2d_fail(const char *msg);  // api.cu

constexpr int BM = 64, BN = 64, BK = 16;

// C[b] (M x N, ldc) = A[b] (M x K, lda) * B[b] (K x N, ldb); 256 threads, 4 x 4 outputs per thread
__global__ void __launch_bounds__(256) k_dgemm(int M, int N, int K, const double *__restrict__ A, int lda, long sA,
                                               const double *__restrict__ B, int ldb, long sB, double *__restrict__ Cm,
                                               int ldc, long sC) {
    __shared__ double As[BK][BM + 1];
    __shared__ double Bs[BK][BN];
    A += (long)blockIdx.z * sA;
    B += (long)blockIdx.z * sB;
    Cm += (long)blockIdx.z * sC;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    double acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
    for (int k0 = 0; k0 < K; k0 += BK) {
        // A tile: BM x BK, read along k (contiguous)
        for (int q = threadIdx.x; q < BM * BK; q += 256) {
            const int mm = q / BK, kk = q % BK;
            const int m = m0 + mm, k = k0 + kk;
            As[kk][mm] = (m < M && k < K) ? A[(long)m * lda + k] : 0.0;
        }
        for (int q = threadIdx.x; q < BK * BN; q += 256) {
            const int kk = q / BN, nn = q % BN;
            const int k = k0 + kk, n = n0 + nn;
            Bs[kk][nn] = (k < K && n < N) ? B[(long)k * ldb + n] : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            double a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[kk][ty + 16 * i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx + 16 * j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty + 16 * i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx + 16 * j;
            if (n < N) Cm[(long)m * ldc + n] = acc[i][j];
        }
    }
}

__device__ __forceinline__ long boundary_node(int q, int nx, int ny) {
    if (q < ny) return q;                                  // i = 0
    if (q < 2 * ny) return (long)(nx - 1) * ny + (q - ny);  // i = nx-1
    q -= 2 * ny;
    if (q < nx - 2) return (long)(q + 1) * ny;  // j = 0
    q -= nx - 2;
    return (long)(q + 1) * ny + (ny - 1);  // j = ny-1
}

// to_row != 0: rows[b][boundary part] = node values and both paddings = 0;  else: node array boundary = row values
__global__ void k_heat2d_boundary(int nx, int ny, int nint, int nint_pad, int pitch, double *__restrict__ phys, long sP,
                                  double *__restrict__ rows, long sR, int to_row) {
    double *P = phys + (long)blockIdx.y * sP;
    double *R = rows + (long)blockIdx.y * sR;
    const int nb = 2 * ny + 2 * (nx - 2);
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < pitch; q += gridDim.x * blockDim.x) {
        if (q >= nint && q < nint_pad) {
            if (to_row) R[q] = 0.0;
        } else if (q >= nint_pad) {
            const int b = q - nint_pad;
            if (b < nb) {
                const long node = boundary_node(b, nx, ny);
                if (to_row)
                    R[q] = P[node];
                else
                    P[node] = R[q];
            } else if (to_row) {
                R[q] = 0.0;
            }
        }
    }
}

int launch_gemm(int M, int N, int K, const double *A, int lda, long sA, const double *B, int ldb, long sB, double *C,
                       int ldc, long sC, int count, cudaStream_t st) {
    dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM, count);
    k_dgemm<<<grid, 256, 0, st>>>(M, N, K, A, lda, sA, B, ldb, sB, C, ldc, sC);
    return cuda_fail(cudaGetLastError(), "heat2d transform");
}

}  // namespace mgb

using namespace mgb;

extern "C" {

int mgb_heat2d_layout(int32_t nx, int32_t ny, int32_t *tile, int32_t *nsys, int32_t *first_boundary_sys, int32_t *pitch) {
    if (nx < 3 || ny < 3 || tile == nullptr || nsys == nullptr || first_boundary_sys == nullptr || pitch == nullptr)
        return heat2d_fail("heat2d layout: need nx, ny >= 3");
    int32_t T = 0, E = 0;
    if (int rc = mgb_team_shape(MGB_APP_HEAT2D, 1, &T, &E)) return rc;
    const long tl = (long)T * E;
    const long nint = (long)(nx - 2) * (ny - 2), nb = 2L * ny + 2L * (nx - 2);
    const long si = (nint + tl - 1) / tl, sb = (nb + tl - 1) / tl;
    if ((si + sb) * tl > 0x7fffffffL) return heat2d_fail("heat2d layout: grid too large");
    *tile = (int32_t)tl;
    *nsys = (int32_t)(si + sb);
    *first_boundary_sys = (int32_t)si;
    *pitch = (int32_t)((si + sb) * tl);
    return MGB_OK;
}

static int heat2d_args(int32_t nx, int32_t ny, const void *a, const void *b, const void *c, const void *d, const void *w,
                       int32_t count, int32_t *tile, int32_t *first_b, int32_t *pitch) {
    if (a == nullptr || b == nullptr || c == nullptr || d == nullptr || w == nullptr || count < 1)
        return heat2d_fail("heat2d transform: bad argument");
    if (device_info() == nullptr) return MGB_ECUDA;
    int32_t nsys;
    return mgb_heat2d_layout(nx, ny, tile, &nsys, first_b, pitch);
}

int mgb_heat2d_to_rows(int32_t nx, int32_t ny, const double *sx_dev, const double *sy_dev, const double *nodes_dev,
                       double *rows_dev, int32_t count, double *work_dev, void *stream) {
    int32_t tile, fb, pitch;
    if (int rc = heat2d_args(nx, ny, sx_dev, sy_dev, nodes_dev, rows_dev, work_dev, count, &tile, &fb, &pitch)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int mx = nx - 2, my = ny - 2;
    const long nn = (long)nx * ny, ni = (long)mx * my;
    // work[b] = U_int[b] * Sy          (mx x my) = (mx x my, lda = ny) (my x my)
    if (int rc = launch_gemm(mx, my, my, nodes_dev + ny + 1, ny, nn, sy_dev, my, 0, work_dev, my, ni, count, st)) return rc;
    // rows[b] = Sx * work[b]
    if (int rc = launch_gemm(mx, my, mx, sx_dev, mx, 0, work_dev, my, ni, rows_dev, my, pitch, count, st)) return rc;
    dim3 grid(64, count);
    k_heat2d_boundary<<<grid, 256, 0, st>>>(nx, ny, (int)ni, fb * tile, pitch, const_cast<double *>(nodes_dev), nn, rows_dev,
                                            pitch, 1);
    return cuda_fail(cudaGetLastError(), "heat2d_to_rows");
}

int mgb_heat2d_from_rows(int32_t nx, int32_t ny, const double *sx_dev, const double *sy_dev, const double *rows_dev,
                         double *nodes_dev, int32_t count, double *work_dev, void *stream) {
    int32_t tile, fb, pitch;
    if (int rc = heat2d_args(nx, ny, sx_dev, sy_dev, rows_dev, nodes_dev, work_dev, count, &tile, &fb, &pitch)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int mx = nx - 2, my = ny - 2;
    const long nn = (long)nx * ny, ni = (long)mx * my;
    if (int rc = launch_gemm(mx, my, mx, sx_dev, mx, 0, rows_dev, my, pitch, work_dev, my, ni, count, st)) return rc;
    if (int rc = launch_gemm(mx, my, my, work_dev, my, ni, sy_dev, my, 0, nodes_dev + ny + 1, ny, nn, count, st)) return rc;
    dim3 grid(64, count);
    k_heat2d_boundary<<<grid, 256, 0, st>>>(nx, ny, (int)ni, fb * tile, pitch, nodes_dev, nn, const_cast<double *>(rows_dev),
                                            pitch, 0);
    return cuda_fail(cudaGetLastError(), "heat2d_from_rows");
}

}  // extern "C"
