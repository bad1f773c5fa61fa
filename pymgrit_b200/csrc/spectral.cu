// spectral.cu -- the sequential coarsest-level solve of Heat1D without its sequential Phi chain.
//
// mgrit.py:459-486 computes u_i = g_i + Phi_i(u_{i-1}), i = 1..N-1, one Phi (a tridiagonal solve, heat_1d.py:198-217)
// after the other: N-1 dependent solves on ONE spatial system, which no amount of batching can spread over the device
// and which every time rank has to wait for (the op-5 chain of mgrit.py:467-484).  Phi_i = (I + dt_i L)^-1 (. + dt_i b_i)
// with the Toeplitz matrix L = (a/dx^2) tridiag(-1, 2, -1), whose eigenvectors are the columns of the orthonormal,
// symmetric sine matrix S and whose eigenvalues lam_k = (a/dx^2) 4 sin^2(pi (k+1) / (2 (n+1))) are known in closed
// form.  With w_i = S g_i (one product for all rows), the recurrence decouples into n scalar ones,
//     uh_i[k] = (uh_{i-1}[k] + sum_q ct_q(i) rxh_q[k]) / (1 + dt_i lam_k) + w_i[k],
// that run in parallel over k, and u_i = S uh_i is again one product for all rows.  The same linear system is solved
// exactly (a direct method, like the reference's SuperLU call); only the order of the rounding errors differs (~1e-15
// relative, tests/test_gpu_parity.py holds it to the 1e-10 of BASELINE.json).
//
// The two products are plain FP64 GEMMs on the CUDA cores (no tensor cores: nothing on the path is low precision).
#include "../../include/mgrit_b200.h"
#include "phi.cuh"
#include "table.h"

namespace mgb {

int heat2d_fail(const char *msg);  // api.cu: records the message, returns MGB_EINVAL

// S[j][k] = sqrt(2/(n+1)) sin(pi (j+1)(k+1)/(n+1)), argument reduced exactly in integers before sinpi
__global__ void k_sine_matrix(int n, double *__restrict__ s, int ld) {
    const long total = (long)n * n;
    const double scale = sqrt(2.0 / (n + 1));
    const long period = 2L * (n + 1);
    for (long q = blockIdx.x * (long)blockDim.x + threadIdx.x; q < total; q += (long)gridDim.x * blockDim.x) {
        const long j = q / n, k = q - j * n;
        long r = ((j + 1) * (k + 1)) % period;  // angle = pi r / (n+1), r in [0, 2(n+1))
        double sign = 1.0;
        if (r > n + 1) {  // sin(pi + x) = -sin(x)
            r -= n + 1;
            sign = -1.0;
        }
        if (2 * r > n + 1) r = n + 1 - r;  // sin(pi - x) = sin(x): argument in [0, pi/2]
        s[j * ld + k] = sign * scale * sinpi((double)r / (double)(n + 1));
    }
}

// C (M x N, ldc) = A (M x K, lda) * B (K x N, ldb), FP64, row-major.  Row 0 of A is read from a_row0 when that is not
// null (the level's u[0] next to its g rows).  TM x 64 tile per CTA, 256 threads, (TM/16) x 4 outputs per thread, K in
// slabs of 16 through shared memory, register-prefetched: the next slab's global loads are in flight while the current
// one is multiplied.  TM is chosen by the host so that the grid covers the SMs (64 for tall A, 32/16 for a few rows).
constexpr int GN = 64, GK = 16;

template <int TM>
__global__ void __launch_bounds__(256) k_rows_gemm(int M, int N, int K, const double *__restrict__ A, int lda,
                                                   const double *__restrict__ a_row0, const double *__restrict__ B, int ldb,
                                                   double *__restrict__ Cm, int ldc, const int *__restrict__ stop) {
    if (stop != nullptr && *stop != 0) return;
    constexpr int RM = TM / 16;                 // output rows per thread
    constexpr int LA = (TM * GK + 255) / 256;   // A elements staged per thread and slab
    __shared__ double As[GK][TM + 4];
    __shared__ double Bs[GK][GN];
    const int m0 = blockIdx.y * TM, n0 = blockIdx.x * GN;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;  // columns tx + 16 j, rows ty + 16 i
    double acc[RM][4];
#pragma unroll
    for (int i = 0; i < RM; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
    double ra[LA], rb[4];
    auto load = [&](int k0) {
#pragma unroll
        for (int q = 0; q < LA; ++q) {
            const int e = threadIdx.x + 256 * q;
            const int mm = e >> 4, kk = e & 15;  // consecutive threads walk along k (contiguous in A)
            const int m = m0 + mm, k = k0 + kk;
            const double *row = (m == 0 && a_row0 != nullptr) ? a_row0 : A + (long)m * lda;
            ra[q] = (mm < TM && m < M && k < K) ? __ldg(row + k) : 0.0;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int e = threadIdx.x + 256 * q;
            const int kk = e >> 6, nn = e & 63;
            const int k = k0 + kk, n = n0 + nn;
            rb[q] = (k < K && n < N) ? __ldg(B + (long)k * ldb + n) : 0.0;
        }
    };
    load(0);
    for (int k0 = 0; k0 < K; k0 += GK) {
#pragma unroll
        for (int q = 0; q < LA; ++q) {
            const int e = threadIdx.x + 256 * q;
            if ((e >> 4) < TM) As[e & 15][e >> 4] = ra[q];
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int e = threadIdx.x + 256 * q;
            Bs[e >> 6][e & 63] = rb[q];
        }
        __syncthreads();
        if (k0 + GK < K) load(k0 + GK);
#pragma unroll
        for (int kk = 0; kk < GK; ++kk) {
            double a[RM], b[4];
#pragma unroll
            for (int i = 0; i < RM; ++i) a[i] = As[kk][ty + 16 * i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx + 16 * j];
#pragma unroll
            for (int i = 0; i < RM; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < RM; ++i) {
        const int m = m0 + ty + 16 * i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx + 16 * j;
            if (n < N) Cm[(long)m * ldc + n] = acc[i][j];
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// The same product  c[r] = a[r] S  as a fast sine transform when n + 1 = N is a power of two (nx = 2^k + 1, the usual
// grids): S is the DST-I, i.e. the imaginary part of the length-2N DFT of the odd extension
//     z = [0, x_0 .. x_{n-1}, 0, -x_{n-1} .. -x_0],   (x S)_j = -sqrt(2/N) Im(Z_{j+1}) / 2.
// One CTA per row: the 2N complex points live in shared memory, log2(2N) radix-2 decimation-in-frequency stages
// (natural order in, bit-reversed out), twiddles exp(-i pi t / N) from a table made once with sincospi (exact argument
// reduction).  O(n log n) instead of O(n^2) per row: 1025 rows of n = 1023 take 0.115 instead of 2.1 GFLOP.
// ---------------------------------------------------------------------------------------------------------------
__global__ void k_dst_twiddles(int N, double2 *__restrict__ w) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= N) return;
    double sn, cs;
    sincospi(-(double)t / (double)N, &sn, &cs);  // exp(-2 pi i t / (2N))
    w[t] = make_double2(cs, sn);
}

__global__ void __launch_bounds__(256) k_rows_dst(const int nrows, const int n, const int log2n2,
                                                  const double *__restrict__ A, const int lda,
                                                  const double *__restrict__ a_row0, const double2 *__restrict__ tw,
                                                  double *__restrict__ Cm, const int ldc, const int *__restrict__ stop) {
    if (stop != nullptr && *stop != 0) return;
    extern __shared__ double2 z[];  // 2N complex
    const int N = n + 1, N2 = 2 * N;
    const double scale = -0.5 * sqrt(2.0 / N);
    for (int r = blockIdx.x; r < nrows; r += gridDim.x) {
        const double *x = (r == 0 && a_row0 != nullptr) ? a_row0 : A + (long)r * lda;
        for (int k = threadIdx.x; k < N; k += blockDim.x) {
            const double v = (k == 0) ? 0.0 : x[k - 1];
            z[k] = make_double2(v, 0.0);
            z[(N2 - k) & (N2 - 1)] = make_double2(k == 0 ? 0.0 : -v, 0.0);
        }
        if (threadIdx.x == 0) z[N] = make_double2(0.0, 0.0);
        __syncthreads();
        for (int s = log2n2 - 1; s >= 0; --s) {
            const int half = 1 << s;
            const int tstride = N >> s;  // twiddle index step: W_{2 half}^pos = W_{2N}^(pos N / half)
            for (int b = threadIdx.x; b < N; b += blockDim.x) {
                const int pos = b & (half - 1);
                const int i = ((b >> s) << (s + 1)) + pos, j = i + half;
                const double2 p = z[i], q = z[j];
                const double2 w = tw[pos * tstride];
                const double dr = p.x - q.x, di = p.y - q.y;
                z[i] = make_double2(p.x + q.x, p.y + q.y);
                z[j] = make_double2(fma(dr, w.x, -di * w.y), fma(dr, w.y, di * w.x));
            }
            __syncthreads();
        }
        double *out = Cm + (long)r * ldc;
        for (int j = threadIdx.x; j < n; j += blockDim.x) {
            const unsigned rev = __brev((unsigned)(j + 1)) >> (32 - log2n2);
            out[j] = scale * z[rev].y;
        }
        __syncthreads();
    }
}

// The scalar recurrences of the sine-space solve, one per mode k, time-parallel:
//   U[i][k] = (U[i-1][k] + sum_q rhs_t[i][q] rxh[q][k]) / (1 + (t[i] - t[i-1]) lam[k]) + G[i][k],   i = 1 .. npts-1
// (G == U: in place on work rows that hold the transformed g; G == nullptr: no g, level 0).  A step is the affine map
// u -> A u + B with A = 1/(1 + dt lam): a CTA holds 32 modes x NCH chunks of `len` consecutive steps;
//   pass 1  every (mode, chunk) thread composes the maps of its chunk from (A, B) = (1, 0) -> shared memory
//   combine every thread pushes the start value U[0][k] through the chunks before its own (<= NCH - 1 FMAs)
//   pass 2  the chunk's steps are rerun from the true start value and stored: the values are those of the sequential
//           recurrence up to the rounding of the chunk start values.
// 2 len + NCH dependent steps instead of npts - 1: 1025 points in 32 chunks -> 96 instead of 1024.
// ends != nullptr (time rank > 0 of a time-parallel solve): the recurrence starts from 0 instead of U[0] (zero_start) and
// the kernel also returns, in ends[0][k] and ends[1][k], its last value and the product of all its step factors: with
// them every rank can work out the true value at its slab start without waiting for its predecessors' recurrences
// (k_spectral_fixup).  On rank 0 ends[1] is not needed; ends[0] is the true last value.
constexpr int kSpectralMaxTerms = 4;
constexpr int UN = 8;   // steps per batch of the fix-up pass
constexpr int kMaxChunks = 32;

template <int NRHS, bool STORE>
__device__ __forceinline__ void sine_chunk(double &u, double &prod, double *__restrict__ U, const double *__restrict__ G,
                                           const int pitch, const int k, const bool act, const int i0, const int i1,
                                           const double *__restrict__ t, const double lk, const double *__restrict__ rhs_t,
                                           const double (&rx)[NRHS > 0 ? NRHS : 1]) {
    constexpr int B = 4;
    int i = i0;
    for (; i + B <= i1; i += B) {
        double gv[B], tv[B + 1], ct[B][NRHS > 0 ? NRHS : 1];
#pragma unroll
        for (int j = 0; j < B; ++j) gv[j] = G ? G[(long)(i + j) * pitch + k] : 0.0;
#pragma unroll
        for (int j = 0; j <= B; ++j) tv[j] = __ldg(t + i - 1 + j);
#pragma unroll
        for (int j = 0; j < B; ++j)
#pragma unroll
            for (int q = 0; q < NRHS; ++q) ct[j][q] = __ldg(rhs_t + (long)(i + j) * NRHS + q);
        double inv[B], sm[B];
#pragma unroll
        for (int j = 0; j < B; ++j) {
            inv[j] = __drcp_rn(fma(__dsub_rn(tv[j + 1], tv[j]), lk, 1.0));
            double acc = 0.0;
#pragma unroll
            for (int q = 0; q < NRHS; ++q) acc = fma(ct[j][q], rx[q], acc);
            sm[j] = acc;
        }
#pragma unroll
        for (int j = 0; j < B; ++j) {
            u = fma(__dadd_rn(u, sm[j]), inv[j], gv[j]);
            if (STORE) {
                if (act) U[(long)(i + j) * pitch + k] = u;
            } else {
                prod *= inv[j];
            }
        }
    }
    for (; i < i1; ++i) {
        const double inv = __drcp_rn(fma(__dsub_rn(__ldg(t + i), __ldg(t + i - 1)), lk, 1.0));
        double acc = 0.0;
#pragma unroll
        for (int q = 0; q < NRHS; ++q) acc = fma(__ldg(rhs_t + (long)i * NRHS + q), rx[q], acc);
        u = fma(__dadd_rn(u, acc), inv, G ? G[(long)i * pitch + k] : 0.0);
        if (STORE) {
            if (act) U[(long)i * pitch + k] = u;
        } else {
            prod *= inv;
        }
    }
}

template <int NRHS>
__global__ void __launch_bounds__(1024) k_sine_solve(double *__restrict__ U, const double *G, const int pitch, const int n,
                                                     const int npts, const double *__restrict__ t,
                                                     const double *__restrict__ lam, const double *__restrict__ rhs_t,
                                                     const double *__restrict__ rxh, double *__restrict__ ends,
                                                     const int zero_start, const int len, const int *__restrict__ stop) {
    if (stop != nullptr && *stop != 0) return;
    __shared__ double sA[kMaxChunks][32], sB[kMaxChunks][32];
    const int tx = threadIdx.x, c = threadIdx.y, nch = blockDim.y;
    const int kreal = blockIdx.x * 32 + tx;
    const bool act = kreal < n;
    const int k = act ? kreal : n - 1;  // idle lanes of the last CTA shadow mode n-1 and store nothing
    const double lk = lam[k];
    double rx[NRHS > 0 ? NRHS : 1];
#pragma unroll
    for (int q = 0; q < NRHS; ++q) rx[q] = rxh[(long)q * pitch + k];
    const int i0 = min(npts, 1 + c * len), i1 = min(npts, i0 + len);
    double a = 1.0, b = 0.0;
    if (nch > 1) {
        sine_chunk<NRHS, false>(b, a, U, G, pitch, k, act, i0, i1, t, lk, rhs_t, rx);
        sA[c][tx] = a;
        sB[c][tx] = b;
        __syncthreads();
    }
    double u = zero_start ? 0.0 : U[k];
    for (int cc = 0; cc < c; ++cc) u = fma(sA[cc][tx], u, sB[cc][tx]);
    double unused = 1.0;
    sine_chunk<NRHS, true>(u, unused, U, G, pitch, k, act, i0, i1, t, lk, rhs_t, rx);
    if (ends != nullptr && act && i1 == npts && (i0 < npts || c == 0)) {
        // the thread that made the last step (or, on a level of one point, chunk 0) reports the end value ...
        ends[k] = u;
        double p = 1.0;  // ... and the product of all step factors
        if (nch > 1) {
            for (int cc = 0; cc < nch; ++cc) p *= sA[cc][tx];
        } else {
            // single chunk: pass 1 was skipped, redo the product
            for (int i = i0; i < i1; ++i) p *= __drcp_rn(fma(__dsub_rn(__ldg(t + i), __ldg(t + i - 1)), lk, 1.0));
        }
        ends[pitch + k] = p;
    }
}

// Time rank `rank` > 0: all[r] = the (last value, factor product) pairs of every rank r (all-gathered, [nranks][2][pitch]).
// The true value at my slab start is s = end_{rank-1}, with end_0 = all[0][0] and end_r = all[r][1] end_{r-1} + all[r][0];
// then W[0] = s and W[i] += (prod_{j <= i} 1 / (1 + dt_j lam)) s.
__global__ void __launch_bounds__(128) k_spectral_fixup(double *__restrict__ W, int pitch, int n, int npts,
                                                        const double *__restrict__ t, const double *__restrict__ lam,
                                                        const double *__restrict__ all, int rank,
                                                        const int *__restrict__ stop) {
    if (stop != nullptr && *stop != 0) return;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const double lk = lam[k];
    double s = all[k];
    for (int r = 1; r < rank; ++r) s = fma(all[(long)(2 * r + 1) * pitch + k], s, all[(long)(2 * r) * pitch + k]);
    W[k] = s;
    double p = s;  // p = s * running product of the step factors
    int i = 1;
    for (; i + UN <= npts; i += UN) {
        double w[UN], tv[UN + 1];
#pragma unroll
        for (int j = 0; j < UN; ++j) w[j] = W[(long)(i + j) * pitch + k];
#pragma unroll
        for (int j = 0; j <= UN; ++j) tv[j] = __ldg(t + i - 1 + j);
#pragma unroll
        for (int j = 0; j < UN; ++j) {
            p *= __drcp_rn(fma(__dsub_rn(tv[j + 1], tv[j]), lk, 1.0));
            W[(long)(i + j) * pitch + k] = w[j] + p;
        }
    }
    for (; i < npts; ++i) {
        p *= __drcp_rn(fma(__dsub_rn(__ldg(t + i), __ldg(t + i - 1)), lk, 1.0));
        W[(long)i * pitch + k] += p;
    }
}

}  // namespace mgb

using namespace mgb;

extern "C" {

int mgb_sine_matrix(int32_t n, double *s_dev, int32_t ld, void *stream) {
    if (n < 1 || ld < n || s_dev == nullptr) return heat2d_fail("sine_matrix: bad argument");
    const DeviceInfo *di = device_info();
    if (di == nullptr) return MGB_ECUDA;
    k_sine_matrix<<<4 * di->sms, 256, 0, (cudaStream_t)stream>>>(n, s_dev, ld);
    return cuda_fail(cudaGetLastError(), "sine_matrix");
}

int mgb_rows_gemm(int32_t m, int32_t n, int32_t k, const double *a_dev, int32_t lda, const double *a_row0_dev,
                  const double *b_dev, int32_t ldb, double *c_dev, int32_t ldc, void *stream) {
    if (m < 0 || n < 1 || k < 1 || a_dev == nullptr || b_dev == nullptr || c_dev == nullptr || lda < k || ldb < n || ldc < n)
        return heat2d_fail("rows_gemm: bad argument");
    const DeviceInfo *di = device_info();
    if (di == nullptr) return MGB_ECUDA;
    if (m == 0) return MGB_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int nb = (n + GN - 1) / GN;
    // the tallest tile whose grid still covers the SMs
    if ((long)((m + 63) / 64) * nb >= di->sms || m > 2048)
        k_rows_gemm<64><<<dim3(nb, (m + 63) / 64), 256, 0, st>>>(m, n, k, a_dev, lda, a_row0_dev, b_dev, ldb, c_dev, ldc, stop_flag());
    else if ((long)((m + 31) / 32) * nb >= di->sms || m > 256)
        k_rows_gemm<32><<<dim3(nb, (m + 31) / 32), 256, 0, st>>>(m, n, k, a_dev, lda, a_row0_dev, b_dev, ldb, c_dev, ldc, stop_flag());
    else
        k_rows_gemm<16><<<dim3(nb, (m + 15) / 16), 256, 0, st>>>(m, n, k, a_dev, lda, a_row0_dev, b_dev, ldb, c_dev, ldc, stop_flag());
    return cuda_fail(cudaGetLastError(), "rows_gemm");
}

int mgb_dst_twiddles(int32_t n, double *tw_dev, void *stream) {
    const int N = n + 1;
    if (n < 1 || (N & (N - 1)) != 0 || tw_dev == nullptr) return heat2d_fail("dst_twiddles: n + 1 must be a power of two");
    if (device_info() == nullptr) return MGB_ECUDA;
    k_dst_twiddles<<<(N + 255) / 256, 256, 0, (cudaStream_t)stream>>>(N, (double2 *)tw_dev);
    return cuda_fail(cudaGetLastError(), "dst_twiddles");
}

int mgb_rows_dst(int32_t m, int32_t n, const double *a_dev, int32_t lda, const double *a_row0_dev, const double *tw_dev,
                 double *c_dev, int32_t ldc, void *stream) {
    const int N = n + 1;
    if (m < 0 || n < 1 || (N & (N - 1)) != 0 || a_dev == nullptr || tw_dev == nullptr || c_dev == nullptr || lda < n ||
        ldc < n)
        return heat2d_fail("rows_dst: bad argument (n + 1 must be a power of two)");
    const DeviceInfo *di = device_info();
    if (di == nullptr) return MGB_ECUDA;
    if (m == 0) return MGB_OK;
    const size_t smem = (size_t)2 * N * sizeof(double2);
    if ((int)smem > di->max_smem_optin) return heat2d_fail("rows_dst: row too long for shared memory");
    static size_t configured = 0;
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(k_rows_dst, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute");
        configured = smem;
    }
    int log2n2 = 0;
    while ((1 << log2n2) < 2 * N) ++log2n2;
    const int grid = m < 8 * di->sms ? m : 8 * di->sms;
    k_rows_dst<<<grid, 256, smem, (cudaStream_t)stream>>>(m, n, log2n2, a_dev, lda, a_row0_dev, (const double2 *)tw_dev,
                                                          c_dev, ldc, stop_flag());
    return cuda_fail(cudaGetLastError(), "rows_dst");
}

static int launch_sine_solve(double *U, const double *G, int pitch, int n, int npts, const double *t, const double *lam,
                             const double *rhs_t, int nrhs, const double *rxh, double *ends, int zero_start, cudaStream_t st) {
    const int steps = npts - 1;
    int nch = steps / 8;  // at least 8 steps per chunk
    if (nch > kMaxChunks) nch = kMaxChunks;
    if (nch < 1) nch = 1;
    const int len = steps > 0 ? (steps + nch - 1) / nch : 1;
    const dim3 grid((n + 31) / 32), block(32, nch);
#define MGB_RECUR(Q)                                                                                                  \
    case Q:                                                                                                           \
        k_sine_solve<Q><<<grid, block, 0, st>>>(U, G, pitch, n, npts, t, lam, rhs_t, rxh, ends, zero_start, len,      \
                                                stop_flag());                                                        \
        break;
    switch (nrhs) {
        MGB_RECUR(0)
        MGB_RECUR(1)
        MGB_RECUR(2)
        MGB_RECUR(3)
        MGB_RECUR(4)
    }
#undef MGB_RECUR
    return cuda_fail(cudaGetLastError(), "sine_solve");
}

int mgb_heat1d_spectral_recur(const mgb_level *lvl, const double *lam_dev, const double *rxhat_dev, double *work_dev,
                              double *ends_dev, int32_t zero_start, void *stream) {
    if (lvl == nullptr || lvl->app != MGB_APP_HEAT1D || lam_dev == nullptr || work_dev == nullptr || lvl->t_dev == nullptr)
        return heat2d_fail("heat1d_spectral_recur: needs a HEAT1D level with its time grid, eigenvalues and work rows");
    if (lvl->rhs_dense_dev != nullptr || lvl->nrhs > kSpectralMaxTerms || (lvl->nrhs > 0 && (rxhat_dev == nullptr || lvl->rhs_t_dev == nullptr)))
        return heat2d_fail("heat1d_spectral_recur: right-hand side must be separable with at most 4 terms");
    if (device_info() == nullptr) return MGB_ECUDA;
    if (lvl->npts < 2 && ends_dev == nullptr) return MGB_OK;
    return launch_sine_solve(work_dev, work_dev, lvl->pitch, lvl->n, lvl->npts, lvl->t_dev, lam_dev, lvl->rhs_t_dev,
                             lvl->nrhs, rxhat_dev, ends_dev, zero_start, (cudaStream_t)stream);
}

int mgb_sine_level_solve(const mgb_level *lvl, const double *lam_dev, const double *rxh_dev, double *ends_dev,
                         int32_t zero_start, void *stream) {
    if (lvl == nullptr || lvl->app != MGB_APP_HEAT1D_SINE || lam_dev == nullptr || lvl->u_dev == nullptr ||
        lvl->t_dev == nullptr)
        return heat2d_fail("sine_level_solve: needs a HEAT1D_SINE level with its time grid and eigenvalues");
    if (lvl->rhs_dense_dev != nullptr || lvl->nrhs > kSpectralMaxTerms || (lvl->nrhs > 0 && (rxh_dev == nullptr || lvl->rhs_t_dev == nullptr)))
        return heat2d_fail("sine_level_solve: right-hand side must be separable with at most 4 terms");
    if (device_info() == nullptr) return MGB_ECUDA;
    if (lvl->npts < 2 && ends_dev == nullptr) return MGB_OK;
    return launch_sine_solve(lvl->u_dev, lvl->g_dev, lvl->pitch, lvl->n, lvl->npts, lvl->t_dev, lam_dev, lvl->rhs_t_dev,
                             lvl->nrhs, rxh_dev, ends_dev, zero_start, (cudaStream_t)stream);
}

int mgb_heat1d_spectral_fixup(const mgb_level *lvl, const double *lam_dev, double *work_dev, const double *all_ends_dev,
                              int32_t rank, void *stream) {
    if (lvl == nullptr || (lvl->app != MGB_APP_HEAT1D && lvl->app != MGB_APP_HEAT1D_SINE) || lam_dev == nullptr ||
        work_dev == nullptr || lvl->t_dev == nullptr || all_ends_dev == nullptr || rank < 1)
        return heat2d_fail("heat1d_spectral_fixup: bad argument");
    if (device_info() == nullptr) return MGB_ECUDA;
    k_spectral_fixup<<<(lvl->n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(work_dev, lvl->pitch, lvl->n, lvl->npts, lvl->t_dev,
                                                                            lam_dev, all_ends_dev, rank, stop_flag());
    return cuda_fail(cudaGetLastError(), "heat1d_spectral_fixup");
}

}  // extern "C"
