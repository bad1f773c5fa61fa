// fourier.cu -- the sequential coarsest-level solve of Advection1D without its sequential Phi chain.
//
// mgrit.py:459-486 computes u_i = g_i + Phi_i(u_{i-1}), i = 1..N-1, one Phi after the other; Phi_i = (I + dt_i L)^-1 with
// the CIRCULANT upwind matrix L = (c/dx)(I - S), (S u)_j = u_{j-1} periodic (advection_1d.py:100-118, 129-143).  A
// circulant matrix is diagonalised by the discrete Fourier transform: with X_k = sum_j x_j exp(-2 pi i j k / n) the solve
// becomes, per frequency k, the complex scalar recurrence
//     U_i[k] = U_{i-1}[k] / (1 + nu_i (1 - exp(-i theta_k))) + G_i[k],     nu_i = (c/dx) dt_i,  theta_k = 2 pi k / n,
// i.e. a composition of affine maps that runs TIME-PARALLEL (chunks of steps composed, combined, rerun -- the scheme of
// spectral.cu's k_sine_solve with complex coefficients).  8192 dependent cyclic solves of cfg 4's coarsest level become
// two transforms of all rows and ~2 N / chunks dependent complex FMAs.  The same linear systems are solved exactly (a
// direct method like the reference's SuperLU call); only the order of the rounding errors differs (tests: <= 1e-12).
//
// n is arbitrary (nx = 4096 -> n = 4095 = 3^2 5 7 13), so the transform is Bluestein's: j k = (j^2 + k^2 - (k-j)^2) / 2,
//     X_k = conj(w_k) sum_j (x_j conj(w_j)) w_{k-j},      w_m = exp(i pi m^2 / n)      (m^2 reduced mod 2n in integers)
// a cyclic convolution of length M = 2^p >= 2n - 1 done with a fast Fourier transform in shared memory: one CTA per PAIR of
// rows (two real rows are the real and imaginary part of one complex transform), M complex doubles (136 KB with padding
// for n <= 4096), decimation in frequency forwards (natural in, bit-reversed out), the chirp's spectrum stored in that
// same bit-reversed order, decimation in time backwards (bit-reversed in, natural out), radix-8 / radix-16 passes in
// registers.  Rows are real, so only k = 0 .. n/2 are kept: [npts][2 (n/2 + 1)] doubles (re, im).
#include "../../include/mgrit_b200.h"
#include "phi.cuh"
#include "table.h"

namespace mgb {

int heat2d_fail(const char *msg);  // api.cu: records the message, returns MGB_EINVAL

namespace fourier {

constexpr int kThreads = 512;   // radix-8 passes hold 8 complex values per thread: 128 registers each

__device__ __forceinline__ double2 cmul(const double2 a, const double2 b) {
    return make_double2(fma(a.x, b.x, -a.y * b.y), fma(a.x, b.y, a.y * b.x));
}
__device__ __forceinline__ double2 cmulc(const double2 a, const double2 b) {  // a * conj(b)
    return make_double2(fma(a.x, b.x, a.y * b.y), fma(a.y, b.x, -a.x * b.y));
}

// tw[t] = exp(-2 pi i t / M), t < M/2;  chirp[j] = exp(i pi j^2 / n), j < n;  b[m] = chirp[|m|] wrapped to length M
__global__ void k_tables(int n, int M, double2 *__restrict__ tw, double2 *__restrict__ chirp, double2 *__restrict__ b) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q < M / 2) {
        double sn, cs;
        sincospi(-2.0 * (double)q / (double)M, &sn, &cs);
        tw[q] = make_double2(cs, sn);
    }
    if (q < M) {
        int m = -1;
        if (q < n) m = q;
        else if (M - q < n) m = M - q;
        double2 v = make_double2(0.0, 0.0);
        if (m >= 0) {
            const long r = ((long)m * m) % (2L * n);
            double sn, cs;
            sincospi((double)r / (double)n, &sn, &cs);
            v = make_double2(cs, sn);
            if (q < n) chirp[q] = v;
        }
        b[q] = v;
    }
}

// ---- the transform of one row in shared memory ------------------------------------------------------------------------
// The radix-2 flow graph (decimation in frequency forwards: natural order in, bit-reversed out; its mirror image
// backwards) is walked in PASSES of RL = 3 or 4 consecutive stages: a thread takes the 2^RL elements
// hi 2^(s_top+1) + a 2^q + lo (q = s_top-RL+1) into registers and runs the RL stages there, so a row of 8192 points
// crosses shared memory 4 times per direction instead of 13.  The passes are bound by the shared-memory pipe (ncu:
// l1tex 85 %, FP64 pipe 15 %, profiles/r02s_fourier_ncu.txt), hence:
//   * one twiddle load per pass and group, W^(lo) of the top stage: the next stage's W^(lo) is its square, and the other
//     twiddles of a stage differ from W^(lo) by a fixed 16th root of unity (W^(a' 2^q + lo) = W^lo exp(-i pi a' / ha));
//     after that the passes wait for L1 misses of the table loads no longer (ncu long scoreboard 5.2 -> ...);
//   * the chirp's spectrum is stored in the order the last forward pass reads it (coalesced over the threads);
//   * element i lives at z[i + i/16]: the 16 elements a thread owns in the last pass (stride 1) and the rows of 16 that a
//     quarter warp touches fall into different banks;
//   * the first forward pass reads its input straight from global memory (the caller's functor), the last forward pass
//     multiplies by the chirp's spectrum on the way out.
__device__ __forceinline__ int pad(const int i) { return i + (i >> 4); }

// w * exp(-i pi k / 8), k = 0 .. 7 (k is a compile-time constant after unrolling)
__device__ __forceinline__ double2 rot16(const double2 w, const int k) {
    constexpr double c1 = 0.92387953251128673848, s1 = 0.38268343236508978178, h = 0.70710678118654752440;
    switch (k) {
        case 0: return w;
        case 1: return cmul(w, make_double2(c1, -s1));
        case 2: return cmul(w, make_double2(h, -h));
        case 3: return cmul(w, make_double2(s1, -c1));
        case 4: return make_double2(w.y, -w.x);
        case 5: return cmul(w, make_double2(-s1, -c1));
        case 6: return cmul(w, make_double2(-h, -h));
        default: return cmul(w, make_double2(-c1, -s1));
    }
}

// stages s_top .. s_top-RL+1 forwards.  src(i): element i of the input (shared memory, or the caller's global data for the
// first pass); mul != nullptr: the results are multiplied by mul[i] on the way out (last pass).
template <int RL, class Src>
__device__ __forceinline__ void dif_pass(double2 *z, const int log2m, const int s_top, const double2 *__restrict__ tw,
                                         const double2 *__restrict__ mul, Src src) {
    constexpr int RN = 1 << RL;
    const int q = s_top - RL + 1, H = 1 << (log2m - 1), groups = 1 << (log2m - RL);
    for (int gi = threadIdx.x; gi < groups; gi += blockDim.x) {
        const int lo = gi & ((1 << q) - 1), hi = gi >> q;
        const int base = (hi << (s_top + 1)) + lo;
        double2 v[RN];
#pragma unroll
        for (int a = 0; a < RN; ++a) v[a] = src(base + (a << q));
        double2 ws = __ldg(tw + lo * (H >> s_top));              // W_(2^(s_top+1))^lo; the next stage's is its square
#pragma unroll
        for (int j = 0; j < RL; ++j) {
            constexpr int kOne = 1;
            const int ha = kOne << (RL - 1 - j);                 // pairs (a, a + ha), a with that bit clear
#pragma unroll
            for (int a = 0; a < RN; ++a) {
                if (a & ha) continue;
                const double2 w = rot16(ws, (a & (ha - 1)) * (8 / ha));
                const double2 x = v[a], y = v[a + ha];
                v[a] = make_double2(x.x + y.x, x.y + y.y);
                v[a + ha] = cmul(make_double2(x.x - y.x, x.y - y.y), w);
            }
            ws = cmul(ws, ws);
        }
#pragma unroll
        for (int a = 0; a < RN; ++a) {
            const int idx = base + (a << q);
            // mul is stored in the order this (last, q = 0) pass reads it: [a][group], coalesced over the threads
            z[pad(idx)] = mul ? cmul(v[a], __ldg(mul + a * groups + gi)) : v[a];
        }
    }
    __syncthreads();
}
// the mirror image: stages q .. q+RL-1 ascending, conjugate twiddles, not scaled
template <int RL>
__device__ __forceinline__ void dit_pass(double2 *z, const int log2m, const int q, const double2 *__restrict__ tw) {
    constexpr int RN = 1 << RL;
    const int s_top = q + RL - 1, H = 1 << (log2m - 1), groups = 1 << (log2m - RL);
    for (int gi = threadIdx.x; gi < groups; gi += blockDim.x) {
        const int lo = gi & ((1 << q) - 1), hi = gi >> q;
        const int base = (hi << (s_top + 1)) + lo;
        double2 v[RN];
#pragma unroll
        for (int a = 0; a < RN; ++a) v[a] = z[pad(base + (a << q))];
        double2 wj[RL];                                          // W^lo of stages s_top, s_top-1, ..: successive squares
        wj[0] = __ldg(tw + lo * (H >> s_top));
#pragma unroll
        for (int j = 1; j < RL; ++j) wj[j] = cmul(wj[j - 1], wj[j - 1]);
#pragma unroll
        for (int j = RL - 1; j >= 0; --j) {
            constexpr int kOne = 1;
            const int ha = kOne << (RL - 1 - j);
            const double2 ws = wj[j];
#pragma unroll
            for (int a = 0; a < RN; ++a) {
                if (a & ha) continue;
                const double2 w = rot16(ws, (a & (ha - 1)) * (8 / ha));
                const double2 x = v[a];
                const double2 y = cmulc(v[a + ha], w);
                v[a] = make_double2(x.x + y.x, x.y + y.y);
                v[a + ha] = make_double2(x.x - y.x, x.y - y.y);
            }
        }
#pragma unroll
        for (int a = 0; a < RN; ++a) z[pad(base + (a << q))] = v[a];
    }
    __syncthreads();
}

// passes of the forward transform from the top stage down: sizes rl[], top stages top[]
struct Plan {
    int count;
    int rl[6], top[6];
};
__device__ __forceinline__ Plan make_plan(const int log2m) {
    Plan p;
    p.count = 0;
    int s = log2m - 1;
    while (s + 1 > 5) {
        p.rl[p.count] = 3;
        p.top[p.count++] = s;
        s -= 3;
    }
    if (s + 1 == 5) {
        p.rl[p.count] = 3;
        p.top[p.count++] = s;
        s -= 3;
    }
    p.rl[p.count] = s + 1;   // 1 .. 4 stages left
    p.top[p.count++] = s;
    return p;
}

template <class Src>
__device__ __forceinline__ void run_dif(double2 *z, const int log2m, const int rl, const int s_top,
                                        const double2 *__restrict__ tw, const double2 *__restrict__ mul, Src src) {
    switch (rl) {
        case 1: dif_pass<1>(z, log2m, s_top, tw, mul, src); break;
        case 2: dif_pass<2>(z, log2m, s_top, tw, mul, src); break;
        case 3: dif_pass<3>(z, log2m, s_top, tw, mul, src); break;
        default: dif_pass<4>(z, log2m, s_top, tw, mul, src); break;
    }
}
__device__ __forceinline__ void run_dit(double2 *z, const int log2m, const int rl, const int q, const double2 *__restrict__ tw) {
    switch (rl) {
        case 1: dit_pass<1>(z, log2m, q, tw); break;
        case 2: dit_pass<2>(z, log2m, q, tw); break;
        case 3: dit_pass<3>(z, log2m, q, tw); break;
        default: dit_pass<4>(z, log2m, q, tw); break;
    }
}

// z <- transform of the input `first(i)` (natural order in, bit-reversed out), times mul on the way out if not null
template <class Src>
__device__ __forceinline__ void fft_dif(double2 *z, const int log2m, const double2 *__restrict__ tw,
                                        const double2 *__restrict__ mul, Src first) {
    const Plan p = make_plan(log2m);
    auto smem = [z](const int i) { return z[pad(i)]; };
    for (int k = 0; k < p.count; ++k) {
        const double2 *m = (k == p.count - 1) ? mul : nullptr;
        if (k == 0)
            run_dif(z, log2m, p.rl[0], p.top[0], tw, m, first);
        else
            run_dif(z, log2m, p.rl[k], p.top[k], tw, m, smem);
    }
}
// the inverse flow graph: bit-reversed order in, natural order out, not scaled (divide by M)
__device__ __forceinline__ void ifft_dit(double2 *z, const int log2m, const double2 *__restrict__ tw) {
    const Plan p = make_plan(log2m);
    for (int k = p.count - 1; k >= 0; --k) run_dit(z, log2m, p.rl[k], p.top[k] - p.rl[k] + 1, tw);
}
// z <- cyclic convolution of the input with the chirp (times M)
template <class Src>
__device__ __forceinline__ void chirp_conv(double2 *z, const int log2m, const double2 *__restrict__ tw,
                                           const double2 *__restrict__ bhat, Src first) {
    fft_dif(z, log2m, tw, bhat, first);
    ifft_dit(z, log2m, tw);
}

__global__ void __launch_bounds__(kThreads) k_bhat(const int log2m, const double2 *__restrict__ tw,
                                                   const double2 *__restrict__ b, double2 *__restrict__ bhat) {
    extern __shared__ double2 z[];
    const int M = 1 << log2m;
    fft_dif(z, log2m, tw, nullptr, [b](const int i) { return b[i]; });
    // stored as the last forward pass reads it: element (group gi, a) = index gi 2^RL + a at [a][gi]
    const Plan pl = make_plan(log2m);
    const int rl = pl.rl[pl.count - 1], groups = M >> rl;
    for (int p = threadIdx.x; p < M; p += blockDim.x) bhat[(p & ((1 << rl) - 1)) * groups + (p >> rl)] = z[pad(p)];
}

// Rows are real: two of them share one complex transform, Z = F(x_a + i x_b), X_a[k] = (Z_k + conj(Z_{n-k})) / 2,
// X_b[k] = (Z_k - conj(Z_{n-k})) / (2i).
// C[r][2k], C[r][2k+1] = re, im of sum_j A[r][j] exp(-2 pi i j k / n), k = 0 .. n/2.  Row 0 of A from a_row0 if not null.
__global__ void __launch_bounds__(kThreads) k_rows_rfft(const int rows, const int n, const int log2m,
                                                        const double *__restrict__ A, const long lda,
                                                        const double *__restrict__ a_row0, const double2 *__restrict__ tw,
                                                        const double2 *__restrict__ chirp, const double2 *__restrict__ bhat,
                                                        double *__restrict__ Cm, const long ldc, const int *__restrict__ stop,
                                                        const int chirp_smem) {
    if (stop != nullptr && *stop != 0) return;
    extern __shared__ double2 z[];
    const int M = 1 << log2m, K = n / 2 + 1;
    const double scale = 0.5 / (double)M;
    // the chirp is read three times per transform: once per CTA into shared memory when it fits behind the row
    const double2 *ch = chirp;
    if (chirp_smem) {
        double2 *cs = z + (M + M / 16 + 1);
        for (int j = threadIdx.x; j < n; j += blockDim.x) cs[j] = chirp[j];
        ch = cs;
        __syncthreads();
    }
    for (int pr = blockIdx.x; 2 * pr < rows; pr += gridDim.x) {
        const int ra = 2 * pr, rb = ra + 1;
        const double *xa = (ra == 0 && a_row0 != nullptr) ? a_row0 : A + (long)ra * lda;
        const double *xb = rb < rows ? A + (long)rb * lda : nullptr;
        chirp_conv(z, log2m, tw, bhat, [=](const int j) {
            if (j >= n) return make_double2(0.0, 0.0);
            return cmulc(make_double2(xa[j], xb ? xb[j] : 0.0), ch[j]);
        });
        double2 *oa = reinterpret_cast<double2 *>(Cm + (long)ra * ldc);
        double2 *ob = xb ? reinterpret_cast<double2 *>(Cm + (long)rb * ldc) : nullptr;
        for (int k = threadIdx.x; k < K; k += blockDim.x) {
            const int km = (k == 0) ? 0 : n - k;
            const double2 zk = cmulc(z[pad(k)], ch[k]), zm = cmulc(z[pad(km)], ch[km]);
            // (Z_k + conj(Z_m)) / 2  and  (Z_k - conj(Z_m)) / (2i) = (Im(Z_k) + Im(Z_m), Re(Z_m) - Re(Z_k)) / 2
            oa[k] = make_double2((zk.x + zm.x) * scale, (zk.y - zm.y) * scale);
            if (ob) ob[k] = make_double2((zk.y + zm.y) * scale, (zm.x - zk.x) * scale);
        }
        __syncthreads();
    }
}

// out[r][j] = (1/n) sum_{k<n} X[r][k] exp(+2 pi i j k / n) with X[r][n-k] = conj(X[r][k]), rows first_row .. rows-1, two
// rows per transform: Y = X_a + i X_b (Hermitian extensions), x_a + i x_b = conj(F(conj(Y))) / n.
__global__ void __launch_bounds__(kThreads) k_rows_irfft(const int rows, const int first_row, const int n, const int log2m,
                                                         const double *__restrict__ Cm, const long ldc,
                                                         const double2 *__restrict__ tw, const double2 *__restrict__ chirp,
                                                         const double2 *__restrict__ bhat, double *__restrict__ out,
                                                         const long ldo, const int *__restrict__ stop, const int chirp_smem) {
    if (stop != nullptr && *stop != 0) return;
    extern __shared__ double2 z[];
    const int M = 1 << log2m, K = n / 2 + 1;
    const double scale = 1.0 / ((double)M * (double)n);
    const double2 *ch = chirp;
    if (chirp_smem) {
        double2 *cs = z + (M + M / 16 + 1);
        for (int j = threadIdx.x; j < n; j += blockDim.x) cs[j] = chirp[j];
        ch = cs;
        __syncthreads();
    }
    for (int pr = blockIdx.x; first_row + 2 * pr < rows; pr += gridDim.x) {
        const int ra = first_row + 2 * pr, rb = ra + 1;
        const double2 *Xa = reinterpret_cast<const double2 *>(Cm + (long)ra * ldc);
        const double2 *Xb = rb < rows ? reinterpret_cast<const double2 *>(Cm + (long)rb * ldc) : nullptr;
        chirp_conv(z, log2m, tw, bhat, [=](const int k) {
            if (k >= n) return make_double2(0.0, 0.0);
            // Hermitian extension: X[k] for k <= n/2, conj(X[n-k]) above
            const bool up = k >= K;
            const int kk = up ? n - k : k;
            double2 a = Xa[kk], b = Xb ? Xb[kk] : make_double2(0.0, 0.0);
            if (up) {
                a.y = -a.y;
                b.y = -b.y;
            }
            // conj(Y) = conj(a + i b) = (a.x - b.y) - i (a.y + b.x)
            return cmulc(make_double2(a.x - b.y, -(a.y + b.x)), ch[k]);
        });
        double *oa = out + (long)ra * ldo;
        double *ob = Xb ? out + (long)rb * ldo : nullptr;
        for (int j = threadIdx.x; j < n; j += blockDim.x) {
            const double2 f = cmulc(z[pad(j)], ch[j]);     // F(conj(Y))_j times M
            oa[j] = f.x * scale;
            if (ob) ob[j] = -f.y * scale;
        }
        __syncthreads();
    }
}

// ---- the complex recurrences, time-parallel (spectral.cu k_sine_solve with complex factors) ----------------------------
constexpr int kFreqs = 16;      // frequencies per CTA: 16 consecutive double2 = 256 B per row access
constexpr int kMaxChunks = 64;  // x chunks of consecutive steps = 1024 threads

// factor of step i for frequency (cx, sx) = (1 - cos theta, sin theta): 1 / (1 + nu cx + i nu sx)
__device__ __forceinline__ double2 step_factor(const double nu, const double cx, const double sx) {
    const double dr = fma(nu, cx, 1.0), di = nu * sx;
    const double inv = __drcp_rn(fma(dr, dr, di * di));
    return make_double2(dr * inv, -di * inv);
}

template <bool STORE>
__device__ __forceinline__ void cplx_chunk(double2 &u, double2 &prod, double2 *__restrict__ W, const long ldw2, const int k,
                                           const bool act, const int i0, const int i1, const double *__restrict__ t,
                                           const double fac, const double cx, const double sx) {
    constexpr int B = 4;
    int i = i0;
    for (; i + B <= i1; i += B) {
        double2 gv[B];
        double tv[B + 1];
#pragma unroll
        for (int j = 0; j < B; ++j) gv[j] = W[(long)(i + j) * ldw2 + k];
#pragma unroll
        for (int j = 0; j <= B; ++j) tv[j] = __ldg(t + i - 1 + j);
        double2 f[B];
#pragma unroll
        for (int j = 0; j < B; ++j) f[j] = step_factor(fac * __dsub_rn(tv[j + 1], tv[j]), cx, sx);
#pragma unroll
        for (int j = 0; j < B; ++j) {
            const double2 m = cmul(u, f[j]);
            u = make_double2(m.x + gv[j].x, m.y + gv[j].y);
            if (STORE) {
                if (act) W[(long)(i + j) * ldw2 + k] = u;
            } else {
                prod = cmul(prod, f[j]);
            }
        }
    }
    for (; i < i1; ++i) {
        const double2 f = step_factor(fac * __dsub_rn(__ldg(t + i), __ldg(t + i - 1)), cx, sx);
        const double2 g = W[(long)i * ldw2 + k];
        const double2 m = cmul(u, f);
        u = make_double2(m.x + g.x, m.y + g.y);
        if (STORE) {
            if (act) W[(long)i * ldw2 + k] = u;
        } else {
            prod = cmul(prod, f);
        }
    }
}

// W[i][k] <- W[i-1][k] f_i[k] + W[i][k], i = 1 .. npts-1, in place (W[0] = transformed start value, W[i] = transformed g_i)
__global__ void __launch_bounds__(1024) k_cplx_solve(double2 *__restrict__ W, const long ldw2, const int n, const int K,
                                                     const int npts, const double *__restrict__ t, const double fac,
                                                     const int len, const int *__restrict__ stop) {
    if (stop != nullptr && *stop != 0) return;
    __shared__ double2 sA[kMaxChunks][kFreqs], sB[kMaxChunks][kFreqs];
    const int tx = threadIdx.x, c = threadIdx.y, nch = blockDim.y;
    const int kreal = blockIdx.x * kFreqs + tx;
    const bool act = kreal < K;
    const int k = act ? kreal : K - 1;  // idle lanes of the last CTA shadow the last frequency and store nothing
    double sn, cs;
    {   // theta = 2 pi k / n, reduced exactly: sincospi(2 k / n)
        sincospi(2.0 * (double)k / (double)n, &sn, &cs);
    }
    const double cx = 1.0 - cs, sx = sn;
    const int i0 = min(npts, 1 + c * len), i1 = min(npts, i0 + len);
    double2 a = make_double2(1.0, 0.0), b = make_double2(0.0, 0.0);
    if (nch > 1) {
        cplx_chunk<false>(b, a, W, ldw2, k, act, i0, i1, t, fac, cx, sx);
        sA[c][tx] = a;
        sB[c][tx] = b;
        __syncthreads();
    }
    double2 u = W[k];
    for (int cc = 0; cc < c; ++cc) {
        const double2 m = cmul(sA[cc][tx], u);
        u = make_double2(m.x + sB[cc][tx].x, m.y + sB[cc][tx].y);
    }
    double2 unused = make_double2(1.0, 0.0);
    cplx_chunk<true>(u, unused, W, ldw2, k, act, i0, i1, t, fac, cx, sx);
}

static int log2_conv_len(int n) {
    int p = 0;
    while ((1L << p) < 2L * n - 1) ++p;
    return p < 1 ? 1 : p;
}

}  // namespace fourier
}  // namespace mgb

using namespace mgb;
using namespace mgb::fourier;

extern "C" {

int mgb_circ_fft_length(int32_t n) {
    if (n < 1) return 0;
    return 1 << log2_conv_len(n);
}

int mgb_circ_fft_tables(int32_t n, double *tw_dev, double *chirp_dev, double *bhat_dev, double *work_dev, void *stream) {
    if (n < 1 || tw_dev == nullptr || chirp_dev == nullptr || bhat_dev == nullptr || work_dev == nullptr)
        return heat2d_fail("circ_fft_tables: bad argument");
    const DeviceInfo *di = device_info();
    if (di == nullptr) return MGB_ECUDA;
    const int p = log2_conv_len(n), M = 1 << p;
    const size_t smem = (size_t)(M + M / 16 + 1) * sizeof(double2);
    if ((int)smem > di->max_smem_optin) return heat2d_fail("circ_fft_tables: row too long for shared memory (n <= 4096)");
    cudaStream_t st = (cudaStream_t)stream;
    k_tables<<<(M + 255) / 256, 256, 0, st>>>(n, M, (double2 *)tw_dev, (double2 *)chirp_dev, (double2 *)work_dev);
    cudaError_t e = cudaFuncSetAttribute(k_bhat, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute");
    k_bhat<<<1, kThreads, smem, st>>>(p, (const double2 *)tw_dev, (const double2 *)work_dev, (double2 *)bhat_dev);
    return cuda_fail(cudaGetLastError(), "circ_fft_tables");
}

static int fft_smem(int n, int *p_out, size_t *smem_out, int *chirp_smem) {
    const DeviceInfo *di = device_info();
    if (di == nullptr) return MGB_ECUDA;
    const int p = log2_conv_len(n);
    size_t smem = (((size_t)1 << p) + ((size_t)1 << p) / 16 + 1) * sizeof(double2);   // element i at z[i + i/16]
    if ((int)smem > di->max_smem_optin) return heat2d_fail("rows_rfft: row too long for shared memory (n <= 4096)");
    *chirp_smem = 0;
    if ((long)(smem + (size_t)n * sizeof(double2)) <= (long)di->max_smem_optin) {   // the chirp behind the row
        smem += (size_t)n * sizeof(double2);
        *chirp_smem = 1;
    }
    *p_out = p;
    *smem_out = smem;
    return MGB_OK;
}

int mgb_rows_rfft(int32_t m, int32_t n, const double *a_dev, int64_t lda, const double *a_row0_dev, const double *tw_dev,
                  const double *chirp_dev, const double *bhat_dev, double *c_dev, int64_t ldc, void *stream) {
    if (m < 0 || n < 1 || a_dev == nullptr || tw_dev == nullptr || chirp_dev == nullptr || bhat_dev == nullptr ||
        c_dev == nullptr || lda < n || ldc < 2 * (n / 2 + 1) || (ldc & 1) || ((uintptr_t)c_dev & 15))
        return heat2d_fail("rows_rfft: bad argument");
    int p, chirp_smem;
    size_t smem;
    if (int rc = fft_smem(n, &p, &smem, &chirp_smem)) return rc;
    if (m == 0) return MGB_OK;
    static size_t configured = 0;
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(k_rows_rfft, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute");
        configured = smem;
    }
    const DeviceInfo *di = device_info();
    const int per_sm = (int)(di->max_smem_optin / smem) > 0 ? (int)(di->max_smem_optin / smem) : 1;
    const int cap = di->sms * (per_sm > 2 ? 2 : per_sm);
    k_rows_rfft<<<(m + 1) / 2 < cap ? (m + 1) / 2 : cap, kThreads, smem, (cudaStream_t)stream>>>(
        m, n, p, a_dev, lda, a_row0_dev, (const double2 *)tw_dev, (const double2 *)chirp_dev, (const double2 *)bhat_dev, c_dev,
        ldc, stop_flag(), chirp_smem);
    return cuda_fail(cudaGetLastError(), "rows_rfft");
}

int mgb_rows_irfft(int32_t m, int32_t first_row, int32_t n, const double *c_dev, int64_t ldc, const double *tw_dev,
                   const double *chirp_dev, const double *bhat_dev, double *out_dev, int64_t ldo, void *stream) {
    if (m < 0 || first_row < 0 || n < 1 || c_dev == nullptr || tw_dev == nullptr || chirp_dev == nullptr ||
        bhat_dev == nullptr || out_dev == nullptr || ldo < n || ldc < 2 * (n / 2 + 1) || (ldc & 1) || ((uintptr_t)c_dev & 15))
        return heat2d_fail("rows_irfft: bad argument");
    int p, chirp_smem;
    size_t smem;
    if (int rc = fft_smem(n, &p, &smem, &chirp_smem)) return rc;
    if (m <= first_row) return MGB_OK;
    static size_t configured = 0;
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(k_rows_irfft, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute");
        configured = smem;
    }
    const DeviceInfo *di = device_info();
    const int per_sm = (int)(di->max_smem_optin / smem) > 0 ? (int)(di->max_smem_optin / smem) : 1;
    const int cap = di->sms * (per_sm > 2 ? 2 : per_sm);
    const int work = (m - first_row + 1) / 2;
    k_rows_irfft<<<work < cap ? work : cap, kThreads, smem, (cudaStream_t)stream>>>(
        m, first_row, n, p, c_dev, ldc, (const double2 *)tw_dev, (const double2 *)chirp_dev, (const double2 *)bhat_dev, out_dev,
        ldo, stop_flag(), chirp_smem);
    return cuda_fail(cudaGetLastError(), "rows_irfft");
}

int mgb_advection1d_spectral_recur(int32_t n, int32_t npts, const double *t_dev, double c_over_dx, double *work_dev,
                                   int64_t ldw, void *stream) {
    if (n < 1 || npts < 1 || t_dev == nullptr || work_dev == nullptr || ldw < 2 * (n / 2 + 1) || (ldw & 1) ||
        ((uintptr_t)work_dev & 15))
        return heat2d_fail("advection1d_spectral_recur: bad argument");
    if (device_info() == nullptr) return MGB_ECUDA;
    if (npts < 2) return MGB_OK;
    const int K = n / 2 + 1, steps = npts - 1;
    int nch = steps / 8;  // at least 8 steps per chunk
    if (nch > kMaxChunks) nch = kMaxChunks;
    if (nch < 1) nch = 1;
    const int len = (steps + nch - 1) / nch;
    const dim3 grid((K + kFreqs - 1) / kFreqs), block(kFreqs, nch);
    k_cplx_solve<<<grid, block, 0, (cudaStream_t)stream>>>((double2 *)work_dev, ldw / 2, n, K, npts, t_dev, c_over_dx, len,
                                                           stop_flag());
    return cuda_fail(cudaGetLastError(), "advection1d_spectral_recur");
}

}  // extern "C"
