// sine_modes.cu -- the hot sweeps of Heat1D in sine space (MGB_APP_HEAT1D_SINE) with ONE THREAD PER MODE.
//
// In sine space Phi is diagonal: mode k of a time point depends on mode k of its predecessor and on nothing else.  The team
// kernels of sweeps.cuh (a row in the registers of one warp, rows staged through shared memory by bulk copies) were built
// for the Toeplitz solve, where the threads of a row must talk to each other; with a diagonal Phi that machinery is pure
// overhead -- 8 to 12 resident warps per SM, a generator / mbarrier hand-shake per row.  Here a thread owns R modes of one
// coarse interval and walks the interval's time points: consecutive threads hold consecutive modes, so every load and store
// of a row is a coalesced 2 KB access per CTA straight from / to HBM, there is no shared memory, a thread needs ~50
// registers, and an SM keeps 32+ warps in flight.  The arithmetic per element is the team kernels', operation by operation
// (fma(ct, rx, x) * inv), so results are bit-identical to them.
//
// Tables in natural mode order, one array per level (struct mgb_level.nat_dev): [2 + nrhs][pitch] =
//   lam_k | 1 / (1 + dt lam_k) for the level's dt (used when the level is uniform in time) | rxh_0 | rxh_1 ...
#include <cstdlib>

#include "../../include/mgrit_b200.h"
#include "phi.cuh"
#include "table.h"

namespace mgb {
namespace modes {

constexpr int TB = 256;  // threads per CTA
constexpr int R = 4;     // modes per thread: mode (chunk + r * TB + tid)
constexpr int MINB = 4;  // resident CTAs per SM the kernels are compiled for (<= 64 registers): the sweeps are chains of
                         // dependent row loads per CTA, so the SM needs many CTAs in flight, not many registers per thread

template <int Q>
struct DiagPhi {
    double d[R];
    double rx[Q > 0 ? Q : 1][R];
    bool recip;

    __device__ __forceinline__ void load(const LevelDev &L, int m0) {
        recip = (L.ndt == 1);
        const double *__restrict__ tab = L.nat + (recip ? L.pitch : 0);
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int m = m0 + r * TB + threadIdx.x;
            const bool ok = m < L.n;
            d[r] = ok ? __ldg(tab + m) : 0.0;
#pragma unroll
            for (int q = 0; q < Q; ++q) rx[q][r] = ok ? __ldg(L.nat + (size_t)(2 + q) * L.pitch + m) : 0.0;
        }
    }
    // the time factors of step i (issued early: the load is a dependent access that a step would otherwise wait for)
    __device__ static __forceinline__ void factors(double (&ct)[Q > 0 ? Q : 1], const LevelDev &L, int i) {
#pragma unroll
        for (int q = 0; q < Q; ++q) ct[q] = __ldg(L.rhs_t + (size_t)i * Q + q);
    }
    // x <- Phi_i(x)   (same operations, in the same order, as Heat1DSine::apply)
    __device__ __forceinline__ void step(double (&x)[R], const LevelDev &L, int i) const {
        double ct[Q > 0 ? Q : 1];
        factors(ct, L, i);
        step(x, L, i, ct);
    }
    __device__ __forceinline__ void step(double (&x)[R], const LevelDev &L, int i, const double (&ct)[Q > 0 ? Q : 1]) const {
#pragma unroll
        for (int q = 0; q < Q; ++q) {
#pragma unroll
            for (int r = 0; r < R; ++r) x[r] = fma(ct[q], rx[q][r], x[r]);
        }
        if (recip) {
#pragma unroll
            for (int r = 0; r < R; ++r) x[r] = x[r] * d[r];
        } else {
            const double dt = __ldg(L.sconst + (size_t)__ldg(L.dtidx + i) * L.cw);
#pragma unroll
            for (int r = 0; r < R; ++r) x[r] = __ddiv_rn(x[r], fma(dt, d[r], 1.0));
        }
    }
};

__device__ __forceinline__ void ldrow(double (&x)[R], const double *__restrict__ base, int i, int pitch, int m0, int n) {
    const double *__restrict__ p = base + (size_t)i * pitch;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int m = m0 + r * TB + threadIdx.x;
        x[r] = (m < n) ? p[m] : 0.0;
    }
}
__device__ __forceinline__ void strow(const double (&x)[R], double *__restrict__ base, int i, int pitch, int m0, int n) {
    double *__restrict__ p = base + (size_t)i * pitch;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int m = m0 + r * TB + threadIdx.x;
        if (m < n) p[m] = x[r];
    }
}
__device__ __forceinline__ void interval(const LevelDev &L, int k, int &s, int &e) {
    s = __ldg(L.cpts + k);
    e = (k + 1 < L.ncpts) ? __ldg(L.cpts + k + 1) : L.npts;
}
// x <- (g[i] +) Phi_i(x)
template <int Q>
__device__ __forceinline__ void advance(double (&x)[R], const DiagPhi<Q> &phi, const LevelDev &L, int i, int m0, bool add_g) {
    phi.step(x, L, i);
    if (add_g && L.g) {
        double g[R];
        ldrow(g, L.g, i, L.pitch, m0, L.n);
#pragma unroll
        for (int r = 0; r < R; ++r) x[r] = g[r] + x[r];
    }
}

// Steps i0 .. i1-1 in a row: x <- (g[i] +) Phi_i(x), stored after every step (STORE) or not at all.  The time factors of 32
// steps come with one coalesced load per warp (lane l holds step base + l; the next 32 are in flight meanwhile) and reach
// the steps by shuffle: a plain load per step is a dependent access that half of all issue slots waited for (ncu, long
// scoreboard 52 %: profiles/r02n_modes_full.txt).
template <int Q, bool STORE, bool HASG>
__device__ __forceinline__ void run_steps_g(double (&x)[R], const DiagPhi<Q> &phi, const LevelDev &L, int i0, int i1, int m0) {
    const int lane = threadIdx.x & 31;
    double nxt[Q > 0 ? Q : 1];
#pragma unroll
    for (int q = 0; q < Q; ++q) nxt[q] = (i0 + lane < i1) ? __ldg(L.rhs_t + (size_t)(i0 + lane) * Q + q) : 0.0;
    for (int base = i0; base < i1; base += 32) {
        double cur[Q > 0 ? Q : 1];
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            cur[q] = nxt[q];
            const int nb = base + 32 + lane;
            nxt[q] = (nb < i1) ? __ldg(L.rhs_t + (size_t)nb * Q + q) : 0.0;
        }
        const int cnt = min(32, i1 - base);
#pragma unroll 4
        for (int s = 0; s < cnt; ++s) {
            const int i = base + s;
            double ct[Q > 0 ? Q : 1];
#pragma unroll
            for (int q = 0; q < Q; ++q) ct[q] = __shfl_sync(0xffffffffu, cur[q], s);
            phi.step(x, L, i, ct);
            if (HASG) {
                double g[R];
                ldrow(g, L.g, i, L.pitch, m0, L.n);
#pragma unroll
                for (int r = 0; r < R; ++r) x[r] = g[r] + x[r];
            }
            if (STORE) strow(x, L.u, i, L.pitch, m0, L.n);
        }
    }
}
// (the test for g rows is taken out of the step loop: the loop of a level-0 chain is shuffles, FMAs and multiplications only)
template <int Q, bool STORE>
__device__ __forceinline__ void run_steps(double (&x)[R], const DiagPhi<Q> &phi, const LevelDev &L, int i0, int i1, int m0) {
    if (L.g)
        run_steps_g<Q, STORE, true>(x, phi, L, i0, i1, m0);
    else
        run_steps_g<Q, STORE, false>(x, phi, L, i0, i1, m0);
}

// One CTA per item: the hardware scheduler keeps every SM full of CTAs in different phases (row loads, arithmetic, stores).
// (Measured: a persistent grid whose CTAs prefetch the first rows of their next item is slower -- 5.3 instead of 4.75 ms per
// solve, profiles/r02m_modes_variants.txt: fewer, longer-lived CTAs overlap less than many short ones.)

// F-relaxation (mgrit.py:312-327); last_only: only the last F-point of an interval is stored.
template <int Q>
__global__ void __launch_bounds__(TB, MINB) k_chain(const LevelDev L, const int last_only) {
    MGB_RETURN_IF_STOPPED(L)
    for (int k = blockIdx.x; k < L.ncpts; k += gridDim.x) {
        int s, e;
        interval(L, k, s, e);
        if (e - s <= 1) continue;
        for (int m0 = 0; m0 < L.n; m0 += TB * R) {
            DiagPhi<Q> phi;
            phi.load(L, m0);
            double x[R];
            ldrow(x, L.u, s, L.pitch, m0, L.n);
            if (last_only) {
                run_steps<Q, false>(x, phi, L, s + 1, e, m0);
                strow(x, L.u, e - 1, L.pitch, m0, L.n);
            } else {
                run_steps<Q, true>(x, phi, L, s + 1, e, m0);
            }
        }
    }
}

// C-relaxation + F-relaxation + FAS restriction in one pass (sweeps.cuh k_down, same formulas); items = C-points j >= 1.
template <int Q>
__global__ void __launch_bounds__(TB, 3) k_down(const LevelDev L, const LevelDev G) {
    MGB_RETURN_IF_STOPPED(L)
    for (int j = 1 + blockIdx.x; j < L.ncpts; j += gridDim.x) {
        const int a = __ldg(L.cpts + j - 1), c = __ldg(L.cpts + j);
        for (int m0 = 0; m0 < L.n; m0 += TB * R) {
            DiagPhi<Q> phi;
            phi.load(L, m0);
            double x[R], yc[R], w[R];
            // time factors of the item's single steps, all in flight before the first row is waited for
            double ct_c[Q > 0 ? Q : 1], ct_a[Q > 0 ? Q : 1], ct_j[Q > 0 ? Q : 1];
            DiagPhi<Q>::factors(ct_c, L, c);
            DiagPhi<Q>::factors(ct_a, L, a == 0 ? c : a);
            DiagPhi<Q>::factors(ct_j, G, j);
            // C-relaxation of c
            ldrow(yc, L.u, c - 1, L.pitch, m0, L.n);
            ldrow(x, L.u, a == 0 ? 0 : a - 1, L.pitch, m0, L.n);
            phi.step(yc, L, c, ct_c);
            if (L.g) {
                double g[R];
                ldrow(g, L.g, c, L.pitch, m0, L.n);
#pragma unroll
                for (int r = 0; r < R; ++r) yc[r] = g[r] + yc[r];
            }
            strow(yc, L.u, c, L.pitch, m0, L.n);
            strow(yc, G.u, j, G.pitch, m0, L.n);  // injection
            // the C-relaxed left end
            if (a != 0) {
                phi.step(x, L, a, ct_a);
                if (L.g) {
                    double g[R];
                    ldrow(g, L.g, a, L.pitch, m0, L.n);
#pragma unroll
                    for (int r = 0; r < R; ++r) x[r] = g[r] + x[r];
                }
            }
            if (j == 1) strow(x, G.u, 0, G.pitch, m0, L.n);  // point 0 (initial condition / ghost) is injected too
            // w = Phi_c(x) with the coarse level's factors
            {
                DiagPhi<Q> cphi;
                cphi.load(G, m0);
#pragma unroll
                for (int r = 0; r < R; ++r) w[r] = x[r];
                cphi.step(w, G, j, ct_j);
            }
            // F-relaxation chain and the fine step into c
            run_steps<Q, false>(x, phi, L, a + 1, c, m0);
            phi.step(x, L, c, ct_c);
            // FAS right-hand side
            if (L.g) {
                double gc[R];
                ldrow(gc, L.g, c, L.pitch, m0, L.n);
#pragma unroll
                for (int r = 0; r < R; ++r) x[r] = (((gc[r] - yc[r]) + x[r]) + yc[r]) - w[r];
            } else {
#pragma unroll
                for (int r = 0; r < R; ++r) x[r] = ((x[r] - yc[r]) + yc[r]) - w[r];
            }
            strow(x, G.g, j, G.pitch, m0, L.n);
        }
    }
}

// Coarse-grid correction (+ the F-relaxation that follows; frelax == 2: only the last F-point is stored)
template <int Q>
__global__ void __launch_bounds__(TB, MINB) k_correct(const LevelDev L, const LevelDev G, const int frelax, const int kfirst) {
    MGB_RETURN_IF_STOPPED(L)
    for (int k = blockIdx.x; k < L.ncpts; k += gridDim.x) {
        int s, e;
        interval(L, k, s, e);
        const bool relax = frelax && (e - s > 1);
        if (k < kfirst && !relax) continue;
        for (int m0 = 0; m0 < L.n; m0 += TB * R) {
            double x[R];
            ldrow(x, L.u, s, L.pitch, m0, L.n);
            if (k >= kfirst) {
                double cu[R];
                ldrow(cu, G.u, k, G.pitch, m0, L.n);
#pragma unroll
                for (int r = 0; r < R; ++r) x[r] = x[r] + (cu[r] - x[r]);
                strow(x, L.u, s, L.pitch, m0, L.n);
            }
            if (!relax) continue;
            DiagPhi<Q> phi;
            phi.load(L, m0);
            if (frelax == 2) {
                run_steps<Q, false>(x, phi, L, s + 1, e, m0);
                strow(x, L.u, e - 1, L.pitch, m0, L.n);
            } else {
                run_steps<Q, true>(x, phi, L, s + 1, e, m0);
            }
        }
    }
}

// out_sq[k] = || (g[c] +) Phi(u[c-1]) - u[c] ||^2, k >= 1; fixed summation order
template <int Q>
__global__ void __launch_bounds__(TB, MINB) k_residual(const LevelDev L, double *__restrict__ out_sq) {
    MGB_RETURN_IF_STOPPED(L)
    __shared__ double part[TB / 32];
    if (blockIdx.x == 0 && threadIdx.x == 0) out_sq[0] = 0.0;
    for (int k = 1 + blockIdx.x; k < L.ncpts; k += gridDim.x) {
        const int c = __ldg(L.cpts + k);
        double acc = 0.0;
        for (int m0 = 0; m0 < L.n; m0 += TB * R) {
            DiagPhi<Q> phi;
            phi.load(L, m0);
            double x[R], y[R];
            ldrow(x, L.u, c - 1, L.pitch, m0, L.n);
            advance<Q>(x, phi, L, c, m0, true);
            ldrow(y, L.u, c, L.pitch, m0, L.n);
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const double d = x[r] - y[r];
                acc = fma(d, d, acc);
            }
        }
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
        __syncthreads();
        if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            double tot = 0.0;
#pragma unroll
            for (int w = 0; w < TB / 32; ++w) tot += part[w];
            out_sq[k] = tot;
        }
    }
}

static int grid_for(int items) {
    const DeviceInfo *di = device_info();
    const long cap = 64L * (di ? di->sms : 148);  // CTAs stride over the items beyond that
    return (int)(items < cap ? (items > 0 ? items : 1) : cap);
}

#define MGB_MODES_DISPATCH(CALL)                  \
    switch (L.nrhs) {                             \
        case 0: { constexpr int Q = 0; CALL; } break; \
        case 1: { constexpr int Q = 1; CALL; } break; \
        default: { constexpr int Q = 2; CALL; } break; \
    }

}  // namespace modes

// ---- entry points used by api.cu ---------------------------------------------------------------------------------
bool sine_modes_ok(const LevelDev &L) {
    const char *e = getenv("MGB_SINE_MODES");  // "0": the team kernels of sweeps.cuh instead (tests compare the two)
    const bool on = !(e && e[0] == '0');
    return on && L.nat != nullptr && L.nrhs >= 0 && L.nrhs <= 2 && L.rhs_dense == nullptr && L.cpts != nullptr &&
           (L.ndt == 1 || L.dtidx != nullptr);
}

// the coarse level of a pair only lends its tables (the coarsest level has no C-point list)
bool sine_modes_coarse_ok(const LevelDev &G, const LevelDev &L) {
    return G.nat != nullptr && G.nrhs == L.nrhs && G.rhs_dense == nullptr && (G.ndt == 1 || G.dtidx != nullptr);
}

int sine_modes_f_relax(const LevelDev &L, int flags, cudaStream_t st) {
    using namespace modes;
    if (L.ncpts < 1) return 0;
    MGB_MODES_DISPATCH((k_chain<Q><<<grid_for(L.ncpts), TB, 0, st>>>(L, flags & 1)))
    return cuda_fail(cudaGetLastError(), "f_relax");
}

int sine_modes_down(const LevelDev &L, const LevelDev &G, cudaStream_t st) {
    using namespace modes;
    if (L.ncpts < 2) return 0;
    MGB_MODES_DISPATCH((k_down<Q><<<grid_for(L.ncpts - 1), TB, 0, st>>>(L, G)))
    return cuda_fail(cudaGetLastError(), "down_sweep");
}

int sine_modes_correct(const LevelDev &L, const LevelDev &G, int frelax, int kfirst, cudaStream_t st) {
    using namespace modes;
    if (L.ncpts < 1) return 0;
    MGB_MODES_DISPATCH((k_correct<Q><<<grid_for(L.ncpts), TB, 0, st>>>(L, G, frelax, kfirst)))
    return cuda_fail(cudaGetLastError(), "error_correction");
}

int sine_modes_residual(const LevelDev &L, double *out_sq, cudaStream_t st) {
    using namespace modes;
    if (L.ncpts < 1) return 0;
    MGB_MODES_DISPATCH((k_residual<Q><<<grid_for(L.ncpts > 1 ? L.ncpts - 1 : 1), TB, 0, st>>>(L, out_sq)))
    return cuda_fail(cudaGetLastError(), "residual_norms");
}

}  // namespace mgb
