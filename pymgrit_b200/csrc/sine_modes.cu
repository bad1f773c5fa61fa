// sine_modes.cu -- the hot sweeps of the applications whose Phi is DIAGONAL in the basis of the level rows, with ONE
// THREAD PER MODE: Heat1D in sine space (MGB_APP_HEAT1D_SINE) and Heat2D (MGB_APP_HEAT2D: rows always hold sine
// coefficients).
//
// In sine space mode k of a time point depends on mode k of its predecessor and on nothing else.  The team kernels of
// sweeps.cuh (a row in the registers of one warp, rows staged through shared memory by bulk copies) were built for the
// Toeplitz solve, where the threads of a row must talk to each other; with a diagonal Phi that machinery is pure
// overhead -- 8 to 12 resident warps per SM, a generator / mbarrier hand-shake per row, and for Heat2D two table tiles
// re-read per work item.  Here a thread owns R modes of one coarse interval and walks the interval's time points:
// consecutive threads hold consecutive modes, so every load and store of a row is a coalesced access per CTA straight
// from / to HBM, there is no shared memory, a thread needs 50-80 registers, and an SM keeps 24-32 warps in flight.  The
// arithmetic per element is the team kernels', operation by operation (fma(ct, rx, x) * inv), so results are
// bit-identical to them.
//
// A CTA owns ONE chunk of TB*R consecutive modes and walks a strided set of coarse intervals; the per-mode tables
// (reciprocals, right-hand-side factors) are loaded once per CTA and stay in registers.  grid = chunks x interval groups.
//
// Heat1D: tables in natural mode order, one array per level (struct mgb_level.nat_dev): [2 + nrhs][pitch] =
//   lam_k | 1 / (1 + dt lam_k) for the level's dt (used when the level is uniform in time) | rxh_0 | rxh_1 ...
// Heat2D: the row-layout arrays of the team kernels (sig_dev, rhs_x_dev); a chunk is one tile of the row, so it holds
//   either sine coefficients or Dirichlet nodes (Phi(x) = the boundary value, whatever x was).
#include <cstdlib>

#include "../../include/mgrit_b200.h"
#include "phi.cuh"
#include "table.h"

namespace mgb {
namespace modes {

// the time factors of step i (issued early: the load is a dependent access that a step would otherwise wait for)
template <int Q>
__device__ __forceinline__ void factors(double (&ct)[Q > 0 ? Q : 1], const LevelDev &L, int i) {
#pragma unroll
    for (int q = 0; q < Q; ++q) ct[q] = __ldg(L.rhs_t + (size_t)i * Q + q);
}

// ---- Heat1D in sine space -------------------------------------------------------------------------------------------
template <int Q_>
struct DiagPhi {
    static constexpr int Q = Q_;
    static constexpr int TB = 256;  // threads per CTA
    static constexpr int R = 4;     // modes per thread: mode (chunk + r * TB + tid)
    static constexpr int MINB = 4;  // resident CTAs per SM the kernels are compiled for (<= 64 registers): the sweeps are
                                    // chains of dependent row loads per CTA, so the SM needs many CTAs in flight
    static constexpr int MINB_DOWN = 3;
    static constexpr bool kSplitNorm = false;  // a CTA of the residual kernel walks all chunks of a row
    static constexpr bool kHoist = false;  // the tables are read per item, where they are used (L1 / L2 hits): held across
                                           // the items of a CTA they cost registers (down-sweep of cfg5 0.258 -> 0.284 ms)
    double d[R];
    double rx[Q > 0 ? Q : 1][R];
    bool recip;

    __device__ static __forceinline__ int row_len(const LevelDev &L) { return L.n; }

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
    // the tables of the coarse level of a pair (the FAS restriction applies one coarse step)
    __device__ __forceinline__ void load_coarse(const LevelDev &G, int m0, const DiagPhi &) { load(G, m0); }

    // x <- Phi_i(x)   (same operations, in the same order, as Heat1DSine::apply)
    __device__ __forceinline__ void step(double (&x)[R], const LevelDev &L, int i, const double (&ct)[Q > 0 ? Q : 1]) {
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

// ---- Heat2D (theta method in sine space; phi.cuh Heat2D, same operations in the same order) -------------------------
// GEN = false: backward Euler on a level that is uniform in time -- a step is Q FMAs and one multiplication per element.
// GEN = true:  Crank-Nicolson / forward Euler and levels with several step sizes: the symbol stays in registers, the
//              reciprocals are remade when the step size changes.
template <int Q_, bool GEN, int TB_, int R_>
struct TilePhi {
    static constexpr int Q = Q_;
    static constexpr int TB = TB_, R = R_;
    static constexpr int MINB = (TB_ * 3 <= 1536 && !GEN && Q_ <= 1) ? 3 : 2;
    static constexpr int MINB_DOWN = 2;
    static constexpr bool kSplitNorm = true;  // one partial sum of squares per (C-point, chunk)
    static constexpr bool kHoist = true;  // the reciprocals are divisions and the tables long: made once per CTA
    double inv[R];  // 1 / (1 + theta dt sig) for the step size dtc; Dirichlet chunk: the boundary values
    double rx[Q > 0 ? Q : 1][R];
    double sg[GEN ? R : 1];
    double dtc, exc;
    bool bnd;

    __device__ static __forceinline__ int row_len(const LevelDev &L) { return L.pitch; }

    __device__ __forceinline__ void make(const LevelDev &L, int m0, int cls) {
        dtc = __ldg(L.sconst + (size_t)cls * L.cw);
        exc = __ldg(L.sconst + (size_t)cls * L.cw + 1);
        bnd = m0 >= L.ip[0] * L.tile;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int m = m0 + r * TB + threadIdx.x;
            const double s = (m < L.pitch) ? __ldg(L.sig + m) : 0.0;
            if (GEN) sg[r] = s;
            inv[r] = bnd ? s : __ddiv_rn(1.0, fma(dtc, s, 1.0));
        }
    }
    __device__ __forceinline__ void load(const LevelDev &L, int m0) {
        make(L, m0, 0);
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int m = m0 + r * TB + threadIdx.x;
#pragma unroll
            for (int q = 0; q < Q; ++q) rx[q][r] = (m < L.pitch) ? __ldg(L.rhs_x + (size_t)q * L.pitch + m) : 0.0;
        }
    }
    // all levels of a hierarchy share the symbol and the right-hand-side factors (checked by the C ABI)
    __device__ __forceinline__ void load_coarse(const LevelDev &G, int m0, const TilePhi &fine) {
        make(G, m0, 0);
#pragma unroll
        for (int r = 0; r < R; ++r) {
#pragma unroll
            for (int q = 0; q < Q; ++q) rx[q][r] = fine.rx[q][r];
        }
    }

    __device__ __forceinline__ void step(double (&x)[R], const LevelDev &L, int i, const double (&ct)[Q > 0 ? Q : 1]) {
        if (bnd) {
#pragma unroll
            for (int r = 0; r < R; ++r) x[r] = inv[r];
            return;
        }
        if (GEN) {
            const int cls = (L.ndt == 1) ? 0 : __ldg(L.dtidx + i);
            const double dt = __ldg(L.sconst + (size_t)cls * L.cw), ex = __ldg(L.sconst + (size_t)cls * L.cw + 1);
            if (dt != dtc) {
#pragma unroll
                for (int r = 0; r < R; ++r) inv[r] = __ddiv_rn(1.0, fma(dt, sg[r], 1.0));
                dtc = dt;
            }
            if (ex != 0.0) {  // (I - (1 - theta) dt L) u first (heat_2d.py:306, 352)
#pragma unroll
                for (int r = 0; r < R; ++r) x[r] = x[r] * fma(-ex, sg[r], 1.0);
            }
#pragma unroll
            for (int q = 0; q < Q; ++q) {
#pragma unroll
                for (int r = 0; r < R; ++r) x[r] = fma(ct[q], rx[q][r], x[r]);
            }
            if (dt != 0.0) {
#pragma unroll
                for (int r = 0; r < R; ++r) x[r] = x[r] * inv[r];
            }
        } else {
#pragma unroll
            for (int q = 0; q < Q; ++q) {
#pragma unroll
                for (int r = 0; r < R; ++r) x[r] = fma(ct[q], rx[q][r], x[r]);
            }
#pragma unroll
            for (int r = 0; r < R; ++r) x[r] = x[r] * inv[r];
        }
    }
};

// ---- rows ------------------------------------------------------------------------------------------------------------
template <class P>
__device__ __forceinline__ void ldrow(double (&x)[P::R], const double *__restrict__ base, int i, int pitch, int m0, int n) {
    const double *__restrict__ p = base + (size_t)i * pitch;
#pragma unroll
    for (int r = 0; r < P::R; ++r) {
        const int m = m0 + r * P::TB + threadIdx.x;
        x[r] = (m < n) ? p[m] : 0.0;
    }
}
template <class P>
__device__ __forceinline__ void strow(const double (&x)[P::R], double *__restrict__ base, int i, int pitch, int m0, int n) {
    double *__restrict__ p = base + (size_t)i * pitch;
#pragma unroll
    for (int r = 0; r < P::R; ++r) {
        const int m = m0 + r * P::TB + threadIdx.x;
        if (m < n) p[m] = x[r];
    }
}
__device__ __forceinline__ void interval(const LevelDev &L, int k, int &s, int &e) {
    s = __ldg(L.cpts + k);
    e = (k + 1 < L.ncpts) ? __ldg(L.cpts + k + 1) : L.npts;
}
// x <- (g[i] +) Phi_i(x)
template <class P>
__device__ __forceinline__ void advance(double (&x)[P::R], P &phi, const LevelDev &L, int i, int m0, int n) {
    double ct[P::Q > 0 ? P::Q : 1];
    factors<P::Q>(ct, L, i);
    phi.step(x, L, i, ct);
    if (L.g) {
        double g[P::R];
        ldrow<P>(g, L.g, i, L.pitch, m0, n);
#pragma unroll
        for (int r = 0; r < P::R; ++r) x[r] = g[r] + x[r];
    }
}

// Steps i0 .. i1-1 in a row: x <- (g[i] +) Phi_i(x), stored after every step (STORE) or not at all.  The time factors of 32
// steps come with one coalesced load per warp (lane l holds step base + l; the next 32 are in flight meanwhile) and reach
// the steps by shuffle: a plain load per step is a dependent access that half of all issue slots waited for (ncu, long
// scoreboard 52 %: profiles/r02n_modes_full.txt).
template <class P, bool STORE, bool HASG>
__device__ __forceinline__ void run_steps_g(double (&x)[P::R], P &phi, const LevelDev &L, int i0, int i1, int m0, int n) {
    constexpr int Q = P::Q, R = P::R;
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
                ldrow<P>(g, L.g, i, L.pitch, m0, n);
#pragma unroll
                for (int r = 0; r < R; ++r) x[r] = g[r] + x[r];
            }
            if (STORE) strow<P>(x, L.u, i, L.pitch, m0, n);
        }
    }
}
// (the test for g rows is taken out of the step loop: the loop of a level-0 chain is shuffles, FMAs and multiplications only)
template <class P, bool STORE>
__device__ __forceinline__ void run_steps(double (&x)[P::R], P &phi, const LevelDev &L, int i0, int i1, int m0, int n) {
    if (L.g)
        run_steps_g<P, STORE, true>(x, phi, L, i0, i1, m0, n);
    else
        run_steps_g<P, STORE, false>(x, phi, L, i0, i1, m0, n);
}

// grid = nch chunks x groups; CTA b owns chunk b % nch and the items (b / nch), (b / nch) + groups, ...
// With one item per CTA the hardware scheduler keeps every SM full of CTAs in different phases (row loads, arithmetic,
// stores).  (Measured: a persistent grid whose CTAs prefetch the first rows of their next item is slower -- 5.3 instead of
// 4.75 ms per solve, profiles/r02m_modes_variants.txt: fewer, longer-lived CTAs overlap less than many short ones.)
#define MGB_MODES_CTA(P, L, nch)                                   \
    const int n = P::row_len(L);                                   \
    const int m0 = (int)(blockIdx.x % (unsigned)nch) * P::TB * P::R; \
    const int kg = blockIdx.x / (unsigned)nch, ng = gridDim.x / (unsigned)nch;

// F-relaxation (mgrit.py:312-327); last_only: only the last F-point of an interval is stored.
template <class P>
__global__ void __launch_bounds__(P::TB, P::MINB) k_chain(const LevelDev L, const int last_only, const int nch) {
    MGB_RETURN_IF_STOPPED(L)
    MGB_MODES_CTA(P, L, nch)
    P phi;
    if (P::kHoist) phi.load(L, m0);
    for (int k = kg; k < L.ncpts; k += ng) {
        int s, e;
        interval(L, k, s, e);
        if (e - s <= 1) continue;
        if (!P::kHoist) phi.load(L, m0);
        double x[P::R];
        ldrow<P>(x, L.u, s, L.pitch, m0, n);
        if (last_only) {
            run_steps<P, false>(x, phi, L, s + 1, e, m0, n);
            strow<P>(x, L.u, e - 1, L.pitch, m0, n);
        } else {
            run_steps<P, true>(x, phi, L, s + 1, e, m0, n);
        }
    }
}

// C-relaxation + F-relaxation + FAS restriction in one pass (sweeps.cuh k_down, same formulas); items = C-points j >= 1.
// (Measured and rejected: a CTA that owns a block of consecutive C-points and keeps the C-relaxed point of item j - 1 as
// the left end of item j -- one row read and one Phi per item less -- is slower on every level but Heat2D's level 0:
// cfg5 4.33 -> 4.63 ms, cfg3 45.9 -> 49.8 ms, profiles/r02q_down_blocked.txt.)
template <class P>
__global__ void __launch_bounds__(P::TB, P::MINB_DOWN) k_down(const LevelDev L, const LevelDev G, const int nch) {
    MGB_RETURN_IF_STOPPED(L)
    MGB_MODES_CTA(P, L, nch)
    constexpr int Q = P::Q, R = P::R;
    P phi, cphi;
    if (P::kHoist) {
        phi.load(L, m0);
        cphi.load_coarse(G, m0, phi);
    }
    for (int j = 1 + kg; j < L.ncpts; j += ng) {
        const int a = __ldg(L.cpts + j - 1), c = __ldg(L.cpts + j);
        if (!P::kHoist) phi.load(L, m0);
        double x[R], yc[R], w[R];
        // time factors of the item's single steps, all in flight before the first row is waited for
        double ct_c[Q > 0 ? Q : 1], ct_a[Q > 0 ? Q : 1], ct_j[Q > 0 ? Q : 1];
        factors<Q>(ct_c, L, c);
        factors<Q>(ct_a, L, a == 0 ? c : a);
        factors<Q>(ct_j, G, j);
        // C-relaxation of c
        ldrow<P>(yc, L.u, c - 1, L.pitch, m0, n);
        ldrow<P>(x, L.u, a == 0 ? 0 : a - 1, L.pitch, m0, n);
        phi.step(yc, L, c, ct_c);
        if (L.g) {
            double g[R];
            ldrow<P>(g, L.g, c, L.pitch, m0, n);
#pragma unroll
            for (int r = 0; r < R; ++r) yc[r] = g[r] + yc[r];
        }
        strow<P>(yc, L.u, c, L.pitch, m0, n);
        strow<P>(yc, G.u, j, G.pitch, m0, n);  // injection
        // the C-relaxed left end
        if (a != 0) {
            phi.step(x, L, a, ct_a);
            if (L.g) {
                double g[R];
                ldrow<P>(g, L.g, a, L.pitch, m0, n);
#pragma unroll
                for (int r = 0; r < R; ++r) x[r] = g[r] + x[r];
            }
        }
        if (j == 1) strow<P>(x, G.u, 0, G.pitch, m0, n);  // point 0 (initial condition / ghost) is injected too
        // w = Phi_c(x) with the coarse level's factors
#pragma unroll
        for (int r = 0; r < R; ++r) w[r] = x[r];
        if (P::kHoist) {
            cphi.step(w, G, j, ct_j);
        } else {
            P cl;
            cl.load_coarse(G, m0, phi);
            cl.step(w, G, j, ct_j);
        }
        // F-relaxation chain and the fine step into c
        run_steps<P, false>(x, phi, L, a + 1, c, m0, n);
        phi.step(x, L, c, ct_c);
        // FAS right-hand side
        if (L.g) {
            double gc[R];
            ldrow<P>(gc, L.g, c, L.pitch, m0, n);
#pragma unroll
            for (int r = 0; r < R; ++r) x[r] = (((gc[r] - yc[r]) + x[r]) + yc[r]) - w[r];
        } else {
#pragma unroll
            for (int r = 0; r < R; ++r) x[r] = ((x[r] - yc[r]) + yc[r]) - w[r];
        }
        strow<P>(x, G.g, j, G.pitch, m0, n);
    }
}

// Coarse-grid correction (+ the F-relaxation that follows; frelax == 2: only the last F-point is stored)
template <class P>
__global__ void __launch_bounds__(P::TB, P::MINB) k_correct(const LevelDev L, const LevelDev G, const int frelax,
                                                            const int kfirst, const int nch) {
    MGB_RETURN_IF_STOPPED(L)
    MGB_MODES_CTA(P, L, nch)
    constexpr int R = P::R;
    P phi;
    if (frelax && P::kHoist) phi.load(L, m0);
    for (int k = kg; k < L.ncpts; k += ng) {
        int s, e;
        interval(L, k, s, e);
        const bool relax = frelax && (e - s > 1);
        if (k < kfirst && !relax) continue;
        double x[R];
        ldrow<P>(x, L.u, s, L.pitch, m0, n);
        if (k >= kfirst) {
            double cu[R];
            ldrow<P>(cu, G.u, k, G.pitch, m0, n);
#pragma unroll
            for (int r = 0; r < R; ++r) x[r] = x[r] + (cu[r] - x[r]);
            strow<P>(x, L.u, s, L.pitch, m0, n);
        }
        if (!relax) continue;
        if (!P::kHoist) phi.load(L, m0);
        if (frelax == 2) {
            run_steps<P, false>(x, phi, L, s + 1, e, m0, n);
            strow<P>(x, L.u, e - 1, L.pitch, m0, n);
        } else {
            run_steps<P, true>(x, phi, L, s + 1, e, m0, n);
        }
    }
}

// out_sq[k] = || (g[c] +) Phi(u[c-1]) - u[c] ||^2, k >= 1; fixed summation order.  kSplitNorm: one partial per chunk at
// out_sq[ncpts + k * nch + chunk] (the layout of sweeps.cuh's k_residual with several systems per row), added by k_sum_chunks.
template <class P>
__global__ void __launch_bounds__(P::TB, P::MINB) k_residual(const LevelDev L, double *__restrict__ out_sq, const int nch) {
    MGB_RETURN_IF_STOPPED(L)
    constexpr int R = P::R, CH = P::TB * P::R;
    const int n = P::row_len(L);
    const int split = P::kSplitNorm ? nch : 1;
    const int part = blockIdx.x % (unsigned)split, kg = blockIdx.x / (unsigned)split, ng = gridDim.x / (unsigned)split;
    const int mbeg = P::kSplitNorm ? part * CH : 0, mend = P::kSplitNorm ? mbeg + CH : n;
    __shared__ double partial[P::TB / 32];
    if (blockIdx.x == 0 && threadIdx.x == 0) out_sq[0] = 0.0;
    for (int k = 1 + kg; k < L.ncpts; k += ng) {
        const int c = __ldg(L.cpts + k);
        double acc = 0.0;
        for (int m0 = mbeg; m0 < mend; m0 += CH) {
            P phi;
            phi.load(L, m0);
            double x[R], y[R];
            ldrow<P>(x, L.u, c - 1, L.pitch, m0, n);
            advance<P>(x, phi, L, c, m0, n);
            ldrow<P>(y, L.u, c, L.pitch, m0, n);
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const double d = x[r] - y[r];
                acc = fma(d, d, acc);
            }
        }
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
        __syncthreads();
        if ((threadIdx.x & 31) == 0) partial[threadIdx.x >> 5] = acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            double tot = 0.0;
#pragma unroll
            for (int w = 0; w < P::TB / 32; ++w) tot += partial[w];
            out_sq[P::kSplitNorm ? L.ncpts + k * nch + part : k] = tot;
        }
    }
}

// out_sq[k] = sum over the chunks of a row, k = 1 .. ncpts-1 (fixed order: deterministic); one warp per C-point
__global__ void k_sum_chunks(double *__restrict__ out_sq, const int ncpts, const int nch, const int *__restrict__ stop) {
    if (stop != nullptr && *stop != 0) return;
    const int k = 1 + blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (k >= ncpts) return;
    const double *part = out_sq + ncpts + (size_t)k * nch;
    double acc = 0.0;
    for (int s = lane; s < nch; s += 32) acc += part[s];
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
    if (lane == 0) out_sq[k] = acc;
}

// chunks x groups CTAs: as many groups as there are items up to `cap` CTAs, beyond that CTAs stride over the items
static int env_int(const char *name, int dflt) {
    const char *e = getenv(name);
    return (e && e[0]) ? atoi(e) : dflt;
}
static int grid_for(int items, int nch, bool tile) {
    const DeviceInfo *di = device_info();
    // Heat2D rows are long (254 chunks at 512 x 512): 8 CTAs per SM walk their intervals with the tables in registers
    // (64: 46.8 ms, 16: 46.6, 8: 45.9, 4: 57.8 ms per cfg3 solve, profiles/r02p_timeline_cfg3_modes_vs_team.txt)
    static const int cap_env = env_int("MGB_MODES_CTAS_PER_SM", 0);
    const int cap_sm = cap_env > 0 ? cap_env : (tile ? 8 : 64);
    const long cap = (long)cap_sm * (di ? di->sms : 148);
    long groups = cap / nch;
    if (groups < 1) groups = 1;
    if (groups > items) groups = items > 0 ? items : 1;
    return (int)(groups * nch);
}

static bool is_tile(const LevelDev &L) { return L.sig != nullptr; }
// Heat2D shapes of a CTA: threads x modes per thread = one tile of the row
constexpr int kTileTB = 384, kTileR = 3;
static bool tile_simple(const LevelDev &L) { return L.ip[1] == 1 && L.ndt == 1; }  // backward Euler, one step size

#define MGB_MODES_Q(QMAX, PH, CALL)                                       \
    switch (L.nrhs) {                                                     \
        case 0: { using P = PH(0); CALL; } break;                         \
        case 1: { using P = PH(1); CALL; } break;                         \
        case 2: { using P = PH(2); CALL; } break;                         \
        default: { using P = PH(QMAX); CALL; } break;                     \
    }
#define MGB_PH_DIAG(q) DiagPhi<q>
#define MGB_PH_TILE_S(q) TilePhi<q, false, kTileTB, kTileR>
#define MGB_PH_TILE_G(q) TilePhi<q, true, kTileTB, kTileR>
// `simple` (Heat2D only): every level the kernel touches is backward Euler with one step size
#define MGB_MODES_DISPATCH(simple, CALL)                                  \
    if (!is_tile(L)) {                                                    \
        const int nch = (L.n + 1023) / 1024;                              \
        MGB_MODES_Q(2, MGB_PH_DIAG, CALL)                                 \
    } else if (simple) {                                                  \
        const int nch = L.pitch / (kTileTB * kTileR);                     \
        MGB_MODES_Q(3, MGB_PH_TILE_S, CALL)                               \
    } else {                                                              \
        const int nch = L.pitch / (kTileTB * kTileR);                     \
        MGB_MODES_Q(3, MGB_PH_TILE_G, CALL)                               \
    }

}  // namespace modes

// ---- entry points used by api.cu ---------------------------------------------------------------------------------
// MGB_SINE_MODES_MASK (experiments): bit 0 f_relax, 1 down_sweep, 2 error_correction, 3 residual_norms; default all
static bool mode_entry_on(int bit) {
    const char *e = getenv("MGB_SINE_MODES_MASK");
    return !(e && e[0]) || ((atoi(e) >> bit) & 1);
}

bool sine_modes_entry(int bit) { return mode_entry_on(bit); }

bool sine_modes_ok(const LevelDev &L) {
    const char *e = getenv("MGB_SINE_MODES");  // "0": the team kernels of sweeps.cuh instead (tests compare the two)
    const bool on = !(e && e[0] == '0');
    if (!on || L.rhs_dense != nullptr || L.cpts == nullptr || !(L.ndt == 1 || L.dtidx != nullptr)) return false;
    if (modes::is_tile(L))  // Heat2D: a chunk must be one tile of the row
        return L.tile == modes::kTileTB * modes::kTileR && L.pitch % L.tile == 0 && L.nrhs >= 0 && L.nrhs <= 3 &&
               L.sconst != nullptr && (L.nrhs == 0 || L.rhs_x != nullptr);
    return L.nat != nullptr && L.nrhs >= 0 && L.nrhs <= 2;
}

// the coarse level of a pair only lends its tables (the coarsest level has no C-point list)
bool sine_modes_coarse_ok(const LevelDev &G, const LevelDev &L) {
    if (G.rhs_dense != nullptr || !(G.ndt == 1 || G.dtidx != nullptr)) return false;
    if (modes::is_tile(L)) return G.sig == L.sig && G.sconst != nullptr && G.nrhs == L.nrhs && (G.nrhs == 0 || G.rhs_t != nullptr);
    return G.nat != nullptr && G.nrhs == L.nrhs;
}

int sine_modes_f_relax(const LevelDev &L, int flags, cudaStream_t st) {
    using namespace modes;
    if (L.ncpts < 1) return 0;
    MGB_MODES_DISPATCH(tile_simple(L), (k_chain<P><<<grid_for(L.ncpts, nch, is_tile(L)), P::TB, 0, st>>>(L, flags & 1, nch)))
    return cuda_fail(cudaGetLastError(), "f_relax");
}

int sine_modes_down(const LevelDev &L, const LevelDev &G, cudaStream_t st) {
    using namespace modes;
    if (L.ncpts < 2) return 0;
    MGB_MODES_DISPATCH(tile_simple(L) && tile_simple(G), (k_down<P><<<grid_for(L.ncpts - 1, nch, is_tile(L)), P::TB, 0, st>>>(L, G, nch)))
    return cuda_fail(cudaGetLastError(), "down_sweep");
}

int sine_modes_correct(const LevelDev &L, const LevelDev &G, int frelax, int kfirst, cudaStream_t st) {
    using namespace modes;
    if (L.ncpts < 1) return 0;
    MGB_MODES_DISPATCH(tile_simple(L), (k_correct<P><<<grid_for(L.ncpts, nch, is_tile(L)), P::TB, 0, st>>>(L, G, frelax, kfirst, nch)))
    return cuda_fail(cudaGetLastError(), "error_correction");
}

int sine_modes_residual(const LevelDev &L, double *out_sq, cudaStream_t st) {
    using namespace modes;
    if (L.ncpts < 1) return 0;
    const int items = L.ncpts > 1 ? L.ncpts - 1 : 1;
    MGB_MODES_DISPATCH(tile_simple(L), (k_residual<P><<<grid_for(items, P::kSplitNorm ? nch : 1, is_tile(L)), P::TB, 0, st>>>(L, out_sq, nch)))
    if (is_tile(L) && L.ncpts > 1) {
        const int nch = L.pitch / (kTileTB * kTileR);
        k_sum_chunks<<<(L.ncpts - 1 + 7) / 8, 256, 0, st>>>(out_sq, L.ncpts, nch, L.stop);
    }
    return cuda_fail(cudaGetLastError(), "residual_norms");
}

}  // namespace mgb
