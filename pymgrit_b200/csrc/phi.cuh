// phi.cuh -- the time integrators Phi as register-resident device functors.
//
//  Heat1D      heat/heat_1d.py:198-217       y = (I + dt L)^-1 (u + dt b(x, t_stop)),  L = (a/dx^2) tridiag(-1,2,-1)
//  Advection1D advection/advection_1d.py:129-143   y = (I + dt L)^-1 u,  L = (c/dx)(I - S_periodic)
//
// Both matrices are Toeplitz, so the solve is done without any per-row table:
//   heat:  (1/r)(I + dt L) = tridiag(-1, delta, -1), delta = 2 + 1/r, = L U + beta e0 e0^T with
//          L = bidiag(alpha; -1), U = bidiag(1; -beta), alpha + beta = delta, alpha beta = 1, beta < 1.
//          L and U are inverted by the two constant-coefficient recurrences
//              y_i = beta (f_i + y_{i-1}),      z_i = y_i + beta z_{i+1},
//          and the rank-one term by Sherman-Morrison with the closed-form h = (LU)^-1 e0.
//   advection: y_i = sigma u_i + rho y_{i-1} (cyclic), closed by y_{n-1} = p_{n-1} / (1 - rho^n).
// Each thread runs the recurrence over its own chunk (split into SUB independent sub-chunks for
// instruction-level parallelism); the chunk-to-chunk carries are one constant-ratio scan over the
// team per direction (warp shuffles; shared memory only across warps).  This is the "Thomas inside
// a lane / parallel scan across lanes" hybrid named in BASELINE.json's north_star.
#pragma once
#include "common.cuh"

namespace mgb {

// Kernel-side view of mgb_level (include/mgrit_b200.h).
struct LevelDev {
    double *u;
    double *g;
    const int *cpts;
    int ncpts;
    int npts;
    int n;
    int pitch;
    int ndt;
    int cw;
    const int *dtidx;
    const double *sconst;
    int nrhs;
    const double *rhs_x;
    const double *rhs_t;
    const double *rhs_dense;
    const double *t;  // [npts] time values (ODE applications)
    double p[8];
    int ip[4];
    // independent spatial systems per row (sweeps.cuh: work items).  n = dofs of ONE system, tile = doubles between two
    // systems of a row = length of the slice a team moves, nrow = dofs of the whole row (n for the 1-D applications)
    int nsys;
    int tile;
    int nrow;
    const double *sig;  // [pitch] per-element symbol of the spatial operator (Heat2D), row layout; NULL otherwise
    const double *diag; // Heat1DSine: [2][E][T] thread-transposed eigenvalues lam_k and reciprocals 1/(1 + dt lam_k)
    const double *nat;  // Heat1DSine: [2 + nrhs][pitch] the same tables in natural mode order (sine_modes.cu), or nullptr
    const int *stop;    // device flag set by mgb_convergence_flag once the stopping criterion is met: sweeps that were
                        // queued ahead of that knowledge return at once (mgb_set_stop_flag); nullptr = not in use
};

// First statement of every sweep kernel.
#define MGB_RETURN_IF_STOPPED(L) \
    if ((L).stop != nullptr && *(L).stop != 0) return;

// Applications without per-element item data
struct NoItem {};

constexpr int kScalarConsts = 24;  // doubles before the per-thread part of a step-constant row

template <int E>
struct SubSplit {
    static constexpr int SUB = (E % 3 == 0) ? 3 : 1;
    static constexpr int SL = E / SUB;
};

// ---------------------------------------------------------------------------------------------
// Heat1D.  Step-constant row layout (doubles), written by mgb_heat1d_step_consts:
//   [0] beta  [1] beta/r  [2] kappa = beta/(1+beta h0)  [3] beta^SL  [4..8] B^1,2,4,8,16 (B = beta^E)
//   [9] B^32  [10..10+SL) beta^(jj+1)   [23] number of scan steps whose power B^(2^k) is >= 1e-30
//   [24 + tid*PT ...): B^lane, B^(31-lane), PH[SUB], QH[SUB]      (PT = 2 + 2 SUB)
// with PH[s] = beta^(tid E + s SL)/(1-beta^2), QH[s] = beta^(2n+1-tid E-(s+1) SL)/(1-beta^2) (0 where
// the sub-chunk holds no valid element).
// ---------------------------------------------------------------------------------------------
template <int T_, int E_>
struct Heat1D {
    using SH = Shape<T_, E_>;
    static constexpr bool kTightChain = false;
    static constexpr bool kCtArg = false;
    static constexpr bool kItemInvariant = true;  // begin_item depends on the level only, not on the work item
    static constexpr int T = T_, E = E_;
    static constexpr int SUB = SubSplit<E_>::SUB, SL = SubSplit<E_>::SL;
    static constexpr int PT = 2 + 2 * SUB;
    static_assert(SL <= 13, "power table does not fit the scalar block");

    struct C {
        double beta, cs, kappa, bsl, Bd[5], B32, pw[SL], blf, blb, PH[SUB], QH[SUB];
        int nscan;
    };
    // doubles of a row slice that carry values (the rest of the tile is padding the pipe reads as 0)
    __device__ static __forceinline__ int row_n(const LevelDev &L) { return L.n; }

    __device__ static __forceinline__ void load_consts(C &c, const double *__restrict__ row, int tid) {
        c.beta = __ldg(row + 0);
        c.cs = __ldg(row + 1);
        c.kappa = __ldg(row + 2);
        c.bsl = __ldg(row + 3);
#pragma unroll
        for (int k = 0; k < 5; ++k) c.Bd[k] = __ldg(row + 4 + k);
        c.B32 = __ldg(row + 9);
#pragma unroll
        for (int j = 0; j < SL; ++j) c.pw[j] = __ldg(row + 10 + j);
        c.nscan = (int)__ldg(row + 23);
        const double *pt = row + kScalarConsts + tid * PT;
        c.blf = __ldg(pt + 0);
        c.blb = __ldg(pt + 1);
#pragma unroll
        for (int s = 0; s < SUB; ++s) {
            c.PH[s] = __ldg(pt + 2 + s);
            c.QH[s] = __ldg(pt + 2 + SUB + s);
        }
    }

    // x <- Phi(x) for the step that produces point i.  (the separable right-hand side is added here;
    // a dense right-hand-side row is added by the caller before.)
    // The first spatial factor of the right-hand side stays in registers for all the steps of a work item: read from
    // the table per step (33 loads through an L1 that the staging slots leave little of) it cost a quarter of the
    // FP64-bound chains.
    struct Item {
        double rx0[E];
    };
    template <class Pipe, class TeamT>
    __device__ static __forceinline__ void begin_item(Item &it, const LevelDev &L, int, Pipe &, TeamT &team) {
        if (L.nrhs > 0) {
            const double *__restrict__ rx = L.rhs_x + team.tid;
#pragma unroll
            for (int j = 0; j < E; ++j) it.rx0[j] = __ldg(rx + j * T);
        }
    }
    // the item's data when the next Phi belongs to another level (FAS restriction: fine step, then coarse step)
    template <class TeamT>
    __device__ static __forceinline__ void retarget_item(Item &it, const LevelDev &from, const LevelDev &to, TeamT &team) {
        if (to.nrhs > 0 && to.rhs_x != from.rhs_x) {
            const double *__restrict__ rx = to.rhs_x + team.tid;
#pragma unroll
            for (int j = 0; j < E; ++j) it.rx0[j] = __ldg(rx + j * T);
        }
    }

    template <class TeamT>
    __device__ static __forceinline__ void apply(double (&x)[E], const C &c, Item &it, const LevelDev &L, int i,
                                                 TeamT &team) {
        const int tid = team.tid;
        // b = u + dt * rhs(x, t_i)                                           heat_1d.py:214
        if (L.nrhs > 0) {
            const double ct = __ldg(L.rhs_t + (size_t)i * L.nrhs);
#pragma unroll
            for (int j = 0; j < E; ++j) x[j] = fma(ct, it.rx0[j], x[j]);
        }
        for (int k = 1; k < L.nrhs; ++k) {
            const double ct = __ldg(L.rhs_t + (size_t)i * L.nrhs + k);
            const double *__restrict__ rx = L.rhs_x + (size_t)k * E * T + tid;
#pragma unroll
            for (int j = 0; j < E; ++j) x[j] = fma(ct, __ldg(rx + j * T), x[j]);
        }
        solve(x, c, L.n, team);
    }

    // x <- (I + r tridiag(-1, 2, -1))^-1 x for the r the constants c were made for; n = unknowns of the system.
    // Elements beyond n must be 0 on entry.
    template <class TeamT>
    __device__ static __forceinline__ void solve(double (&x)[E], const C &c, const int n, TeamT &team) {
        const int tid = team.tid;
        const int nv = n - tid * E;
        // f = b * beta / r, forward recurrence y = beta (f + y_prev) inside each sub-chunk
        double e[SUB];
#pragma unroll
        for (int s = 0; s < SUB; ++s) e[s] = 0.0;
#pragma unroll
        for (int jj = 0; jj < SL; ++jj) {
#pragma unroll
            for (int s = 0; s < SUB; ++s) {
                const int j = s * SL + jj;
                e[s] = fma(c.beta, e[s], x[j] * c.cs);
                x[j] = e[s];
            }
        }
        double a = e[0];
#pragma unroll
        for (int s = 1; s < SUB; ++s) a = fma(c.bsl, a, e[s]);
        double in = team.scan_fwd(a, c.Bd, c.B32, c.blf, c.nscan);
        // add the inflow; everything beyond element n-1 must stay exactly 0 for the backward recurrence
        if (n % E == 0) {
            // no thread holds a partially valid chunk (e.g. n = 1023 = 31 x 33): threads beyond n have zero data, so
            // cutting their inflow keeps them at 0 without a select per element
            in = (nv > 0) ? in : 0.0;
#pragma unroll
            for (int s = 0; s < SUB; ++s) {
#pragma unroll
                for (int jj = 0; jj < SL; ++jj) {
                    const int j = s * SL + jj;
                    x[j] = fma(c.pw[jj], in, x[j]);
                }
                in = fma(c.bsl, in, e[s]);
            }
        } else {
#pragma unroll
            for (int s = 0; s < SUB; ++s) {
#pragma unroll
                for (int jj = 0; jj < SL; ++jj) {
                    const int j = s * SL + jj;
                    const double y = fma(c.pw[jj], in, x[j]);
                    x[j] = (j < nv) ? y : 0.0;
                }
                in = fma(c.bsl, in, e[s]);
            }
        }
        // backward recurrence z = y + beta z_next
        double f[SUB];
#pragma unroll
        for (int s = 0; s < SUB; ++s) f[s] = 0.0;
#pragma unroll
        for (int jj = SL - 1; jj >= 0; --jj) {
#pragma unroll
            for (int s = 0; s < SUB; ++s) {
                const int j = s * SL + jj;
                f[s] = fma(c.beta, f[s], x[j]);
                x[j] = f[s];
            }
        }
        double a2 = f[SUB - 1];
#pragma unroll
        for (int s = SUB - 2; s >= 0; --s) a2 = fma(c.bsl, a2, f[s]);
        double inb[SUB];
        inb[SUB - 1] = team.scan_bwd(a2, c.Bd, c.B32, c.blb, c.nscan);
#pragma unroll
        for (int s = SUB - 2; s >= 0; --s) inb[s] = fma(c.bsl, inb[s + 1], f[s + 1]);
        // Sherman-Morrison:  x = z - gamma h,  gamma = kappa * z_0
        const double z0 = team.bcast(fma(c.pw[SL - 1], inb[0], x[0]), 0);
        const double gamma = c.kappa * z0;
#pragma unroll
        for (int s = 0; s < SUB; ++s) {
            const double Bc = fma(gamma, c.QH[s], inb[s]);
            const double Ac = -gamma * c.PH[s];
#pragma unroll
            for (int jj = 0; jj < SL; ++jj) {
                const int j = s * SL + jj;
                x[j] = fma(c.pw[SL - 1 - jj], Bc, fma(c.pw[jj], Ac, x[j]));
            }
        }
    }
};

// ---------------------------------------------------------------------------------------------
// Heat1D2Pts: the two-point (pair) states of heat/heat_1d_2pts_bdf1.py:90-117 and heat/heat_1d_2pts_bdf2.py:92-138.
// A time point holds the values at t_i ("first") and t_i + dtau ("second"); one Phi is two tridiagonal solves
//     s1   = a1 first + b1 second + sum_k ct1_k(i) X_k        tmp1 = (I + r1 tridiag(-1,2,-1))^-1 s1
//     s2   = a2 second + b2 tmp1  + sum_k ct2_k(i) X_k        tmp2 = (I + r2 tridiag(-1,2,-1))^-1 s2
//     Phi(first, second) = (tmp1, tmp2)
// BDF1: a1 = a2 = 0, b1 = b2 = 1, r1 = (dt - dtau) a/dx^2, r2 = dtau a/dx^2, ct1 = (dt - dtau) T_k(t_i),
//       ct2 = dtau T_k(t_i + dtau).
// BDF2: the reference solves (L + coeff I) y = rhs - coeffm2 u_{-2} + coeffm1 u_{-1}; divided by coeff this is the same
//       form with r = (a/dx^2)/coeff, a = -coeffm2/coeff, b = coeffm1/coeff, ct = T_k / coeff (host: heat_1d_2pts.py).
// The method is a property of the level's constants, so a BDF2 fine level over BDF1 coarse levels
// (examples/example_heat_1d_bdf2.py) runs through the same kernels.
// Row layout (private to the engine, values enter and leave through Heat1D2Pts.values_to_rows / rows_to_values): thread
// tid owns row[tid E, tid E + E) = first[tid EH, +EH) | second[tid EH, +EH) | one zero, E = 2 EH + 1 (odd stride).
// Step-constant row: [Heat1D constants for r1 (width W)] [Heat1D constants for r2 (W)] [a1 b1 a2 b2 skip1 0 0 0],
// W = kScalarConsts + T (2 + 2 SUB(EH)); skip1 != 0: r1 = 0, the first solve is the identity (dt = dtau).
// Right-hand-side time table: [npts][2 nrhs] = ct1_0.. ct1_{q-1} ct2_0 .. ct2_{q-1}.
// ---------------------------------------------------------------------------------------------
template <int T_, int E_>
struct Heat1D2Pts {
    using SH = Shape<T_, E_>;
    static constexpr bool kTightChain = false;
    static constexpr bool kCtArg = false;
    static constexpr bool kItemInvariant = true;  // begin_item depends on the level only, not on the work item
    static constexpr int T = T_, E = E_;
    static constexpr int EH = (E_ - 1) / 2;
    using H = Heat1D<T_, EH>;
    static constexpr int W = kScalarConsts + T_ * H::PT;

    struct C {
        typename H::C s1, s2;
        double a1, b1, a2, b2;
        bool skip1;
    };
    __device__ static __forceinline__ int row_n(const LevelDev &L) { return L.tile; }

    __device__ static __forceinline__ void load_consts(C &c, const double *__restrict__ row, int tid) {
        H::load_consts(c.s1, row, tid);
        H::load_consts(c.s2, row + W, tid);
        c.a1 = __ldg(row + 2 * W + 0);
        c.b1 = __ldg(row + 2 * W + 1);
        c.a2 = __ldg(row + 2 * W + 2);
        c.b2 = __ldg(row + 2 * W + 3);
        c.skip1 = __ldg(row + 2 * W + 4) != 0.0;
    }

    struct Item {
        double rx0[EH];
    };
    __device__ static __forceinline__ void load_rx0(Item &it, const LevelDev &L, int tid) {
        const double *__restrict__ rx = L.rhs_x + tid;
#pragma unroll
        for (int j = 0; j < EH; ++j) it.rx0[j] = __ldg(rx + j * T);
    }
    template <class Pipe, class TeamT>
    __device__ static __forceinline__ void begin_item(Item &it, const LevelDev &L, int, Pipe &, TeamT &team) {
        if (L.nrhs > 0) load_rx0(it, L, team.tid);
    }
    template <class TeamT>
    __device__ static __forceinline__ void retarget_item(Item &it, const LevelDev &from, const LevelDev &to, TeamT &team) {
        if (to.nrhs > 0 && to.rhs_x != from.rhs_x) load_rx0(it, to, team.tid);
    }

    // y += sum_k ct_k X_k, ct = L.rhs_t[i][half * nrhs + k]
    __device__ static __forceinline__ void add_rhs(double (&y)[EH], const Item &it, const LevelDev &L, int i, int half,
                                                   int tid) {
        if (L.nrhs <= 0) return;
        const double *__restrict__ ctp = L.rhs_t + ((size_t)i * 2 + half) * L.nrhs;
        const double ct0 = __ldg(ctp);
#pragma unroll
        for (int j = 0; j < EH; ++j) y[j] = fma(ct0, it.rx0[j], y[j]);
        for (int k = 1; k < L.nrhs; ++k) {
            const double ct = __ldg(ctp + k);
            const double *__restrict__ rx = L.rhs_x + (size_t)k * EH * T + tid;
#pragma unroll
            for (int j = 0; j < EH; ++j) y[j] = fma(ct, __ldg(rx + j * T), y[j]);
        }
    }

    template <class TeamT>
    __device__ static __forceinline__ void apply(double (&x)[E], const C &c, Item &it, const LevelDev &L, int i,
                                                 TeamT &team) {
        const int tid = team.tid;
        const int nv = L.n - tid * EH;
        double p[EH], q[EH];
#pragma unroll
        for (int j = 0; j < EH; ++j) p[j] = fma(c.a1, x[j], c.b1 * x[EH + j]);
        add_rhs(p, it, L, i, 0, tid);
#pragma unroll
        for (int j = 0; j < EH; ++j) p[j] = (j < nv) ? p[j] : 0.0;
        if (!c.skip1) H::solve(p, c.s1, L.n, team);
#pragma unroll
        for (int j = 0; j < EH; ++j) q[j] = fma(c.a2, x[EH + j], c.b2 * p[j]);
        add_rhs(q, it, L, i, 1, tid);
#pragma unroll
        for (int j = 0; j < EH; ++j) q[j] = (j < nv) ? q[j] : 0.0;
        H::solve(q, c.s2, L.n, team);
#pragma unroll
        for (int j = 0; j < EH; ++j) {
            x[j] = (j < nv) ? p[j] : 0.0;
            x[EH + j] = (j < nv) ? q[j] : 0.0;
        }
        x[2 * EH] = 0.0;
    }
};

// ---------------------------------------------------------------------------------------------
// Advection1D.  Step-constant row layout:
//   [0] rho  [1] sigma  [2] 1/(1-rho^n)  [3] rho^SL  [4..8] B^1,2,4,8,16 (B = rho^E)  [9] B^32
//   [10..10+SL) rho^(jj+1)   [23] number of scan steps (as for Heat1D)
//   [24 + tid*PT ...): B^lane, (unused), RH[SUB] = rho^(tid E + s SL), (unused)[SUB]
// ---------------------------------------------------------------------------------------------
template <int T_, int E_>
struct Advection1D {
    using SH = Shape<T_, E_>;
    static constexpr bool kTightChain = false;
    static constexpr bool kCtArg = false;
    static constexpr bool kItemInvariant = true;  // begin_item depends on the level only, not on the work item
    static constexpr int T = T_, E = E_;
    static constexpr int SUB = SubSplit<E_>::SUB, SL = SubSplit<E_>::SL;
    static constexpr int PT = 2 + 2 * SUB;

    struct C {
        double rho, sig, dinv, rsl, Bd[5], B32, pw[SL], blf, RH[SUB];
        int nscan;
    };
    __device__ static __forceinline__ int row_n(const LevelDev &L) { return L.n; }

    __device__ static __forceinline__ void load_consts(C &c, const double *__restrict__ row, int tid) {
        c.rho = __ldg(row + 0);
        c.sig = __ldg(row + 1);
        c.dinv = __ldg(row + 2);
        c.rsl = __ldg(row + 3);
#pragma unroll
        for (int k = 0; k < 5; ++k) c.Bd[k] = __ldg(row + 4 + k);
        c.B32 = __ldg(row + 9);
#pragma unroll
        for (int j = 0; j < SL; ++j) c.pw[j] = __ldg(row + 10 + j);
        c.nscan = (int)__ldg(row + 23);
        const double *pt = row + kScalarConsts + tid * PT;
        c.blf = __ldg(pt + 0);
#pragma unroll
        for (int s = 0; s < SUB; ++s) c.RH[s] = __ldg(pt + 2 + s);
    }

    using Item = NoItem;
    template <class Pipe, class TeamT>
    __device__ static __forceinline__ void begin_item(Item &, const LevelDev &, int, Pipe &, TeamT &) {}
    template <class TeamT>
    __device__ static __forceinline__ void retarget_item(Item &, const LevelDev &, const LevelDev &, TeamT &) {}

    template <class TeamT>
    __device__ static __forceinline__ void apply(double (&x)[E], const C &c, Item &, const LevelDev &L, int i,
                                                 TeamT &team) {
        const int tid = team.tid;
        const int last = L.n - 1;
        const int jstar = last - tid * E;  // position of element n-1 in this thread's chunk (if 0 <= jstar < E)
        double e[SUB];
#pragma unroll
        for (int s = 0; s < SUB; ++s) e[s] = 0.0;
#pragma unroll
        for (int jj = 0; jj < SL; ++jj) {
#pragma unroll
            for (int s = 0; s < SUB; ++s) {
                const int j = s * SL + jj;
                e[s] = fma(c.rho, e[s], x[j] * c.sig);
                x[j] = e[s];
            }
        }
        double a = e[0];
#pragma unroll
        for (int s = 1; s < SUB; ++s) a = fma(c.rsl, a, e[s]);
        double in = team.scan_fwd(a, c.Bd, c.B32, c.blf, c.nscan);
        // solution with zero inflow at element 0; pick out its last element
        double ylast = 0.0;
        double ins[SUB];
#pragma unroll
        for (int s = 0; s < SUB; ++s) {
            ins[s] = in;
#pragma unroll
            for (int jj = 0; jj < SL; ++jj) {
                const int j = s * SL + jj;
                const double y = fma(c.pw[jj], in, x[j]);
                ylast = (j == jstar) ? y : ylast;
            }
            in = fma(c.rsl, in, e[s]);
        }
        const double Y = team.bcast(ylast, last / E) * c.dinv;  // y_{n-1} of the cyclic system
#pragma unroll
        for (int s = 0; s < SUB; ++s) {
            const double cf = fma(c.RH[s], Y, ins[s]);
#pragma unroll
            for (int jj = 0; jj < SL; ++jj) {
                const int j = s * SL + jj;
                x[j] = fma(c.pw[jj], cf, x[j]);
            }
        }
    }
};

// ---------------------------------------------------------------------------------------------
// Heat2D (backward Euler), heat/heat_2d.py:289-320, 360-366, in sine space.
//
// The reference solves (I + dt L) y = b with the 5-point Laplacian L on the interior nodes and identity rows on the
// boundary nodes (whose values are the Dirichlet data, whatever the input was).  L = fx Tx (x) I + fy I (x) Ty with
// Toeplitz tridiag(-1, 2, -1) factors, which the orthonormal sine transforms Sx, Sy diagonalise exactly.  Every other
// operation of MGRIT (sums, differences, injection, 2-norms) is linear or orthogonally invariant, so the whole
// hierarchy is kept in sine space: a level row holds  [Sx U_interior Sy | padding | boundary values | padding]  and
//     Phi(x)_e = (x_e + sum_k ct_k(i) rx_k,e) / (1 + dt_i sig_e)            interior coefficient e = (k, l)
//     Phi(x)_e = bc_e                                                        boundary node
// with sig_e = fx lambda_k + fy mu_l, rx = the transformed spatial factors of the right-hand side (plus the constant
// coupling of the interior to the boundary data), ct_k(i) = dt_i T_k(t_i).  Rows are cut into tiles of T*E elements;
// a tile is one "system" (no coupling between elements), tiles >= ip[0] hold boundary nodes and `sig` holds bc there.
// Transforms happen only where values enter or leave the solver (csrc/heat2d.cu).
// The theta method of heat_2d.py:322-366 (FE theta = 0, CN 1/2, BE 1) only changes the two factors:
//     Phi(x)_e = ((1 - (1-theta) dt_i sig_e) x_e + sum_k ct_k(i) rx_k,e) / (1 + theta dt_i sig_e),
// with ct_k(i) = dt_i (theta T_k(t_i) + (1-theta) T_k(t_{i-1})) made on the host.
//   step-constant row: [0] theta dt   [1] (1 - theta) dt
// ---------------------------------------------------------------------------------------------
constexpr int kHeat2DMaxTerms = 3;

template <int T_, int E_>
struct Heat2D {
    using SH = Shape<T_, E_>;
    static constexpr bool kTightChain = true;   // chain() below
    static constexpr bool kCtArg = false;
    static constexpr bool kItemInvariant = false;  // begin_item depends on the level only, not on the work item
    static constexpr int T = T_, E = E_;
    static constexpr int QMAX = kHeat2DMaxTerms;

    struct C {
        double dt;   // theta * dt: the implicit part, (1 + dt sig) y = ...
        double ex;   // (1 - theta) * dt: the explicit part, ... = (1 - ex sig) x + rhs   (0 for backward Euler)
    };
    struct Item {
        double sig[E];       // symbol of the spatial operator (boundary tiles: the Dirichlet values)
        double rx[QMAX][E];
        double inv[E];       // 1 / (1 + theta dt sig) for the dt in dtc ...
        double fac[E];       // ... and 1 - (1 - theta) dt sig (Crank-Nicolson / forward Euler)
        double dtc, exc;     // the step the two arrays were made for (NaN: none yet)
        bool boundary;
    };
    __device__ static __forceinline__ int row_n(const LevelDev &L) { return L.n; }

    __device__ static __forceinline__ void load_consts(C &c, const double *__restrict__ row, int) {
        c.dt = __ldg(row);
        c.ex = __ldg(row + 1);
    }

    template <class Pipe, class TeamT>
    __device__ static __forceinline__ void begin_item(Item &it, const LevelDev &L, int sys, Pipe &pipe, TeamT &team) {
        pipe.pop(it.sig, team);
#pragma unroll
        for (int k = 0; k < QMAX; ++k)
            if (k < L.nrhs) pipe.pop(it.rx[k], team);
        it.boundary = sys >= L.ip[0];
        it.dtc = it.exc = nan("");
    }
    // all levels share the symbol and the right-hand-side factors (checked by the C ABI)
    template <class TeamT>
    __device__ static __forceinline__ void retarget_item(Item &, const LevelDev &, const LevelDev &, TeamT &) {}

    __device__ static __forceinline__ void factors(Item &it, const C &c) {
        // The two per-element factors depend on the step size only: they are made once per (item, dt) -- one division per
        // element -- and every step of that size is then a multiplication (a division per element and step used to be
        // most of the arithmetic of a sweep).
        if (c.dt != it.dtc || c.ex != it.exc) {
#pragma unroll
            for (int j = 0; j < E; ++j) {
                it.inv[j] = __ddiv_rn(1.0, fma(c.dt, it.sig[j], 1.0));
                it.fac[j] = fma(-c.ex, it.sig[j], 1.0);
            }
            it.dtc = c.dt;
            it.exc = c.ex;
        }
    }

    // Steps i0 .. i1-1 without g / dense rows and nothing stored in between, backward Euler on a uniform level with at
    // most one right-hand-side term: as Heat1DSine::chain (one coalesced load of 32 time factors per warp, a step is a
    // shuffle, E FMAs and E multiplications).
    __device__ static __forceinline__ bool chain_ok(const C &c, const LevelDev &L) {
        return c.ex == 0.0 && c.dt != 0.0 && L.ndt == 1 && L.nrhs <= 1;
    }
    __device__ static __forceinline__ double chain_prefetch(const LevelDev &L, int i0, int i1, int lane) {
        return (L.nrhs == 1 && i0 + lane < i1) ? __ldg(L.rhs_t + i0 + lane) : 0.0;
    }
    template <class TeamT>
    __device__ static __forceinline__ void chain(double (&x)[E], const C &c, Item &it, const LevelDev &L, int i0, int i1,
                                                 TeamT &team, double first) {
        if (i1 <= i0) return;
        if (it.boundary) {
#pragma unroll
            for (int j = 0; j < E; ++j) x[j] = it.sig[j];
            return;
        }
        factors(it, c);
        if (L.nrhs == 0) {
            for (int i = i0; i < i1; ++i) {
#pragma unroll
                for (int j = 0; j < E; ++j) x[j] = x[j] * it.inv[j];
            }
            return;
        }
        double nxt = first;
        for (int base = i0; base < i1; base += 32) {
            const double cur = nxt;
            const int nb = base + 32 + team.lane;
            nxt = (nb < i1) ? __ldg(L.rhs_t + nb) : 0.0;
            const int cnt = min(32, i1 - base);
#pragma unroll 1
            for (int s = 0; s < cnt; ++s) {
                const double ct = __shfl_sync(0xffffffffu, cur, s);
#pragma unroll
                for (int j = 0; j < E; ++j) x[j] = fma(ct, it.rx[0][j], x[j]) * it.inv[j];
            }
        }
    }

    template <class TeamT>
    __device__ static __forceinline__ void apply(double (&x)[E], const C &c, Item &it, const LevelDev &L, int i,
                                                 TeamT &team) {
        if (it.boundary) {
#pragma unroll
            for (int j = 0; j < E; ++j) x[j] = it.sig[j];
            return;
        }
        factors(it, c);
        if (c.ex != 0.0) {  // Crank-Nicolson / forward Euler: (I - (1 - theta) dt L) u first (heat_2d.py:306, 352)
#pragma unroll
            for (int j = 0; j < E; ++j) x[j] = x[j] * it.fac[j];
        }
#pragma unroll
        for (int k = 0; k < QMAX; ++k) {
            if (k < L.nrhs) {
                const double ct = __ldg(L.rhs_t + (size_t)i * L.nrhs + k);
#pragma unroll
                for (int j = 0; j < E; ++j) x[j] = fma(ct, it.rx[k][j], x[j]);
            }
        }
        if (c.dt != 0.0) {
#pragma unroll
            for (int j = 0; j < E; ++j) x[j] = x[j] * it.inv[j];
        }
    }
};

// ---------------------------------------------------------------------------------------------
// Heat1DSine: Heat1D (heat/heat_1d.py:198-217) with the level rows kept in sine space.
//
// L = (a/dx^2) tridiag(-1, 2, -1) is diagonalised by the orthonormal sine matrix S (S = S^T = S^-1, spectral.cu):
// L = S diag(lam) S.  Every operation of MGRIT on level rows (sums, differences, injection, 2-norms, ghost copies) is
// linear or orthogonally invariant, so a hierarchy whose rows hold  x = u S  instead of u runs the same algorithm, and
//     Phi(x)_k = (x_k + sum_q ct_q(i) rxh_q[k]) / (1 + dt_i lam_k)
// is one FMA and one multiplication per unknown instead of the 7.6 FP64 instructions of the Toeplitz solve: the sweeps
// become HBM-bound.  It is the same direct solve of the same linear system as the reference's spsolve, with a different
// order of rounding errors (tests: <= 1e-12 relative against the tridiagonal kernels).  Values are transformed only
// where they enter or leave (initial condition, spatial right-hand-side factors, Mgrit.u[l][i].get_values()).
//   diag (per level): [0][j*T + tid] = lam_k, [1][j*T + tid] = 1/(1 + dt lam_k) for the level's dt, k = tid*E + j
//   step-constant row: [0] dt  [1] != 0: the level is uniform in time, use the reciprocals (else divide per step)
//   rhs_x: [nrhs][E][T] thread-transposed rxh_q = X_q S
// ---------------------------------------------------------------------------------------------
template <int T_, int E_>
struct Heat1DSine {
    using SH = Shape<T_, E_>;
    static constexpr int T = T_, E = E_;
    static constexpr bool kItemInvariant = true;

    struct C {
        double dt;
        bool recip;
    };
    struct Item {
        double d[E];        // 1/(1 + dt lam) (uniform level) or lam, of the level the item was begun for
        double rx0[E];      // first spatial right-hand-side factor
        const double *dsrc; // the table d was loaded from: a step of another level (the coarse step of the FAS restriction)
                            // reads its factors straight from that level's table instead of evicting these registers
    };
    __device__ static __forceinline__ int row_n(const LevelDev &L) { return L.n; }

    __device__ static __forceinline__ void load_consts(C &c, const double *__restrict__ row, int) {
        c.dt = __ldg(row);
        c.recip = __ldg(row + 1) != 0.0;
    }
    __device__ static __forceinline__ const double *dtable(const LevelDev &L) { return L.diag + (L.ndt == 1 ? T * E : 0); }
    template <class Pipe, class TeamT>
    __device__ static __forceinline__ void begin_item(Item &it, const LevelDev &L, int, Pipe &, TeamT &team) {
        it.dsrc = dtable(L);
        const double *__restrict__ d = it.dsrc + team.tid;
#pragma unroll
        for (int j = 0; j < E; ++j) it.d[j] = __ldg(d + j * T);
        if (L.nrhs > 0) {
            const double *__restrict__ rx = L.rhs_x + team.tid;
#pragma unroll
            for (int j = 0; j < E; ++j) it.rx0[j] = __ldg(rx + j * T);
        }
    }
    template <class TeamT>
    __device__ static __forceinline__ void retarget_item(Item &it, const LevelDev &from, const LevelDev &to, TeamT &team) {
        if (to.nrhs > 0 && to.rhs_x != from.rhs_x) {
            const double *__restrict__ rx = to.rhs_x + team.tid;
#pragma unroll
            for (int j = 0; j < E; ++j) it.rx0[j] = __ldg(rx + j * T);
        }
    }

    // Time factor of the first right-hand-side term for step i, loaded ahead of the step (the load is a dependent L2
    // access of several hundred cycles that the few resident warps cannot hide once Phi is this cheap).
    static constexpr bool kCtArg = true;
    __device__ static __forceinline__ double ct_load(const LevelDev &L, int i) {
        return L.nrhs > 0 ? __ldg(L.rhs_t + (size_t)i * L.nrhs) : 0.0;
    }

    // Steps i0 .. i1-1 of a level without g rows and without a dense right-hand side, nothing stored in between (the
    // F-relaxations of a down-sweep).  Uniform level, at most one right-hand-side term: the time factors of 32 steps come
    // with ONE coalesced load per warp (lane l holds step base + l; the next 32 are in flight while these are used) and a
    // step is a shuffle, E FMAs and E multiplications -- nothing else in the loop.  Same arithmetic as apply().
    // chain_prefetch() issues the first of those loads; call it before waiting for the item's rows.
    static constexpr bool kTightChain = true;
    __device__ static __forceinline__ bool chain_ok(const C &c, const LevelDev &L) {
        return c.recip && L.ndt == 1 && L.nrhs <= 1;
    }
    __device__ static __forceinline__ double chain_prefetch(const LevelDev &L, int i0, int i1, int lane) {
        return (L.nrhs == 1 && i0 + lane < i1) ? __ldg(L.rhs_t + i0 + lane) : 0.0;
    }
    template <class TeamT>
    __device__ static __forceinline__ void chain(double (&x)[E], const C &c, Item &it, const LevelDev &L, int i0, int i1,
                                                 TeamT &team, double first) {
        if (L.nrhs == 0) {
            for (int i = i0; i < i1; ++i) {
#pragma unroll
                for (int j = 0; j < E; ++j) x[j] = x[j] * it.d[j];
            }
            return;
        }
        const double *__restrict__ ctp = L.rhs_t;
        double nxt = first;
        for (int base = i0; base < i1; base += 32) {
            const double cur = nxt;
            const int nb = base + 32 + team.lane;
            nxt = (nb < i1) ? __ldg(ctp + nb) : 0.0;
            const int cnt = min(32, i1 - base);
#pragma unroll 1
            for (int s = 0; s < cnt; ++s) {
                const double ct = __shfl_sync(0xffffffffu, cur, s);
#pragma unroll
                for (int j = 0; j < E; ++j) x[j] = fma(ct, it.rx0[j], x[j]) * it.d[j];
            }
        }
    }

    template <class TeamT>
    __device__ static __forceinline__ void apply_ct(double (&x)[E], const C &c, Item &it, const LevelDev &L, int i,
                                                    TeamT &team, const double ct0) {
        if (L.nrhs > 0) {
#pragma unroll
            for (int j = 0; j < E; ++j) x[j] = fma(ct0, it.rx0[j], x[j]);
        }
        for (int k = 1; k < L.nrhs; ++k) {
            const double ct = __ldg(L.rhs_t + (size_t)i * L.nrhs + k);
            const double *__restrict__ rx = L.rhs_x + (size_t)k * E * T + team.tid;
#pragma unroll
            for (int j = 0; j < E; ++j) x[j] = fma(ct, __ldg(rx + j * T), x[j]);
        }
        const double *dt = dtable(L);
        if (dt == it.dsrc) {
            if (c.recip) {
#pragma unroll
                for (int j = 0; j < E; ++j) x[j] = x[j] * it.d[j];
            } else {
#pragma unroll
                for (int j = 0; j < E; ++j) x[j] = __ddiv_rn(x[j], fma(c.dt, it.d[j], 1.0));
            }
        } else {  // a step of another level than the item's: its factors straight from the table
            const double *__restrict__ d = dt + team.tid;
            if (c.recip) {
#pragma unroll
                for (int j = 0; j < E; ++j) x[j] = x[j] * __ldg(d + j * T);
            } else {
#pragma unroll
                for (int j = 0; j < E; ++j) x[j] = __ddiv_rn(x[j], fma(c.dt, __ldg(d + j * T), 1.0));
            }
        }
    }
    template <class TeamT>
    __device__ static __forceinline__ void apply(double (&x)[E], const C &c, Item &it, const LevelDev &L, int i,
                                                 TeamT &team) {
        apply_ct(x, c, it, L, i, team, ct_load(L, i));
    }
};

}  // namespace mgb
