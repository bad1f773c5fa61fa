// sweeps.cuh -- the batched MGRIT sweeps, templated on the Phi functor (phi.cuh).
// One CTA = one team; CTAs stride over the coarse intervals of the level (persistent grid sized to
// the SM count by the host).  Every kernel is written as: a generator that lists the input rows in
// pop order (run ahead by thread 0 to prefetch), and a body that pops rows, applies Phi out of
// registers and pushes result rows.
#pragma once
#include "phi.cuh"

namespace mgb {

extern __shared__ __align__(128) unsigned char g_smem[];

template <int E>
__device__ __forceinline__ void vadd(double (&x)[E], const double (&y)[E]) {  // x = y + x
#pragma unroll
    for (int j = 0; j < E; ++j) x[j] = y[j] + x[j];
}

// Registers -> a row slice in HBM with plain stores (used once per sweep, for the injection of point 0: not worth a slot)
template <int E>
__device__ __forceinline__ void store_direct(const double (&x)[E], double *dst, int tid, int tile) {
#pragma unroll
    for (int j = 0; j < E; ++j)
        if (tid * E + j < tile) dst[tid * E + j] = x[j];
}

// first/last point of the interval that starts at C-point k
__device__ __forceinline__ void interval_of(const LevelDev &L, int k, int &s, int &e) {
    if (L.cpts == nullptr) {  // whole level as one interval (forward solve)
        s = 0;
        e = L.npts;
    } else {
        s = __ldg(L.cpts + k);
        e = (k + 1 < L.ncpts) ? __ldg(L.cpts + k + 1) : L.npts;
    }
}

// Work items.  A row of a level holds `nsys` independent spatial systems of `tile` doubles each (1 for the 1-D
// applications; the sine-space tiles of Heat2D).  Work item w of a sweep = (interval or C-point kbase + w / nsys,
// system w % nsys); `soff` is the offset of the system inside a row.
struct ItemPos {
    int k, sys;
    size_t soff;
};
__device__ __forceinline__ ItemPos item_pos(const LevelDev &L, int w, int kbase = 0) {
    ItemPos ip;
    if (L.nsys <= 1) {
        ip.k = kbase + w;
        ip.sys = 0;
        ip.soff = 0;
    } else {
        const int q = w / L.nsys;
        ip.k = kbase + q;
        ip.sys = w - q * L.nsys;
        ip.soff = (size_t)ip.sys * L.tile;
    }
    return ip;
}

// Item prologue: per-element Phi data an item keeps in registers (Heat2D: the symbol tile, then the tiles of the
// separable right-hand-side factors).  They come through the row pipe like any other row, ahead of the item's rows.
__device__ __forceinline__ int n_prologue(const LevelDev &L) { return L.sig ? 1 + L.nrhs : 0; }
__device__ __forceinline__ const double *prologue_row(const LevelDev &L, int q, size_t soff) {
    return (q == 0 ? L.sig : L.rhs_x + (size_t)(q - 1) * L.pitch) + soff;
}

template <class Phi>
__device__ __forceinline__ void step_consts(typename Phi::C &c, const LevelDev &L, int i, int tid) {
    if (L.ndt > 1) Phi::load_consts(c, L.sconst + (size_t)__ldg(L.dtidx + i) * L.cw, tid);
}

// x <- (g_i +) Phi_i(x): pops the dense right-hand-side row and the g row if the level has them
template <class Phi, class Pipe, class TeamT>
__device__ __forceinline__ void advance(double (&x)[Phi::E], typename Phi::C &c, typename Phi::Item &it,
                                        const LevelDev &L, int i, Pipe &pipe, TeamT &team, bool add_g = true) {
    MGB_T0
    step_consts<Phi>(c, L, i, team.tid);
    if (L.rhs_dense) {
        double b[Phi::E];
        pipe.pop(b, team);
        vadd(x, b);
    }
    MGB_T(8)
    Phi::apply(x, c, it, L, i, team);
    MGB_T(9)
    if (add_g && L.g) {
        double gg[Phi::E];
        pipe.pop(gg, team);
        vadd(x, gg);
    }
    MGB_T(10)
}

// advance() with the step's first time factor already in a register (Phi::ct_load issued earlier); functors without
// that hook ignore it.
template <class Phi, class Pipe, class TeamT>
__device__ __forceinline__ void advance_ct(double (&x)[Phi::E], typename Phi::C &c, typename Phi::Item &it,
                                           const LevelDev &L, int i, Pipe &pipe, TeamT &team, bool add_g, double ct) {
    if constexpr (Phi::kCtArg) {
        step_consts<Phi>(c, L, i, team.tid);
        if (L.rhs_dense) {
            double b[Phi::E];
            pipe.pop(b, team);
            vadd(x, b);
        }
        Phi::apply_ct(x, c, it, L, i, team, ct);
        if (add_g && L.g) {
            double gg[Phi::E];
            pipe.pop(gg, team);
            vadd(x, gg);
        }
    } else {
        advance<Phi>(x, c, it, L, i, pipe, team, add_g);
    }
}
template <class Phi>
__device__ __forceinline__ double ct_ahead(const LevelDev &L, int i) {
    if constexpr (Phi::kCtArg) return Phi::ct_load(L, i);
    return 0.0;
}
template <class Phi>
__device__ __forceinline__ double chain_ahead(const LevelDev &L, int i0, int i1, int lane) {
    if constexpr (Phi::kTightChain) return Phi::chain_prefetch(L, i0, i1, lane);
    return 0.0;
}

// x <- Phi_{i1-1}( ... Phi_{i0}(x)) on a level without g rows and dense right-hand-side rows (nothing to pop, nothing
// stored in between): the functor's tight loop where it has one (Heat1DSine), else step by step.
template <class Phi, class Pipe, class TeamT>
__device__ __forceinline__ void run_chain(double (&x)[Phi::E], typename Phi::C &c, typename Phi::Item &it,
                                          const LevelDev &L, int i0, int i1, Pipe &pipe, TeamT &team, double first) {
    if constexpr (Phi::kTightChain) {
        if (Phi::chain_ok(c, L)) {
            Phi::chain(x, c, it, L, i0, i1, team, first);  // first = chain_ahead(L, i0, i1, lane)
            return;
        }
    }
    for (int i = i0; i < i1; ++i) advance<Phi>(x, c, it, L, i, pipe, team, false);
}
__device__ __forceinline__ bool plain_steps(const LevelDev &L) { return L.g == nullptr && L.rhs_dense == nullptr; }

// rows `advance` pops for step i, in order: returns false when stage runs past them
struct StepRows {
    // stage 0: dense rhs row, stage 1: g row
    __device__ static __forceinline__ bool next(const LevelDev &L, int i, int &stage, const double *&p, size_t soff,
                                                bool with_g = true) {
        if (stage == 0) {
            stage = 1;
            if (L.rhs_dense) {
                p = L.rhs_dense + (size_t)i * L.pitch + soff;
                return true;
            }
        }
        if (stage == 1) {
            stage = 2;
            if (with_g && L.g) {
                p = L.g + (size_t)i * L.pitch + soff;
                return true;
            }
        }
        return false;
    }
};

// ------------------------------------------------------------------------------------------------
// F-relaxation / forward solve (mgrit.py:312-327, 471-481):  for each interval [s, e):
//   x = u[s];  for i = s+1 .. e-1:  x = (g[i] +) Phi_i(x);  u[i] = x
// With last_only != 0 only the last point of each interval is stored: in the down-sweep of a cycle the other F-points
// are dead values (nothing reads them before the F-relaxation after the coarse-grid correction rewrites them), so the
// kernel computes the same chain and skips the dead stores.
// ------------------------------------------------------------------------------------------------
struct GenChain {
    LevelDev L;
    int w, nw, stride;
    int s, e, i, stage, pro;
    size_t soff;
    __device__ GenChain(const LevelDev &L_, int first, int nw_, int stride_)
        : L(L_), w(first), nw(nw_), stride(stride_), s(0), e(0), i(0), stage(-2), pro(0), soff(0) {}
    __device__ bool next(const double *&p) {
        for (;;) {
            if (w >= nw) return false;
            if (stage == -2) {
                const ItemPos ip = item_pos(L, w);
                interval_of(L, ip.k, s, e);
                if (e - s <= 1) {
                    w += stride;
                    continue;
                }
                soff = ip.soff;
                pro = 0;
                stage = -1;
            }
            if (stage == -1) {
                if (pro < n_prologue(L)) {
                    p = prologue_row(L, pro++, soff);
                    return true;
                }
                p = L.u + (size_t)s * L.pitch + soff;
                i = s + 1;
                stage = 0;
                return true;
            }
            if (i >= e) {
                w += stride;
                stage = -2;
                continue;
            }
            if (StepRows::next(L, i, stage, p, soff)) return true;
            ++i;
            stage = 0;
        }
    }
};

// (Measured: compiling the last-only variant for 12 instead of 8 warps per SM does not help -- the FP64 pipe is the
// limiter, and the smaller L1 and the spills cost what the extra warps gain; profiles/r01c_last_only_occupancy.txt.)
template <class Phi>
__global__ void __launch_bounds__(Phi::T) k_chain(const LevelDev L, const int nw, const int last_only, const int nin) {
    MGB_RETURN_IF_STOPPED(L)
    using SH = typename Phi::SH;
    Team<Phi::T> team(g_smem);
    RowPipe<SH, GenChain> pipe(g_smem, nin, L.tile, Phi::row_n(L), GenChain(L, blockIdx.x, nw, gridDim.x));
    pipe.start(team);
    typename Phi::C c;
    Phi::load_consts(c, L.sconst, team.tid);
    typename Phi::Item it;
    bool have_item = false;
    for (int w = blockIdx.x; w < nw; w += gridDim.x) {
        const ItemPos ip = item_pos(L, w);
        int s, e;
        interval_of(L, ip.k, s, e);
        if (e - s <= 1) continue;
        if (!Phi::kItemInvariant || !have_item) Phi::begin_item(it, L, ip.sys, pipe, team);
        have_item = true;
        const bool tight = last_only && plain_steps(L);
        const double first = tight ? chain_ahead<Phi>(L, s + 1, e, team.lane) : 0.0;
        double x[Phi::E];
        pipe.pop(x, team);
        if (tight) {
            run_chain<Phi>(x, c, it, L, s + 1, e, pipe, team, first);
            pipe.push(x, L.u + (size_t)(e - 1) * L.pitch + ip.soff, team);
            continue;
        }
        for (int i = s + 1; i < e; ++i) {
            advance<Phi>(x, c, it, L, i, pipe, team);
            if (!last_only || i == e - 1) pipe.push(x, L.u + (size_t)i * L.pitch + ip.soff, team);
        }
    }
    pipe.finish(team);
}

// ------------------------------------------------------------------------------------------------
// AT-MGRIT, local coarse grids on the coarsest level (core/at_mgrit.py:75-86, one time rank): every point p >= 1 is the
// end of its own chain of at most k-1 steps that starts from the PREVIOUS iterate,
//   x = old[max(0, p-k+1)];  for i = max(1, p-k+2) .. p:  x = (g[i] +) Phi_i(x);   u[p] = x.
// The chains are independent (they read `old`, a copy of u made before the launch), one work item per point.
// ------------------------------------------------------------------------------------------------
struct GenWindow {
    LevelDev L;
    const double *old;
    int k, w, nw, stride;
    int s, e, i, stage, pro;
    size_t soff;
    __device__ GenWindow(const LevelDev &L_, const double *old_, int k_, int first, int nw_, int stride_)
        : L(L_), old(old_), k(k_), w(first), nw(nw_), stride(stride_), s(0), e(0), i(0), stage(-2), pro(0), soff(0) {}
    __device__ bool next(const double *&p) {
        for (;;) {
            if (w >= nw) return false;
            if (stage == -2) {
                const ItemPos ip = item_pos(L, w, 1);
                e = ip.k + 1;  // one past the point this item produces
                s = ip.k - k + 1;
                if (s < 0) s = 0;
                soff = ip.soff;
                pro = 0;
                stage = -1;
            }
            if (stage == -1) {
                if (pro < n_prologue(L)) {
                    p = prologue_row(L, pro++, soff);
                    return true;
                }
                p = old + (size_t)s * L.pitch + soff;
                i = s + 1;
                stage = 0;
                return true;
            }
            if (i >= e) {
                w += stride;
                stage = -2;
                continue;
            }
            if (StepRows::next(L, i, stage, p, soff)) return true;
            ++i;
            stage = 0;
        }
    }
};

template <class Phi>
__global__ void __launch_bounds__(Phi::T) k_window(const LevelDev L, const double *__restrict__ old, const int k,
                                                   const int nw, const int nin) {
    MGB_RETURN_IF_STOPPED(L)
    using SH = typename Phi::SH;
    Team<Phi::T> team(g_smem);
    RowPipe<SH, GenWindow> pipe(g_smem, nin, L.tile, Phi::row_n(L), GenWindow(L, old, k, blockIdx.x, nw, gridDim.x));
    pipe.start(team);
    typename Phi::C c;
    Phi::load_consts(c, L.sconst, team.tid);
    for (int w = blockIdx.x; w < nw; w += gridDim.x) {
        const ItemPos ip = item_pos(L, w, 1);
        const int pt = ip.k;
        int s = pt - k + 1;
        if (s < 0) s = 0;
        typename Phi::Item it;
        Phi::begin_item(it, L, ip.sys, pipe, team);
        double x[Phi::E];
        pipe.pop(x, team);
        for (int i = s + 1; i <= pt; ++i) advance<Phi>(x, c, it, L, i, pipe, team);
        pipe.push(x, L.u + (size_t)pt * L.pitch + ip.soff, team);
    }
    pipe.finish(team);
}

// ------------------------------------------------------------------------------------------------
// C-relaxation (mgrit.py:354-368): for each C-point c = cpts[k], k >= 1:
//   u[c] = ((g[c] +) Phi_c(u[c-1])) * w + u[c] * (1 - w)
// ------------------------------------------------------------------------------------------------
// With non-uniform coarsening two C-points can be adjacent; the reference visits C-points in ascending order, so
// the second one sees the first one's new value.  A run of adjacent C-points is therefore one work item, owned by
// the team of its first point ("leader") and walked sequentially out of registers.
__device__ __forceinline__ bool c_run_leader(const LevelDev &L, int k) {
    return k == 1 || __ldg(L.cpts + k) - __ldg(L.cpts + k - 1) > 1;
}
__device__ __forceinline__ bool c_run_continues(const LevelDev &L, int kk) {
    return kk + 1 < L.ncpts && __ldg(L.cpts + kk + 1) == __ldg(L.cpts + kk) + 1;
}

struct GenCRelax {
    LevelDev L;
    int w, nw, stride, stage, kk, pro;
    size_t soff;
    bool weighted;
    int kbase;
    __device__ GenCRelax(const LevelDev &L_, int first, int nw_, int stride_, bool weighted_, int kbase_)
        : L(L_), w(first), nw(nw_), stride(stride_), stage(-2), kk(0), pro(0), soff(0), weighted(weighted_), kbase(kbase_) {}
    __device__ bool next(const double *&p) {
        for (;;) {
            if (w >= nw) return false;
            if (stage == -2) {
                const ItemPos ip = item_pos(L, w, kbase);
                if (!c_run_leader(L, ip.k) && ip.k != kbase) {
                    w += stride;
                    continue;
                }
                kk = ip.k;
                soff = ip.soff;
                pro = 0;
                stage = -1;
            }
            if (stage == -1) {
                if (pro < n_prologue(L)) {
                    p = prologue_row(L, pro++, soff);
                    return true;
                }
                p = L.u + (size_t)(__ldg(L.cpts + kk) - 1) * L.pitch + soff;
                stage = 0;
                return true;
            }
            const int c = __ldg(L.cpts + kk);
            if (stage < 2 && StepRows::next(L, c, stage, p, soff)) return true;
            if (stage == 2) {
                stage = 3;
                if (weighted) {
                    p = L.u + (size_t)c * L.pitch + soff;
                    return true;
                }
            }
            if (c_run_continues(L, kk)) {
                ++kk;
                stage = 0;
                continue;
            }
            w += stride;
            stage = -2;
        }
    }
};

// kbase: first C-point relaxed (1 for the whole level; ncpts-1 to relax only the last C-point, which a time rank does
// ahead of the fused down-sweep so that the next rank's ghost can travel first).  A run that starts before kbase is cut
// there: the caller guarantees that point cpts[kbase]-1 is current.
template <class Phi>
__global__ void __launch_bounds__(Phi::T) k_c_relax(const LevelDev L, const double wgt, const int nw, const int nin,
                                                    const int kbase) {
    MGB_RETURN_IF_STOPPED(L)
    using SH = typename Phi::SH;
    const bool weighted = (wgt != 1.0);
    Team<Phi::T> team(g_smem);
    RowPipe<SH, GenCRelax> pipe(g_smem, nin, L.tile, Phi::row_n(L), GenCRelax(L, blockIdx.x, nw, gridDim.x, weighted, kbase));
    pipe.start(team);
    typename Phi::C c;
    Phi::load_consts(c, L.sconst, team.tid);
    for (int w = blockIdx.x; w < nw; w += gridDim.x) {
        const ItemPos ip = item_pos(L, w, kbase);
        if (!c_run_leader(L, ip.k) && ip.k != kbase) continue;
        typename Phi::Item it;
        Phi::begin_item(it, L, ip.sys, pipe, team);
        double x[Phi::E];
        pipe.pop(x, team);
        for (int kk = ip.k;; ++kk) {
            const int cp = __ldg(L.cpts + kk);
            advance<Phi>(x, c, it, L, cp, pipe, team);
            if (weighted) {
                double old[Phi::E];
                pipe.pop(old, team);
                const double w1 = 1.0 - wgt;
#pragma unroll
                for (int j = 0; j < Phi::E; ++j) x[j] = __dadd_rn(__dmul_rn(x[j], wgt), __dmul_rn(old[j], w1));
            }
            pipe.push(x, L.u + (size_t)cp * L.pitch + ip.soff, team);
            if (!c_run_continues(L, kk)) break;
        }
    }
    pipe.finish(team);
}

// ------------------------------------------------------------------------------------------------
// FAS restriction (mgrit.py:497-547, identity transfer): for each C-point j >= 1, c = cpts[j]:
//   r   = Phi_f(u[c-1]) - u[c]                 (level 0)      or  (g[c] - u[c]) + Phi_f(u[c-1])   (level > 0)
//   G.u[j] = u[c]
//   G.g[j] = (r + u[c]) - Phi_c(u[cpts[j-1]])
// ------------------------------------------------------------------------------------------------
struct GenFas {
    LevelDev L, G;
    int w, nw, stride, stage, sub, j, pro;
    size_t soff;
    __device__ GenFas(const LevelDev &L_, const LevelDev &G_, int first, int nw_, int stride_)
        : L(L_), G(G_), w(first), nw(nw_), stride(stride_), stage(-2), sub(0), j(0), pro(0), soff(0) {}
    __device__ bool next(const double *&p) {
        for (;;) {
            if (w >= nw) return false;
            if (stage == -2) {
                const ItemPos ip = item_pos(L, w, 1);
                j = ip.k;
                soff = ip.soff;
                pro = 0;
                stage = -1;
            }
            const int c = __ldg(L.cpts + j);
            switch (stage) {
                case -1:
                    if (pro < n_prologue(L)) {
                        p = prologue_row(L, pro++, soff);
                        return true;
                    }
                    stage = 0;
                    break;
                case 0:  // u[c-1]
                    p = L.u + (size_t)(c - 1) * L.pitch + soff;
                    stage = 1;
                    sub = 0;
                    return true;
                case 1:  // dense rhs row of the fine step (no g: it is combined by hand below)
                    stage = 2;
                    if (StepRows::next(L, c, sub, p, soff, false)) return true;
                    break;
                case 2:  // u[c]
                    p = L.u + (size_t)c * L.pitch + soff;
                    stage = 3;
                    return true;
                case 3:  // g[c]
                    stage = 4;
                    if (L.g) {
                        p = L.g + (size_t)c * L.pitch + soff;
                        return true;
                    }
                    break;
                case 4:  // v[j-1] = u[cpts[j-1]]
                    p = L.u + (size_t)__ldg(L.cpts + j - 1) * L.pitch + soff;
                    stage = 5;
                    sub = 0;
                    return true;
                case 5:  // dense rhs row of the coarse step
                    stage = 6;
                    if (StepRows::next(G, j, sub, p, soff, false)) return true;
                    break;
                default:
                    w += stride;
                    stage = -2;
            }
        }
    }
};

template <class Phi>
__global__ void __launch_bounds__(Phi::T) k_fas_residual(const LevelDev L, const LevelDev G, const int nw, const int nin) {
    MGB_RETURN_IF_STOPPED(L)
    using SH = typename Phi::SH;
    constexpr int E = Phi::E;
    Team<Phi::T> team(g_smem);
    RowPipe<SH, GenFas> pipe(g_smem, nin, L.tile, Phi::row_n(L), GenFas(L, G, blockIdx.x, nw, gridDim.x));
    pipe.start(team);
    typename Phi::C c;  // reloaded before each Phi: fine and coarse steps use different constants
    for (int w = blockIdx.x; w < nw; w += gridDim.x) {
        const ItemPos ip = item_pos(L, w, 1);
        const int j = ip.k;
        const int cp = __ldg(L.cpts + j);
        typename Phi::Item it;  // per-element data: the same spatial operator on both levels
        Phi::begin_item(it, L, ip.sys, pipe, team);
        double x[E], y[E];
        pipe.pop(x, team);
        Phi::load_consts(c, L.sconst + (L.ndt > 1 ? (size_t)__ldg(L.dtidx + cp) * L.cw : 0), team.tid);
        advance<Phi>(x, c, it, L, cp, pipe, team, false);            // Phi_f(u[c-1])
        pipe.pop(y, team);                                           // u[c]
        pipe.push(y, G.u + (size_t)j * G.pitch + ip.soff, team);     // injection
        if (L.g) {
            double gg[E];
            pipe.pop(gg, team);
#pragma unroll
            for (int q = 0; q < E; ++q) x[q] = ((gg[q] - y[q]) + x[q]) + y[q];
        } else {
#pragma unroll
            for (int q = 0; q < E; ++q) x[q] = (x[q] - y[q]) + y[q];
        }
        pipe.pop(y, team);                                           // v[j-1]
        if (j == 1) store_direct(y, G.u + ip.soff, team.tid, L.tile);  // point 0 (initial condition / ghost) is injected too
        Phi::retarget_item(it, L, G, team);
        Phi::load_consts(c, G.sconst + (G.ndt > 1 ? (size_t)__ldg(G.dtidx + j) * G.cw : 0), team.tid);
        advance<Phi>(y, c, it, G, j, pipe, team, false);             // Phi_c(v[j-1])
#pragma unroll
        for (int q = 0; q < E; ++q) x[q] = x[q] - y[q];
        pipe.push(x, G.g + (size_t)j * G.pitch + ip.soff, team);
    }
    pipe.finish(team);
}

// ------------------------------------------------------------------------------------------------
// FAS restriction with a spatial grid transfer R (mgrit.py:488-549, transfer != GridTransferCopy).  The fine and the
// coarse level have different spatial sizes, so the sweep is split around the transfer:
//   k_residual_rows (fine shape):    out[j] = Phi_f(u[c_j-1]) - u[c_j]               (level 0)
//                                    out[j] = (g[c_j] - u[c_j]) + Phi_f(u[c_j-1])    (level > 0),  j >= 1
//   row-wise transfer kernels:       G.u[j] = R(u[c_j]), V = copy of G.u, RR[j] = R(out[j])        (csrc/transfer.cu)
//   k_fas_rhs (coarse shape):        G.g[j] = (RR[j] + V[j]) - Phi_c(V[j-1]),  j >= 1
// ------------------------------------------------------------------------------------------------
struct GenResRows {
    LevelDev L;
    int w, nw, stride, stage, sub, j, pro;
    size_t soff;
    __device__ GenResRows(const LevelDev &L_, int first, int nw_, int stride_)
        : L(L_), w(first), nw(nw_), stride(stride_), stage(-2), sub(0), j(0), pro(0), soff(0) {}
    __device__ bool next(const double *&p) {
        for (;;) {
            if (w >= nw) return false;
            if (stage == -2) {
                const ItemPos ip = item_pos(L, w, 1);
                j = ip.k;
                soff = ip.soff;
                pro = 0;
                stage = -1;
            }
            const int c = __ldg(L.cpts + j);
            switch (stage) {
                case -1:
                    if (pro < n_prologue(L)) {
                        p = prologue_row(L, pro++, soff);
                        return true;
                    }
                    stage = 0;
                    break;
                case 0:  // u[c-1]
                    p = L.u + (size_t)(c - 1) * L.pitch + soff;
                    stage = 1;
                    sub = 0;
                    return true;
                case 1:  // dense rhs row of the fine step
                    stage = 2;
                    if (StepRows::next(L, c, sub, p, soff, false)) return true;
                    break;
                case 2:  // u[c]
                    p = L.u + (size_t)c * L.pitch + soff;
                    stage = 3;
                    return true;
                case 3:  // g[c]
                    stage = 4;
                    if (L.g) {
                        p = L.g + (size_t)c * L.pitch + soff;
                        return true;
                    }
                    break;
                default:
                    w += stride;
                    stage = -2;
            }
        }
    }
};

template <class Phi>
__global__ void __launch_bounds__(Phi::T) k_residual_rows(const LevelDev L, double *__restrict__ out, const int nw,
                                                          const int nin) {
    MGB_RETURN_IF_STOPPED(L)
    using SH = typename Phi::SH;
    constexpr int E = Phi::E;
    Team<Phi::T> team(g_smem);
    RowPipe<SH, GenResRows> pipe(g_smem, nin, L.tile, Phi::row_n(L), GenResRows(L, blockIdx.x, nw, gridDim.x));
    pipe.start(team);
    typename Phi::C c;
    Phi::load_consts(c, L.sconst, team.tid);
    for (int w = blockIdx.x; w < nw; w += gridDim.x) {
        const ItemPos ip = item_pos(L, w, 1);
        const int j = ip.k;
        const int cp = __ldg(L.cpts + j);
        typename Phi::Item it;
        Phi::begin_item(it, L, ip.sys, pipe, team);
        double x[E], y[E];
        pipe.pop(x, team);
        advance<Phi>(x, c, it, L, cp, pipe, team, false);  // Phi_f(u[c-1])
        pipe.pop(y, team);                                 // u[c]
        if (L.g) {
            double gg[E];
            pipe.pop(gg, team);
#pragma unroll
            for (int q = 0; q < E; ++q) x[q] = (gg[q] - y[q]) + x[q];
        } else {
#pragma unroll
            for (int q = 0; q < E; ++q) x[q] = x[q] - y[q];
        }
        pipe.push(x, out + (size_t)j * L.pitch + ip.soff, team);
    }
    pipe.finish(team);
}

struct GenFasRhs {
    LevelDev G;
    const double *V, *RR;
    int w, nw, stride, stage, sub, j, pro;
    size_t soff;
    __device__ GenFasRhs(const LevelDev &G_, const double *V_, const double *RR_, int first, int nw_, int stride_)
        : G(G_), V(V_), RR(RR_), w(first), nw(nw_), stride(stride_), stage(-2), sub(0), j(0), pro(0), soff(0) {}
    __device__ bool next(const double *&p) {
        for (;;) {
            if (w >= nw) return false;
            if (stage == -2) {
                const ItemPos ip = item_pos(G, w, 1);
                j = ip.k;
                soff = ip.soff;
                pro = 0;
                stage = -1;
            }
            switch (stage) {
                case -1:
                    if (pro < n_prologue(G)) {
                        p = prologue_row(G, pro++, soff);
                        return true;
                    }
                    stage = 0;
                    break;
                case 0:  // V[j-1]
                    p = V + (size_t)(j - 1) * G.pitch + soff;
                    stage = 1;
                    sub = 0;
                    return true;
                case 1:  // dense rhs row of the coarse step
                    stage = 2;
                    if (StepRows::next(G, j, sub, p, soff, false)) return true;
                    break;
                case 2:  // RR[j]
                    p = RR + (size_t)j * G.pitch + soff;
                    stage = 3;
                    return true;
                case 3:  // V[j]
                    p = V + (size_t)j * G.pitch + soff;
                    stage = 4;
                    return true;
                default:
                    w += stride;
                    stage = -2;
            }
        }
    }
};

template <class Phi>
__global__ void __launch_bounds__(Phi::T) k_fas_rhs(const LevelDev G, const double *__restrict__ V,
                                                    const double *__restrict__ RR, const int nw, const int nin) {
    MGB_RETURN_IF_STOPPED(G)
    using SH = typename Phi::SH;
    constexpr int E = Phi::E;
    Team<Phi::T> team(g_smem);
    RowPipe<SH, GenFasRhs> pipe(g_smem, nin, G.tile, Phi::row_n(G), GenFasRhs(G, V, RR, blockIdx.x, nw, gridDim.x));
    pipe.start(team);
    typename Phi::C c;
    Phi::load_consts(c, G.sconst, team.tid);
    for (int w = blockIdx.x; w < nw; w += gridDim.x) {
        const ItemPos ip = item_pos(G, w, 1);
        const int j = ip.k;
        typename Phi::Item it;
        Phi::begin_item(it, G, ip.sys, pipe, team);
        double x[E], y[E], z[E];
        pipe.pop(x, team);
        advance<Phi>(x, c, it, G, j, pipe, team, false);  // Phi_c(V[j-1])
        pipe.pop(y, team);                                // R(fine residual)
        pipe.pop(z, team);                                // V[j]
#pragma unroll
        for (int q = 0; q < E; ++q) x[q] = (y[q] + z[q]) - x[q];
        pipe.push(x, G.g + (size_t)j * G.pitch + ip.soff, team);
    }
    pipe.finish(team);
}

// ------------------------------------------------------------------------------------------------
// Down-sweep of a cycle in one pass (mgrit.py:277-281 + 497-547 for cf_iter's last round, weight 1): C-relaxation,
// the F-relaxation that follows it and the FAS restriction.  Work item j = 1 .. ncpts-1, a = cpts[j-1], c = cpts[j]:
//     yc   = (g[c] +) Phi_f(u[c-1])                  C-relaxation of c (from the F-point as it is before this sweep)
//     u[c] = yc,  G.u[j] = yc                        stored; injection
//     x    = a == 0 ? u[0] : (g[a] +) Phi_f(u[a-1])  the C-relaxed left end, recomputed (its owner is item j-1)
//     w    = Phi_c(x)                                coarse step from v[j-1] = new u[a]
//     x    = Phi_f(F-relaxation chain from x to c-1) the new last F-point, never stored: nothing reads it before the
//                                                    F-relaxation after the correction rewrites the interval
//     G.g[j] = ((x - yc) + yc) - w                   (level 0)    or  (((g[c] - yc) + x) + yc) - w   (level > 0)
// Same arithmetic, operation by operation, as k_c_relax + k_chain + k_fas_residual (bit-identical results), but per
// interval it reads 2 rows (1 distinct) and writes 3 instead of 5 + 4, and the level is swept once instead of 3 times.
// C-point rows are only written and F-point rows only read here, so items do not depend on each other.
// Requires every interval to hold at least one F-point (no adjacent C-points) and weight 1; the host checks.
// yc waits in the output slot and w in the stash slot, so that only one row (plus Phi's item data) lives in registers.
// (Measured: routing the five Phi applications through one call site -- a loop over the item's steps -- to shrink the
// code is slower, 1.79 ms against 1.41 ms on level 0 of cfg5: the run-time choice of level and constants costs more
// than the instruction-cache misses of the straight-line version.)
// ------------------------------------------------------------------------------------------------
struct GenDown {
    LevelDev L, G;
    int w, nw, stride, stage, sub, j, a, c, i, pro;
    size_t soff;
    __device__ GenDown(const LevelDev &L_, const LevelDev &G_, int first, int nw_, int stride_)
        : L(L_), G(G_), w(first), nw(nw_), stride(stride_), stage(-2), sub(0), j(0), a(0), c(0), i(0), pro(0), soff(0) {}
    __device__ bool next(const double *&p) {
        for (;;) {
            if (w >= nw) return false;
            switch (stage) {
                case -2: {
                    const ItemPos ip = item_pos(L, w, 1);
                    j = ip.k;
                    soff = ip.soff;
                    a = __ldg(L.cpts + j - 1);
                    c = __ldg(L.cpts + j);
                    pro = 0;
                    stage = -1;
                    break;
                }
                case -1:
                    if (pro < n_prologue(L)) {
                        p = prologue_row(L, pro++, soff);
                        return true;
                    }
                    stage = 0;
                    break;
                case 0:  // u[c-1] for the C-relaxation of c
                    p = L.u + (size_t)(c - 1) * L.pitch + soff;
                    stage = 1;
                    sub = 0;
                    return true;
                case 1:  // its dense rhs row and g[c]
                    if (StepRows::next(L, c, sub, p, soff)) return true;
                    stage = 2;
                    break;
                case 2:  // left end: u[0], or u[a-1] and the rows of step a
                    stage = 3;
                    sub = (a == 0) ? 2 : 0;
                    p = L.u + (size_t)(a == 0 ? 0 : a - 1) * L.pitch + soff;
                    return true;
                case 3:
                    if (StepRows::next(L, a, sub, p, soff)) return true;
                    stage = 4;
                    sub = 0;
                    break;
                case 4: {  // dense rhs row of the coarse step j
                    const bool got = StepRows::next(G, j, sub, p, soff, false);
                    stage = 5;
                    i = a + 1;
                    sub = 0;
                    if (got) return true;
                    break;
                }
                case 5:  // F-relaxation chain a+1 .. c-1
                    if (i >= c) {
                        stage = 6;
                        sub = 0;
                        break;
                    }
                    if (StepRows::next(L, i, sub, p, soff)) return true;
                    ++i;
                    sub = 0;
                    break;
                case 6:  // fine step into c: dense rhs row only
                    stage = 7;
                    if (StepRows::next(L, c, sub, p, soff, false)) return true;
                    break;
                case 7:  // g[c] for the FAS right-hand side
                    stage = 8;
                    if (L.g) {
                        p = L.g + (size_t)c * L.pitch + soff;
                        return true;
                    }
                    break;
                default:
                    w += stride;
                    stage = -2;
            }
        }
    }
};

template <class Phi>
__global__ void __launch_bounds__(Phi::T) k_down(const LevelDev L, const LevelDev G, const int nw, const int nin) {
    MGB_RETURN_IF_STOPPED(L)
    using SH = typename Phi::SH;
    constexpr int E = Phi::E;
    Team<Phi::T> team(g_smem);
    RowPipe<SH, GenDown> pipe(g_smem, nin, L.tile, Phi::row_n(L), GenDown(L, G, blockIdx.x, nw, gridDim.x));
    pipe.start(team);
    typename Phi::C cf, cc;  // fine and coarse step constants (reloaded per step on non-uniform grids)
    Phi::load_consts(cf, L.sconst, team.tid);
    Phi::load_consts(cc, G.sconst, team.tid);
    double *const stash = pipe.stash_slot() + team.tid * E;
    const double *const outs = pipe.out_slot() + team.tid * E;
    typename Phi::Item it;
    bool have_item = false;
    for (int w = blockIdx.x; w < nw; w += gridDim.x) {
        const ItemPos ip = item_pos(L, w, 1);
        const int j = ip.k;
        const int a = __ldg(L.cpts + j - 1), cp = __ldg(L.cpts + j);
        if (!Phi::kItemInvariant || !have_item) Phi::begin_item(it, L, ip.sys, pipe, team);
        have_item = true;
        // the time factors of the item's single steps and of the start of its chain: all in flight before the first row
        // is waited for
        const bool tight = plain_steps(L);
        const double ct_c = ct_ahead<Phi>(L, cp), ct_a = a != 0 ? ct_ahead<Phi>(L, a) : 0.0, ct_j = ct_ahead<Phi>(G, j);
        const double first = tight ? chain_ahead<Phi>(L, a + 1, cp + 1, team.lane) : 0.0;
        double x[E];
        // C-relaxation of c; the output slot keeps yc until the last push of this item
        pipe.pop(x, team);
        advance_ct<Phi>(x, cf, it, L, cp, pipe, team, true, ct_c);
        pipe.push(x, L.u + (size_t)cp * L.pitch + ip.soff, team);
        pipe.push_again(G.u + (size_t)j * G.pitch + ip.soff);
        // the C-relaxed left end
        pipe.pop(x, team);
        if (a != 0) advance_ct<Phi>(x, cf, it, L, a, pipe, team, true, ct_a);
        if (j == 1) store_direct(x, G.u + ip.soff, team.tid, L.tile);  // point 0 (initial condition / ghost) is injected too
        // w = Phi_c(x) into the stash, x back
#pragma unroll
        for (int q = 0; q < E; ++q) stash[q] = x[q];
        Phi::retarget_item(it, L, G, team);
        advance_ct<Phi>(x, cc, it, G, j, pipe, team, false, ct_j);
        Phi::retarget_item(it, G, L, team);
#pragma unroll
        for (int q = 0; q < E; ++q) {
            const double t = stash[q];
            stash[q] = x[q];
            x[q] = t;
        }
        // F-relaxation chain and the fine step into c
        if (tight) {
            run_chain<Phi>(x, cf, it, L, a + 1, cp + 1, pipe, team, first);
        } else {
            for (int i = a + 1; i < cp; ++i) advance<Phi>(x, cf, it, L, i, pipe, team);
            advance<Phi>(x, cf, it, L, cp, pipe, team, false);
        }
        // FAS right-hand side
        if (L.g) {
            const double *gg = pipe.pop_ptr() + team.tid * E;
            const int nv = Phi::row_n(L) - team.tid * E;
#pragma unroll
            for (int q = 0; q < E; ++q) {
                const double r = (((gg[q] - outs[q]) + x[q]) + outs[q]) - stash[q];
                x[q] = (q < nv) ? r : 0.0;  // the slot holds the raw row: keep the padding at zero
            }
            pipe.pop_release(team);
        } else {
#pragma unroll
            for (int q = 0; q < E; ++q) x[q] = ((x[q] - outs[q]) + outs[q]) - stash[q];
        }
        pipe.push(x, G.g + (size_t)j * G.pitch + ip.soff, team);
    }
    pipe.finish(team);
}

// ------------------------------------------------------------------------------------------------
// Coarse-grid correction (mgrit.py:722-726) fused with the F-relaxation that follows (mgrit.py:287):
//   for interval k:  if k >= kfirst: u[c] = u[c] + (G.u[k] - u[c]);   then, if f_relax, the chain from u[c]
//   (frelax == 2: only the last F-point of the interval is stored, see mgb_error_correction).
// kfirst = 1 on time rank 0 (point 0 is the initial condition).  On the other ranks point 0 is the ghost copy of the
// previous rank's last C-point and kfirst = 0: the ghost is corrected here with the same arithmetic its owner uses
// (the coarse ghost row is already current), so the F-relaxation of the first interval starts from the corrected
// value without waiting for an exchange.
// ------------------------------------------------------------------------------------------------
struct GenCorrect {
    LevelDev L, G;
    int w, nw, stride;
    int k, s, e, i, stage, pro;
    size_t soff;
    bool frelax;
    int kfirst;
    __device__ GenCorrect(const LevelDev &L_, const LevelDev &G_, int first, int nw_, int stride_, bool fr, int kf)
        : L(L_), G(G_), w(first), nw(nw_), stride(stride_), k(0), s(0), e(0), i(0), stage(-3), pro(0), soff(0),
          frelax(fr), kfirst(kf) {}
    __device__ bool next(const double *&p) {
        for (;;) {
            if (w >= nw) return false;
            if (stage == -3) {
                const ItemPos ip = item_pos(L, w);
                k = ip.k;
                interval_of(L, k, s, e);
                const bool chain = frelax && (e - s > 1);
                if (k < kfirst && !chain) {
                    w += stride;
                    continue;
                }
                soff = ip.soff;
                pro = 0;
                stage = -2;
            }
            if (stage == -2) {
                const bool chain = frelax && (e - s > 1);
                if (chain && pro < n_prologue(L)) {
                    p = prologue_row(L, pro++, soff);
                    return true;
                }
                p = L.u + (size_t)s * L.pitch + soff;
                stage = -1;
                return true;
            }
            if (stage == -1) {
                stage = 0;
                i = s + 1;
                if (k >= kfirst) {
                    p = G.u + (size_t)k * G.pitch + soff;
                    return true;
                }
            }
            if (!frelax || i >= e) {
                w += stride;
                stage = -3;
                continue;
            }
            if (StepRows::next(L, i, stage, p, soff)) return true;
            ++i;
            stage = 0;
        }
    }
};

template <class Phi>
__global__ void __launch_bounds__(Phi::T) k_correct(const LevelDev L, const LevelDev G, const int frelax, const int kfirst,
                                                    const int nw, const int nin) {
    MGB_RETURN_IF_STOPPED(L)
    using SH = typename Phi::SH;
    constexpr int E = Phi::E;
    Team<Phi::T> team(g_smem);
    RowPipe<SH, GenCorrect> pipe(g_smem, nin, L.tile, Phi::row_n(L), GenCorrect(L, G, blockIdx.x, nw, gridDim.x, frelax != 0, kfirst));
    pipe.start(team);
    typename Phi::C c;
    Phi::load_consts(c, L.sconst, team.tid);
    typename Phi::Item it;
    bool have_item = false;
    for (int w = blockIdx.x; w < nw; w += gridDim.x) {
        const ItemPos ip = item_pos(L, w);
        const int k = ip.k;
        int s, e;
        interval_of(L, k, s, e);
        const bool chain = frelax && (e - s > 1);
        if (k < kfirst && !chain) continue;
        if (chain && (!Phi::kItemInvariant || !have_item)) {
            Phi::begin_item(it, L, ip.sys, pipe, team);
            have_item = true;
        }
        const bool tight = chain && frelax == 2 && plain_steps(L);
        const double first = tight ? chain_ahead<Phi>(L, s + 1, e, team.lane) : 0.0;
        double x[E];
        pipe.pop(x, team);
        if (k >= kfirst) {
            double cu[E];
            pipe.pop(cu, team);
#pragma unroll
            for (int q = 0; q < E; ++q) x[q] = x[q] + (cu[q] - x[q]);
            pipe.push(x, L.u + (size_t)s * L.pitch + ip.soff, team);
        }
        if (tight) {
            run_chain<Phi>(x, c, it, L, s + 1, e, pipe, team, first);
            pipe.push(x, L.u + (size_t)(e - 1) * L.pitch + ip.soff, team);
        } else if (chain) {
            for (int i = s + 1; i < e; ++i) {
                advance<Phi>(x, c, it, L, i, pipe, team);
                if (frelax == 1 || i == e - 1) pipe.push(x, L.u + (size_t)i * L.pitch + ip.soff, team);
            }
        }
    }
    pipe.finish(team);
}

// ------------------------------------------------------------------------------------------------
// Residual at C-points (mgrit.py:405-413): out[k] = || Phi(u[c-1]) - u[c] ||^2  (level 0: no g).
// With several systems per row each item writes its partial sum to out_sq[ncpts + k * nsys + sys] and
// k_sum_systems adds them up in a fixed order.
// ------------------------------------------------------------------------------------------------
struct GenResidual {
    LevelDev L;
    int w, nw, stride, stage, k, pro;
    size_t soff;
    __device__ GenResidual(const LevelDev &L_, int first, int nw_, int stride_)
        : L(L_), w(first), nw(nw_), stride(stride_), stage(-2), k(0), pro(0), soff(0) {}
    __device__ bool next(const double *&p) {
        for (;;) {
            if (w >= nw) return false;
            if (stage == -2) {
                const ItemPos ip = item_pos(L, w, 1);
                k = ip.k;
                soff = ip.soff;
                pro = 0;
                stage = -1;
            }
            const int c = __ldg(L.cpts + k);
            if (stage == -1) {
                if (pro < n_prologue(L)) {
                    p = prologue_row(L, pro++, soff);
                    return true;
                }
                p = L.u + (size_t)(c - 1) * L.pitch + soff;
                stage = 0;
                return true;
            }
            if (stage < 2 && StepRows::next(L, c, stage, p, soff)) return true;
            if (stage == 2) {
                stage = 3;
                p = L.u + (size_t)c * L.pitch + soff;
                return true;
            }
            w += stride;
            stage = -2;
        }
    }
};

template <class Phi>
__global__ void __launch_bounds__(Phi::T) k_residual(const LevelDev L, double *__restrict__ out_sq, const int nw,
                                                     const int nin) {
    MGB_RETURN_IF_STOPPED(L)
    using SH = typename Phi::SH;
    constexpr int E = Phi::E;
    Team<Phi::T> team(g_smem);
    RowPipe<SH, GenResidual> pipe(g_smem, nin, L.tile, Phi::row_n(L), GenResidual(L, blockIdx.x, nw, gridDim.x));
    pipe.start(team);
    typename Phi::C c;
    Phi::load_consts(c, L.sconst, team.tid);
    if (blockIdx.x == 0 && team.tid == 0) out_sq[0] = 0.0;
    typename Phi::Item it;
    bool have_item = false;
    for (int w = blockIdx.x; w < nw; w += gridDim.x) {
        const ItemPos ip = item_pos(L, w, 1);
        const int cp = __ldg(L.cpts + ip.k);
        if (!Phi::kItemInvariant || !have_item) Phi::begin_item(it, L, ip.sys, pipe, team);
        have_item = true;
        const double ct_c = ct_ahead<Phi>(L, cp);
        double x[E], y[E];
        pipe.pop(x, team);
        advance_ct<Phi>(x, c, it, L, cp, pipe, team, true, ct_c);
        pipe.pop(y, team);
        double acc = 0.0;
        const int nv = Phi::row_n(L) - team.tid * E;
#pragma unroll
        for (int q = 0; q < E; ++q) {
            const double r = x[q] - y[q];
            acc = (q < nv) ? fma(r, r, acc) : acc;
        }
        acc = team.sum(acc);
        if (team.tid == 0) out_sq[L.nsys <= 1 ? ip.k : L.ncpts + ip.k * L.nsys + ip.sys] = acc;
    }
    pipe.finish(team);
}

// out_sq[k] = sum over the systems of a row, k = 1 .. ncpts-1 (fixed order: deterministic)
template <class Phi>
__global__ void k_sum_systems(double *__restrict__ out_sq, const int ncpts, const int nsys) {
    const int k = 1 + blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= ncpts) return;
    const double *part = out_sq + ncpts + (size_t)k * nsys;
    double acc = 0.0;
    for (int s = 0; s < nsys; ++s) acc += part[s];
    out_sq[k] = acc;
}

// ------------------------------------------------------------------------------------------------
// One Phi for Application.step: out = Phi_point(in)   (rows in the level's layout)
// ------------------------------------------------------------------------------------------------
struct GenStep {
    LevelDev L;
    const double *in;
    int point, w, nw, stride, stage, pro;
    size_t soff;
    __device__ GenStep(const LevelDev &L_, const double *in_, int point_, int first, int nw_, int stride_)
        : L(L_), in(in_), point(point_), w(first), nw(nw_), stride(stride_), stage(-2), pro(0), soff(0) {}
    __device__ bool next(const double *&p) {
        for (;;) {
            if (w >= nw) return false;
            if (stage == -2) {
                soff = item_pos(L, w).soff;
                pro = 0;
                stage = -1;
            }
            if (stage == -1) {
                if (pro < n_prologue(L)) {
                    p = prologue_row(L, pro++, soff);
                    return true;
                }
                p = in + soff;
                stage = 0;
                return true;
            }
            if (stage == 0) {
                stage = 1;
                if (L.rhs_dense) {
                    p = L.rhs_dense + (size_t)point * L.pitch + soff;
                    return true;
                }
            }
            w += stride;
            stage = -2;
        }
    }
};

template <class Phi>
__global__ void __launch_bounds__(Phi::T) k_step(const LevelDev L, const int point, const double *in, double *out,
                                                 const int nw, const int nin) {
    using SH = typename Phi::SH;
    Team<Phi::T> team(g_smem);
    RowPipe<SH, GenStep> pipe(g_smem, nin, L.tile, Phi::row_n(L), GenStep(L, in, point, blockIdx.x, nw, gridDim.x));
    pipe.start(team);
    typename Phi::C c;
    Phi::load_consts(c, L.sconst, team.tid);
    for (int w = blockIdx.x; w < nw; w += gridDim.x) {
        const ItemPos ip = item_pos(L, w);
        typename Phi::Item it;
        Phi::begin_item(it, L, ip.sys, pipe, team);
        double x[Phi::E];
        pipe.pop(x, team);
        advance<Phi>(x, c, it, L, point, pipe, team, false);
        pipe.push(x, out + ip.soff, team);
    }
    pipe.finish(team);
}

}  // namespace mgb
