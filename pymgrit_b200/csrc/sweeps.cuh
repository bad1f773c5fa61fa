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

template <class Phi>
__device__ __forceinline__ void step_consts(typename Phi::C &c, const LevelDev &L, int i, int tid) {
    if (L.ndt > 1) Phi::load_consts(c, L.sconst + (size_t)__ldg(L.dtidx + i) * L.cw, tid);
}

// x <- (g_i +) Phi_i(x): pops the dense right-hand-side row and the g row if the level has them
template <class Phi, class Pipe, class TeamT>
__device__ __forceinline__ void advance(double (&x)[Phi::E], typename Phi::C &c, const LevelDev &L, int i, Pipe &pipe,
                                        TeamT &team, bool add_g = true) {
    step_consts<Phi>(c, L, i, team.tid);
    if (L.rhs_dense) {
        double b[Phi::E];
        pipe.pop(b, team);
        vadd(x, b);
    }
    Phi::apply(x, c, L, i, team);
    if (add_g && L.g) {
        double gg[Phi::E];
        pipe.pop(gg, team);
        vadd(x, gg);
    }
}

// rows `advance` pops for step i, in order: returns false when stage runs past them
struct StepRows {
    // stage 0: dense rhs row, stage 1: g row
    __device__ static __forceinline__ bool next(const LevelDev &L, int i, int &stage, const double *&p, bool with_g = true) {
        if (stage == 0) {
            stage = 1;
            if (L.rhs_dense) {
                p = L.rhs_dense + (size_t)i * L.pitch;
                return true;
            }
        }
        if (stage == 1) {
            stage = 2;
            if (with_g && L.g) {
                p = L.g + (size_t)i * L.pitch;
                return true;
            }
        }
        return false;
    }
};

// ------------------------------------------------------------------------------------------------
// F-relaxation / forward solve (mgrit.py:312-327, 471-481):  for each interval [s, e):
//   x = u[s];  for i = s+1 .. e-1:  x = (g[i] +) Phi_i(x);  u[i] = x
// ------------------------------------------------------------------------------------------------
struct GenChain {
    LevelDev L;
    int item, nitems, stride;
    int s, e, i, stage;
    __device__ GenChain(const LevelDev &L_, int first, int nitems_, int stride_)
        : L(L_), item(first), nitems(nitems_), stride(stride_), s(0), e(0), i(0), stage(-1) {}
    __device__ bool next(const double *&p) {
        for (;;) {
            if (item >= nitems) return false;
            if (stage < 0) {
                interval_of(L, item, s, e);
                if (e - s <= 1) {
                    item += stride;
                    continue;
                }
                p = L.u + (size_t)s * L.pitch;
                i = s + 1;
                stage = 0;
                return true;
            }
            if (i >= e) {
                item += stride;
                stage = -1;
                continue;
            }
            if (StepRows::next(L, i, stage, p)) return true;
            ++i;
            stage = 0;
        }
    }
};

template <class Phi>
__global__ void __launch_bounds__(Phi::T) k_chain(const LevelDev L, const int nitems, const int nin) {
    using SH = typename Phi::SH;
    Team<Phi::T> team(g_smem);
    RowPipe<SH, GenChain> pipe(g_smem, nin, L.pitch, L.n, GenChain(L, blockIdx.x, nitems, gridDim.x));
    pipe.start(team);
    typename Phi::C c;
    Phi::load_consts(c, L.sconst, team.tid);
    for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
        int s, e;
        interval_of(L, item, s, e);
        if (e - s <= 1) continue;
        double x[Phi::E];
        pipe.pop(x, team);
        for (int i = s + 1; i < e; ++i) {
            advance<Phi>(x, c, L, i, pipe, team);
            pipe.push(x, L.u + (size_t)i * L.pitch, team);
        }
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
    int item, nitems, stride, stage, kk;
    bool weighted;
    __device__ GenCRelax(const LevelDev &L_, int first, int nitems_, int stride_, bool weighted_)
        : L(L_), item(first), nitems(nitems_), stride(stride_), stage(-1), kk(0), weighted(weighted_) {}
    __device__ bool next(const double *&p) {
        for (;;) {
            if (item >= nitems) return false;
            if (stage < 0) {
                if (!c_run_leader(L, item)) {
                    item += stride;
                    continue;
                }
                kk = item;
                p = L.u + (size_t)(__ldg(L.cpts + kk) - 1) * L.pitch;
                stage = 0;
                return true;
            }
            const int c = __ldg(L.cpts + kk);
            if (stage < 2 && StepRows::next(L, c, stage, p)) return true;
            if (stage == 2) {
                stage = 3;
                if (weighted) {
                    p = L.u + (size_t)c * L.pitch;
                    return true;
                }
            }
            if (c_run_continues(L, kk)) {
                ++kk;
                stage = 0;
                continue;
            }
            item += stride;
            stage = -1;
        }
    }
};

template <class Phi>
__global__ void __launch_bounds__(Phi::T) k_c_relax(const LevelDev L, const double w, const int nin) {
    using SH = typename Phi::SH;
    const bool weighted = (w != 1.0);
    Team<Phi::T> team(g_smem);
    RowPipe<SH, GenCRelax> pipe(g_smem, nin, L.pitch, L.n, GenCRelax(L, 1 + blockIdx.x, L.ncpts, gridDim.x, weighted));
    pipe.start(team);
    typename Phi::C c;
    Phi::load_consts(c, L.sconst, team.tid);
    for (int k = 1 + blockIdx.x; k < L.ncpts; k += gridDim.x) {
        if (!c_run_leader(L, k)) continue;
        double x[Phi::E];
        pipe.pop(x, team);
        for (int kk = k;; ++kk) {
            const int cp = __ldg(L.cpts + kk);
            advance<Phi>(x, c, L, cp, pipe, team);
            if (weighted) {
                double old[Phi::E];
                pipe.pop(old, team);
                const double w1 = 1.0 - w;
#pragma unroll
                for (int j = 0; j < Phi::E; ++j) x[j] = __dadd_rn(__dmul_rn(x[j], w), __dmul_rn(old[j], w1));
            }
            pipe.push(x, L.u + (size_t)cp * L.pitch, team);
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
    int item, nitems, stride, stage, sub;
    __device__ GenFas(const LevelDev &L_, const LevelDev &G_, int first, int nitems_, int stride_)
        : L(L_), G(G_), item(first), nitems(nitems_), stride(stride_), stage(0), sub(0) {}
    __device__ bool next(const double *&p) {
        for (;;) {
            if (item >= nitems) return false;
            const int c = __ldg(L.cpts + item);
            switch (stage) {
                case 0:  // u[c-1]
                    p = L.u + (size_t)(c - 1) * L.pitch;
                    stage = 1;
                    sub = 0;
                    return true;
                case 1:  // dense rhs row of the fine step (no g: it is combined by hand below)
                    stage = 2;
                    if (StepRows::next(L, c, sub, p, false)) return true;
                    break;
                case 2:  // u[c]
                    p = L.u + (size_t)c * L.pitch;
                    stage = 3;
                    return true;
                case 3:  // g[c]
                    stage = 4;
                    if (L.g) {
                        p = L.g + (size_t)c * L.pitch;
                        return true;
                    }
                    break;
                case 4:  // v[j-1] = u[cpts[j-1]]
                    p = L.u + (size_t)__ldg(L.cpts + item - 1) * L.pitch;
                    stage = 5;
                    sub = 0;
                    return true;
                case 5:  // dense rhs row of the coarse step
                    stage = 6;
                    if (StepRows::next(G, item, sub, p, false)) return true;
                    break;
                default:
                    item += stride;
                    stage = 0;
            }
        }
    }
};

template <class Phi>
__global__ void __launch_bounds__(Phi::T) k_fas_residual(const LevelDev L, const LevelDev G, const int nin) {
    using SH = typename Phi::SH;
    constexpr int E = Phi::E;
    Team<Phi::T> team(g_smem);
    RowPipe<SH, GenFas> pipe(g_smem, nin, L.pitch, L.n, GenFas(L, G, 1 + blockIdx.x, L.ncpts, gridDim.x));
    pipe.start(team);
    typename Phi::C c;  // reloaded before each Phi: fine and coarse steps use different constants
    for (int j = 1 + blockIdx.x; j < L.ncpts; j += gridDim.x) {
        const int cp = __ldg(L.cpts + j);
        double x[E], y[E];
        pipe.pop(x, team);
        Phi::load_consts(c, L.sconst + (L.ndt > 1 ? (size_t)__ldg(L.dtidx + cp) * L.cw : 0), team.tid);
        advance<Phi>(x, c, L, cp, pipe, team, false);  // Phi_f(u[c-1])
        pipe.pop(y, team);                               // u[c]
        pipe.push(y, G.u + (size_t)j * G.pitch, team);   // injection
        if (L.g) {
            double gg[E];
            pipe.pop(gg, team);
#pragma unroll
            for (int q = 0; q < E; ++q) x[q] = ((gg[q] - y[q]) + x[q]) + y[q];
        } else {
#pragma unroll
            for (int q = 0; q < E; ++q) x[q] = (x[q] - y[q]) + y[q];
        }
        pipe.pop(y, team);                               // v[j-1]
        Phi::load_consts(c, G.sconst + (G.ndt > 1 ? (size_t)__ldg(G.dtidx + j) * G.cw : 0), team.tid);
        advance<Phi>(y, c, G, j, pipe, team, false);     // Phi_c(v[j-1])
#pragma unroll
        for (int q = 0; q < E; ++q) x[q] = x[q] - y[q];
        pipe.push(x, G.g + (size_t)j * G.pitch, team);
    }
    pipe.finish(team);
}

// ------------------------------------------------------------------------------------------------
// Coarse-grid correction (mgrit.py:722-726) fused with the F-relaxation that follows (mgrit.py:287):
//   for interval k:  if k >= kfirst: u[c] = u[c] + (G.u[k] - u[c]);   then, if f_relax, the chain from u[c].
// kfirst = 1 on time rank 0 (point 0 is the initial condition).  On the other ranks point 0 is the ghost copy of the
// previous rank's last C-point and kfirst = 0: the ghost is corrected here with the same arithmetic its owner uses
// (the coarse ghost row is already current), so the F-relaxation of the first interval starts from the corrected
// value without waiting for an exchange.
// ------------------------------------------------------------------------------------------------
struct GenCorrect {
    LevelDev L, G;
    int item, nitems, stride;
    int s, e, i, stage;
    bool frelax;
    int kfirst;
    __device__ GenCorrect(const LevelDev &L_, const LevelDev &G_, int first, int nitems_, int stride_, bool fr, int kf)
        : L(L_), G(G_), item(first), nitems(nitems_), stride(stride_), s(0), e(0), i(0), stage(-2), frelax(fr), kfirst(kf) {}
    __device__ bool next(const double *&p) {
        for (;;) {
            if (item >= nitems) return false;
            if (stage == -2) {
                interval_of(L, item, s, e);
                const bool chain = frelax && (e - s > 1);
                if (item < kfirst && !chain) {
                    item += stride;
                    continue;
                }
                p = L.u + (size_t)s * L.pitch;
                stage = -1;
                return true;
            }
            if (stage == -1) {
                stage = 0;
                i = s + 1;
                if (item >= kfirst) {
                    p = G.u + (size_t)item * G.pitch;
                    return true;
                }
            }
            if (!frelax || i >= e) {
                item += stride;
                stage = -2;
                continue;
            }
            if (StepRows::next(L, i, stage, p)) return true;
            ++i;
            stage = 0;
        }
    }
};

template <class Phi>
__global__ void __launch_bounds__(Phi::T) k_correct(const LevelDev L, const LevelDev G, const int frelax, const int kfirst,
                                                    const int nin) {
    using SH = typename Phi::SH;
    constexpr int E = Phi::E;
    Team<Phi::T> team(g_smem);
    RowPipe<SH, GenCorrect> pipe(g_smem, nin, L.pitch, L.n, GenCorrect(L, G, blockIdx.x, L.ncpts, gridDim.x, frelax != 0, kfirst));
    pipe.start(team);
    typename Phi::C c;
    Phi::load_consts(c, L.sconst, team.tid);
    for (int k = blockIdx.x; k < L.ncpts; k += gridDim.x) {
        int s, e;
        interval_of(L, k, s, e);
        const bool chain = frelax && (e - s > 1);
        if (k < kfirst && !chain) continue;
        double x[E];
        pipe.pop(x, team);
        if (k >= kfirst) {
            double cu[E];
            pipe.pop(cu, team);
#pragma unroll
            for (int q = 0; q < E; ++q) x[q] = x[q] + (cu[q] - x[q]);
            pipe.push(x, L.u + (size_t)s * L.pitch, team);
        }
        if (chain) {
            for (int i = s + 1; i < e; ++i) {
                advance<Phi>(x, c, L, i, pipe, team);
                pipe.push(x, L.u + (size_t)i * L.pitch, team);
            }
        }
    }
    pipe.finish(team);
}

// ------------------------------------------------------------------------------------------------
// Residual at C-points (mgrit.py:405-413): out[k] = || Phi(u[c-1]) - u[c] ||^2  (level 0: no g)
// ------------------------------------------------------------------------------------------------
struct GenResidual {
    LevelDev L;
    int item, nitems, stride, stage;
    __device__ GenResidual(const LevelDev &L_, int first, int nitems_, int stride_)
        : L(L_), item(first), nitems(nitems_), stride(stride_), stage(-1) {}
    __device__ bool next(const double *&p) {
        for (;;) {
            if (item >= nitems) return false;
            const int c = __ldg(L.cpts + item);
            if (stage < 0) {
                p = L.u + (size_t)(c - 1) * L.pitch;
                stage = 0;
                return true;
            }
            if (stage < 2 && StepRows::next(L, c, stage, p)) return true;
            if (stage == 2) {
                stage = 3;
                p = L.u + (size_t)c * L.pitch;
                return true;
            }
            item += stride;
            stage = -1;
        }
    }
};

template <class Phi>
__global__ void __launch_bounds__(Phi::T) k_residual(const LevelDev L, double *__restrict__ out_sq, const int nin) {
    using SH = typename Phi::SH;
    constexpr int E = Phi::E;
    Team<Phi::T> team(g_smem);
    RowPipe<SH, GenResidual> pipe(g_smem, nin, L.pitch, L.n, GenResidual(L, 1 + blockIdx.x, L.ncpts, gridDim.x));
    pipe.start(team);
    typename Phi::C c;
    Phi::load_consts(c, L.sconst, team.tid);
    if (blockIdx.x == 0 && team.tid == 0) out_sq[0] = 0.0;
    for (int k = 1 + blockIdx.x; k < L.ncpts; k += gridDim.x) {
        const int cp = __ldg(L.cpts + k);
        double x[E], y[E];
        pipe.pop(x, team);
        advance<Phi>(x, c, L, cp, pipe, team);
        pipe.pop(y, team);
        double acc = 0.0;
        const int nv = L.n - team.tid * E;
#pragma unroll
        for (int q = 0; q < E; ++q) {
            const double r = x[q] - y[q];
            acc = (q < nv) ? fma(r, r, acc) : acc;
        }
        acc = team.sum(acc);
        if (team.tid == 0) out_sq[k] = acc;
    }
    pipe.finish(team);
}

// ------------------------------------------------------------------------------------------------
// One Phi for Application.step: out = Phi_point(in)
// ------------------------------------------------------------------------------------------------
struct GenStep {
    LevelDev L;
    const double *in;
    int point, stage;
    __device__ GenStep(const LevelDev &L_, const double *in_, int point_) : L(L_), in(in_), point(point_), stage(-1) {}
    __device__ bool next(const double *&p) {
        if (stage < 0) {
            p = in;
            stage = 0;
            return true;
        }
        if (stage == 0) {
            stage = 1;
            if (L.rhs_dense) {
                p = L.rhs_dense + (size_t)point * L.pitch;
                return true;
            }
        }
        return false;
    }
};

template <class Phi>
__global__ void __launch_bounds__(Phi::T) k_step(const LevelDev L, const int point, const double *in, double *out,
                                                 const int nin) {
    using SH = typename Phi::SH;
    Team<Phi::T> team(g_smem);
    RowPipe<SH, GenStep> pipe(g_smem, nin, L.pitch, L.n, GenStep(L, in, point));
    pipe.start(team);
    typename Phi::C c;
    Phi::load_consts(c, L.sconst, team.tid);
    double x[Phi::E];
    pipe.pop(x, team);
    advance<Phi>(x, c, L, point, pipe, team, false);
    pipe.push(x, out, team);
    pipe.finish(team);
}

}  // namespace mgb
