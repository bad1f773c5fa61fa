// api.cu -- the C ABI of include/mgrit_b200.h: argument checks, shape dispatch, host-side constant
// tables, and the small row-wise kernels (injection, jump, temporal norm, Vector arithmetic).
#include <cmath>
#include <cstdio>
#include <cstring>
#include <mutex>

#include "../../include/mgrit_b200.h"
#include "phi.cuh"
#include "table.h"

namespace mgb {

// ---- error state ---------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

static int fail(int code, const char *fmt, const char *a = "", long b = 0, long c = 0) {
    snprintf(g_err, sizeof g_err, fmt, a, b, c);
    return code;
}

int heat2d_fail(const char *msg) { return fail(MGB_EINVAL, "%s", msg); }

int cuda_fail(cudaError_t e, const char *what) {
    if (e == cudaSuccess) return MGB_OK;
    snprintf(g_err, sizeof g_err, "%s: %s", what, cudaGetErrorString(e));
    return MGB_ECUDA;
}

// levels with at least this many points are streams of rows when every F-point is stored: the team kernels' bulk stores
// write them best; shorter levels are latency-bound and run as one thread per mode (sine_modes.cu)
constexpr int kStreamPoints = 1 << 17;

static const int *g_stop = nullptr;
const int *stop_flag() { return g_stop; }

const DeviceInfo *device_info() {
    static DeviceInfo info;
    static int state = 0;  // 0 unknown, 1 ok, 2 failed
    static std::mutex mu;
    std::lock_guard<std::mutex> lock(mu);
    if (state == 0) {
        int dev = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e == cudaSuccess) e = cudaDeviceGetAttribute(&info.sms, cudaDevAttrMultiProcessorCount, dev);
        if (e == cudaSuccess) e = cudaDeviceGetAttribute(&info.max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        if (e != cudaSuccess) {
            cuda_fail(e, "no usable CUDA device (this library has no CPU fallback)");
            state = 2;
        } else {
            state = 1;
        }
    }
    return state == 1 ? &info : nullptr;
}

// ---- compiled shapes --------------------------------------------------------------------------------
#define MGB_APP_Heat1D MGB_APP_HEAT1D
#define MGB_APP_Advection1D MGB_APP_ADVECTION1D
#define MGB_APP_Heat2D MGB_APP_HEAT2D
#define MGB_APP_Heat1D2Pts MGB_APP_HEAT1D_2PTS
#define MGB_APP_Heat1DSine MGB_APP_HEAT1D_SINE
#define MGB_SHAPE(APP, T, E) const SweepTable *mgb_table_##APP##_##T##_##E();
#include "shapes.inc"
#undef MGB_SHAPE
const SweepTable *mgb_table_tiny(int app);  // tiny.cu

struct ShapeEntry {
    int app, T, E;
    const SweepTable *(*get)();
};
static const ShapeEntry g_shapes[] = {
#define MGB_SHAPE(APP, T, E) {MGB_APP_##APP, T, E, &mgb_table_##APP##_##T##_##E},
#include "shapes.inc"
#undef MGB_SHAPE
};

static const SweepTable *find_table(int app, int T, int E) {
    if (app == MGB_APP_DAHLQUIST || app == MGB_APP_BRUSSELATOR) return mgb_table_tiny(app);
    for (const ShapeEntry &s : g_shapes)
        if (s.app == app && s.T == T && s.E == E) return s.get();
    return nullptr;
}

static int check_level(const mgb_level *l, const SweepTable **tab, LevelDev *out) {
    if (l == nullptr) return fail(MGB_EINVAL, "null level%s");
    if (l->n < 1 || l->pitch < l->n || l->npts < 1 || l->u_dev == nullptr)
        return fail(MGB_EINVAL, "bad level geometry%s (n=%ld, npts=%ld)", "", l->n, l->npts);
    const bool tiny = (l->app == MGB_APP_DAHLQUIST || l->app == MGB_APP_BRUSSELATOR);
    const bool multi = (l->app == MGB_APP_HEAT2D);
    if (multi) {
        const long tl = (long)l->team_threads * l->chunk;
        if (l->nsys < 1 || (long)l->nsys * tl != l->pitch || l->n != l->pitch)
            return fail(MGB_EINVAL, "heat2d rows must be whole tiles%s (pitch %ld, nsys %ld)", "", l->pitch, l->nsys);
        if (l->sig_dev == nullptr || l->sconst_dev == nullptr || l->ndt < 1 || (l->ndt > 1 && l->dtidx_dev == nullptr))
            return fail(MGB_EINVAL, "missing heat2d symbol or step-constant table%s");
        if (l->nrhs < 0 || l->nrhs > kHeat2DMaxTerms || (l->nrhs > 0 && (l->rhs_x_dev == nullptr || l->rhs_t_dev == nullptr)))
            return fail(MGB_EINVAL, "heat2d takes at most 3 separable right-hand-side terms%s (got %ld)", "", l->nrhs);
    } else if (l->app == MGB_APP_HEAT1D_2PTS) {
        const long h = (l->chunk - 1) / 2;
        if (l->chunk < 3 || l->chunk % 2 == 0 || (long)l->team_threads * h < l->n ||
            (long)l->team_threads * l->chunk != l->pitch)
            return fail(MGB_EINVAL, "two-point rows%s: pitch %ld must be team_threads * chunk and cover n = %ld", "",
                        l->pitch, l->n);
        if (l->sconst_dev == nullptr || l->ndt < 1 || (l->ndt > 1 && l->dtidx_dev == nullptr))
            return fail(MGB_EINVAL, "missing step-constant table%s");
        if (l->nrhs > 0 && (l->rhs_x_dev == nullptr || l->rhs_t_dev == nullptr))
            return fail(MGB_EINVAL, "missing right-hand-side tables%s");
        if (l->rhs_dense_dev != nullptr)
            return fail(MGB_EINVAL, "two-point heat levels take a separable right-hand side only%s");
    } else if (l->app == MGB_APP_HEAT1D_SINE) {
        if (l->pitch % 2) return fail(MGB_EINVAL, "pitch must be even%s (got %ld)", "", l->pitch);
        if ((long)l->team_threads * l->chunk < l->pitch)
            return fail(MGB_EINVAL, "team shape%s %ld x %ld does not cover a row", "", l->team_threads, l->chunk);
        if (l->diag_dev == nullptr || l->sconst_dev == nullptr || l->ndt < 1 || (l->ndt > 1 && l->dtidx_dev == nullptr))
            return fail(MGB_EINVAL, "missing eigenvalue or step-constant table of a sine-space level%s");
        if (l->nrhs > 0 && (l->rhs_x_dev == nullptr || l->rhs_t_dev == nullptr))
            return fail(MGB_EINVAL, "missing right-hand-side tables%s");
        if (l->rhs_dense_dev != nullptr)
            return fail(MGB_EINVAL, "sine-space heat levels take a separable right-hand side only%s");
    } else if (!tiny) {
        if (l->pitch % 2) return fail(MGB_EINVAL, "pitch must be even%s (got %ld)", "", l->pitch);
        if ((long)l->team_threads * l->chunk < l->pitch)
            return fail(MGB_EINVAL, "team shape%s %ld x %ld does not cover a row", "", l->team_threads, l->chunk);
        if (l->sconst_dev == nullptr || l->ndt < 1 || (l->ndt > 1 && l->dtidx_dev == nullptr))
            return fail(MGB_EINVAL, "missing step-constant table%s");
        if (l->nrhs > 0 && (l->rhs_x_dev == nullptr || l->rhs_t_dev == nullptr))
            return fail(MGB_EINVAL, "missing right-hand-side tables%s");
    }
    *tab = find_table(l->app, l->team_threads, l->chunk);
    if (*tab == nullptr)
        return fail(MGB_ENOSHAPE, "no kernels compiled for app%s %ld with team %ld", "", l->app, l->team_threads);
    out->u = l->u_dev;
    out->g = l->g_dev;
    out->cpts = l->cpts_dev;
    out->ncpts = l->cpts_dev ? l->ncpts : 0;
    out->npts = l->npts;
    out->n = l->n;
    out->pitch = l->pitch;
    out->ndt = l->ndt;
    out->cw = l->cw;
    out->dtidx = l->dtidx_dev;
    out->sconst = l->sconst_dev;
    out->nrhs = l->nrhs;
    out->rhs_x = l->rhs_x_dev;
    out->rhs_t = l->rhs_t_dev;
    out->rhs_dense = l->rhs_dense_dev;
    out->t = l->t_dev;
    for (int k = 0; k < 8; ++k) out->p[k] = l->p[k];
    for (int k = 0; k < 4; ++k) out->ip[k] = l->ip[k];
    out->nsys = multi ? l->nsys : 1;
    out->tile = multi ? l->team_threads * l->chunk : l->pitch;
    out->nrow = (l->app == MGB_APP_HEAT1D_2PTS) ? l->pitch : l->n;  // doubles of a row the row-wise helpers touch
    out->sig = multi ? l->sig_dev : nullptr;
    out->diag = (l->app == MGB_APP_HEAT1D_SINE) ? l->diag_dev : nullptr;
    out->nat = (l->app == MGB_APP_HEAT1D_SINE) ? l->nat_dev : nullptr;
    out->stop = g_stop;
    if (multi) out->n = out->tile;
    if (tiny && l->t_dev == nullptr) return fail(MGB_EINVAL, "ODE applications need the time grid t_dev%s");
    return MGB_OK;
}

static int check_pair(const mgb_level *f, const mgb_level *c) {
    if (f->app != c->app || f->n != c->n || f->pitch != c->pitch || f->team_threads != c->team_threads ||
        f->chunk != c->chunk)
        return fail(MGB_EINVAL, "fine and coarse level differ in application or spatial size%s");
    if (f->cpts_dev == nullptr || c->npts < f->ncpts)
        return fail(MGB_EINVAL, "coarse level has fewer points than the fine level has C-points%s");
    if (c->g_dev == nullptr) return fail(MGB_EINVAL, "coarse level needs a g array%s");
    if (f->app == MGB_APP_HEAT2D) {
        // one spatial operator for the whole hierarchy: the item data loaded for the fine step serves the coarse step
        if (f->nsys != c->nsys || f->sig_dev != c->sig_dev || f->ip[0] != c->ip[0])
            return fail(MGB_EINVAL, "heat2d levels must share the symbol table%s");
        if (c->nrhs != 0 && (c->nrhs != f->nrhs || c->rhs_x_dev != f->rhs_x_dev))
            return fail(MGB_EINVAL, "heat2d levels must share the right-hand-side factors%s");
    }
    return MGB_OK;
}

// ---- row-wise helper kernels ------------------------------------------------------------------------
__global__ void k_inject_up(double *__restrict__ fu, const double *__restrict__ cu, const int *__restrict__ cpts,
                            int ncpts, int n, int pitch) {
    for (int j = 1 + blockIdx.x; j < ncpts; j += gridDim.x) {
        double *dst = fu + (size_t)cpts[j] * pitch;
        const double *src = cu + (size_t)j * pitch;
        for (int q = threadIdx.x; q < n; q += blockDim.x) dst[q] = src[q];
    }
}

__device__ __forceinline__ double block_sum(double v) {
    __shared__ double s[32];
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) s[warp] = v;
    __syncthreads();
    double tot = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += s[w];
    return tot;
}

// out_sq[j] = ||u[c_j] - last[c_j]||^2 (j >= 1); then last <- u everywhere (mgrit.py:379-384)
__global__ void k_jump(const double *__restrict__ u, double *__restrict__ last, const int *__restrict__ cpts, int ncpts,
                       int npts, int n, int pitch, double *__restrict__ out_sq) {
    if (blockIdx.x == 0 && threadIdx.x == 0) out_sq[0] = 0.0;
    for (int j = 1 + blockIdx.x; j < ncpts; j += gridDim.x) {
        const size_t off = (size_t)cpts[j] * pitch;
        double acc = 0.0;
        for (int q = threadIdx.x; q < n; q += blockDim.x) {
            const double d = u[off + q] - last[off + q];
            acc = fma(d, d, acc);
        }
        acc = block_sum(acc);
        if (threadIdx.x == 0) out_sq[j] = acc;
    }
}

__global__ void k_copy_rows(const double *__restrict__ src, double *__restrict__ dst, size_t count) {
    for (size_t q = blockIdx.x * (size_t)blockDim.x + threadIdx.x; q < count; q += (size_t)gridDim.x * blockDim.x)
        dst[q] = src[q];
}

// single CTA, fixed-order reduction -> deterministic.  Four independent accumulators per thread keep enough loads in
// flight that the single CTA is not latency-bound on a 2 MB input.
__device__ __forceinline__ double tn_map(double v, int mode) { return mode == MGB_TNORM_TWO ? v : sqrt(v); }
__device__ __forceinline__ double tn_comb(double a, double b, int mode) { return mode == MGB_TNORM_INF ? fmax(a, b) : a + b; }

// hist != nullptr (one time rank: nothing to reduce between the norm and the stopping test): the norm also goes to hist[0]
// and the flag goes up if the stopping test of mgrit.py:626 holds -- k_convergence_flag in the same launch
__global__ void k_temporal_norm(const double *__restrict__ sq, int count, int mode, double *__restrict__ out,
                                const int *__restrict__ stop, double tol, double *__restrict__ hist, int *__restrict__ flag) {
    if (stop != nullptr && *stop != 0) return;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    const int T = blockDim.x;
    int q = threadIdx.x;
    for (; q + 3 * T < count; q += 4 * T) {
        const double v0 = sq[q], v1 = sq[q + T], v2 = sq[q + 2 * T], v3 = sq[q + 3 * T];
        a0 = tn_comb(a0, tn_map(v0, mode), mode);
        a1 = tn_comb(a1, tn_map(v1, mode), mode);
        a2 = tn_comb(a2, tn_map(v2, mode), mode);
        a3 = tn_comb(a3, tn_map(v3, mode), mode);
    }
    for (; q < count; q += T) a0 = tn_comb(a0, tn_map(sq[q], mode), mode);
    double acc = tn_comb(tn_comb(a0, a1, mode), tn_comb(a2, a3, mode), mode);
    __shared__ double s[1024];
    s[threadIdx.x] = acc;
    __syncthreads();
    for (int d = blockDim.x >> 1; d >= 1; d >>= 1) {
        if ((int)threadIdx.x < d) s[threadIdx.x] = tn_comb(s[threadIdx.x], s[threadIdx.x + d], mode);
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        out[0] = s[0];
        if (hist != nullptr) {
            hist[0] = s[0];
            const double conv = (mode == MGB_TNORM_TWO) ? sqrt(s[0]) : s[0];
            if (conv < tol) *flag = 1;
        }
    }
}

__global__ void k_axpby(int n, double a, const double *__restrict__ x, double b, const double *__restrict__ y,
                        double *__restrict__ out) {
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < n; q += gridDim.x * blockDim.x) {
        // products and sum rounded separately, like numpy's x*a + y*b
        const double xa = (a == 1.0) ? x[q] : __dmul_rn(x[q], a);
        const double yb = (b == 0.0 || y == nullptr) ? 0.0 : ((b == 1.0) ? y[q] : (b == -1.0 ? -y[q] : __dmul_rn(y[q], b)));
        out[q] = (b == 0.0 || y == nullptr) ? xa : __dadd_rn(xa, yb);
    }
}

__global__ void k_sumsq(int n, const double *__restrict__ x, double *__restrict__ out) {
    double acc = 0.0;
    for (int q = threadIdx.x; q < n; q += blockDim.x) acc = fma(x[q], x[q], acc);
    acc = block_sum(acc);
    if (threadIdx.x == 0) out[0] = acc;
}

// hist[0] = the (rank-reduced) temporal norm; the flag goes up if the stopping test of mgrit.py:626 holds
__global__ void k_convergence_flag(const double *__restrict__ partial, int mode, double tol, double *__restrict__ hist,
                                   int *__restrict__ flag) {
    if (*flag != 0) return;  // queued after the criterion was met: leave the history alone
    const double v = partial[0];
    hist[0] = v;
    const double conv = (mode == MGB_TNORM_TWO) ? sqrt(v) : v;
    if (conv < tol) *flag = 1;
}

__global__ void k_set_flag(int *flag, int value) { *flag = value; }

// ---- host-side step constants -------------------------------------------------------------------------
static long double lpow(long double b, long e) { return e <= 0 ? 1.0L : powl(b, (long double)e); }

// doubling steps of the cross-lane scan whose ratio B^(2^k) still matters (>= 1e-30)
static int scan_steps(long double B) {
    int n = 0;
    while (n < 5 && lpow(B, 1L << n) >= 1e-30L) ++n;
    return n;
}

}  // namespace mgb

using namespace mgb;

extern "C" {

int mgb_abi_version(void) { return MGB_ABI_VERSION; }

const char *mgb_last_error(void) { return g_err; }

int mgb_team_shape(int32_t app, int32_t n, int32_t *team_threads, int32_t *chunk) {
    if (n < 1 || team_threads == nullptr || chunk == nullptr) return fail(MGB_EINVAL, "bad argument%s");
    if (app == MGB_APP_DAHLQUIST || app == MGB_APP_BRUSSELATOR) {
        *team_threads = 1;
        *chunk = n;
        return MGB_OK;
    }
    if (app == MGB_APP_HEAT2D) {  // elementwise in sine space: one tile shape for every grid size
        for (const ShapeEntry &s : g_shapes)
            if (s.app == app) {
                *team_threads = s.T;
                *chunk = s.E;
                return MGB_OK;
            }
        return fail(MGB_ENOSHAPE, "no heat2d kernels compiled%s");
    }
    const int pitch = n + (n & 1);
    long best = -1;
    for (const ShapeEntry &s : g_shapes) {
        // two-point rows: a thread owns (E - 1) / 2 elements of each of the two time points
        const long cover = (app == MGB_APP_HEAT1D_2PTS) ? (long)s.T * ((s.E - 1) / 2) : (long)s.T * s.E;
        if (s.app != app || cover < (app == MGB_APP_HEAT1D_2PTS ? n : pitch)) continue;
        // fewest threads first, then the smallest chunk
        const long key = (long)s.T * 1000 + s.E;
        if (best < 0 || key < best) {
            best = key;
            *team_threads = s.T;
            *chunk = s.E;
        }
    }
    if (best < 0) return fail(MGB_ENOSHAPE, "no kernel shape for app%s %ld with n = %ld", "", app, n);
    return MGB_OK;
}

int mgb_step_consts_width(int32_t app, int32_t team_threads, int32_t chunk) {
    if (app == MGB_APP_HEAT2D || app == MGB_APP_HEAT1D_SINE) return 8;  // [0] dt
    if (app == MGB_APP_HEAT1D_2PTS) return 2 * mgb_heat1d_2pts_half_width(team_threads, chunk) + 8;
    const int sub = (chunk % 3 == 0) ? 3 : 1;
    return kScalarConsts + team_threads * (2 + 2 * sub);
}

int mgb_heat1d_2pts_half_width(int32_t team_threads, int32_t chunk) {
    const int h = (chunk - 1) / 2;
    const int sub = (h % 3 == 0) ? 3 : 1;
    return kScalarConsts + team_threads * (2 + 2 * sub);
}

int mgb_heat1d_step_consts(double r_, int32_t n, int32_t T, int32_t E, double *out) {
    if (!(r_ > 0.0) || n < 1 || T < 32 || E < 1 || out == nullptr) return fail(MGB_EINVAL, "bad argument%s");
    const int SUB = (E % 3 == 0) ? 3 : 1, SL = E / SUB, PT = 2 + 2 * SUB;
    if (SL > 13) return fail(MGB_EINVAL, "chunk too long%s");
    const long double r = r_;
    // delta = 2 + 1/r; beta = smaller root of b^2 - delta b + 1;  delta^2 - 4 = (4 + 1/r)/r
    const long double delta = 2.0L + 1.0L / r;
    const long double beta = 2.0L / (delta + sqrtl((4.0L + 1.0L / r) / r));
    const long double om = 1.0L - beta * beta;
    const long double h0 = (beta - lpow(beta, 2L * n + 1)) / om;
    const long double B = lpow(beta, E);
    memset(out, 0, sizeof(double) * (size_t)(kScalarConsts + T * PT));
    out[0] = (double)beta;
    out[1] = (double)(beta / r);
    out[2] = (double)(beta / (1.0L + beta * h0));
    out[3] = (double)lpow(beta, SL);
    for (int k = 0; k < 5; ++k) out[4 + k] = (double)lpow(B, 1L << k);
    out[9] = (double)lpow(B, 32);
    for (int j = 0; j < SL; ++j) out[10 + j] = (double)lpow(beta, j + 1);
    out[23] = (double)scan_steps(B);
    for (int t = 0; t < T; ++t) {
        double *pt = out + kScalarConsts + t * PT;
        const int lane = t & 31;
        pt[0] = (double)lpow(B, lane);
        pt[1] = (double)lpow(B, 31 - lane);
        for (int s = 0; s < SUB; ++s) {
            const long base = (long)t * E + (long)s * SL;  // first element of the sub-chunk
            if (base >= n) continue;                      // no valid element: leave zeros
            pt[2 + s] = (double)(lpow(beta, base) / om);
            const long ex = 2L * n + 1 - base - SL;       // >= 0 whenever the sub-chunk holds a valid element
            pt[2 + SUB + s] = (double)((ex >= 0 ? lpow(beta, ex) : 1.0L / lpow(beta, -ex)) / om);
        }
    }
    return MGB_OK;
}

int mgb_advection1d_step_consts(double nu_, int32_t n, int32_t T, int32_t E, double *out) {
    if (!(nu_ > 0.0) || n < 1 || T < 32 || E < 1 || out == nullptr) return fail(MGB_EINVAL, "bad argument%s");
    const int SUB = (E % 3 == 0) ? 3 : 1, SL = E / SUB, PT = 2 + 2 * SUB;
    if (SL > 13) return fail(MGB_EINVAL, "chunk too long%s");
    const long double nu = nu_;
    const long double rho = nu / (1.0L + nu), sig = 1.0L / (1.0L + nu);
    const long double B = lpow(rho, E);
    memset(out, 0, sizeof(double) * (size_t)(kScalarConsts + T * PT));
    out[0] = (double)rho;
    out[1] = (double)sig;
    out[2] = (double)(1.0L / (1.0L - lpow(rho, n)));
    out[3] = (double)lpow(rho, SL);
    for (int k = 0; k < 5; ++k) out[4 + k] = (double)lpow(B, 1L << k);
    out[9] = (double)lpow(B, 32);
    for (int j = 0; j < SL; ++j) out[10 + j] = (double)lpow(rho, j + 1);
    out[23] = (double)scan_steps(B);
    for (int t = 0; t < T; ++t) {
        double *pt = out + kScalarConsts + t * PT;
        pt[0] = (double)lpow(B, t & 31);
        for (int s = 0; s < SUB; ++s) pt[2 + s] = (double)lpow(rho, (long)t * E + (long)s * SL);
    }
    return MGB_OK;
}

#define MGB_PROLOGUE(lvl)                \
    const SweepTable *tab = nullptr;     \
    LevelDev L;                          \
    if (int rc = check_level(lvl, &tab, &L)) return rc; \
    cudaStream_t st = (cudaStream_t)stream;

int mgb_f_relax(const mgb_level *lvl, int32_t flags, void *stream) {
    MGB_PROLOGUE(lvl)
    if (L.cpts == nullptr) return fail(MGB_EINVAL, "f_relax needs the C-point table%s");
    // one thread per mode for the chains that store one point; with every F-point stored the sweep is a stream of rows and
    // the team kernel's bulk stores are the better writer (88 % against 85 % of the HBM peak, profiles/r02p_*)
    if (sine_modes_entry(0) && sine_modes_ok(L) && ((flags & MGB_F_RELAX_LAST_ONLY) || L.npts < kStreamPoints)) return sine_modes_f_relax(L, flags, st);
    return tab->f_relax(L, flags, st);
}

int mgb_c_relax(const mgb_level *lvl, double weight, void *stream) {
    MGB_PROLOGUE(lvl)
    if (L.cpts == nullptr) return fail(MGB_EINVAL, "c_relax needs the C-point table%s");
    return tab->c_relax(L, weight, 0, st);
}

int mgb_c_relax_last(const mgb_level *lvl, double weight, void *stream) {
    MGB_PROLOGUE(lvl)
    if (L.cpts == nullptr) return fail(MGB_EINVAL, "c_relax needs the C-point table%s");
    return tab->c_relax(L, weight, 1, st);
}

int mgb_fas_residual(const mgb_level *fine, const mgb_level *coarse, void *stream) {
    MGB_PROLOGUE(fine)
    const SweepTable *tab2 = nullptr;
    LevelDev G;
    if (int rc = check_level(coarse, &tab2, &G)) return rc;
    if (int rc = check_pair(fine, coarse)) return rc;
    return tab->fas_residual(L, G, st);
}

int mgb_down_sweep(const mgb_level *fine, const mgb_level *coarse, void *stream) {
    MGB_PROLOGUE(fine)
    const SweepTable *tab2 = nullptr;
    LevelDev G;
    if (int rc = check_level(coarse, &tab2, &G)) return rc;
    if (int rc = check_pair(fine, coarse)) return rc;
    if (tab->down == nullptr) return fail(MGB_ENOSHAPE, "no fused down-sweep for this application%s");
    if (sine_modes_entry(1) && sine_modes_ok(L) && sine_modes_coarse_ok(G, L)) return sine_modes_down(L, G, st);
    return tab->down(L, G, st);
}

int mgb_error_correction(const mgb_level *fine, const mgb_level *coarse, int32_t flags, void *stream) {
    MGB_PROLOGUE(fine)
    const SweepTable *tab2 = nullptr;
    LevelDev G;
    if (int rc = check_level(coarse, &tab2, &G)) return rc;
    if (int rc = check_pair(fine, coarse)) return rc;
    const int frelax = (flags & MGB_CORRECT_F_RELAX) ? ((flags & MGB_CORRECT_LAST_ONLY) ? 2 : 1) : 0;
    if (sine_modes_entry(2) && sine_modes_ok(L) && (frelax != 1 || L.npts < kStreamPoints)) return sine_modes_correct(L, G, frelax, (flags & MGB_CORRECT_GHOST) ? 0 : 1, st);
    return tab->correct(L, G, frelax, (flags & MGB_CORRECT_GHOST) ? 0 : 1, st);
}

int mgb_forward_solve(const mgb_level *lvl, void *stream) {
    MGB_PROLOGUE(lvl)
    return tab->forward_solve(L, st);
}

int mgb_local_coarse_solve(const mgb_level *lvl, const double *old_dev, int32_t k, void *stream) {
    MGB_PROLOGUE(lvl)
    if (old_dev == nullptr || k < 1 || old_dev == L.u) return fail(MGB_EINVAL, "local_coarse_solve: needs k >= 1 and a copy of u%s");
    if (tab->window == nullptr) return fail(MGB_ENOSHAPE, "no local coarse-grid solve for this application%s");
    return tab->window(L, old_dev, k, st);
}

int mgb_residual_norms(const mgb_level *lvl, double *out_sq_dev, void *stream) {
    MGB_PROLOGUE(lvl)
    if (L.cpts == nullptr || out_sq_dev == nullptr) return fail(MGB_EINVAL, "residual_norms needs C-points and an output%s");
    if (sine_modes_entry(3) && sine_modes_ok(L)) return sine_modes_residual(L, out_sq_dev, st);
    return tab->residual(L, out_sq_dev, st);
}

int mgb_residual_rows(const mgb_level *lvl, double *out_rows_dev, void *stream) {
    MGB_PROLOGUE(lvl)
    if (L.cpts == nullptr || out_rows_dev == nullptr) return fail(MGB_EINVAL, "residual_rows needs C-points and an output%s");
    if (tab->residual_rows == nullptr) return fail(MGB_ENOSHAPE, "no spatial-transfer sweeps for this application%s");
    return tab->residual_rows(L, out_rows_dev, st);
}

int mgb_fas_coarse_rhs(const mgb_level *coarse, const double *v_dev, const double *rres_dev, int32_t nrows, void *stream) {
    MGB_PROLOGUE(coarse)
    if (v_dev == nullptr || rres_dev == nullptr || nrows < 0 || nrows > L.npts || L.g == nullptr)
        return fail(MGB_EINVAL, "fas_coarse_rhs: bad argument%s");
    if (tab->fas_rhs == nullptr) return fail(MGB_ENOSHAPE, "no spatial-transfer sweeps for this application%s");
    return tab->fas_rhs(L, v_dev, rres_dev, nrows, st);
}

int mgb_jump_norms(const mgb_level *lvl, double *last_dev, double *out_sq_dev, void *stream) {
    MGB_PROLOGUE(lvl)
    (void)tab;
    if (L.cpts == nullptr || last_dev == nullptr || out_sq_dev == nullptr) return fail(MGB_EINVAL, "bad argument%s");
    const DeviceInfo *di = device_info();
    if (di == nullptr) return MGB_ECUDA;
    const int threads = L.nrow >= 256 ? 256 : 32 * ((L.nrow + 31) / 32);
    int grid = L.ncpts - 1 < 8 * di->sms ? L.ncpts - 1 : 8 * di->sms;
    if (grid < 1) grid = 1;
    k_jump<<<grid, threads, 0, st>>>(L.u, last_dev, L.cpts, L.ncpts, L.npts, L.nrow, L.pitch, out_sq_dev);
    const size_t count = (size_t)L.npts * L.pitch;
    k_copy_rows<<<4 * di->sms, 256, 0, st>>>(L.u, last_dev, count);
    return cuda_fail(cudaGetLastError(), "jump_norms");
}

int mgb_temporal_norm(const double *sq_dev, int32_t count, int32_t t_norm, double *out_dev, void *stream) {
    if (sq_dev == nullptr || out_dev == nullptr || count < 0 || t_norm < 1 || t_norm > 3)
        return fail(MGB_EINVAL, "bad argument%s");
    if (device_info() == nullptr) return MGB_ECUDA;
    k_temporal_norm<<<1, 1024, 0, (cudaStream_t)stream>>>(sq_dev, count, t_norm, out_dev, g_stop, 0.0, nullptr, nullptr);
    return cuda_fail(cudaGetLastError(), "temporal_norm");
}

int mgb_temporal_norm_flag(const double *sq_dev, int32_t count, int32_t t_norm, double *out_dev, double tol, double *hist_dev,
                           int32_t *flag_dev, void *stream) {
    if (sq_dev == nullptr || out_dev == nullptr || hist_dev == nullptr || flag_dev == nullptr || count < 0 || t_norm < 1 ||
        t_norm > 3)
        return fail(MGB_EINVAL, "bad argument%s");
    if (device_info() == nullptr) return MGB_ECUDA;
    // the kernel's own stop test (flag_dev is the registered stop flag while a solve is queued ahead) keeps the history
    // of an iteration that was queued after the criterion was met untouched, like k_convergence_flag
    k_temporal_norm<<<1, 1024, 0, (cudaStream_t)stream>>>(sq_dev, count, t_norm, out_dev, flag_dev, tol, hist_dev, flag_dev);
    return cuda_fail(cudaGetLastError(), "temporal_norm_flag");
}

int mgb_set_stop_flag(const int32_t *flag_dev) {
    g_stop = flag_dev;
    return MGB_OK;
}

int mgb_write_flag(int32_t *flag_dev, int32_t value, void *stream) {
    if (flag_dev == nullptr) return fail(MGB_EINVAL, "bad argument%s");
    if (device_info() == nullptr) return MGB_ECUDA;
    k_set_flag<<<1, 1, 0, (cudaStream_t)stream>>>(flag_dev, value);
    return cuda_fail(cudaGetLastError(), "write_flag");
}

int mgb_convergence_flag(const double *norm_dev, int32_t t_norm, double tol, double *hist_dev, int32_t *flag_dev,
                         void *stream) {
    if (norm_dev == nullptr || hist_dev == nullptr || flag_dev == nullptr || t_norm < 1 || t_norm > 3)
        return fail(MGB_EINVAL, "bad argument%s");
    if (device_info() == nullptr) return MGB_ECUDA;
    k_convergence_flag<<<1, 1, 0, (cudaStream_t)stream>>>(norm_dev, t_norm, tol, hist_dev, flag_dev);
    return cuda_fail(cudaGetLastError(), "convergence_flag");
}

int mgb_inject_up(const mgb_level *fine, const mgb_level *coarse, void *stream) {
    MGB_PROLOGUE(fine)
    (void)tab;
    if (coarse == nullptr || coarse->u_dev == nullptr || L.cpts == nullptr || coarse->npts < L.ncpts ||
        coarse->pitch != L.pitch)
        return fail(MGB_EINVAL, "bad level pair%s");
    const DeviceInfo *di = device_info();
    if (di == nullptr) return MGB_ECUDA;
    if (L.ncpts < 2) return MGB_OK;
    const int threads = L.nrow >= 256 ? 256 : 32 * ((L.nrow + 31) / 32);
    const int grid = L.ncpts - 1 < 8 * di->sms ? L.ncpts - 1 : 8 * di->sms;
    k_inject_up<<<grid, threads, 0, st>>>(L.u, coarse->u_dev, L.cpts, L.ncpts, L.nrow, L.pitch);
    return cuda_fail(cudaGetLastError(), "inject_up");
}

int mgb_step(const mgb_level *lvl, int32_t point, const double *in_dev, double *out_dev, void *stream) {
    MGB_PROLOGUE(lvl)
    if (point < 1 || point >= L.npts || in_dev == nullptr || out_dev == nullptr)
        return fail(MGB_EINVAL, "bad step%s (point %ld of %ld)", "", point, L.npts);
    return tab->step(L, point, in_dev, out_dev, st);
}

int mgb_vec_axpby(int32_t n, double a, const double *x_dev, double b, const double *y_dev, double *out_dev, void *stream) {
    if (n < 1 || x_dev == nullptr || out_dev == nullptr) return fail(MGB_EINVAL, "bad argument%s");
    if (device_info() == nullptr) return MGB_ECUDA;
    const int threads = 256;
    int grid = (n + threads - 1) / threads;
    if (grid > 1184) grid = 1184;
    k_axpby<<<grid, threads, 0, (cudaStream_t)stream>>>(n, a, x_dev, b, y_dev, out_dev);
    return cuda_fail(cudaGetLastError(), "vec_axpby");
}

int mgb_vec_sumsq(int32_t n, const double *x_dev, double *out_dev, void *stream) {
    if (n < 1 || x_dev == nullptr || out_dev == nullptr) return fail(MGB_EINVAL, "bad argument%s");
    if (device_info() == nullptr) return MGB_ECUDA;
    k_sumsq<<<1, 1024, 0, (cudaStream_t)stream>>>(n, x_dev, out_dev);
    return cuda_fail(cudaGetLastError(), "vec_sumsq");
}

}  // extern "C"
