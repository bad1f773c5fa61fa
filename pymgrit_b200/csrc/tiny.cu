// tiny.cu -- the same sweeps for the ODE applications (1 or 2 unknowns per time point):
//   Dahlquist   dahlquist/dahlquist.py:88-111    scalar, BE / FE / TR / MR
//   Brusselator brusselator/brusselator.py:105-132   classical RK4, 2 unknowns
// One thread owns one coarse interval and walks it sequentially; rows are 8 or 16 bytes, so there is
// nothing to stage.  Products and sums are rounded separately (no FMA contraction) so that the
// arithmetic is the reference's numpy arithmetic operation by operation.
#include "../../include/mgrit_b200.h"
#include "phi.cuh"
#include "table.h"

namespace mgb {

#define MUL(a, b) __dmul_rn((a), (b))
#define ADD(a, b) __dadd_rn((a), (b))
#define SUB(a, b) __dsub_rn((a), (b))
#define DIV(a, b) __ddiv_rn((a), (b))

struct DahlquistPhi {
    static constexpr int N = 1;
    __device__ static void apply(double (&u)[1], const LevelDev &L, int i) {
        const double dt = SUB(L.t[i], L.t[i - 1]);
        const double z = MUL(dt, L.p[0]);
        switch (L.ip[0]) {
            case MGB_DAHLQUIST_BE: u[0] = MUL(DIV(1.0, SUB(1.0, z)), u[0]); break;
            case MGB_DAHLQUIST_FE: u[0] = MUL(ADD(1.0, z), u[0]); break;
            case MGB_DAHLQUIST_TR: u[0] = MUL(DIV(ADD(1.0, DIV(z, 2.0)), SUB(1.0, DIV(z, 2.0))), u[0]); break;
            default: {  // MR, dahlquist.py:107-109
                const double k1 = MUL(DIV(-1.0, SUB(1.0, DIV(z, 2.0))), u[0]);
                u[0] = ADD(u[0], MUL(dt, k1));
            }
        }
    }
};

struct BrusselatorPhi {
    static constexpr int N = 2;
    __device__ static void f(const double (&y)[2], double (&o)[2]) {  // brusselator.py:68-81, a = 1, b = 3
        const double y0sq = MUL(y[0], y[0]);
        const double q = MUL(y0sq, y[1]);
        o[0] = SUB(ADD(1.0, q), MUL(4.0, y[0]));
        o[1] = SUB(MUL(3.0, y[0]), q);
    }
    __device__ static void apply(double (&u)[2], const LevelDev &L, int i) {
        const double dt = SUB(L.t[i], L.t[i - 1]);
        const double h2 = DIV(dt, 2.0);
        double k1[2], k2[2], k3[2], k4[2], w[2];
        f(u, k1);
        w[0] = ADD(u[0], MUL(h2, k1[0])); w[1] = ADD(u[1], MUL(h2, k1[1]));
        f(w, k2);
        w[0] = ADD(u[0], MUL(h2, k2[0])); w[1] = ADD(u[1], MUL(h2, k2[1]));
        f(w, k3);
        w[0] = ADD(u[0], MUL(dt, k3[0])); w[1] = ADD(u[1], MUL(dt, k3[1]));
        f(w, k4);
        const double h6 = DIV(dt, 6.0);
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const double s = ADD(ADD(ADD(k1[q], MUL(2.0, k2[q])), MUL(2.0, k3[q])), k4[q]);
            u[q] = ADD(u[q], MUL(h6, s));
        }
    }
};

template <class P>
__device__ __forceinline__ void ld(double (&x)[P::N], const double *base, int i, int pitch) {
#pragma unroll
    for (int q = 0; q < P::N; ++q) x[q] = base[(size_t)i * pitch + q];
}
template <class P>
__device__ __forceinline__ void st(const double (&x)[P::N], double *base, int i, int pitch) {
#pragma unroll
    for (int q = 0; q < P::N; ++q) base[(size_t)i * pitch + q] = x[q];
}

template <class P>
__device__ __forceinline__ void advance_tiny(double (&x)[P::N], const LevelDev &L, int i, bool add_g = true) {
    P::apply(x, L, i);
    if (add_g && L.g) {
#pragma unroll
        for (int q = 0; q < P::N; ++q) x[q] = ADD(L.g[(size_t)i * L.pitch + q], x[q]);
    }
}

__device__ __forceinline__ void tiny_interval(const LevelDev &L, int k, int &s, int &e) {
    if (L.cpts == nullptr) {
        s = 0;
        e = L.npts;
    } else {
        s = L.cpts[k];
        e = (k + 1 < L.ncpts) ? L.cpts[k + 1] : L.npts;
    }
}

template <class P>
__global__ void kt_chain(const LevelDev L, int nitems) {
    MGB_RETURN_IF_STOPPED(L)
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < nitems; k += gridDim.x * blockDim.x) {
        int s, e;
        tiny_interval(L, k, s, e);
        if (e - s <= 1) continue;
        double x[P::N];
        ld<P>(x, L.u, s, L.pitch);
        for (int i = s + 1; i < e; ++i) {
            advance_tiny<P>(x, L, i);
            st<P>(x, L.u, i, L.pitch);
        }
    }
}

template <class P>
__global__ void kt_c_relax(const LevelDev L, double w) {
    MGB_RETURN_IF_STOPPED(L)
    for (int k = 1 + blockIdx.x * blockDim.x + threadIdx.x; k < L.ncpts; k += gridDim.x * blockDim.x) {
        // a run of adjacent C-points (non-uniform coarsening) is walked by the thread of its first point, in
        // ascending order like the reference's loop (mgrit.py:356)
        if (k > 1 && L.cpts[k] - L.cpts[k - 1] == 1) continue;
        double x[P::N], old[P::N];
        ld<P>(x, L.u, L.cpts[k] - 1, L.pitch);
        for (int kk = k;; ++kk) {
            const int c = L.cpts[kk];
            ld<P>(old, L.u, c, L.pitch);
            advance_tiny<P>(x, L, c);
            // (.)*w + u*(1-w) is evaluated even for w = 1 by the reference (mgrit.py:360-363)
#pragma unroll
            for (int q = 0; q < P::N; ++q) x[q] = ADD(MUL(x[q], w), MUL(old[q], SUB(1.0, w)));
            st<P>(x, L.u, c, L.pitch);
            if (!(kk + 1 < L.ncpts && L.cpts[kk + 1] == c + 1)) break;
        }
    }
}

template <class P>
__global__ void kt_fas(const LevelDev L, const LevelDev G) {
    MGB_RETURN_IF_STOPPED(L)
    for (int j = 1 + blockIdx.x * blockDim.x + threadIdx.x; j < L.ncpts; j += gridDim.x * blockDim.x) {
        const int c = L.cpts[j];
        double x[P::N], y[P::N], v[P::N];
        ld<P>(x, L.u, c - 1, L.pitch);
        advance_tiny<P>(x, L, c, false);
        ld<P>(y, L.u, c, L.pitch);
        st<P>(y, G.u, j, G.pitch);
#pragma unroll
        for (int q = 0; q < P::N; ++q) {
            if (L.g)
                x[q] = ADD(ADD(SUB(L.g[(size_t)c * L.pitch + q], y[q]), x[q]), y[q]);
            else
                x[q] = ADD(SUB(x[q], y[q]), y[q]);
        }
        ld<P>(v, L.u, L.cpts[j - 1], L.pitch);
        if (j == 1) st<P>(v, G.u, 0, G.pitch);  // point 0 (initial condition / ghost) is injected like any other C-point
        advance_tiny<P>(v, G, j, false);
#pragma unroll
        for (int q = 0; q < P::N; ++q) x[q] = SUB(x[q], v[q]);
        st<P>(x, G.g, j, G.pitch);
    }
}

template <class P>
__global__ void kt_correct(const LevelDev L, const LevelDev G, int frelax, int kfirst) {
    MGB_RETURN_IF_STOPPED(L)
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < L.ncpts; k += gridDim.x * blockDim.x) {
        int s, e;
        tiny_interval(L, k, s, e);
        double x[P::N];
        ld<P>(x, L.u, s, L.pitch);
        if (k >= kfirst) {
#pragma unroll
            for (int q = 0; q < P::N; ++q) x[q] = ADD(x[q], SUB(G.u[(size_t)k * G.pitch + q], x[q]));
            st<P>(x, L.u, s, L.pitch);
        }
        if (frelax)
            for (int i = s + 1; i < e; ++i) {
                advance_tiny<P>(x, L, i);
                st<P>(x, L.u, i, L.pitch);
            }
    }
}

template <class P>
__global__ void kt_residual(const LevelDev L, double *out_sq) {
    MGB_RETURN_IF_STOPPED(L)
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < L.ncpts; k += gridDim.x * blockDim.x) {
        if (k == 0) {
            out_sq[0] = 0.0;
            continue;
        }
        const int c = L.cpts[k];
        double x[P::N], y[P::N];
        ld<P>(x, L.u, c - 1, L.pitch);
        advance_tiny<P>(x, L, c);
        ld<P>(y, L.u, c, L.pitch);
        double acc = 0.0;
#pragma unroll
        for (int q = 0; q < P::N; ++q) {
            const double r = SUB(x[q], y[q]);
            acc = ADD(acc, MUL(r, r));
        }
        out_sq[k] = acc;
    }
}

template <class P>
__global__ void kt_step(const LevelDev L, int point, const double *in, double *out) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        double x[P::N];
#pragma unroll
        for (int q = 0; q < P::N; ++q) x[q] = in[q];
        P::apply(x, L, point);
#pragma unroll
        for (int q = 0; q < P::N; ++q) out[q] = x[q];
    }
}

template <class P>
struct TinyLaunch {
    static int grid(int items) {
        int g = (items + 127) / 128;
        return g < 1 ? 1 : (g > 1184 ? 1184 : g);
    }
    static int f_relax(const LevelDev &L, int /*flags*/, cudaStream_t st) {
        if (device_info() == nullptr) return MGB_ECUDA;
        kt_chain<P><<<grid(L.ncpts), 128, 0, st>>>(L, L.ncpts);
        return cuda_fail(cudaGetLastError(), "f_relax");
    }
    static int forward_solve(const LevelDev &L0, cudaStream_t st) {
        if (device_info() == nullptr) return MGB_ECUDA;
        LevelDev L = L0;
        L.cpts = nullptr;
        L.ncpts = 0;
        kt_chain<P><<<1, 32, 0, st>>>(L, 1);
        return cuda_fail(cudaGetLastError(), "forward_solve");
    }
    static int c_relax(const LevelDev &L, double w, int last_only, cudaStream_t st) {
        if (last_only) return 2;  // MGB_ENOSHAPE: only used ahead of the fused down-sweep, which the ODE kernels do not have
        if (device_info() == nullptr) return MGB_ECUDA;
        kt_c_relax<P><<<grid(L.ncpts), 128, 0, st>>>(L, w);
        return cuda_fail(cudaGetLastError(), "c_relax");
    }
    static int fas(const LevelDev &L, const LevelDev &G, cudaStream_t st) {
        if (device_info() == nullptr) return MGB_ECUDA;
        kt_fas<P><<<grid(L.ncpts), 128, 0, st>>>(L, G);
        return cuda_fail(cudaGetLastError(), "fas_residual");
    }
    static int correct(const LevelDev &L, const LevelDev &G, int fr, int kfirst, cudaStream_t st) {
        if (device_info() == nullptr) return MGB_ECUDA;
        kt_correct<P><<<grid(L.ncpts), 128, 0, st>>>(L, G, fr, kfirst);
        return cuda_fail(cudaGetLastError(), "error_correction");
    }
    static int residual(const LevelDev &L, double *out, cudaStream_t st) {
        if (device_info() == nullptr) return MGB_ECUDA;
        kt_residual<P><<<grid(L.ncpts), 128, 0, st>>>(L, out);
        return cuda_fail(cudaGetLastError(), "residual_norms");
    }
    static int step(const LevelDev &L, int point, const double *in, double *out, cudaStream_t st) {
        if (device_info() == nullptr) return MGB_ECUDA;
        kt_step<P><<<1, 32, 0, st>>>(L, point, in, out);
        return cuda_fail(cudaGetLastError(), "step");
    }
    static const SweepTable *table() {
        static const SweepTable t = {1, P::N, &f_relax, &forward_solve, &c_relax, &fas, &correct, &residual, &step};
        return &t;
    }
};

const SweepTable *mgb_table_tiny(int app) {
    return app == MGB_APP_DAHLQUIST ? TinyLaunch<DahlquistPhi>::table() : TinyLaunch<BrusselatorPhi>::table();
}

}  // namespace mgb
