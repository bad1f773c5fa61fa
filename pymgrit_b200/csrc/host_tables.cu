// host_tables.cu -- the two O(nt) passes of the setup that are pure arithmetic on host arrays, as native multi-threaded
// code: with 4 ms of device work per solve the constructor's host time is half of the end-to-end time, and NumPy pieces
// on Python threads fight for the interpreter lock with the thread that is building the other levels
// (profiles/r02t_e2e_breakdown_n1.txt).  Called through ctypes, which drops the lock for the duration of the call.
// No device code here; results are bit-identical to the NumPy expressions they replace (one IEEE operation per element).
#include <algorithm>
#include <cstdint>
#include <thread>
#include <vector>

#include "../../include/mgrit_b200.h"

namespace {

template <class F>
void split_range(int64_t n, int threads, F fn) {
    if (threads < 1) threads = 1;
    const int64_t min_piece = 1 << 15;
    int nt = (int)std::min<int64_t>(threads, std::max<int64_t>(1, n / min_piece));
    if (nt <= 1) {
        fn(0, (int64_t)0, n);
        return;
    }
    std::vector<std::thread> pool;
    pool.reserve(nt - 1);
    const int64_t per = (n + nt - 1) / nt;
    for (int k = 1; k < nt; ++k) {
        const int64_t a = std::min(n, k * per), b = std::min(n, a + per);
        try {
            pool.emplace_back([=] { fn(k, a, b); });
        } catch (...) {   // no thread to be had (resource limits): the piece runs here -- nothing may cross the C ABI
            fn(k, a, b);
        }
    }
    fn(0, (int64_t)0, std::min(n, per));
    for (auto &th : pool) th.join();
}

}  // namespace

extern "C" {

// dt[0] = 0, dt[i] = t[i] - t[i-1]; *lo, *hi = smallest / largest step (0, 0 for n < 2)      (core/device_level.py time_steps)
int mgb_host_time_steps(const double *t, int64_t n, double *dt, double *lo, double *hi, int32_t threads) {
    if (t == nullptr || dt == nullptr || lo == nullptr || hi == nullptr || n < 0) return MGB_EINVAL;
    *lo = *hi = 0.0;
    if (n == 0) return MGB_OK;
    dt[0] = 0.0;
    if (n == 1) return MGB_OK;
    const int cap = 64;
    double los[cap], his[cap];
    if (threads > cap) threads = cap;
    for (int k = 0; k < cap; ++k) {
        los[k] = 1e300;
        his[k] = -1e300;
    }
    split_range(n - 1, threads, [&](int k, int64_t a, int64_t b) {  // steps into points a+1 .. b
        // blocks that stay in L1: a subtraction loop the compiler vectorises, then min / max with independent accumulators
        double l[4] = {1e300, 1e300, 1e300, 1e300}, h[4] = {-1e300, -1e300, -1e300, -1e300};
        for (int64_t i0 = a + 1; i0 <= b; i0 += 2048) {
            const int64_t i1 = std::min<int64_t>(b + 1, i0 + 2048);
            const double *__restrict__ tp = t;
            double *__restrict__ dp = dt;
            for (int64_t i = i0; i < i1; ++i) dp[i] = tp[i] - tp[i - 1];
            int64_t i = i0;
            for (; i + 4 <= i1; i += 4) {
                for (int j = 0; j < 4; ++j) {
                    const double d = dp[i + j];
                    l[j] = d < l[j] ? d : l[j];
                    h[j] = d > h[j] ? d : h[j];
                }
            }
            for (; i < i1; ++i) {
                const double d = dp[i];
                l[0] = d < l[0] ? d : l[0];
                h[0] = d > h[0] ? d : h[0];
            }
        }
        los[k] = std::min(std::min(l[0], l[1]), std::min(l[2], l[3]));
        his[k] = std::max(std::max(h[0], h[1]), std::max(h[2], h[3]));
    });
    double l = 1e300, h = -1e300;
    for (int k = 0; k < cap; ++k) {
        l = std::min(l, los[k]);
        h = std::max(h, his[k]);
    }
    *lo = l;
    *hi = h;
    return MGB_OK;
}

// out[i] = i * step + start, i < n (two roundings per element, as NumPy's linspace computes arange(n) * step + start:
// a multiplication pass and an addition pass per block, so that no compiler contracts them into one fused operation)
int mgb_host_affine_ramp(double start, double step, int64_t n, double *out, int32_t threads) {
    if (out == nullptr || n < 0) return MGB_EINVAL;
    split_range(n, threads, [&](int, int64_t a, int64_t b) {
        double *__restrict__ o = out;
        for (int64_t i0 = a; i0 < b; i0 += 2048) {
            const int64_t i1 = std::min<int64_t>(b, i0 + 2048);
            for (int64_t i = i0; i < i1; ++i) o[i] = (double)i * step;
            volatile double s = start;            // the addition is a separate, visible operation
            const double sv = s;
            for (int64_t i = i0; i < i1; ++i) o[i] = o[i] + sv;
        }
    });
    return MGB_OK;
}

// out[i][k] = src[k][i] (* scale[i]), i < n, k < q: the time factors of a separable right-hand side, [q][n] as they were
// evaluated -> [n][q] as the kernels read them (core/rhs_tables.py RhsSplit.coefficients)
int mgb_host_scale_rows(const double *src, int64_t ld_src, int32_t q, int64_t n, const double *scale, double *out,
                        int32_t threads) {
    if (src == nullptr || out == nullptr || q < 1 || n < 0 || ld_src < n) return MGB_EINVAL;
    split_range(n, threads, [&](int, int64_t a, int64_t b) {
        if (q == 1) {
            const double *__restrict__ s0 = src;
            double *__restrict__ o = out;
            if (scale) {
                const double *__restrict__ sc = scale;
                for (int64_t i = a; i < b; ++i) o[i] = s0[i] * sc[i];
            } else {
                for (int64_t i = a; i < b; ++i) o[i] = s0[i];
            }
            return;
        }
        for (int64_t i = a; i < b; ++i) {
            const double s = scale ? scale[i] : 1.0;
            for (int k = 0; k < q; ++k) out[i * q + k] = scale ? src[(int64_t)k * ld_src + i] * s : src[(int64_t)k * ld_src + i];
        }
    });
    return MGB_OK;
}

}  // extern "C"
