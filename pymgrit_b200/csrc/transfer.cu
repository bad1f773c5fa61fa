// transfer.cu -- spatial grid transfer between Heat1D levels, row-wise over all C-points of a level in one launch:
// full-weighting restriction and linear interpolation on the interior points of a 1-D grid with homogeneous Dirichlet
// boundaries (the GridTransfer the reference's users write for examples/example_spatial_coarsening.py:18-79, the only
// non-identity transfer in its tree).  n_fine = 2 n_coarse + 1.  Sums are taken in that example's order.
#include "../../include/mgrit_b200.h"
#include "table.h"

namespace mgb {
int heat2d_fail(const char *msg);  // api.cu: records msg, returns MGB_EINVAL

// dst[j][i] = (src[r_j][2i] / 4 + src[r_j][2i+1] / 2) + src[r_j][2i+2] / 4,  r_j = index ? index[j] : j
__global__ void k_restrict_fw(const int nrows, const double *__restrict__ src, const int src_pitch,
                              const int *__restrict__ index, const int n_coarse, double *__restrict__ dst,
                              const int dst_pitch) {
    for (int j = blockIdx.x; j < nrows; j += gridDim.x) {
        const double *s = src + (size_t)(index ? index[j] : j) * src_pitch;
        double *d = dst + (size_t)j * dst_pitch;
        for (int i = threadIdx.x; i < dst_pitch; i += blockDim.x) {
            double v = 0.0;
            if (i < n_coarse) v = __dadd_rn(__dadd_rn(s[2 * i] * 0.25, s[2 * i + 1] * 0.5), s[2 * i + 2] * 0.25);
            d[i] = v;  // the padding of the row is kept at zero
        }
    }
}

// e = a[j] - b[j] (or a[j]);  E = P e:  E[2i+1] = e[i],  E[2i] = e[i-1] / 2 + e[i] / 2 (missing neighbours = 0);
// dst[r_j] = accumulate ? dst[r_j] + E : E   for j = first .. nrows-1
__global__ void k_interp_linear(const int nrows, const int first, const double *__restrict__ a,
                                const double *__restrict__ b, const int c_pitch, const int n_coarse,
                                double *__restrict__ dst, const int dst_pitch, const int *__restrict__ index,
                                const int accumulate) {
    const int n_fine = 2 * n_coarse + 1;
    for (int j = first + blockIdx.x; j < nrows; j += gridDim.x) {
        const double *pa = a + (size_t)j * c_pitch;
        const double *pb = b ? b + (size_t)j * c_pitch : nullptr;
        double *d = dst + (size_t)(index ? index[j] : j) * dst_pitch;
        for (int x = threadIdx.x; x < n_fine; x += blockDim.x) {
            const int i = x >> 1;
            double v;
            if (x & 1) {
                v = pb ? __dsub_rn(pa[i], pb[i]) : pa[i];
            } else {
                const double lo = (i >= 1) ? (pb ? __dsub_rn(pa[i - 1], pb[i - 1]) : pa[i - 1]) : 0.0;
                const double hi = (i < n_coarse) ? (pb ? __dsub_rn(pa[i], pb[i]) : pa[i]) : 0.0;
                v = __dadd_rn(0.5 * lo, 0.5 * hi);
            }
            d[x] = accumulate ? __dadd_rn(d[x], v) : v;
        }
    }
}

}  // namespace mgb

using namespace mgb;

extern "C" {

int mgb_heat1d_restrict_rows(int32_t nrows, const double *src_dev, int32_t src_pitch, const int32_t *src_index_dev,
                             int32_t n_fine, double *dst_dev, int32_t dst_pitch, void *stream) {
    if (nrows < 0 || src_dev == nullptr || dst_dev == nullptr || n_fine < 3 || n_fine % 2 == 0 || src_pitch < n_fine ||
        dst_pitch < (n_fine - 1) / 2)
        return heat2d_fail("restrict_rows: bad argument (n_fine must be 2 n_coarse + 1)");
    const DeviceInfo *di = device_info();
    if (di == nullptr) return MGB_ECUDA;
    if (nrows == 0) return MGB_OK;
    const int n_coarse = (n_fine - 1) / 2;
    const int threads = dst_pitch >= 256 ? 256 : 32 * ((dst_pitch + 31) / 32);
    const int grid = nrows < 8 * di->sms ? nrows : 8 * di->sms;
    k_restrict_fw<<<grid, threads, 0, (cudaStream_t)stream>>>(nrows, src_dev, src_pitch, src_index_dev, n_coarse, dst_dev,
                                                               dst_pitch);
    return cuda_fail(cudaGetLastError(), "restrict_rows");
}

int mgb_heat1d_interp_rows(int32_t nrows, int32_t first, const double *a_dev, const double *b_dev, int32_t c_pitch,
                           int32_t n_coarse, double *dst_dev, int32_t dst_pitch, const int32_t *dst_index_dev,
                           int32_t accumulate, void *stream) {
    if (nrows < 0 || first < 0 || a_dev == nullptr || dst_dev == nullptr || n_coarse < 1 || c_pitch < n_coarse ||
        dst_pitch < 2 * n_coarse + 1)
        return heat2d_fail("interp_rows: bad argument");
    const DeviceInfo *di = device_info();
    if (di == nullptr) return MGB_ECUDA;
    if (nrows - first <= 0) return MGB_OK;
    const int n_fine = 2 * n_coarse + 1;
    const int threads = n_fine >= 256 ? 256 : 32 * ((n_fine + 31) / 32);
    const int grid = nrows - first < 8 * di->sms ? nrows - first : 8 * di->sms;
    k_interp_linear<<<grid, threads, 0, (cudaStream_t)stream>>>(nrows, first, a_dev, b_dev, c_pitch, n_coarse, dst_dev,
                                                                 dst_pitch, dst_index_dev, accumulate);
    return cuda_fail(cudaGetLastError(), "interp_rows");
}

}  // extern "C"
