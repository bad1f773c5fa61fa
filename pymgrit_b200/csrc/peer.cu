// peer.cu -- the ghost-row exchange between time ranks as two small kernels over peer memory (NVLink / NVSwitch).
//
// The reference sends the last C-point of a rank to its successor as a pickled vector (message kinds 0 and 4,
// mgrit.py:305-310, 510-517, 693-713).  Here every rank owns a "mailbox" in symmetric memory that its predecessor can
// address directly: the sender stores the row into the receiver's mailbox slot over NVLink and then publishes a sequence
// number (system-scope release); the receiver's kernel spins on that number (acquire), copies the row into its ghost row
// and acknowledges.  No host involvement, no rendezvous: ~2 us of latency on top of the stores instead of an NCCL
// send/recv pair.  Two slots per level (sequence parity) and the acknowledgement keep the sender from overwriting a row
// that has not been consumed.
#include "../../include/mgrit_b200.h"
#include "phi.cuh"
#include "table.h"

namespace mgb {

int heat2d_fail(const char *msg);  // api.cu: records the message, returns MGB_EINVAL

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Every wait on a peer is bounded: a rank that died or took another code path must not hang the other GPUs for ever.
// After g_peer_timeout_ns the waiting kernel records which wait gave up in g_peer_error and goes on (its data is then
// meaningless); the host reads the word with mgb_peer_status() when the solve is over and raises.
__device__ unsigned long long g_peer_timeout_ns = 30ull * 1000ull * 1000ull * 1000ull;
__device__ int g_peer_error = 0;

__device__ __forceinline__ unsigned long long now_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// spin until *p >= want (acquire, system scope); false after the timeout
__device__ __forceinline__ bool wait_ge(const unsigned long long *p, unsigned long long want, int kind) {
    if (ld_acquire_sys(p) >= want) return true;
    const unsigned long long t0 = now_ns();
    while (ld_acquire_sys(p) < want) {
        if (now_ns() - t0 > g_peer_timeout_ns) {
            atomicExch(&g_peer_error, kind);
            return false;
        }
    }
    return true;
}

// CTA-wide row copy: 16-byte accesses when both rows are 16-byte aligned (the PDE applications), 8-byte otherwise (the
// one- and two-component ODE rows)
__device__ __forceinline__ void copy_row(const double *__restrict__ src, double *__restrict__ dst, int count) {
    if (((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0) {
        const double2 *s2 = reinterpret_cast<const double2 *>(src);
        double2 *d2 = reinterpret_cast<double2 *>(dst);
        for (int q = threadIdx.x; q < count / 2; q += blockDim.x) d2[q] = s2[q];
        if ((count & 1) && threadIdx.x == 0) dst[count - 1] = src[count - 1];
    } else {
        for (int q = threadIdx.x; q < count; q += blockDim.x) dst[q] = src[q];
    }
}

// src (local) -> dst (peer mailbox slot), then *flag (peer) = seq.  Before touching the slot: wait until the receiver has
// acknowledged sequence number seq - 2 (the previous use of this slot) in *ack (local).
__global__ void __launch_bounds__(256) k_put_row(const double *__restrict__ src, double *__restrict__ dst, int count,
                                                 unsigned long long *flag, const unsigned long long *ack,
                                                 unsigned long long seq) {
    if (threadIdx.x == 0 && seq > 2) wait_ge(ack, seq - 2, MGB_PEER_WAIT_ACK);
    __syncthreads();
    copy_row(src, dst, count);
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) st_release_sys(flag, seq);
}

// wait for *flag (local) == seq, copy the mailbox slot into dst (the ghost row), then *ack (peer, the sender's) = seq
__global__ void __launch_bounds__(256) k_wait_row(const double *__restrict__ slot, double *__restrict__ dst, int count,
                                                  const unsigned long long *flag, unsigned long long *ack,
                                                  unsigned long long seq) {
    if (threadIdx.x == 0) wait_ge(flag, seq, MGB_PEER_WAIT_ROW);
    __syncthreads();
    copy_row(slot, dst, count);
    __syncthreads();
    if (threadIdx.x == 0) st_release_sys(ack, seq);
}

// The same for one-to-many: CTA b stores src into peer b's slot and publishes seq in peer b's flag (after peer b has
// acknowledged seq - 2); the matching wait is CTA b spinning on the flag source b wrote and acknowledging seq - 1, i.e.
// the data of the PREVIOUS round, which every kernel queued before this one has finished reading.
constexpr int kMaxPeers = 16;
struct PeerList {
    unsigned long long a[kMaxPeers];  // put: peer slot address   | wait: my flag address
    unsigned long long b[kMaxPeers];  // put: peer flag address   | wait: peer ack address
    unsigned long long c[kMaxPeers];  // put: my ack address      | wait: unused
    int n;
};

__global__ void __launch_bounds__(256) k_put_rows(const double *__restrict__ src, int count, const PeerList pl,
                                                  unsigned long long seq) {
    const int p = blockIdx.x;
    if (threadIdx.x == 0 && seq > 2)
        wait_ge(reinterpret_cast<const unsigned long long *>(pl.c[p]), seq - 2, MGB_PEER_WAIT_GATHER_ACK);
    __syncthreads();
    copy_row(src, reinterpret_cast<double *>(pl.a[p]), count);
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) st_release_sys(reinterpret_cast<unsigned long long *>(pl.b[p]), seq);
}

__global__ void __launch_bounds__(32) k_wait_flags(const PeerList pl, unsigned long long seq) {
    const int p = blockIdx.x;
    if (threadIdx.x == 0) {
        wait_ge(reinterpret_cast<const unsigned long long *>(pl.a[p]), seq, MGB_PEER_WAIT_GATHER);
        st_release_sys(reinterpret_cast<unsigned long long *>(pl.b[p]), seq - 1);
    }
}

static int fill(PeerList &pl, int n, const uint64_t *a, const uint64_t *b, const uint64_t *c) {
    if (n < 0 || n > kMaxPeers || (n > 0 && (a == nullptr || b == nullptr))) return 1;
    pl.n = n;
    for (int k = 0; k < n; ++k) {
        pl.a[k] = a[k];
        pl.b[k] = b[k];
        pl.c[k] = c ? c[k] : 0;
        if (pl.a[k] == 0 || pl.b[k] == 0) return 1;
    }
    return 0;
}

}  // namespace mgb

using namespace mgb;

extern "C" {

int mgb_peer_status(int32_t *error_out, int32_t reset) {
    if (error_out == nullptr) return heat2d_fail("peer_status: bad argument");
    if (device_info() == nullptr) return MGB_ECUDA;
    int err = 0;
    cudaError_t e = cudaMemcpyFromSymbol(&err, g_peer_error, sizeof(int));
    if (e == cudaSuccess && reset && err != 0) {
        const int zero = 0;
        e = cudaMemcpyToSymbol(g_peer_error, &zero, sizeof(int));
    }
    *error_out = err;
    return cuda_fail(e, "peer_status");
}

int mgb_peer_set_timeout(double seconds) {
    if (!(seconds > 0.0)) return heat2d_fail("peer_set_timeout: bad argument");
    if (device_info() == nullptr) return MGB_ECUDA;
    const unsigned long long ns = (unsigned long long)(seconds * 1e9);
    return cuda_fail(cudaMemcpyToSymbol(g_peer_timeout_ns, &ns, sizeof(ns)), "peer_set_timeout");
}

int mgb_peer_put_row(const double *src_dev, double *peer_slot_dev, int32_t count, void *peer_flag_dev, const void *my_ack_dev,
                     uint64_t seq, void *stream) {
    if (src_dev == nullptr || peer_slot_dev == nullptr || peer_flag_dev == nullptr || my_ack_dev == nullptr || count < 1 ||
        seq < 1)
        return heat2d_fail("peer_put_row: bad argument");
    if (device_info() == nullptr) return MGB_ECUDA;
    k_put_row<<<1, 256, 0, (cudaStream_t)stream>>>(src_dev, peer_slot_dev, count, (unsigned long long *)peer_flag_dev,
                                                   (const unsigned long long *)my_ack_dev, seq);
    return cuda_fail(cudaGetLastError(), "peer_put_row");
}

int mgb_peer_wait_row(const double *my_slot_dev, double *dst_dev, int32_t count, const void *my_flag_dev, void *peer_ack_dev,
                      uint64_t seq, void *stream) {
    if (my_slot_dev == nullptr || dst_dev == nullptr || my_flag_dev == nullptr || peer_ack_dev == nullptr || count < 1 ||
        seq < 1)
        return heat2d_fail("peer_wait_row: bad argument");
    if (device_info() == nullptr) return MGB_ECUDA;
    k_wait_row<<<1, 256, 0, (cudaStream_t)stream>>>(my_slot_dev, dst_dev, count, (const unsigned long long *)my_flag_dev,
                                                    (unsigned long long *)peer_ack_dev, seq);
    return cuda_fail(cudaGetLastError(), "peer_wait_row");
}

int mgb_peer_put_rows(const double *src_dev, int32_t count, int32_t npeers, const uint64_t *peer_slot_ptrs,
                      const uint64_t *peer_flag_ptrs, const uint64_t *my_ack_ptrs, uint64_t seq, void *stream) {
    PeerList pl;
    if (src_dev == nullptr || count < 1 || seq < 1 || my_ack_ptrs == nullptr || fill(pl, npeers, peer_slot_ptrs, peer_flag_ptrs, my_ack_ptrs))
        return heat2d_fail("peer_put_rows: bad argument");
    if (device_info() == nullptr) return MGB_ECUDA;
    if (npeers == 0) return MGB_OK;
    k_put_rows<<<npeers, 256, 0, (cudaStream_t)stream>>>(src_dev, count, pl, seq);
    return cuda_fail(cudaGetLastError(), "peer_put_rows");
}

int mgb_peer_wait_flags(int32_t npeers, const uint64_t *my_flag_ptrs, const uint64_t *peer_ack_ptrs, uint64_t seq,
                        void *stream) {
    PeerList pl;
    if (seq < 1 || fill(pl, npeers, my_flag_ptrs, peer_ack_ptrs, nullptr)) return heat2d_fail("peer_wait_flags: bad argument");
    if (device_info() == nullptr) return MGB_ECUDA;
    if (npeers == 0) return MGB_OK;
    k_wait_flags<<<npeers, 32, 0, (cudaStream_t)stream>>>(pl, seq);
    return cuda_fail(cudaGetLastError(), "peer_wait_flags");
}

}  // extern "C"
