// common.cuh -- device plumbing shared by every sweep kernel: PTX wrappers for mbarrier / bulk (TMA)
// copies, the team (threads that own one spatial system), its constant-ratio scans, and the row pipe
// that streams time-point rows  HBM -> shared memory -> registers -> shared memory -> HBM.
//
// Design: one CTA = one team of T threads = one coarse interval at a time.
// Thread `tid` keeps elements [tid*E, tid*E+E) of the current time point in registers across all the
// Phi applications of the interval.  Rows travel with one cp.async.bulk (UBLKCP) per row into a
// shared-memory slot and are read out with stride-E LDS.64, which is bank-conflict free because E is
// odd; results go back through a slot and one bulk store.  No thread ever issues a strided global
// access, and no row is read from HBM twice.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace mgb {

// Optional cycle probes (scripts/micro/chain_cycles.cu defines MGB_CYCLES): where a team's time goes.
#ifdef MGB_CYCLES
__device__ long long g_cyc[24];
#define MGB_T0 long long _t0 = clock64();
#define MGB_T(k)                                  \
    {                                             \
        const long long _t1 = clock64();          \
        if (threadIdx.x == 0) g_cyc[k] += _t1 - _t0; \
        _t0 = _t1;                                \
    }
#else
#define MGB_T0
#define MGB_T(k)
#endif

// ------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}

__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}

// global -> shared bulk copy (TMA engine, SASS UBLKCP), completion signalled on an mbarrier.
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void *src_gmem, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
                 "l"(src_gmem), "r"(bytes), "r"(bar)
                 : "memory");
}

// shared -> global bulk copy, tracked by the thread's bulk async-group.
__device__ __forceinline__ void bulk_s2g(void *dst_gmem, uint32_t src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(src_smem), "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }

template <int N>
__device__ __forceinline__ void bulk_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}

__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// make generic-proxy writes to shared memory visible to the async proxy (before a bulk store reads them)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ double shfl_up_d(double v, int d) { return __shfl_up_sync(0xffffffffu, v, d); }
__device__ __forceinline__ double shfl_down_d(double v, int d) { return __shfl_down_sync(0xffffffffu, v, d); }
__device__ __forceinline__ double shfl_xor_d(double v, int d) { return __shfl_xor_sync(0xffffffffu, v, d); }
__device__ __forceinline__ double shfl_idx_d(double v, int l) { return __shfl_sync(0xffffffffu, v, l); }

// ------------------------------------------------------------------------------------------------
// Shared-memory layout of one CTA (dynamic):
//   [0, 64)        mbarriers of the input slots (8 x 8 B)
//   [64, 128)      unused
//   [128, 1152)    team scratch: 4 banks of 32 doubles (cross-warp scan / reduce / broadcast)
//   [1152, ...)    row slots, SLOT_BYTES each: NIN input slots, then NOUT output slots
// ------------------------------------------------------------------------------------------------
constexpr int kMaxIn = 8;
constexpr int kHeaderBytes = 1152;

template <int T_, int E_>
struct Shape {
    static constexpr int T = T_;
    static constexpr int E = E_;
    static constexpr int W = T_ / 32;
    static constexpr int SLOT_BYTES = ((T_ * E_ * 8 + 127) / 128) * 128;
    static_assert(T_ % 32 == 0 && T_ <= 1024, "team must be whole warps");
    static_assert(E_ % 2 == 1, "chunk must be odd (bank-conflict-free stride)");
};

// The team: T threads that own one spatial system.  All collectives are called by every thread.
template <int T>
struct Team {
    static constexpr int W = T / 32;
    int tid, lane, warp;
    double *scratch;  // 4 x 32 doubles in shared memory
    int flip;

    __device__ __forceinline__ Team(unsigned char *smem) {
        tid = threadIdx.x;
        lane = tid & 31;
        warp = tid >> 5;
        scratch = reinterpret_cast<double *>(smem + 128);
        flip = 0;
    }

    __device__ __forceinline__ void sync() const {
        if (T == 32)
            __syncwarp();
        else
            __syncthreads();
    }

    // Exclusive forward scan with constant ratio B:  returns sum_{i < tid} B^(tid-1-i) a_i.
    // Bd = {B, B^2, B^4, B^8, B^16}; B32 = B^32; blane = B^lane.
    // Only the first `nsteps` doubling steps are done: the host sets nsteps so that the dropped powers B^(2^k) are
    // below 1e-30 (a lane that far away cannot contribute to any digit of the result).
    __device__ __forceinline__ double scan_fwd(double a, const double (&Bd)[5], double B32, double blane, int nsteps = 5) {
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            if (k < nsteps) {
                const int d = 1 << k;
                const double t = shfl_up_d(a, d);
                if (lane >= d) a = fma(Bd[k], t, a);
            }
        }
        double excl = shfl_up_d(a, 1);
        if (lane == 0) excl = 0.0;
        if (W > 1) {
            double *s = scratch + (flip & 1) * 32;
            flip ^= 1;
            if (lane == 31) s[warp] = a;
            __syncthreads();
            double carry = 0.0;
            for (int v = 0; v < warp; ++v) carry = fma(B32, carry, s[v]);
            excl = fma(blane, carry, excl);
        }
        return excl;
    }

    // Exclusive backward scan:  returns sum_{i > tid} B^(i-tid-1) a_i.   blane = B^(31-lane).
    __device__ __forceinline__ double scan_bwd(double a, const double (&Bd)[5], double B32, double blane, int nsteps = 5) {
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            if (k < nsteps) {
                const int d = 1 << k;
                const double t = shfl_down_d(a, d);
                if (lane + d < 32) a = fma(Bd[k], t, a);
            }
        }
        double excl = shfl_down_d(a, 1);
        if (lane == 31) excl = 0.0;
        if (W > 1) {
            double *s = scratch + (flip & 1) * 32;
            flip ^= 1;
            if (lane == 0) s[warp] = a;
            __syncthreads();
            double carry = 0.0;
            for (int v = W - 1; v > warp; --v) carry = fma(B32, carry, s[v]);
            excl = fma(blane, carry, excl);
        }
        return excl;
    }

    // Value held by thread `src`, for everyone.
    __device__ __forceinline__ double bcast(double v, int src) {
        if (W == 1) return shfl_idx_d(v, src);
        double *s = scratch + 64 + (flip & 1) * 32;
        flip ^= 1;
        if (tid == src) s[0] = v;
        __syncthreads();
        return s[0];
    }

    // Sum over the team in a fixed order (deterministic): xor tree inside a warp, warps in order.
    __device__ __forceinline__ double sum(double v) {
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) v += shfl_xor_d(v, d);
        if (W == 1) return v;
        double *s = scratch + (flip & 1) * 32;
        flip ^= 1;
        if (lane == 0) s[warp] = v;
        __syncthreads();
        double tot = 0.0;
        for (int w = 0; w < W; ++w) tot += s[w];
        return tot;
    }
};

// ------------------------------------------------------------------------------------------------
// Row pipe.  `Gen` yields the addresses of the input rows in the order the kernel pops them; thread 0
// runs it NIN rows ahead of the consumer.  NOUT = 1: a slot is rewritten only after the previous bulk
// store has finished READING it (cp.async.bulk.wait_group.read), which by then is long done.
// ------------------------------------------------------------------------------------------------
template <class SH, class Gen>
struct RowPipe {
    static constexpr int T = SH::T;
    static constexpr int E = SH::E;
    unsigned char *smem;
    uint32_t bar0;       // shared address of mbarrier 0
    uint32_t slot0;      // shared address of input slot 0
    int nin;
    uint32_t row_bytes;  // pitch * 8
    int n;
    int cons;            // rows popped so far
    int cons_slot;
    uint32_t cons_parity;
    Gen gen;             // meaningful in thread 0 only
    int tid;

    __device__ __forceinline__ RowPipe(unsigned char *smem_, int nin_, int pitch, int n_, const Gen &g)
        : smem(smem_), nin(nin_), row_bytes(pitch * 8u), n(n_), cons(0), cons_slot(0), cons_parity(0), gen(g) {
        tid = threadIdx.x;
        bar0 = smem_u32(smem);
        slot0 = smem_u32(smem + kHeaderBytes);
    }

    __device__ __forceinline__ double *in_slot(int s) const {
        return reinterpret_cast<double *>(smem + kHeaderBytes + (size_t)s * SH::SLOT_BYTES);
    }
    __device__ __forceinline__ double *out_slot() const {
        return reinterpret_cast<double *>(smem + kHeaderBytes + (size_t)nin * SH::SLOT_BYTES);
    }

    __device__ __forceinline__ void issue(int s) {
        const double *p;
        if (gen.next(p)) {
            const uint32_t bar = bar0 + 8u * s;
            mbar_arrive_expect_tx(bar, row_bytes);
            bulk_g2s(slot0 + (uint32_t)s * SH::SLOT_BYTES, p, row_bytes, bar);
        }
    }

    // Called by every thread once, before the first pop.
    template <class TeamT>
    __device__ __forceinline__ void start(TeamT &team) {
        if (tid == 0) {
            for (int s = 0; s < nin; ++s) mbar_init(bar0 + 8u * s, 1);
            mbar_fence_init();
        }
        team.sync();
        if (tid == 0)
            for (int s = 0; s < nin; ++s) issue(s);
    }

    // Next input row -> registers (elements beyond n read as 0).
    template <class TeamT>
    __device__ __forceinline__ void pop(double (&x)[E], TeamT &team) {
        MGB_T0
        mbar_wait(bar0 + 8u * cons_slot, cons_parity);
        MGB_T(0)
        const double *s = in_slot(cons_slot) + tid * E;
        const int nv = n - tid * E;
#pragma unroll
        for (int j = 0; j < E; ++j) {
            const double v = s[j];
            x[j] = (j < nv) ? v : 0.0;
        }
        team.sync();
        MGB_T(1)
        if (tid == 0) issue(cons_slot);
        MGB_T(2)
        ++cons;
        if (++cons_slot == nin) {
            cons_slot = 0;
            cons_parity ^= 1u;
        }
    }

    // Next input row left in its shared-memory slot: the caller reads its own elements (s[tid*E + j], elements beyond n
    // are whatever the row padding holds) and then calls pop_release().
    __device__ __forceinline__ const double *pop_ptr() {
        mbar_wait(bar0 + 8u * cons_slot, cons_parity);
        return in_slot(cons_slot);
    }
    template <class TeamT>
    __device__ __forceinline__ void pop_release(TeamT &team) {
        team.sync();
        if (tid == 0) issue(cons_slot);
        ++cons;
        if (++cons_slot == nin) {
            cons_slot = 0;
            cons_parity ^= 1u;
        }
    }

    // A team-private scratch row behind the output slot (kernels that ask for it size the shared memory accordingly).
    __device__ __forceinline__ double *stash_slot() const {
        return reinterpret_cast<double *>(smem + kHeaderBytes + (size_t)(nin + 1) * SH::SLOT_BYTES);
    }

    // The row that the last push() left in the output slot, stored once more (another destination).
    __device__ __forceinline__ void push_again(double *dst) {
        if (tid == 0) {
            bulk_s2g(dst, smem_u32(out_slot()), row_bytes);
            bulk_commit();
        }
    }

    // Registers -> output row in HBM.
    template <class TeamT>
    __device__ __forceinline__ void push(const double (&x)[E], double *dst, TeamT &team) {
        MGB_T0
        if (tid == 0) bulk_wait_read<0>();
        team.sync();
        MGB_T(3)
        double *s = out_slot() + tid * E;
#pragma unroll
        for (int j = 0; j < E; ++j) s[j] = x[j];
        MGB_T(4)
        fence_proxy_async();
        team.sync();
        MGB_T(5)
        if (tid == 0) {
            bulk_s2g(dst, smem_u32(out_slot()), row_bytes);
            bulk_commit();
        }
        MGB_T(6)
    }

    template <class TeamT>
    __device__ __forceinline__ void finish(TeamT &team) {
        if (tid == 0) bulk_wait_all();
        team.sync();
    }
};

}  // namespace mgb
