// launch.cuh -- host-side launchers for the team kernels of sweeps.cuh, one table per (Phi, T, E).
#pragma once
#include <cstdlib>
#include <mutex>

#include "sweeps.cuh"
#include "table.h"

namespace mgb {

template <class Phi>
struct Launch {
    using SH = typename Phi::SH;

    static size_t smem_bytes(int nin) { return kHeaderBytes + (size_t)(nin + 1) * SH::SLOT_BYTES; }

    // persistent grid: as many CTAs as fit on the device at once, but no more than there are items.  The shared-memory
    // opt-in and the occupancy query are done once per (kernel, slot count, device) and remembered: on the coarse levels
    // a sweep runs for ~10 us, so two runtime calls per launch are not free.
    template <class K>
    static int grid_for(K kernel, int nitems, int nin, int *grid) {
        const size_t smem = smem_bytes(nin);
        const DeviceInfo *di = device_info();
        if (di == nullptr) return 3;
        if ((int)smem > di->max_smem_optin) return 2;
        struct Entry {
            const void *fn;
            int nin, dev, per_sm;
            size_t smem;
        };
        static Entry cache[64];
        static int ncache = 0;
        static std::mutex mu;
        int dev = 0;
        cudaGetDevice(&dev);
        int per_sm = -1;
        size_t opted = 0;  // largest dynamic shared memory this kernel has been opted in for (the limit only ever grows)
        {
            std::lock_guard<std::mutex> lock(mu);
            for (int q = 0; q < ncache; ++q)
                if (cache[q].fn == (const void *)kernel && cache[q].dev == dev) {
                    if (cache[q].nin == nin) per_sm = cache[q].per_sm;
                    if (cache[q].smem > opted) opted = cache[q].smem;
                }
        }
        if (per_sm < 0) {
            cudaError_t e = cudaSuccess;
            if (smem > opted) e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute");
            per_sm = 0;
            e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, Phi::T, smem);
            if (e != cudaSuccess) return cuda_fail(e, "cudaOccupancyMaxActiveBlocksPerMultiprocessor");
            if (per_sm < 1) return 2;
            std::lock_guard<std::mutex> lock(mu);
            if (ncache < 64) cache[ncache++] = Entry{(const void *)kernel, nin, dev, per_sm, smem};
        }
        const long cap = (long)per_sm * di->sms;
        *grid = (int)(nitems < cap ? nitems : cap);
        if (*grid < 1) *grid = 1;
        return 0;
    }

    static int rows_extra(const LevelDev &L) { return (L.g ? 1 : 0) + (L.rhs_dense ? 1 : 0); }
    // Input slots of a sweep that pops `base` rows per item plus a row per step on levels with g / dense right-hand-side
    // rows.  (Measured: three more slots for those per-step rows do not help -- level 1 of the headline workload got 20 %
    // slower, profiles/r02i_timeline.txt: the chains there wait for the dependent loads of Phi, not for the bulk copies.)
    static int slots(const LevelDev &L, int base) { return base + rows_extra(L); }
    static int nsys(const LevelDev &L) { return L.nsys > 1 ? L.nsys : 1; }

    // flags & 1: store only the last point of every interval (the other F-points are dead in a down-sweep)
    static int f_relax(const LevelDev &L, int flags, cudaStream_t st) {
        if (L.ncpts < 1) return 0;
        const int nin = slots(L, 2);
        const int nw = L.ncpts * nsys(L);
        int grid;
        if (int rc = grid_for(k_chain<Phi>, nw, nin, &grid)) return rc;
        k_chain<Phi><<<grid, Phi::T, smem_bytes(nin), st>>>(L, nw, flags & 1, nin);
        return cuda_fail(cudaGetLastError(), "f_relax");
    }

    static int forward_solve(const LevelDev &L0, cudaStream_t st) {
        LevelDev L = L0;
        L.cpts = nullptr;  // one interval [0, npts)
        L.ncpts = 0;
        if (L.npts < 2) return 0;
        const int nin = 4, nw = nsys(L);
        int grid;
        if (int rc = grid_for(k_chain<Phi>, nw, nin, &grid)) return rc;
        k_chain<Phi><<<grid, Phi::T, smem_bytes(nin), st>>>(L, nw, 0, nin);
        return cuda_fail(cudaGetLastError(), "forward_solve");
    }

    // last_only: relax the last C-point of the level only
    static int c_relax(const LevelDev &L, double w, int last_only, cudaStream_t st) {
        if (L.ncpts < 2) return 0;
        const int kbase = last_only ? L.ncpts - 1 : 1;
        const int nin = 3, nw = (L.ncpts - kbase) * nsys(L);
        int grid;
        if (int rc = grid_for(k_c_relax<Phi>, nw, nin, &grid)) return rc;
        k_c_relax<Phi><<<grid, Phi::T, smem_bytes(nin), st>>>(L, w, nw, nin, kbase);
        return cuda_fail(cudaGetLastError(), "c_relax");
    }

    static int fas_residual(const LevelDev &L, const LevelDev &G, cudaStream_t st) {
        if (L.ncpts < 2) return 0;
        const int nin = 4, nw = (L.ncpts - 1) * nsys(L);
        int grid;
        if (int rc = grid_for(k_fas_residual<Phi>, nw, nin, &grid)) return rc;
        k_fas_residual<Phi><<<grid, Phi::T, smem_bytes(nin), st>>>(L, G, nw, nin);
        return cuda_fail(cudaGetLastError(), "fas_residual");
    }

    // C-relaxation + F-relaxation + FAS restriction in one pass (k_down); one more slot: the stash
    static int down(const LevelDev &L, const LevelDev &G, cudaStream_t st) {
        if (L.ncpts < 2) return 0;
        const int nin = slots(L, 1), nw = (L.ncpts - 1) * nsys(L);
        int grid;
        if (int rc = grid_for(k_down<Phi>, nw, nin + 1, &grid)) return rc;
        k_down<Phi><<<grid, Phi::T, smem_bytes(nin + 1), st>>>(L, G, nw, nin);
        return cuda_fail(cudaGetLastError(), "down_sweep");
    }

    static int correct(const LevelDev &L, const LevelDev &G, int frelax, int kfirst, cudaStream_t st) {
        if (L.ncpts < 1) return 0;
        // every F-point stored: the sweep is a stream of rows, three input slots keep the loads ahead; last point only: two
        // rows in per interval, fewer slots leave room for more resident teams
        const int nin = slots(L, frelax == 2 ? 2 : 3), nw = L.ncpts * nsys(L);
        int grid;
        if (int rc = grid_for(k_correct<Phi>, nw, nin, &grid)) return rc;
        k_correct<Phi><<<grid, Phi::T, smem_bytes(nin), st>>>(L, G, frelax, kfirst, nw, nin);
        return cuda_fail(cudaGetLastError(), "error_correction");
    }

    // out_sq holds ncpts doubles; with several systems per row, ncpts * (1 + nsys) (partials behind the results)
    static int residual(const LevelDev &L, double *out_sq, cudaStream_t st) {
        if (L.ncpts < 1) return 0;
        const int nin = 3, nw = (L.ncpts > 1 ? L.ncpts - 1 : 0) * nsys(L);
        int grid;
        if (int rc = grid_for(k_residual<Phi>, nw > 0 ? nw : 1, nin, &grid)) return rc;
        k_residual<Phi><<<grid, Phi::T, smem_bytes(nin), st>>>(L, out_sq, nw, nin);
        if (nsys(L) > 1 && L.ncpts > 1) k_sum_systems<Phi><<<(L.ncpts + 127) / 128, 128, 0, st>>>(out_sq, L.ncpts, L.nsys);
        return cuda_fail(cudaGetLastError(), "residual_norms");
    }

    // in / out: rows in the level's layout
    static int step(const LevelDev &L, int point, const double *in, double *out, cudaStream_t st) {
        const int nin = 2, nw = nsys(L);
        int grid;
        if (int rc = grid_for(k_step<Phi>, nw, nin, &grid)) return rc;
        k_step<Phi><<<grid, Phi::T, smem_bytes(nin), st>>>(L, point, in, out, nw, nin);
        return cuda_fail(cudaGetLastError(), "step");
    }

    // out: [ncpts][pitch] rows, row j = fine residual at C-point j (row 0 untouched)
    static int residual_rows(const LevelDev &L, double *out, cudaStream_t st) {
        if (L.ncpts < 2) return 0;
        const int nin = 3, nw = (L.ncpts - 1) * nsys(L);
        int grid;
        if (int rc = grid_for(k_residual_rows<Phi>, nw, nin, &grid)) return rc;
        k_residual_rows<Phi><<<grid, Phi::T, smem_bytes(nin), st>>>(L, out, nw, nin);
        return cuda_fail(cudaGetLastError(), "residual_rows");
    }

    // G.g[j] = (RR[j] + V[j]) - Phi_c(V[j-1]), j = 1 .. nrows-1
    static int fas_rhs(const LevelDev &G, const double *V, const double *RR, int nrows, cudaStream_t st) {
        if (nrows < 2) return 0;
        const int nin = 4, nw = (nrows - 1) * nsys(G);
        int grid;
        if (int rc = grid_for(k_fas_rhs<Phi>, nw, nin, &grid)) return rc;
        k_fas_rhs<Phi><<<grid, Phi::T, smem_bytes(nin), st>>>(G, V, RR, nw, nin);
        return cuda_fail(cudaGetLastError(), "fas_coarse_rhs");
    }

    // u[p] = end of the chain of at most k-1 steps from old[max(0, p-k+1)], p = 1 .. npts-1
    static int window(const LevelDev &L0, const double *old, int k, cudaStream_t st) {
        LevelDev L = L0;
        L.cpts = nullptr;
        L.ncpts = 0;
        if (L.npts < 2) return 0;
        const int nin = slots(L, 2), nw = (L.npts - 1) * nsys(L);
        int grid;
        if (int rc = grid_for(k_window<Phi>, nw, nin, &grid)) return rc;
        k_window<Phi><<<grid, Phi::T, smem_bytes(nin), st>>>(L, old, k, nw, nin);
        return cuda_fail(cudaGetLastError(), "local_coarse_solve");
    }

    static const SweepTable *table() {
        static const SweepTable t = {Phi::T,   Phi::E, &f_relax, &forward_solve,  &c_relax, &fas_residual, &correct,
                                     &residual, &step,  &down,    &residual_rows, &fas_rhs,  &window};
        return &t;
    }
};

}  // namespace mgb
