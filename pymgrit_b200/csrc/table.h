// table.h -- one dispatch table per compiled kernel shape (application, threads per system, chunk).
#pragma once
#include <cuda_runtime.h>

namespace mgb {

struct LevelDev;

struct DeviceInfo {
    int sms;
    int max_smem_optin;
};
const DeviceInfo *device_info();                 // api.cu; nullptr (and an error message) without a CUDA device
int cuda_fail(cudaError_t e, const char *what);  // api.cu: records the message, returns MGB_ECUDA or 0
const int *stop_flag();                          // api.cu: the device flag of mgb_set_stop_flag, or nullptr

// sine_modes.cu: the hot sweeps of a HEAT1D_SINE level with one thread per mode
bool sine_modes_ok(const LevelDev &L);
bool sine_modes_entry(int bit);   // experiments: MGB_SINE_MODES_MASK
bool sine_modes_coarse_ok(const LevelDev &G, const LevelDev &L);
int sine_modes_f_relax(const LevelDev &L, int flags, cudaStream_t st);
int sine_modes_down(const LevelDev &L, const LevelDev &G, cudaStream_t st);
int sine_modes_correct(const LevelDev &L, const LevelDev &G, int frelax, int kfirst, cudaStream_t st);
int sine_modes_residual(const LevelDev &L, double *out_sq, cudaStream_t st);

struct SweepTable {
    int team_threads;
    int chunk;
    int (*f_relax)(const LevelDev &, int flags, cudaStream_t);
    int (*forward_solve)(const LevelDev &, cudaStream_t);
    int (*c_relax)(const LevelDev &, double, int last_only, cudaStream_t);
    int (*fas_residual)(const LevelDev &, const LevelDev &, cudaStream_t);
    int (*correct)(const LevelDev &, const LevelDev &, int, int, cudaStream_t);
    int (*residual)(const LevelDev &, double *, cudaStream_t);
    int (*step)(const LevelDev &, int, const double *, double *, cudaStream_t);
    int (*down)(const LevelDev &, const LevelDev &, cudaStream_t);  // nullptr: not available for this application
    // FAS restriction with a spatial grid transfer, split around the transfer (nullptr: not available)
    int (*residual_rows)(const LevelDev &, double *, cudaStream_t);
    int (*fas_rhs)(const LevelDev &, const double *, const double *, int, cudaStream_t);
    // AT-MGRIT local coarse grids (nullptr: not available)
    int (*window)(const LevelDev &, const double *, int, cudaStream_t);
};

}  // namespace mgb
