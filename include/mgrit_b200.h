/*
 * mgrit_b200.h -- C ABI of libmgrit_b200.so: the batched MGRIT sweeps of the pymgrit hot path as
 * hand-written sm_100a CUDA kernels.
 *
 * The reference (pymgrit v1.0.6, /root/reference) is pure Python: the interface this library sits
 * behind is not an FFI but the Python plugin API  Application.step / Vector / Mgrit  (SURVEY.md 8b).
 * Every entry point below replaces one Python loop of src/pymgrit/core/mgrit.py that calls
 * Application.step once per time point; the citation on each declaration names that loop.  The
 * Python host package pymgrit_b200 binds these symbols with ctypes (pymgrit_b200/_lib.py); the
 * binding a reference maintainer would add is shown in INTEGRATION.md.
 *
 * Conventions
 *   - plain C types only; all pointers named *_dev are CUDA device pointers owned by the caller
 *     (the Python host allocates them as torch tensors), everything else is host memory;
 *   - every level of the time-grid hierarchy is one row-major array u[npts][pitch] of doubles, one
 *     row per time point (row = the reference's Vector.get_values()), pitch >= n, pitch even, so
 *     that rows are 16-byte aligned for bulk (TMA) copies;  g has the same shape (FAS right-hand
 *     side, levels > 0 only).  The FAS copy v of the reference (mgrit.py:520) is never stored: with
 *     injection it is bit-identical to the fine level's C-point rows while it is live;
 *   - point 0 of a level is never written by a sweep: it is the initial condition on time rank 0
 *     and the ghost row received from the previous rank otherwise (mgrit.py:781-790);
 *   - cpts_dev lists the C-points of the level in ascending order and always starts with 0;
 *   - all launches are asynchronous on `stream` (a cudaStream_t passed as void*); no allocation,
 *     no synchronisation, no host<->device copy happens inside a sweep;
 *   - return value 0 = ok, otherwise an MGB_E* code; mgb_last_error() gives the message.  Nothing
 *     throws across the ABI.  There is no CPU fallback: without a CUDA device every launch fails.
 */
#ifndef MGRIT_B200_H
#define MGRIT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MGB_ABI_VERSION 10

/* application kinds (which Phi) */
#define MGB_APP_HEAT1D 1      /* heat/heat_1d.py:198-217   backward Euler, Toeplitz tridiagonal solve      */
#define MGB_APP_ADVECTION1D 2 /* advection/advection_1d.py:129-143   implicit upwind, cyclic bidiagonal     */
#define MGB_APP_DAHLQUIST 3   /* dahlquist/dahlquist.py:88-111       scalar BE / FE / TR / MR                */
#define MGB_APP_BRUSSELATOR 4 /* brusselator/brusselator.py:105-132  classical RK4 on 2 components           */
#define MGB_APP_HEAT2D 5      /* heat/heat_2d.py:322-366 (BE branch) 5-point Laplacian, Dirichlet data;     */
                              /* rows are kept in sine space (see mgb_heat2d_to_rows)                         */
#define MGB_APP_HEAT1D_2PTS 6 /* heat/heat_1d_2pts_bdf1.py:90-117, heat_1d_2pts_bdf2.py:92-138: a time point is    */
                              /* the pair (u(t), u(t + dtau)), Phi = two Toeplitz tridiagonal solves; BDF1 or     */
                              /* BDF2 is a property of the level's step constants (see "two-point rows" below)    */

#define MGB_APP_HEAT1D_SINE 7 /* heat/heat_1d.py:198-217 with the level rows kept in sine space (rows hold u S, S the   */
                              /* orthonormal sine matrix that diagonalises the Toeplitz operator): Phi is one FMA and  */
                              /* one multiplication per unknown; values are transformed where they enter or leave     */
                              /* (mgb_rows_dst / mgb_rows_gemm).  See "Heat1D in sine space" below                    */

/* error codes */
#define MGB_OK 0
#define MGB_EINVAL 1     /* bad argument                                     */
#define MGB_ENOSHAPE 2   /* no kernel instantiation for this problem size    */
#define MGB_ECUDA 3      /* CUDA runtime error (launch failed, no device...) */

/* temporal norm of the convergence criterion (mgrit.py:182) */
#define MGB_TNORM_ONE 1
#define MGB_TNORM_TWO 2
#define MGB_TNORM_INF 3

/* Dahlquist methods (dahlquist.py:98-110) */
#define MGB_DAHLQUIST_BE 0
#define MGB_DAHLQUIST_FE 1
#define MGB_DAHLQUIST_TR 2
#define MGB_DAHLQUIST_MR 3

/*
 * One level of the hierarchy as seen by one time rank.  Mirrors the per-level state the reference
 * keeps in Mgrit.u/g/t/index_local_c (mgrit.py:150-172, 840-858) plus what its Application object
 * holds for Phi (heat_1d.py:149-175 etc.).
 */
typedef struct mgb_level {
    int32_t app;            /* MGB_APP_*                                                          */
    int32_t n;              /* spatial dofs per time point                                        */
    int32_t pitch;          /* row pitch in doubles                                               */
    int32_t npts;           /* time points held (incl. point 0 = initial condition / ghost)       */
    double *u_dev;          /* [npts][pitch] solution                                             */
    double *g_dev;          /* [npts][pitch] FAS right-hand side, NULL on level 0                 */
    const int32_t *cpts_dev;/* [ncpts] C-point indices, ascending, cpts[0] == 0; NULL on coarsest */
    int32_t ncpts;
    int32_t team_threads;   /* kernel shape for this n: threads per system ...                    */
    int32_t chunk;          /* ... and contiguous elements per thread (mgb_team_shape)            */

    /* Phi data.  The step that produces point i uses dt_i = t[i] - t[i-1]. */
    int32_t ndt;            /* distinct dt values on this level (1 for a uniform grid)            */
    int32_t cw;             /* doubles per row of the step-constant table (mgb_step_consts_width) */
    const int32_t *dtidx_dev; /* [npts] row of sconst for the step into point i; NULL if ndt == 1 */
    const double *sconst_dev; /* [ndt][cw] per-dt constants of Phi (mgb_*_step_consts)            */

    /* right-hand side of the PDE, b(x, t_i) * dt_i = sum_k rhs_t[i][k] * rhs_x[k][x]  (heat only) */
    int32_t nrhs;           /* number of separable terms; 0 = homogeneous                         */
    int32_t nsys;           /* independent systems per row: 0/1 for the 1-D applications; HEAT2D: */
                            /* tiles of pitch/nsys doubles (mgb_heat2d_layout)                    */
    const double *rhs_x_dev;/* [nrhs][chunk][team_threads]: spatial factors, thread-transposed    */
                            /* (HEAT2D: [nrhs][pitch] in row layout, at most 3 terms)             */
    const double *rhs_t_dev;/* [npts][nrhs]: time factors already multiplied by dt_i              */
    const double *rhs_dense_dev; /* [npts][pitch] b(x,t_i)*dt_i for non-separable data, or NULL   */

    const double *t_dev;    /* [npts] time values of the level's points (needed by the ODE apps)  */
    double p[8];            /* application scalars (dahlquist: p[0] = lambda)                     */
    int32_t ip[4];          /* application integers (dahlquist: ip[0] = method; heat2d: ip[0] =   */
                            /* first tile of boundary nodes, ip[1] = 1 for backward Euler: lets   */
                            /* the one-thread-per-mode sweeps drop the explicit factor)           */
    const double *sig_dev;  /* HEAT2D: [pitch] fx*lambda_k + fy*mu_l per sine coefficient, the    */
                            /* Dirichlet value per boundary node, 0 in the padding; else NULL     */
    const double *diag_dev; /* HEAT1D_SINE: [2][chunk][team_threads] thread-transposed: the       */
                            /* eigenvalues lam_k of (a/dx^2) tridiag(-1,2,-1), then 1/(1+dt lam_k)*/
                            /* for the level's dt (used when ndt == 1); else NULL                 */
    const double *nat_dev;  /* HEAT1D_SINE, optional: [2 + nrhs][pitch] in natural mode order: lam_k,     */
                            /* 1/(1+dt lam_k), X_q S.  With it mgb_f_relax, mgb_down_sweep,               */
                            /* mgb_error_correction and mgb_residual_norms run as one thread per mode     */
                            /* (csrc/sine_modes.cu: no shared-memory staging, coalesced rows, same values)*/
} mgb_level;

int mgb_abi_version(void);
const char *mgb_last_error(void);

/* Shape of the kernel instantiation used for n dofs: threads per system and elements per thread.
 * Returns MGB_ENOSHAPE if n is larger than the largest compiled shape. */
int mgb_team_shape(int32_t app, int32_t n, int32_t *team_threads, int32_t *chunk);

/* Width (doubles) of one row of the step-constant table for a shape. */
int mgb_step_consts_width(int32_t app, int32_t team_threads, int32_t chunk);

/* Host helpers: fill one row of the step-constant table.
 *   heat1d:      r  = dt * a / dx^2   (heat_1d.py:185, 213)
 *   advection1d: nu = dt * c / dx     (advection_1d.py:108, 140) */
int mgb_heat1d_step_consts(double r, int32_t n, int32_t team_threads, int32_t chunk, double *out);
int mgb_advection1d_step_consts(double nu, int32_t n, int32_t team_threads, int32_t chunk, double *out);

/* Two-point rows (MGB_APP_HEAT1D_2PTS).  n = unknowns of ONE of the two time points, chunk = 2 h + 1 with h the elements
 * of each time point a thread owns, pitch = team_threads * chunk and a row is laid out per thread:
 *     row[tid * chunk + j] = first[tid * h + j],  row[tid * chunk + h + j] = second[tid * h + j]  (j < h),  one 0.
 * Phi(first, second) = (tmp1, tmp2),  tmp1 = S(r1) (a1 first + b1 second + rhs1),  tmp2 = S(r2) (a2 second + b2 tmp1 + rhs2),
 * S(r) = (I + r tridiag(-1, 2, -1))^-1.  One step-constant row (width mgb_step_consts_width) is
 *     [mgb_heat1d_step_consts(r1, n, team_threads, h)] [mgb_heat1d_step_consts(r2, ...)] [a1 b1 a2 b2 skip1 0 0 0]
 * (skip1 != 0: r1 = 0, the first solve is the identity and its constants are ignored), rhs_x_dev is
 * [nrhs][h][team_threads] and rhs_t_dev is [npts][2 nrhs]: the time factors of rhs1, then those of rhs2.
 * mgb_heat1d_2pts_half_width returns the width of one of the two Heat1D blocks. */
int mgb_heat1d_2pts_half_width(int32_t team_threads, int32_t chunk);

/* ---- sweeps (one launch covers every coarse interval of the level) -------------------------- */

/* F-relaxation, mgrit.py:292-333: for every C-point c and every F-point i that follows it,
 * u[i] = (g[i] +) Phi(u[i-1]).  flags:
 *   MGB_F_RELAX_LAST_ONLY  store only the last F-point of every interval.  In the down-sweep of a cycle
 *                          (mgrit.py:274-281) nothing reads the other F-points before the F-relaxation after the
 *                          coarse-grid correction (mgrit.py:287) overwrites them, so their stores are skipped;
 *                          the values that are stored are bit-identical. */
#define MGB_F_RELAX_LAST_ONLY 1
int mgb_f_relax(const mgb_level *lvl, int32_t flags, void *stream);

/* C-relaxation, mgrit.py:335-370: for every C-point c != 0,
 * u[c] = w * ((g[c] +) Phi(u[c-1])) + (1 - w) * u[c]. */
int mgb_c_relax(const mgb_level *lvl, double weight, void *stream);
/* The same for the last C-point of the level only.  A time rank calls it ahead of mgb_down_sweep so that its last row
 * -- the next rank's ghost C-point, reference message kind 0 (mgrit.py:305-310) -- can travel before the fused pass
 * (which recomputes the same value). */
int mgb_c_relax_last(const mgb_level *lvl, double weight, void *stream);

/* FAS restriction, mgrit.py:488-549 with GridTransferCopy (grid_transfer_copy.py:25-47): for every
 * C-point j, coarse.u[j] = fine.u[c_j]; for j >= 1
 *   coarse.g[j] = (Phi_f(fine.u[c_j - 1]) - fine.u[c_j] (+ fine.g[c_j])) + fine.u[c_j]
 *                 - Phi_c(fine.u[c_{j-1}]). */
int mgb_fas_residual(const mgb_level *fine, const mgb_level *coarse, void *stream);

/* The down-sweep of a cycle in one pass, mgrit.py:277-281 + 488-549: exactly mgb_c_relax(fine, 1.0), then
 * mgb_f_relax(fine, MGB_F_RELAX_LAST_ONLY), then mgb_fas_residual(fine, coarse) -- bit-identical values in fine.u at the
 * C-points, coarse.u and coarse.g -- except that the last F-point of every interval is not stored (nothing reads it before
 * the F-relaxation after the coarse-grid correction rewrites the interval).  Per interval 2 rows are read and 3 written
 * instead of 5 + 4.  Requires weight 1 and at least one F-point between any two C-points (no adjacent C-points); the
 * caller checks (pymgrit_b200/core/mgrit.py).  MGB_ENOSHAPE for the ODE applications (no team kernels). */
int mgb_down_sweep(const mgb_level *fine, const mgb_level *coarse, void *stream);

/* Coarse-grid correction, mgrit.py:715-726: fine.u[c_j] += coarse.u[j] - v[j], j >= 1.  flags:
 *   MGB_CORRECT_F_RELAX  also run the F-relaxation of mgrit.py:287 out of the same registers (one launch);
 *   MGB_CORRECT_GHOST    point 0 is the ghost copy of the previous time rank's last C-point (time rank > 0):
 *                        correct it too, exactly as its owner does, so that no exchange is needed before the
 *                        F-relaxation of the first interval.  On time rank 0 point 0 is the initial condition
 *                        and is never corrected (mgrit.py:723). */
/*   MGB_CORRECT_LAST_ONLY  (with MGB_CORRECT_F_RELAX) store only the last F-point of every interval: it is all that the
 *                        residual (mgrit.py:405-413) and the next C-relaxation read.  The solver uses it on level 0 while
 *                        it iterates and materialises the F-points with one mgb_f_relax when the iteration stops -- the
 *                        values are the ones the reference holds (F-points are Phi chains from the C-points). */
#define MGB_CORRECT_F_RELAX 1
#define MGB_CORRECT_GHOST 2
#define MGB_CORRECT_LAST_ONLY 4
int mgb_error_correction(const mgb_level *fine, const mgb_level *coarse, int32_t flags, void *stream);

/* FAS restriction with a spatial grid transfer R != identity (mgrit.py:488-549 with a user GridTransfer,
 * core/grid_transfer.py:31-55; examples/example_spatial_coarsening.py).  Fine and coarse level differ in spatial size, so
 * the sweep is split around the transfer:
 *   mgb_residual_rows(fine, out)      out[j] = Phi_f(u[c_j-1]) - u[c_j] (level 0) or (g[c_j] - u[c_j]) + Phi_f(u[c_j-1]),
 *                                     j >= 1; out is [ncpts][pitch] in the fine level's row layout
 *   mgb_heat1d_restrict_rows          coarse.u[j] = R(fine.u[c_j]) (all j), v = copy of coarse.u, rres[j] = R(out[j])
 *   mgb_fas_coarse_rhs(coarse, v, rres, nrows)   coarse.g[j] = (rres[j] + v[j]) - Phi_c(v[j-1]),  j = 1 .. nrows-1
 * v (mgrit.py:520) must be stored here: it is no longer a copy of fine-level rows. */
int mgb_residual_rows(const mgb_level *lvl, double *out_rows_dev, void *stream);
int mgb_fas_coarse_rhs(const mgb_level *coarse, const double *v_dev, const double *rres_dev, int32_t nrows, void *stream);

/* The spatial transfer of examples/example_spatial_coarsening.py:18-79 for Heat1D rows, n_fine = 2 n_coarse + 1 interior
 * points, homogeneous Dirichlet boundaries; one launch covers all rows.
 *   restrict (full weighting):  dst[j][i] = (s[2i]/4 + s[2i+1]/2) + s[2i+2]/4,  s = src[src_index ? src_index[j] : j],
 *                               j < nrows; the padding of dst rows is zeroed
 *   interp (linear):            e = a[j] - b[j] (a[j] if b is NULL);  E[2i+1] = e[i], E[2i] = e[i-1]/2 + e[i]/2;
 *                               dst[r] = accumulate ? dst[r] + E : E,  r = dst_index ? dst_index[j] : j,  first <= j < nrows
 * With accumulate this is the coarse-grid correction of mgrit.py:715-726 (a = coarse.u, b = v, dst = fine.u, dst_index =
 * the fine level's C-points), without it the interpolation of nested iteration (mgrit.py:559-563). */
int mgb_heat1d_restrict_rows(int32_t nrows, const double *src_dev, int32_t src_pitch, const int32_t *src_index_dev,
                             int32_t n_fine, double *dst_dev, int32_t dst_pitch, void *stream);
int mgb_heat1d_interp_rows(int32_t nrows, int32_t first, const double *a_dev, const double *b_dev, int32_t c_pitch,
                           int32_t n_coarse, double *dst_dev, int32_t dst_pitch, const int32_t *dst_index_dev,
                           int32_t accumulate, void *stream);

/* Sequential solve on the coarsest level, mgrit.py:459-486: u[i] = (g[i] +) Phi(u[i-1]), i = 1..npts-1. */
int mgb_forward_solve(const mgb_level *lvl, void *stream);

/* AT-MGRIT, core/at_mgrit.py:75-86 (one time rank): instead of the sequential solve, every point p >= 1 of the coarsest
 * level is the end of its own local coarse grid of at most k points started from the previous iterate,
 *   x = old[max(0, p-k+1)];  for i = max(1, p-k+2) .. p:  x = (g[i] +) Phi(x);   u[p] = x,
 * all points in one launch.  old_dev: a copy of u made before the call ([npts][pitch]). */
int mgb_local_coarse_solve(const mgb_level *lvl, const double *old_dev, int32_t k, void *stream);

/* Space-time residual at the C-points of level 0, mgrit.py:387-413: out_sq_dev[j] = ||Phi(u[c_j-1]) -
 * u[c_j]||_2^2 for j >= 1 (out_sq_dev[0] = 0).  out_sq_dev holds ncpts doubles, or ncpts * (1 + nsys) when the
 * level has nsys > 1 systems per row (the per-system partial sums are kept behind the results). */
int mgb_residual_norms(const mgb_level *lvl, double *out_sq_dev, void *stream);

/* Jump criterion, mgrit.py:372-385: out_sq_dev[j] = ||u[c_j] - last[c_j]||_2^2 for j >= 1, then
 * last <- u for every point. */
int mgb_jump_norms(const mgb_level *lvl, double *last_dev, double *out_sq_dev, void *stream);

/* Temporal norm of mgrit.py:430-431 over sq_dev[0..count): writes the rank-local partial
 * (TWO: sum of squares, ONE: sum of roots, INF: max of roots) to out_dev[0]. */
int mgb_temporal_norm(const double *sq_dev, int32_t count, int32_t t_norm, double *out_dev, void *stream);

/* Stopping test without a host round trip per iteration (mgrit.py:626: `if self.conv[iteration + 1] < self.tol: break`).
 * The reference decides on the host after every iteration; a GPU solver that does so idles while the host reads one double
 * and queues the next cycle.  Instead the host queues one iteration ahead:
 *   mgb_convergence_flag  after the residual norm of an iteration has been reduced to norm_dev[0] (the 2-norm's sum of
 *                         squares, or the 1-/inf-norm), stores it in hist_dev[0] and raises flag_dev[0] if the norm
 *                         (its square root for MGB_TNORM_TWO) is below tol;
 *   mgb_set_stop_flag     makes every sweep launched from now on return at once while that flag is up, so the cycle
 *                         that was queued before the host learned of convergence changes nothing (NULL: off);
 *   mgb_write_flag        sets the flag from the stream (cleared before the next solve).
 * The values the host reads one iteration late are the same numbers; the iteration count and the iterates are
 * unchanged.  Ghost-row exchanges between time ranks are not skipped (their sequence numbers stay in step). */
int mgb_set_stop_flag(const int32_t *flag_dev);
int mgb_write_flag(int32_t *flag_dev, int32_t value, void *stream);
int mgb_convergence_flag(const double *norm_dev, int32_t t_norm, double tol, double *hist_dev, int32_t *flag_dev,
                         void *stream);
/* One time rank: mgb_temporal_norm and mgb_convergence_flag in one launch (nothing to reduce in between). */
int mgb_temporal_norm_flag(const double *sq_dev, int32_t count, int32_t t_norm, double *out_dev, double tol, double *hist_dev,
                           int32_t *flag_dev, void *stream);

/* Nested iteration, mgrit.py:559-563: fine.u[c_j] = coarse.u[j] for j >= 1. */
int mgb_inject_up(const mgb_level *fine, const mgb_level *coarse, void *stream);

/* One application of Phi for Application.step (heat_1d.py:198 etc.): out = Phi(in) for the step that
 * produces point `point` of the level (its dt and right-hand side).  in/out are rows in the level's layout. */
int mgb_step(const mgb_level *lvl, int32_t point, const double *in_dev, double *out_dev, void *stream);

/* ---- Heat2D: node arrays <-> level rows (heat/heat_2d.py:20-136, 322-366) --------------------- */
/* Row layout for an (nx, ny) node grid: (nx-2)(ny-2) sine coefficients, zero padding to a whole tile, the
 * 2 ny + 2 (nx-2) boundary values (row i = 0, row i = nx-1, column j = 0, column j = ny-1), zero padding.
 * tile = doubles per system, nsys = tiles per row, first_boundary_sys = first tile of boundary values,
 * pitch = nsys * tile. */
int mgb_heat2d_layout(int32_t nx, int32_t ny, int32_t *tile, int32_t *nsys, int32_t *first_boundary_sys, int32_t *pitch);

/* rows[b] = layout( Sx * nodes[b]_interior * Sy, nodes[b]_boundary ), b < count.  sx_dev [(nx-2)^2] and sy_dev
 * [(ny-2)^2] are the orthonormal sine matrices S[j][k] = sqrt(2/(n+1)) sin(pi (j+1)(k+1)/(n+1)); nodes_dev is
 * [count][nx*ny] (VectorHeat2D.get_values(), heat_2d.py:105-110), rows_dev [count][pitch], work_dev
 * [count][(nx-2)(ny-2)] scratch. */
int mgb_heat2d_to_rows(int32_t nx, int32_t ny, const double *sx_dev, const double *sy_dev, const double *nodes_dev,
                       double *rows_dev, int32_t count, double *work_dev, void *stream);
/* The inverse: nodes[b] from rows[b]. */
int mgb_heat2d_from_rows(int32_t nx, int32_t ny, const double *sx_dev, const double *sy_dev, const double *rows_dev,
                         double *nodes_dev, int32_t count, double *work_dev, void *stream);

/* ---- host-side passes of the setup (csrc/host_tables.cu; no device, multi-threaded, bit-identical to NumPy) --------- */
/* dt[0] = 0, dt[i] = t[i] - t[i-1]; lo / hi = the smallest / largest step. */
int mgb_host_time_steps(const double *t, int64_t n, double *dt, double *lo, double *hi, int32_t threads);
/* out[i] = i * step + start (product and sum rounded separately, like numpy.linspace's arange(n) * step + start). */
int mgb_host_affine_ramp(double start, double step, int64_t n, double *out, int32_t threads);
/* out[i][k] = src[k][i] * scale[i] (scale may be NULL): [q][n] rows with stride ld_src -> [n][q]. */
int mgb_host_scale_rows(const double *src, int64_t ld_src, int32_t q, int64_t n, const double *scale, double *out,
                        int32_t threads);

/* ---- Advection1D: the coarsest-level solve in Fourier space (csrc/fourier.cu) ------------------------------------ */
/* mgrit.py:459-486 on a level of Advection1D (advection_1d.py:129-143): Phi_i = (I + dt_i (c/dx)(I - S))^-1 is circulant,
 * so u_i = g_i + Phi_i(u_{i-1}) decouples under the discrete Fourier transform into n/2 + 1 complex scalar recurrences
 *     W[i][k] = W[i-1][k] / (1 + nu_i (1 - exp(-2 pi i k / n))) + W[i][k],   nu_i = c_over_dx (t[i] - t[i-1]),
 * which run time-parallel in one launch.  Three calls replace the chain of npts - 1 dependent cyclic solves:
 *     work  = rfft([u[0]; g[1..]])     mgb_rows_rfft    (Bluestein: any n <= 4096; row r -> [2 (n/2+1)] doubles re, im)
 *     recurrences in place on work      mgb_advection1d_spectral_recur
 *     u[i]  = irfft(work[i]), i >= 1    mgb_rows_irfft
 * Tables (made once per n by mgb_circ_fft_tables; M = mgb_circ_fft_length(n) = 2^p >= 2n - 1): tw [M/2], chirp [n],
 * bhat [M] complex doubles each; work_dev: M complex doubles of scratch.  ld* in doubles, ldc and ldw even, c_dev and
 * work_dev 16-byte aligned. */
int mgb_circ_fft_length(int32_t n);
int mgb_circ_fft_tables(int32_t n, double *tw_dev, double *chirp_dev, double *bhat_dev, double *work_dev, void *stream);
int mgb_rows_rfft(int32_t m, int32_t n, const double *a_dev, int64_t lda, const double *a_row0_dev, const double *tw_dev,
                  const double *chirp_dev, const double *bhat_dev, double *c_dev, int64_t ldc, void *stream);
int mgb_rows_irfft(int32_t m, int32_t first_row, int32_t n, const double *c_dev, int64_t ldc, const double *tw_dev,
                   const double *chirp_dev, const double *bhat_dev, double *out_dev, int64_t ldo, void *stream);
int mgb_advection1d_spectral_recur(int32_t n, int32_t npts, const double *t_dev, double c_over_dx, double *work_dev,
                                   int64_t ldw, void *stream);

/* ---- Heat1D: the coarsest-level solve in sine space (csrc/spectral.cu) ------------------------------------- */
/* mgrit.py:459-486 runs npts-1 dependent Phi applications on one spatial system; every time rank waits for that chain
 * (mgrit.py:467-484).  For HEAT1D the spatial operator is Toeplitz: with the rows transformed by the orthonormal sine
 * matrix S (S = S^T = S^-1) the chain decouples into n independent scalar recurrences.  Usage for a level `lvl`:
 *     work[0]  = u[0] S;  work[i] = g[i] S (i >= 1; 0 without g)        mgb_rows_gemm (one launch, a_row0 = u[0])
 *     work[i]  = (work[i-1] + sum_q rhs_t[i][q] rxhat[q]) / (1 + dt_i lam) + work[i]      mgb_heat1d_spectral_recur
 *     u[i]     = work[i] S                                              mgb_rows_gemm
 * Between time ranks only work[npts-1] -> next rank's work[0] travels. */

/* s_dev[j*ld + k] = sqrt(2/(n+1)) sin(pi (j+1)(k+1)/(n+1)), j, k < n. */
int mgb_sine_matrix(int32_t n, double *s_dev, int32_t ld, void *stream);
/* c (m x n, ldc) = a (m x k, lda) * b (k x n, ldb): plain FP64 product on the CUDA cores, row-major.  If a_row0_dev is
 * not NULL, row 0 of a is read from there (the level's u[0] next to its g rows). */
int mgb_rows_gemm(int32_t m, int32_t n, int32_t k, const double *a_dev, int32_t lda, const double *a_row0_dev,
                  const double *b_dev, int32_t ldb, double *c_dev, int32_t ldc, void *stream);
/* The same product c = a S (k = n) as a fast sine transform, for n + 1 a power of two: one CTA per row, radix-2 FFT of
 * the odd extension in shared memory, O(n log n) per row.  tw_dev: n + 1 complex doubles filled once by
 * mgb_dst_twiddles. */
int mgb_dst_twiddles(int32_t n, double *tw_dev, void *stream);
int mgb_rows_dst(int32_t m, int32_t n, const double *a_dev, int32_t lda, const double *a_row0_dev, const double *tw_dev,
                 double *c_dev, int32_t ldc, void *stream);
/* The n scalar recurrences, in place on work_dev [npts][pitch].  lam_dev [n] = eigenvalues of the spatial operator
 * (a/dx^2) tridiag(-1, 2, -1); rxhat_dev [nrhs][pitch] = the spatial right-hand-side factors times S (rhs_t_dev of the
 * level supplies the time factors).  Needs a separable right-hand side (rhs_dense_dev == NULL), nrhs <= 4. */
int mgb_heat1d_spectral_recur(const mgb_level *lvl, const double *lam_dev, const double *rxhat_dev, double *work_dev,
                              double *ends_dev, int32_t zero_start, void *stream);
/* Time-parallel form (replaces the rank-to-rank chain of mgrit.py:467-484).  Every rank runs mgb_heat1d_spectral_recur
 * with ends_dev [2][pitch] != NULL -- rank 0 from its true first row, the others with zero_start = 1 -- which returns the
 * last value and the product of the step factors per mode; the pairs are all-gathered ([nranks][2][pitch]) and rank r > 0
 * calls mgb_heat1d_spectral_fixup: work[0] = true value at the slab start, work[i] += (product of the first i factors)
 * times it.  One collective of 16 KB per rank instead of nranks - 1 dependent sends. */
int mgb_heat1d_spectral_fixup(const mgb_level *lvl, const double *lam_dev, double *work_dev, const double *all_ends_dev,
                              int32_t rank, void *stream);

/* ---- Heat1D in sine space (MGB_APP_HEAT1D_SINE) ------------------------------------------------------------------ */
/* Level rows hold x = u S.  Per level: diag_dev (above), sconst_dev rows [dt, use_reciprocals, 0...] (width 8), rhs_x_dev =
 * [nrhs][chunk][team_threads] thread-transposed X_q S, rhs_t_dev as for HEAT1D.  All sweeps above work on such levels.
 * The sequential solve of mgrit.py:459-486 on such a level needs no transforms at all and is done time-parallel:
 * mgb_sine_level_solve cuts the level's steps into chunks, runs every chunk's scalar recurrences from zero, combines the
 * chunk results (value, product of the step factors) in order and reruns the chunks from their true start values:
 *     u[i][k] = (u[i-1][k] + sum_q rhs_t[i][q] rxh[q][k]) / (1 + (t[i] - t[i-1]) lam[k]) + g[i][k],   i = 1 .. npts-1
 * lam_dev [n], rxh_dev [nrhs][pitch] in natural order, lvl->t_dev the level's time grid.  ends_dev / zero_start as for
 * mgb_heat1d_spectral_recur (time ranks > 0 start from zero and are fixed up by mgb_heat1d_spectral_fixup on u). */
int mgb_sine_level_solve(const mgb_level *lvl, const double *lam_dev, const double *rxh_dev, double *ends_dev,
                         int32_t zero_start, void *stream);

/* ---- The batched path: applications without fused team kernels (csrc/generic.cu) ---------------------------------- */
/* The reference's plug-in is Application.step (core/application.py:99).  An application that brings its own batched device
 * Phi (pymgrit_b200.BatchedApplication.step_rows: many (row, step) pairs per call) runs the same sweeps: the host walks
 * the positions inside a coarse interval, every position is one batched Phi over all intervals of the level, and the
 * row-wise parts of mgrit.py:312-327, 354-368, 497-547, 722-726 are these two entry points.  Index arrays are int32 device
 * arrays of `count` row numbers (NULL: row k).
 *   mgb_rows_lincomb  out[oi[k]] = (a x[xi[k]] + b y[yi[k]]) + c z[zi[k]]  on n doubles per row (y, z may be NULL; every
 *                     product and sum rounded on its own, like the reference's NumPy expressions)
 *   mgb_rows_sumsq    out_sq[k] = sum_q x[xi[k]][q]^2   (Vector.norm squared; fixed summation order) */
int mgb_rows_lincomb(int32_t count, int32_t n, double *out_dev, int32_t out_pitch, const int32_t *out_idx_dev, double a,
                     const double *x_dev, int32_t x_pitch, const int32_t *x_idx_dev, double b, const double *y_dev,
                     int32_t y_pitch, const int32_t *y_idx_dev, double c, const double *z_dev, int32_t z_pitch,
                     const int32_t *z_idx_dev, void *stream);
int mgb_rows_sumsq(int32_t count, int32_t n, const double *x_dev, int32_t pitch, const int32_t *x_idx_dev,
                   double *out_sq_dev, void *stream);

/* Allen-Cahn, IMEX branch of allen_cahn/allen_cahn.py:191-197, batched: for k < count
 *   rhs = u + dt[k] (1/eps^2) u (1 - u^nu),  dst[di[k]] = (I - dt[k] L)^-1 rhs,   u = src[si[k]] (nx x nx nodes, row-major)
 * with the periodic 5-point Laplacian L = L1 (x) I + I (x) L1, L1 = -Q diag(mu) Q^T: q_dev [nx*nx] the real orthonormal
 * Fourier basis (columns), qt_dev its transpose, mu_dev [nx] = (4/dx^2) sin^2(pi m / nx) per column; work1/2_dev
 * [count][nx*nx] scratch.  Four batched FP64 products per call (spsolve in the reference: the same linear system). */
int mgb_allen_cahn_imex_rows(int32_t nx, int32_t count, const double *src_dev, int32_t src_pitch,
                             const int32_t *src_idx_dev, double *dst_dev, int32_t dst_pitch, const int32_t *dst_idx_dev,
                             const double *dt_dev, double inv_eps2, int32_t nu, const double *q_dev, const double *qt_dev,
                             const double *mu_dev, double *work1_dev, double *work2_dev, void *stream);

/* ---- Ghost rows between time ranks over peer memory (csrc/peer.cu) ----------------------------------------- */
/* Replaces the reference's kind-0/4 messages (mgrit.py:305-310, 510-517, 693-713: pickled vector, isend/recv) by direct
 * stores into the successor's mailbox over NVLink.  The caller owns a symmetric allocation per rank (same layout on
 * every rank, peer-mapped): per level two row slots, two flags and one acknowledgement word (uint64, zero-initialised).
 * The k-th exchange of a level uses seq = k >= 1 and slot / flag k & 1 on both sides.
 *   sender:   mgb_peer_put_row(my last row, successor's slot, count, successor's flag, MY ack word, seq)
 *             waits (on the device) until the successor has acknowledged seq - 2, stores the row, publishes seq;
 *   receiver: mgb_peer_wait_row(my slot, my ghost row, count, my flag, PREDECESSOR's ack word, seq)
 *             waits (on the device) for seq, copies the row, acknowledges.
 * Both are ordinary asynchronous launches on `stream`; neither touches the host. */
/* Every device-side wait on a peer gives up after a timeout (30 s unless mgb_peer_set_timeout changes it): the kernel
 * records which wait it was and goes on, so a rank that died or took another code path cannot hang the other GPUs.
 * mgb_peer_status copies the word to the host (synchronising; the solver calls it when a solve is over): 0 = every wait
 * was answered, else one of MGB_PEER_WAIT_*; reset != 0 clears it. */
#define MGB_PEER_WAIT_ROW 1        /* a ghost row did not arrive                              */
#define MGB_PEER_WAIT_ACK 2        /* the successor never consumed the previous ghost row      */
#define MGB_PEER_WAIT_GATHER 3     /* rows of the coarsest-level gather did not arrive         */
#define MGB_PEER_WAIT_GATHER_ACK 4 /* a rank never consumed the previous gather                */
int mgb_peer_status(int32_t *error_out, int32_t reset);
int mgb_peer_set_timeout(double seconds);
int mgb_peer_put_row(const double *src_dev, double *peer_slot_dev, int32_t count, void *peer_flag_dev, const void *my_ack_dev,
                     uint64_t seq, void *stream);
int mgb_peer_wait_row(const double *my_slot_dev, double *dst_dev, int32_t count, const void *my_flag_dev, void *peer_ack_dev,
                      uint64_t seq, void *stream);

/* One-to-many form, used by the sine-space coarsest solve instead of an all-gather: rank r stores its (last value,
 * factor product) rows into the gather buffer of every rank above it.  The pointer arrays are HOST arrays of npeers
 * (<= 16) device addresses: for put, the destination slot and flag in each peer's buffer and the acknowledgement word
 * that peer writes in mine; for wait, the flags the sources write in my buffer and the acknowledgement words in theirs.
 * The wait acknowledges seq - 1: the rows of the previous round, which everything queued before it has finished reading
 * (the rows of this round are read in place by the next kernel, mgb_heat1d_spectral_fixup). */
int mgb_peer_put_rows(const double *src_dev, int32_t count, int32_t npeers, const uint64_t *peer_slot_ptrs,
                      const uint64_t *peer_flag_ptrs, const uint64_t *my_ack_ptrs, uint64_t seq, void *stream);
int mgb_peer_wait_flags(int32_t npeers, const uint64_t *my_flag_ptrs, const uint64_t *peer_ack_ptrs, uint64_t seq,
                        void *stream);

/* ---- Vector arithmetic (core/vector.py:38-110) ---------------------------------------------- */
/* out = a*x + b*y on n doubles */
int mgb_vec_axpby(int32_t n, double a, const double *x_dev, double b, const double *y_dev, double *out_dev,
                  void *stream);
/* out_dev[0] = sum of squares of x (Vector.norm = sqrt of it, heat_1d.py:62-68) */
int mgb_vec_sumsq(int32_t n, const double *x_dev, double *out_dev, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* MGRIT_B200_H */
