/*
 * lbm_b200.h -- C ABI of the B200-native D2Q9 BGK lattice-Boltzmann step.
 *
 * This is the drop-in boundary of the repository: a plain-C shared library
 * (liblbm_b200.so; nvcc, sm_100a) with no torch / pybind11 / C++ types in any
 * signature.  Everything above it (the `_lbkernels` / `PyLB` mirror of the
 * reference API, the Lattice object, bench.py) binds these symbols with
 * ctypes; INTEGRATION.md shows the binding a maintainer of the reference adds.
 *
 * Two groups of entry points:
 *
 *  (1) lbk_*  -- stateless, HOST buffers in / out.  One symbol per overload the
 *      reference registers in c/_lbkernels.cpp:36-48, plus PyLB.stream
 *      (PyLB/Streaming.py:33-46).  Same argument meaning, in-place semantics
 *      and (lack of) length checks as the reference; the work runs on the GPU
 *      (H2D -> kernel -> D2H inside the call).
 *
 *  (2) lb_*   -- a device-resident lattice whose whole time step
 *      (cavity_opt2.py:272-277: communicate + stream_and_bounce_back + collide)
 *      is ONE fused kernel launch.  This has no counterpart in the reference
 *      (which moves host arrays through three calls per step); it is what the
 *      re-hosted simulators and the benchmark drive.
 *
 * All functions return 0 on success and a negative lb_status on failure;
 * lb_last_error() returns a thread-local message for the last failure.
 * Reference citations are file:line relative to the upstream repository root.
 */
#ifndef LBM_B200_H
#define LBM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LB_ABI_VERSION 1
#define LB_API __attribute__((visibility("default")))

typedef enum lb_status {
    LB_OK = 0,
    LB_ERR_INVALID = -1,      /* bad argument / unsupported configuration      */
    LB_ERR_CUDA = -2,         /* a CUDA runtime call failed (see lb_last_error) */
    LB_ERR_NO_DEVICE = -3,    /* no usable CUDA device: there is NO CPU fallback */
    LB_ERR_STATE = -4,        /* call order violated (e.g. step before connect) */
    LB_ERR_HALO_TIMEOUT = -5  /* a neighbour's halo flag never arrived          */
} lb_status;

typedef enum lb_dtype { LB_F32 = 0, LB_F64 = 1 } lb_dtype;

/* Boundary handling of the fused step.  Predicates are evaluated on GLOBAL cell
 * coordinates, so any decomposition reproduces the single-block result bit for
 * bit.                                                       */
typedef enum lb_boundary {
    LB_PERIODIC = 0,          /* shear_wave_opt2.py:95-97: stream; collide                 */
    LB_CAVITY = 1,            /* cavity_opt2.py:109-177: 4 half-way bounce-back walls + lid */
    LB_CAVITY_XPERIODIC = 2,  /* same with `if True:` -> `if False:` (cavity_opt2.py:147)   */
    LB_SF_COUETTE = 3,        /* simple_flows/PoiseuilleFlow.py:93-111 (wall layers)        */
    LB_SF_POISEUILLE = 4,     /* simple_flows/PoiseuilleFlow.py:129-148                     */
    LB_SF_SLIDING_LID = 5,    /* simple_flows/slidingLid.py:68-108                          */
    LB_SF_TABLE = 6           /* simple_flows family with a per-cell boundary TABLE (lb_set_boundary_table):
                                 slidingLidMPI.py:180-204, experimantal_flows/obstacle_canal.py:413-458 */
} lb_boundary;

/* Arithmetic of the collision.
 * EXACT: every operation individually rounded, in the expression order of
 *        c/d2q9.h:59-81,121-131, IEEE division, no FMA contraction -> the
 *        kernel is bit-identical to the reference's baseline-x86-64 build.
 * FAST : FMA contraction allowed, divisions by constants become multiplies,
 *        one reciprocal of rho per cell.  Deviation from EXACT is documented
 *        in DESIGN.md and asserted < 1e-12 relative after 1000 steps.       */
typedef enum lb_arith { LB_ARITH_EXACT = 0, LB_ARITH_FAST = 1 } lb_arith;

/* Direction slots of the 8 neighbours of a block (dx, dy):
 * 0:(-1,0) 1:(+1,0) 2:(0,-1) 3:(0,+1) 4:(-1,-1) 5:(-1,+1) 6:(+1,-1) 7:(+1,+1) */
#define LB_NUM_DIRS 8

typedef struct lb_config {
    int32_t device;       /* CUDA device ordinal                                        */
    int32_t dtype;        /* lb_dtype                                                   */
    int32_t boundary;     /* lb_boundary                                                */
    int32_t arith;        /* lb_arith                                                   */
    int64_t gnx, gny;     /* global lattice (cavity_opt2.py:53-54 nx, ny)               */
    int64_t x0, y0;       /* global coordinate of this block's cell (0,0)               */
    int64_t lnx, lny;     /* this block's extent WITHOUT ghosts (cavity_opt2.py:231-258) */
    double omega;         /* relaxation parameter (cavity_opt2.py:66)                   */
    double u_wall;        /* lid / moving-wall velocity u0 (cavity_opt2.py:109) or uw   */
    double rho_in;        /* LB_SF_POISEUILLE only (PoiseuilleFlow.py:134)              */
    double rho_out;       /* LB_SF_POISEUILLE only (PoiseuilleFlow.py:135)              */
} lb_config;

typedef struct lb_lattice lb_lattice;   /* opaque */

/* Everything a neighbour (same process, other process, other GPU) needs in
 * order to push ghost populations straight into this block's memory.        */
typedef struct lb_export {
    uint8_t ipc_mem_handle[64];  /* cudaIpcMemHandle_t of the block's single allocation */
    uint64_t local_base;         /* device address (valid inside the exporting process) */
    int64_t pid;                 /* exporting process                                   */
    int32_t device;              /* exporting device ordinal                            */
    int32_t dtype;
    int64_t lnx, lny;            /* extents without ghosts                              */
    int64_t pitch;               /* row pitch in elements                               */
    int64_t pop_stride;          /* population stride in elements                       */
    int64_t buf_bytes;           /* bytes of one f buffer (there are two, A then B)     */
    int64_t ycol_offset;         /* byte offset of the ghost-column arrays (A then B)   */
    int64_t ycol_bytes;          /* bytes of one ghost-column array                     */
    int64_t frame_offset;        /* byte offset of the temporal-blocking frame storage  */
    int64_t state_offset;        /* byte offset of the device state block (flags)       */
    int64_t total_bytes;
} lb_export;

LB_API const char *lb_last_error(void);
LB_API int lb_abi_version(void);
/* sizeof(lb_config) / sizeof(lb_export) as compiled into the library, for bindings to verify their struct
 * declarations against (a mismatch would silently corrupt the halo wiring).                       */
LB_API int64_t lb_sizeof_config(void);
LB_API int64_t lb_sizeof_export(void);
LB_API int lb_device_count(void);

/* ---- (2) device-resident lattice ------------------------------------------------ */
LB_API int lb_create(const lb_config *cfg, lb_lattice **out);
/* lb_create with creation flags.
 * LB_CREATE_INPLACE: ONE copy of the populations instead of the A/B pair (half the footprint: 32768^2 fp64 in
 *   77 GB instead of 155 GB), advanced in place with the AA pattern -- even steps gather from the neighbours and
 *   store each result back where its input came from, odd steps work on the cell's own slots -- at the same
 *   144 B per cell per step and bit-identical to the A/B kernels.  One self-connected block, periodic / cavity
 *   boundaries, single-step kernel only (no temporal blocking, probe, stream-only or host step).            */
#define LB_CREATE_INPLACE 1
LB_API int lb_create_ex(const lb_config *cfg, int flags, lb_lattice **out);
LB_API int lb_destroy(lb_lattice *lat);
/* Launch all work of this lattice on `cuda_stream` (a cudaStream_t / CUstream,
 * e.g. torch.cuda.current_stream().cuda_stream); 0 = the lattice's own stream. */
LB_API int lb_set_stream(lb_lattice *lat, void *cuda_stream);
/* The stream the lattice currently launches on (to share it with other blocks). */
LB_API void *lb_get_stream(lb_lattice *lat);
LB_API int lb_sync(lb_lattice *lat);

/* Halo wiring (replaces Create_cart/Shift + communicate(), cavity_opt2.py:179-229). */
LB_API int lb_get_export(lb_lattice *lat, lb_export *out);
/* Connect direction slot `dir` to a neighbour.  A neighbour with the same pid is
 * addressed directly (same device, or another device with peer access enabled);
 * any other is opened with cudaIpcOpenMemHandle.  Passing the lattice's own
 * export closes a periodic ring on itself.                                     */
LB_API int lb_connect(lb_lattice *lat, int dir, const lb_export *neighbour);
/* Push the edge populations of the CURRENT state into the neighbours' ghost
 * layers (needed once after init/upload; the fused step does it every step).
 * The caller must make sure all neighbours have finished writing their state
 * (stream sync + rank barrier) before and after.                               */
LB_API int lb_halo_refresh(lb_lattice *lat);

/* State in / out.  Host arrays are C-contiguous (9, lnx, lny), the reference's
 * f_ikl layout (cavity_opt2.py:79-83), ghosts excluded.                        */
LB_API int lb_upload_f(lb_lattice *lat, const void *host_f);
LB_API int lb_download_f(lb_lattice *lat, void *host_f);
/* Rows [k_lo, k_hi) of the current state to / from a C-contiguous host array (9, k_hi - k_lo, lny): strip
 * checks against the oracle at sizes whose full field does not fit a host, block-wise checkpoints.
 * After lb_upload_rows the caller refreshes the halos (lb_halo_refresh) like after lb_upload_f.        */
LB_API int lb_download_rows(lb_lattice *lat, int64_t k_lo, int64_t k_hi, void *host_rows);
LB_API int lb_upload_rows(lb_lattice *lat, int64_t k_lo, int64_t k_hi, const void *host_rows);
/* 64-bit digest of the current state, computed on the device: sum over populations and real cells of
 * mix(bit pattern, GLOBAL cell index) modulo 2^64.  The digests of the blocks of ANY decomposition add up
 * (mod 2^64) to the digest of the undecomposed lattice, so "1 GPU == N GPUs" and "two steps per pass ==
 * single steps" can be checked bit for bit at sizes that are never gathered (cavity_opt2.py has no
 * counterpart; the gate is BASELINE.json's bit-exact decomposition requirement).                       */
LB_API int lb_checksum(lb_lattice *lat, uint64_t *digest);
/* f = feq(rho, ux, uy) with the arithmetic of c/d2q9.h:59-81 (cavity_opt2.py:265-269).
 * Host arrays of lnx*lny values, or NULL for rho = 1 / u = 0.                  */
LB_API int lb_init_equilibrium(lb_lattice *lat, const void *rho, const void *ux, const void *uy);

/* Advance `nsteps` time steps; asynchronous on the lattice's stream.           */
LB_API int lb_step(lb_lattice *lat, int64_t nsteps);
/* Stream + boundary handling WITHOUT the collision (PyLB.stream /
 * stream_and_bounce_back as stand-alone operations, cavity_opt2.py:109-177).   */
LB_API int lb_stream_only(lb_lattice *lat, int64_t nsteps);
/* ONE time step with HOST input and output (the reference's calling convention: the loop body
 * of cavity_opt2.py:275-277 applied to a host f_ikl): rows are uploaded, updated and downloaded
 * slab by slab on three streams so that H2D, compute and D2H overlap.  host_out may alias
 * host_in; both are C-contiguous (9, lnx, lny).  Single self-connected block; returns when
 * host_out is complete.  Pinned host memory is required for real overlap.              */
LB_API int lb_step_host(lb_lattice *lat, const void *host_in, void *host_out, int nslabs);
/* The same on a DECOMPOSED lattice (one block per rank / GPU), three phases per step with the ranks' own
 * barrier between them -- the host-side schedule that replaces communicate() (cavity_opt2.py:179-210) when the
 * state lives in host memory:
 *   1. lb_step_host_begin(lat, host_in)   the block's rim (rows 0 / lnx-1, columns 0 / lny-1) goes to the device
 *      [barrier]
 *   2. lb_halo_refresh(lat); lb_sync(lat) every rank pushes its rim into its neighbours' ghosts over NVLink
 *      [barrier]
 *   3. lb_step_host(lat, host_in, host_out, nslabs)   slab-pipelined H2D / compute / D2H as above.        */
LB_API int lb_step_host_begin(lb_lattice *lat, const void *host_in);
/* Same, bracketed by CUDA events on the launching stream; returns elapsed ms.  */
LB_API int lb_step_timed(lb_lattice *lat, int64_t nsteps, float *elapsed_ms);
LB_API int64_t lb_steps_done(lb_lattice *lat);
/* 0 if healthy, else an lb_status (halo timeout detected inside a kernel).     */
LB_API int lb_health(lb_lattice *lat);

/* rho = sum_i f_i, u = (f^T c)/rho (cavity_opt2.py:280-281) to host arrays of
 * lnx*lny values (any pointer may be NULL).                                    */
LB_API int lb_moments(lb_lattice *lat, void *rho, void *ux, void *uy);
/* Shear-wave probe (shear_wave_opt2.py:99): after every step append
 *   sum_k uy(k, l_probe) * uy_k[k] * 2/gnx      (this block's k-range only)
 * to a device-side series of `capacity` values.  uy_k: host array of lnx values. */
LB_API int lb_probe_shear_enable(lb_lattice *lat, int64_t l_probe_global, const void *uy_k, int64_t capacity);
LB_API int lb_probe_shear_read(lb_lattice *lat, void *out, int64_t n);

/* Tuning: rows of the lattice handled by one CTA (default 4 for fp64, 8 for fp32).                  */
LB_API int lb_set_rows_per_tile(lb_lattice *lat, int rows);
/* How long a rim CTA waits for a neighbour's halo flag before the lattice is marked failed
 * (lb_health -> LB_ERR_HALO_TIMEOUT) instead of hanging the GPU.  Default 20 s.            */
LB_API int lb_set_halo_timeout_ms(lb_lattice *lat, int64_t ms);
/* Time steps per pass over HBM.  2 = temporal blocking: lb_step advances pairs of steps with one read and
 * one write of the lattice (bit-identical results; periodic / cavity boundaries, blocks of at least 16 x 16
 * cells; odd remainders and everything else use the single-step kernel).  1 = always the single-step kernel.
 * 0 (default) = automatic: temporal blocking when the block offers at least 512 fused tiles of 16 rows (about
 * 1400^2 cells), else the single-step kernel.  The mode is a COLLECTIVE property of a decomposition: blocks that
 * exchange halos must all use the same mode (latticeboltzmann_b200.distributed decides it for the whole
 * world and sets 1 or 2 explicitly).  rows_per_tile > 0 overrides the fused tile height (default 16, 32 or 48 by block size;
 * lb_temporal_rows returns the height in use).
 * Environment override: LBM_TEMPORAL=0|1|2.                                                        */
LB_API int lb_set_temporal(lb_lattice *lat, int steps_per_pass, int rows_per_tile);
/* One phase (1, 2, 3) of a temporal-blocking double step, for drivers that run several blocks on ONE
 * stream: phase 3 of a block waits for phase 1 of its neighbours, so the phases of all blocks must be
 * interleaved.  lb_temporal_active tells whether lb_step uses double steps for this lattice.       */
LB_API int lb_double_step_phase(lb_lattice *lat, int phase);
LB_API int lb_temporal_active(lb_lattice *lat);
LB_API int lb_temporal_rows(lb_lattice *lat);
/* lb_step replays a CUDA graph of 64 fused steps for long runs (default on).                */
LB_API int lb_set_use_graph(lb_lattice *lat, int on);
/* L2-resident lattices (a single self-connected block of at most 2^20 cells that does not use temporal
 * blocking): lb_step(n >= 2) advances all n steps in ONE cooperative launch with a grid-wide barrier per step
 * -- shear probe, Couette's collide-first order and Poiseuille's pressure columns run inside it -- instead of
 * one (simple_flows: three) launches per step.  Bit-identical; default on; LBM_RESIDENT=0|1 overrides.       */
LB_API int lb_set_resident(lb_lattice *lat, int on);
/* LB_SF_TABLE: the walls of the simple_flows family as a per-cell gather table.  The reference writes them as a
 * sequence of overlapping slice assignments on the streamed array (slidingLidMPI.py:180-204, obstacle_canal.py:
 * 413-458); replayed symbolically on the host (latticeboltzmann_b200/boundary_table.py) that sequence becomes, for
 * each of the `n` listed cells (flat index k*lny + l), the pre-stream source of every population (flat element
 * index i*lnx*lny + k*lny + l into the (9, lnx, lny) state, src[9*j + i]) and an additive constant (add[9*j + i]):
 *     post[i, cell_j] = pre[src[9 j + i]] + add[9 j + i]
 * All other cells stream periodically; then moments and collision (simple_flows arithmetic, fp64, one block).
 * The table is copied to the device; calling it again replaces the table.                                        */
LB_API int lb_set_boundary_table(lb_lattice *lat, int64_t n, const int64_t *cells, const int64_t *src, const double *add);
/* Geometry queries (elements). */
LB_API int64_t lb_pitch(lb_lattice *lat);
LB_API int64_t lb_pop_stride(lb_lattice *lat);
LB_API int lb_kernel_launches(lb_lattice *lat, int64_t *count);

/* ---- (1) stateless reference-API entry points, HOST buffers ---------------------- */
/* c/_lbkernels.cpp:36-37,43-44  equilibrium(rho, ux, uy) -> 9 values            */
LB_API int lbk_equilibrium1_f32(float rho, float ux, float uy, float *out9);
LB_API int lbk_equilibrium1_f64(double rho, double ux, double uy, double *out9);
/* c/_lbkernels.cpp:38-39,45-46  equilibrium(rho[N], ux[N], uy[N], f[9,N])       */
LB_API int lbk_equilibriumn_f32(const float *rho, const float *ux, const float *uy, float *f, int64_t n);
LB_API int lbk_equilibriumn_f64(const double *rho, const double *ux, const double *uy, double *f, int64_t n);
/* c/_lbkernels.cpp:40-41,47-48  collide(f[9,N], omega), in place                */
LB_API int lbk_collide_f32(float *f, int64_t n, float omega);
LB_API int lbk_collide_f64(double *f, int64_t n, double omega);
/* PyLB/Streaming.py:33-46       stream(f[9,nx,ny]), in place periodic roll      */
LB_API int lbk_stream_f32(float *f, int64_t nx, int64_t ny);
LB_API int lbk_stream_f64(double *f, int64_t nx, int64_t ny);
/* The opt2 loop body on a host array (cavity_opt2.py:275-277 on one rank /
 * shear_wave_opt2.py:96-97): nsteps x { stream (+ bounce back) ; collide },
 * f[9,nx,ny] in place.  boundary is LB_PERIODIC, LB_CAVITY or LB_CAVITY_XPERIODIC. */
LB_API int lbk_step_host_f32(float *f, int64_t nx, int64_t ny, int boundary, float omega, float u0, int64_t nsteps);
LB_API int lbk_step_host_f64(double *f, int64_t nx, int64_t ny, int boundary, double omega, double u0, int64_t nsteps);

/* Self-test hook: evaluates the kernel's 3-operation correctly-rounded x/9 and x/6 on the given
 * bit patterns and counts results that differ from the IEEE division (must be 0).          */
LB_API int lbk_selftest_div_const_f64(const uint64_t *bits, int64_t n, int64_t *mismatches);
LB_API int lbk_selftest_div_const_f32(const uint32_t *bits, int64_t n, int64_t *mismatches);

#ifdef __cplusplus
}
#endif
#endif /* LBM_B200_H */
