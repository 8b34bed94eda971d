/*
 * libfdtd2d -- C ABI of the B200-native 2D FDTD (TM-mode, Yee grid) time-stepping engine.
 *
 * This is the drop-in boundary for ONE path of the reference project
 * skunnavakkam/fdtd-2d: the leapfrog loop of python-src/fdtd.py:30-34, whose arithmetic is
 * python-src/main.py:12-76 (update_Ez, update_Hx_Hy) and :182-195 (sources).  The reference has no
 * FFI of its own (it is plain numpy); the entry points below are what a ctypes/cffi binding of that
 * path binds.  Each group cites the reference interface it replaces.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes; no C++/torch types cross this boundary.
 *   - Every function returns 0 on success or a negative FDTD2D_E* code; fdtd2d_last_error() returns a
 *     thread-local human-readable message for the last failure on the calling thread.
 *   - A handle owns its device buffers and is not thread-safe.  All calls are ordered on the handle's
 *     CUDA stream; only *_download*, *_read* and fdtd2d_sync block the host.
 *   - Host arrays use the reference's three shapes (main.py:79-85), C-contiguous, in the handle's
 *     dtype, one set per batch grid (batch-major):  Ez[R][C], Hx[R][C-1], Hy[R-1][C].
 *   - There is NO CPU fallback: every compute entry point fails with FDTD2D_ECUDA when no CUDA device
 *     is usable.
 */
#ifndef FDTD2D_H
#define FDTD2D_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FDTD2D_ABI_VERSION 1

/* dtype codes */
#define FDTD2D_F32 0
#define FDTD2D_F64 1

/* error codes */
#define FDTD2D_OK 0
#define FDTD2D_EINVAL (-1) /* bad argument (shape < 11, null pointer, k out of range, ...) */
#define FDTD2D_ECUDA (-2)  /* CUDA runtime error / no device */
#define FDTD2D_ENOMEM (-3) /* device or host allocation failed */
#define FDTD2D_ESTATE (-4) /* call sequence error (e.g. step before coefficients are set) */

/* phase bits for fdtd2d_step_phases (used by the per-function parity entry points) */
#define FDTD2D_PHASE_H 1   /* main.py:66-76  */
#define FDTD2D_PHASE_E 2   /* main.py:12-63 (interior + Mur + corners) */
#define FDTD2D_PHASE_SRC 4 /* fdtd.py:34 (source add) and probe sampling */

typedef struct fdtd2d_sim fdtd2d_sim;

/* ---- library ------------------------------------------------------------------------------ */
int fdtd2d_abi_version(void);
const char* fdtd2d_last_error(void);
int fdtd2d_device_count(int* count);

/* ---- handle: replaces grid_init (main.py:79-85) -------------------------------------------- */
/* `batch` independent rows x cols grids in `dtype` on CUDA device `device`; zero state.
 * rows, cols >= 11 (the staged Mur/corner dataflow equals the reference only from 11). */
int fdtd2d_create(fdtd2d_sim** out, int rows, int cols, int dtype, int device, int batch);
/* One y-slab of a global_rows x cols grid: this handle owns global rows [row_begin, row_end) and
 * keeps `halo` ghost rows on each side that has a neighbour slab. batch = 1. */
int fdtd2d_create_slab(fdtd2d_sim** out, int global_rows, int cols, int row_begin, int row_end, int halo,
                       int dtype, int device);
int fdtd2d_destroy(fdtd2d_sim* s);
/* Order this handle's work on an external CUDA stream (cudaStream_t passed as void*; NULL is the
 * legacy default stream).  fdtd2d_reset_stream goes back to the handle's own non-blocking stream. */
int fdtd2d_set_stream(fdtd2d_sim* s, void* cuda_stream);
int fdtd2d_reset_stream(fdtd2d_sim* s);
int fdtd2d_sync(fdtd2d_sim* s);

/* geometry queries: local_rows includes ghost rows; row0 = global index of local row 0 */
int fdtd2d_geometry(const fdtd2d_sim* s, int* local_rows, int* cols, int* row0, int* global_rows, int* batch,
                    int* dtype, size_t* pitch_elems);

/* ---- state: the three arrays grid_init returns / the kernels mutate -------------------------- */
/* Host -> device.  For a slab handle the arrays cover the LOCAL rows (ghost rows included):
 * Ez[Rl][C], Hx[Rl][C-1], Hy[Rl or Rl-1][C] (Rl-1 only when the slab holds the global last row). */
int fdtd2d_upload_state(fdtd2d_sim* s, const void* Ez, const void* Hx, const void* Hy);
int fdtd2d_download_state(fdtd2d_sim* s, void* Ez, void* Hx, void* Hy);
int fdtd2d_zero_state(fdtd2d_sim* s);

/* ---- materials: replaces the per-step dt/(eps*dx), dt/(mu*dx), Mur coef (main.py:27,30-31,70,74) */
/* Host-precomputed maps in the run dtype, (R, C) each, batch-major; mur_coef: one scalar per grid. */
int fdtd2d_set_coeffs(fdtd2d_sim* s, const void* ce, const void* ch, const void* mur_coef);
/* eps/mu maps ((R, C) each, run dtype, batch-major) -> device forms ce = dt/(eps*dx), ch = dt/(mu*dx)
 * and the Mur coefficient from cell (0,0) with the reference's operation order, bit-identically. */
int fdtd2d_set_materials(fdtd2d_sim* s, const void* eps, const void* mu, double dt, double dx);
/* Mur coefficient(s) only (one scalar per grid, run dtype).  Needed by slab handles that do not hold
 * global cell (0,0): fdtd2d_set_materials leaves their coefficient unset. */
int fdtd2d_set_mur_coef(fdtd2d_sim* s, const void* mur_coef);
/* Synthetic medium generated on the device (for grids too large for host numpy):
 * eps = eps0*(1 + span*u), u in [0,1) from a counter-based hash of (seed, grid, global row, col);
 * mu = mu0.  eps0 = 8.85418e-12, mu0 = 4*pi*1e-7 (main.py:100-101). See fdtd2d_hash_uniform. */
int fdtd2d_set_materials_random(fdtd2d_sim* s, uint64_t seed, double span, double dt, double dx);
/* Structure image -> materials on the device (main.py:109-123): gray is the uint8 "L" image already resized to
 * (R_local, C) per grid (PIL does that on the host, as in the reference); the device forms
 * eps = (1 + (black_point - 1) * (1 - gray/255)) * eps0 in float64, casts to the run dtype, sets mu = mu0 and
 * then the coefficient maps exactly as fdtd2d_set_materials does.  1 byte per cell crosses PCIe. */
int fdtd2d_set_materials_gray(fdtd2d_sim* s, const unsigned char* gray, double black_point, double dt, double dx);
/* Random two-phase media of the dataset generator (diffusion_training.py:54-93), one per batch grid, generated
 * on the device: u = fdtd2d_hash_uniform(seed, grid, row, col) blurred with the grid's 15 x 15 kernel
 * weights[grid][15*15] (float32, zero padding, row-major accumulation, no FMA), eps = blur > 0.5 ? eps_hi : eps_lo,
 * mu uniform; then the coefficient maps as fdtd2d_set_materials.  eps_out (optional, host, (R, C) per grid in the
 * run dtype) receives the permittivity maps.  Whole grids only (not slabs). */
int fdtd2d_generate_materials_blobs(fdtd2d_sim* s, uint64_t seed, const float* weights, double eps_lo, double eps_hi, double mu,
                                    double dt, double dx, void* eps_out);
/* Host-side definition of the generator above (so a CPU checker can rebuild the same map). */
double fdtd2d_hash_uniform(uint64_t seed, uint32_t grid, uint32_t row, uint32_t col);
/* Read back the coefficient maps ((R_local, C), run dtype, batch-major) and per-grid Mur coefs. */
int fdtd2d_download_coeffs(fdtd2d_sim* s, void* ce, void* ch, void* mur_coef);

/* ---- sources: replaces `Ez += ricker(rows, cols, r, c, i*dt, fc)` (fdtd.py:34, main.py:182-195) -- */
/* n_cells source cells; cell q lives in grid grid[q] at GLOBAL (row[q], col[q]) and adds
 * tables[wave[q]][i] (float64) at step index i:  Ez = (T)((double)Ez + amp)  -- the reference's
 * float64 add then cast.  tables is [n_waves][n_steps], evaluated by the host with numpy exactly as
 * main.py:183-184/193-194 do.  Steps with index >= n_steps add nothing.  n_cells = 0 clears. */
int fdtd2d_set_sources(fdtd2d_sim* s, int n_cells, const int32_t* grid, const int32_t* row, const int32_t* col,
                       const int32_t* wave, int n_waves, int n_steps, const double* tables);

/* ---- probes (field readout at fixed cells after every step) ------------------------------------ */
/* capacity_steps rows of n_probes samples are kept on the device. */
int fdtd2d_set_probes(fdtd2d_sim* s, int n_probes, const int32_t* grid, const int32_t* row, const int32_t* col,
                      int capacity_steps);
/* out[n_steps][n_probes] in the run dtype, for step indices first_step .. first_step+n_steps-1. */
int fdtd2d_read_probes(fdtd2d_sim* s, void* out, int64_t first_step, int n_steps);

/* ---- time stepping: replaces the loop body fdtd.py:31-34 ------------------------------------- */
/* n_steps leapfrog steps (H -> Ez+Mur+corners -> source -> probe sample), k_temporal steps per HBM
 * round trip (1 <= k <= FDTD2D_MAX_K; for slabs k <= halo and the caller exchanges halos every k).
 * k_temporal = 0 picks the library default (fp64: 4; fp32: 8).  k = 12 has a row-streaming wavefront instance for
 * grids with uniform permeability; on B200 it is no faster than k = 8 (latency-bound at 255 registers), so it is not
 * chosen automatically. */
#define FDTD2D_MAX_K 12
int fdtd2d_step(fdtd2d_sim* s, int n_steps, int k_temporal);
/* One pass applying only the phases in `phases` (FDTD2D_PHASE_*); with PHASE_H alone it is
 * update_Hx_Hy (main.py:66-76), with PHASE_E alone update_Ez (main.py:12-63).  Does not advance the
 * step counter unless PHASE_SRC is included. */
int fdtd2d_step_phases(fdtd2d_sim* s, int phases);
int fdtd2d_get_step_index(const fdtd2d_sim* s, int64_t* step);
int fdtd2d_set_step_index(fdtd2d_sim* s, int64_t step);
/* Select the tile kernel: 0 = automatic, 1 = generic shared-memory tiles only,
 * 2 = register-resident tiles everywhere (TMA-fed plain tiles + edge-capable tiles),
 * 3 = register-resident plain tiles + shared-memory generic kernel for edge/source/probe tiles,
 * 4 = cluster-resident kernel (fp32 grids of 16..256 columns and 16..384 rows stay on chip for a whole
 *     fdtd2d_step call, one thread-block cluster per grid; k_temporal does not apply) or FDTD2D_EINVAL.
 * Automatic = 4 when the grid is eligible, else 2. */
int fdtd2d_set_kernel_variant(fdtd2d_sim* s, int variant);
/* Number of stepping passes so far: one per HBM round trip of the fields (k leapfrog steps of the tile / wavefront
 * kernels, or a whole fdtd2d_step call of the cluster-resident kernel).  bench.py's roofline divides by it. */
int fdtd2d_pass_count(const fdtd2d_sim* s, int64_t* passes);
/* Host-only planning helper (no GPU needed; exported so that the CPU tests cover it): how the row-streaming wavefront
 * kernel cuts `n_stretches` vertical stretches of rows[i] rows (ring[i] != 0: a strip that carries the left / right Mur
 * ring, whose rows cost about twice as much) into runs for `warps` independent warps.  parts[i] receives the number of
 * (nearly equal) runs of stretch i, *run_rows the plain run length chosen.  No reference counterpart. */
int fdtd2d_plan_wave_runs(int n_stretches, const int32_t* rows, const uint8_t* ring, int warps, int cap_rows, int k,
                          int32_t* parts, int32_t* run_rows);
/* Number of kernel launches issued by this handle so far (for bench.py's gpu_launches). */
int fdtd2d_launch_count(const fdtd2d_sim* s, int64_t* launches);

/* ---- field readout as an image: replaces capture_snapshot (main.py:153-179) -------------------------- */
/* gray: the uint8 permittivity background (main.py:157-165), (R_local, C) per grid, computed once on the
 * host; lut: 256 x 3 float64 colormap entries (the reference uses matplotlib's "seismic").  */
int fdtd2d_set_snapshot_background(fdtd2d_sim* s, const unsigned char* gray, const double* lut);
/* Renders Ez of grid `grid` on the device (clip to [vmin, vmax], colormap, alpha 0.7 over the background,
 * uint8) and copies the (R_local, C, 3) RGB frame to out_rgb.  Blocks until the frame is on the host. */
int fdtd2d_render_snapshot(fdtd2d_sim* s, int grid, double vmin, double vmax, unsigned char* out_rgb);

/* ---- multi-GPU y-slabs -------------------------------------------------------------- */
/* Device pointers and byte counts of the halo blocks of the CURRENT state, for NCCL send/recv or
 * peer copies driven by the host layer.  field: 0 = Ez, 1 = Hx, 2 = Hy.  side: 0 = top (towards
 * smaller rows), 1 = bottom.  send_ptr = the `halo` owned rows next to that side; recv_ptr = the
 * ghost rows on that side.  Blocks are contiguous: halo * pitch elements. */
int fdtd2d_halo_block(fdtd2d_sim* s, int field, int side, void** send_ptr, void** recv_ptr, size_t* nbytes);
/* Split pass for overlapping the halo exchange with compute.  fdtd2d_pass_begin launches only the tiles
 * whose stores produce the rows a neighbour needs (or touch ghost rows) for a k-step pass;
 * fdtd2d_halo_block_next gives the halo blocks of the state being WRITTEN by that pass, so the caller
 * can start the exchange on another stream as soon as pass_begin's work is done; fdtd2d_pass_end launches
 * the remaining tiles, makes the new state current and advances the step index by k. */
int fdtd2d_pass_begin(fdtd2d_sim* s, int k);
int fdtd2d_pass_end(fdtd2d_sim* s);
int fdtd2d_halo_block_next(fdtd2d_sim* s, int field, int side, void** send_ptr, void** recv_ptr, size_t* nbytes);
/* Raw device pointer of a field of the current state (local_rows x pitch elements per grid). */
int fdtd2d_device_field(fdtd2d_sim* s, int field, void** ptr);

#ifdef __cplusplus
}
#endif
#endif /* FDTD2D_H */
