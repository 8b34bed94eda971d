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
#define FDTD2D_EINVAL (-1) /* bad argument (shape < 6, null pointer, k out of range, ...) */
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
 * rows, cols >= 6: the smallest grid on which the reference's own boundary code (main.py:33-61) indexes inside its
 * arrays.  Below 11 the five-deep Mur strips of opposite sides overlap and the step is executed statement by
 * statement in the reference's order (one CTA per grid); from 11 on the staged dataflow of SURVEY A.3 applies. */
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
/* The stream the handle's work is currently ordered on (cudaStream_t as void*). */
int fdtd2d_get_stream(const fdtd2d_sim* s, void** cuda_stream);
/* Waits for the handle's stream (and its copy stream); on a slab with peer links also until both neighbours have
 * delivered the ghost rows of the current state, and reports a halo wait that timed out inside a kernel (FDTD2D_ESTATE). */
int fdtd2d_sync(fdtd2d_sim* s);

/* ---- tuning options (no reference counterpart) ----------------------------------------------------------- */
/* A handle copies its options when it is created; the defaults come from the environment variables FDTD2D_<KEY>
 * (upper case), read once at that moment.  fdtd2d_set_option changes one option of this handle and drops its cached
 * plans.  Keys: "wavefront" (1), "wave_min_tiles" (-1 = automatic), "ring_min_tiles" (-1), "ring_strips" (1),
 * "wave_run_rows" (640), "edge_reserve" (-1 = automatic: whole grids whose edge tiles are 2 .. 25 % of a pass leave some SMs to
 * them and launch the wavefront first; 0 = never, n = that many SMs), "ring_cost" (0 = automatic; percent of a plain row that
 * a ring-strip row costs when runs are balanced), "auto_k12" (0), "uniform_ch" (1), "resident" (1), "resident_cfg" (5), "resident_cluster" (0),
 * "resident_trim" (-1), "resident_rows" (0), "tma_pair" (0), "f64_k" (0 = automatic), "fuse" (0; 1 = two k = 8 passes per launch, the second fed from L2: an
 * experiment that is bit-exact but slower on B200; -1 = on for large grids), "stage" (0; 4 / 5 = k = 12 passes on the staged
 * wavefront, the other experiment), "debug" (0).  Every setting is covered by the parity tests: options change which kernel
 * runs, never a result bit -- with ONE exception, "measure_skip" (0), a timing aid that leaves parts of a pass out (1 = no
 * edge tiles, 2 = no runs / plain tiles, 4 = edge tiles not overlapped with the runs): fields are wrong while it is set. */
int fdtd2d_set_option(fdtd2d_sim* s, const char* key, int value);
int fdtd2d_get_option(const fdtd2d_sim* s, const char* key, int* value);

/* geometry queries: local_rows includes ghost rows; row0 = global index of local row 0 */
int fdtd2d_geometry(const fdtd2d_sim* s, int* local_rows, int* cols, int* row0, int* global_rows, int* batch,
                    int* dtype, size_t* pitch_elems);

/* ---- state: the three arrays grid_init returns / the kernels mutate -------------------------- */
/* Host -> device.  For a slab handle the arrays cover the LOCAL rows (ghost rows included):
 * Ez[Rl][C], Hx[Rl][C-1], Hy[Rl or Rl-1][C] (Rl-1 only when the slab holds the global last row). */
int fdtd2d_upload_state(fdtd2d_sim* s, const void* Ez, const void* Hx, const void* Hy);
int fdtd2d_download_state(fdtd2d_sim* s, void* Ez, void* Hx, void* Hy);
/* grid_init (main.py:79-85) on the device: all three fields zero, step index 0.  Queued on the handle's stream (never
 * blocks a handle that was synchronised after its last step); on a slab with peer links it first waits, if need be, until
 * the neighbours' last ghost rows have arrived, and clears the current field set only (the other set's ghost rows are the
 * neighbours' to write). */
int fdtd2d_zero_state(fdtd2d_sim* s);
/* Non-blocking forms for PINNED host arrays: the copy is ordered after the stepping work issued so far and before the
 * stepping work issued afterwards, on the handle's own copy stream, and the call returns at once -- one host thread
 * can keep the copies of one handle overlapped with the kernels of another.  The host arrays must stay valid until
 * fdtd2d_copy_wait (or fdtd2d_sync) returns. */
int fdtd2d_upload_state_async(fdtd2d_sim* s, const void* Ez, const void* Hx, const void* Hy);
int fdtd2d_download_state_async(fdtd2d_sim* s, void* Ez, void* Hx, void* Hy);
int fdtd2d_copy_wait(fdtd2d_sim* s);

/* ---- materials: replaces the per-step dt/(eps*dx), dt/(mu*dx), Mur coef (main.py:27,30-31,70,74) */
/* Host-precomputed maps in the run dtype, (R, C) each, batch-major; mur_coef: one scalar per grid. */
int fdtd2d_set_coeffs(fdtd2d_sim* s, const void* ce, const void* ch, const void* mur_coef);
/* eps/mu maps ((R, C) each, run dtype, batch-major) -> device forms ce = dt/(eps*dx), ch = dt/(mu*dx)
 * and the Mur coefficient from cell (0,0) with the reference's operation order, bit-identically. */
int fdtd2d_set_materials(fdtd2d_sim* s, const void* eps, const void* mu, double dt, double dx);
/* fdtd2d_set_materials without blocking the host (pinned eps / mu, see fdtd2d_upload_state_async).  The upload, the
 * coefficient kernels and the uniform-permeability check all run on the handle's copy stream (its next stepping work
 * waits for them), so on a compute stream shared with other handles nothing of this call queues behind their kernels. */
int fdtd2d_set_materials_async(fdtd2d_sim* s, const void* eps, const void* mu, double dt, double dx);
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
/* Structure drawing on the device: replaces RegionDrawer (region_drawer.py:5-87, PIL ImageDraw on an "L" canvas) together
 * with the image -> permittivity mapping of material_init (main.py:109-123), so that a structure for a grid too large for
 * the host never leaves the GPU.  The canvas has one uint8 cell per grid cell (x = column, y = GLOBAL row; a slab handle
 * draws its local rows of the global picture), starts white (255 = eps0) and is drawn black (0 = black_point * eps0).
 *   fdtd2d_canvas_rect     inclusive rectangle: what ImageDraw.line(width) paints for a horizontal / vertical segment;
 *   fdtd2d_canvas_ellipse  width <= 0: ImageDraw.ellipse(box, fill) bit for bit (Pillow's integer quarter walk);
 *                          width > 0: the ring between that ellipse and the one of the box shrunk by `width` on each side
 *                          (ImageDraw.ellipse(box, outline, width) up to a few cells along the inner edge);
 *   fdtd2d_canvas_segment  slanted thick segment: the cells whose centre lies in the rectangle of that width;
 *   fdtd2d_canvas_apply    canvas -> eps = (1 + (black_point - 1)(1 - g/255)) eps0 (float64, cast to the run dtype),
 *                          mu = mu0, then the coefficient maps and the Mur coefficient as fdtd2d_set_materials. */
int fdtd2d_canvas_clear(fdtd2d_sim* s, int value);
int fdtd2d_canvas_rect(fdtd2d_sim* s, int grid, int x0, int y0, int x1, int y1, int value);
int fdtd2d_canvas_ellipse(fdtd2d_sim* s, int grid, int x0, int y0, int x1, int y1, int width, int value);
int fdtd2d_canvas_segment(fdtd2d_sim* s, int grid, double x0, double y0, double x1, double y1, double width, int value);
int fdtd2d_canvas_download(fdtd2d_sim* s, unsigned char* gray);
int fdtd2d_canvas_apply(fdtd2d_sim* s, double black_point, double dt, double dx);
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
 * round trip (1 <= k <= FDTD2D_MAX_K; for slabs k <= halo, and without peer links the caller exchanges halos every k).
 * k_temporal = 0 picks the library default (fp32: 8; fp64: 8 where the wavefront kernel takes the grid, else 4).  k = 12 has a row-streaming wavefront instance for
 * grids with uniform permeability; on B200 it is no faster than k = 8 (latency-bound at 255 registers), so it is not
 * chosen automatically. */
#define FDTD2D_MAX_K 12
int fdtd2d_step(fdtd2d_sim* s, int n_steps, int k_temporal);
/* One pass applying only the phases in `phases` (FDTD2D_PHASE_*); with PHASE_H alone it is
 * update_Hx_Hy (main.py:66-76), with PHASE_E alone update_Ez (main.py:12-63).  Does not advance the
 * step counter unless PHASE_SRC is included. */
int fdtd2d_step_phases(fdtd2d_sim* s, int phases);
/* How far the handle can be stepped before its source tables / probe traces run out: number of source cells, steps
 * per waveform table (sources add nothing from that step index on), rows of the probe trace (0 = no probes). */
int fdtd2d_source_steps(const fdtd2d_sim* s, int* n_cells, int* n_steps, int64_t* probe_capacity);
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
/* Host-only: how many SMs the wavefront kernel of a k-step pass leaves to the pass's `n_edge` edge tiles when its stretches
 * hold `wave_rows` rows (ring-strip rows weighted by their cost); 0 = none (edge tiles below 2 % or above 25 % of the pass).  No
 * reference counterpart. */
int fdtd2d_plan_edge_reserve(int64_t n_edge, int64_t wave_rows, int sm_count, int k);
/* Host-only: how the cluster-resident kernels would split a whole fp32 grid of rows x cols (cfg: option resident_cfg,
 * cluster: option resident_cluster, 0 = as few CTAs as fit).  out[4] = kernel shape used (5 = the packed kernel), CTAs per
 * grid, rows of the first band, rows of the other bands (the last band takes what is left); all -1: not eligible. */
int fdtd2d_plan_resident(int rows, int cols, int cfg, int cluster, int32_t* out);
/* Host-only: the whole plan of a k-step pass for a geometry, without a handle or a GPU (the CPU tests check that every
 * owned cell is produced exactly once, for whole grids and slabs, fp32 and fp64).  geom[15] = dtype, batch, global rows,
 * cols, row_begin, row_end, halo, k, SM count, kernel variant, wave_min_tiles, ring_min_tiles, wavefront, ring_strips,
 * uniform permeability (1 / 0).  src / probe: n x (grid, global row, col).  plan[16] receives tile rows, tile columns,
 * core rows, core columns, column halo, first owned local row, local rows, edge tiles, of which band tiles, TMA tiles,
 * wavefront runs, of which band runs, ring strips present, band tasks top / bottom, pitch; tile_kind[tiles] (optional)
 * 0 edge / 1 wavefront / 2 ring strip / 3 TMA per tile; tasks[runs][8] (optional) the runs: grid, first column, first
 * row, end row, first / end stored column of the strip, ring side, band. */
int fdtd2d_plan_host(const int32_t* geom, int n_src, const int32_t* src, int n_probe, const int32_t* probe, int32_t* plan,
                     int32_t* tile_kind, int cap_tiles, int32_t* tasks, int cap_tasks);
/* Host-only: the task lists of a FUSED pair of k = 8 passes of a whole fp32 grid (geom as above, rows = the whole grid;
 * fuse: -1 automatic, 0 off, 1 on).  counts[4] = runs in the fused launch, runs left to the second launch, 16-row blocks
 * per tile column, runs of a single pass.  fused / deferred (optional): 12 words per run -- grid, first column, first
 * row, end row, first / end stored column, ring side, band, phase (0 first pass, 1 second), tile column stored, first
 * and last tile column read. */
int fdtd2d_plan_host_fused(const int32_t* geom, int fuse, int n_src, const int32_t* src, int n_probe, const int32_t* probe, int32_t* counts,
                           int32_t* fused, int cap_fused, int32_t* deferred, int cap_deferred);
/* What a k-step pass of this handle consists of (builds and caches the plan; needs the materials): info[0..12] =
 * tile rows, tile columns, core rows, core columns, edge tiles, of which band tiles, TMA tiles, wavefront runs, of
 * which band runs, ring strips present, band tasks next to the top / bottom neighbour, SMs the wavefront kernel leaves
 * to the edge tiles. */
#define FDTD2D_PLAN_INFO_WORDS 13
int fdtd2d_plan_info(fdtd2d_sim* s, int k, int32_t* info, int n_info);
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
/* Peer links: the halo exchange INSIDE the stepping kernels.  Each slab exports a blob (its geometry, device
 * pointers and CUDA IPC handles of its two field sets and its flag block); after fdtd2d_peer_attach of the blobs of
 * its neighbours (side 0 = top, 1 = bottom; the same or another process on the same box) the tasks of a pass that
 * produce the `halo` rows next to a neighbour store them straight into that neighbour's ghost rows (peer stores over
 * NVLink), the last one raises the neighbour's flag, and the tasks that read ghost rows wait for their own flag -- no
 * host work and no collective per pass, and fdtd2d_step may advance any number of steps.  Rules: attach every
 * neighbour before the first step; every slab steps the same sequence of passes; one slab per GPU runs its passes on
 * its own stream, slabs that SHARE a GPU must share one stream and be stepped pass by pass in turn (a kernel that
 * waited for a kernel queued behind it on the same GPU would never finish; the wait gives up after 2 s and the next
 * fdtd2d_sync reports it); synchronise all slabs (fdtd2d_sync on each, then a barrier between the processes) before
 * uploading a new state, detaching or destroying. */
#define FDTD2D_PEER_BLOB_BYTES 640
int fdtd2d_peer_export(fdtd2d_sim* s, void* blob);
int fdtd2d_peer_attach(fdtd2d_sim* s, int side, const void* blob);
int fdtd2d_peer_detach(fdtd2d_sim* s);
/* out[0] = attached sides (bit 0 top, bit 1 bottom), out[1] = passes stepped with a peer attached, out[2], out[3] =
 * newest state delivered by the top / bottom neighbour, out[4] != 0: a halo wait timed out. */
int fdtd2d_peer_status(fdtd2d_sim* s, uint32_t* out);
/* Raw device pointer of a field of the current state (local_rows x pitch elements per grid). */
int fdtd2d_device_field(fdtd2d_sim* s, int field, void** ptr);

#ifdef __cplusplus
}
#endif
#endif /* FDTD2D_H */
