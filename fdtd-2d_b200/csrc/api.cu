// libfdtd2d: C-ABI implementation (include/fdtd2d.h) over the sm_100a kernels.
// Host-side logic only: handle lifetime, padded device layout, tile / run planning, launches, the peer links of y-slabs.
// There is no CPU compute path in this file or anywhere in the library.
#include <algorithm>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <thread>
#include <type_traits>
#include <vector>

#include <unistd.h>

#include "../../include/fdtd2d.h"
#include "common.cuh"
#include "options.h"
#include "tile_edge.cuh"
#include "tile_tma.cuh"
#include "tile_generic.cuh"
#include "grid_resident.cuh"
#include "grid_resident_x2.cuh"
namespace fdtd2d {
// resident_x2.cu (a translation unit of its own: the build compiles the two side by side) holds the twelve instantiations of
// the packed cluster-resident kernel and this launcher; api.cu never names the kernel template.
cudaError_t resident_x2_launch(bool uch, int rl, const cudaLaunchConfig_t* cfg, const PassParams<float>& p, float ch_uniform, bool set_smem_attr,
                               size_t smem, int* max_active_clusters);
}  // namespace fdtd2d
#include "strip_wave.cuh"
#include "strip_stage.cuh"
#include "grid_small.cuh"
#include "structure.cuh"

using namespace fdtd2d;

// ------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------
static thread_local std::string g_err;

static int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CUDA_TRY(expr)                                                                      \
    do {                                                                                    \
        cudaError_t e_ = (expr);                                                            \
        if (e_ != cudaSuccess)                                                              \
            return fail(e_ == cudaErrorMemoryAllocation ? FDTD2D_ENOMEM : FDTD2D_ECUDA,     \
                        "%s failed: %s", #expr, cudaGetErrorString(e_));                   \
    } while (0)

#define REQUIRE(cond, ...)                                \
    do {                                                  \
        if (!(cond)) return fail(FDTD2D_EINVAL, __VA_ARGS__); \
    } while (0)

// ------------------------------------------------------------------------------------------------
// tile geometry
// ------------------------------------------------------------------------------------------------
constexpr int G_TH = 36, G_TW = 128;  // the shared-memory generic kernel (per-function passes, variant 1)
// threads per generic CTA: tiles that fill most of an SM's shared memory (1 CTA/SM) get 1024 threads
template <typename T, int TH> constexpr int generic_nt() { return 6 * TH * G_TW * sizeof(T) > 113 * 1024 ? 1024 : 512; }
// register-resident tile kernels: MR rows per thread x NW warps -> (MR * NW) x 128 tiles
constexpr int F_MR = 4, F_NW = 16, F_TH = F_MR * F_NW;  // fp32: 64 x 128 (edge tiles, persistent TMA-fed plain tiles)
constexpr int D_MR = 2, D_NW = 16, D_TH = D_MR * D_NW;  // fp64: 32 x 128 (edge tiles)
constexpr int TILE_TW = 128;
constexpr int MIN_LAST = 8;  // smallest core extent allowed for the last tile row/column (ring safety)

struct TilePlan {
    int k = 0, hx = 0, CH = 0, CW = 0, tiles_y = 0, tiles_x = 0;
};

// One cached pass configuration for k steps: the tile grid, which tiles need the edge-capable kernel (Mur ring, sources,
// probes, array edges, a slab's band) and how the rest is covered -- runs of the wavefront kernel or the TMA tile kernel.
struct PassPlan {
    bool valid = false;
    TilePlan tp;
    int n_edge = 0, n_edge_band = 0;  // edge tiles; the first n_edge_band hold band rows of a slab
    int n_fast = 0;                   // plain tiles of the TMA kernel (never band tiles)
    int* d_edge = nullptr;
    int* d_fast = nullptr;
    int n_wave = 0, n_wave_band = 0;  // wavefront runs; the first n_wave_band are band runs
    WaveTask* d_wave = nullptr;
    bool wave_ring = false;           // the list holds runs of the strips with the left / right Mur ring
    int band_expected[2] = {0, 0};    // band tasks (runs + edge tiles) next to the top / bottom neighbour
    int* d_ticket = nullptr;          // next run to hand out (zero between launches: the last draw of a launch resets it)
    int reserve_sms = 0;              // SMs the wavefront kernel leaves to the edge tiles (its runs are cut for the others)
    // fused double pass (strip_wave.cuh): the runs of two consecutive passes in one ticket order; the phase-1 pieces next to
    // edge tiles wait for a second, short launch
    int n_fused = 0, n_deferred = 0, fuse_nblk = 0;
    WaveTask* d_fused = nullptr;
    WaveTask* d_deferred = nullptr;
    unsigned* d_fuse_flags = nullptr;
    size_t fuse_flag_bytes = 0;
};

// one neighbour slab as this process sees it
struct PeerLink {
    bool attached = false, ipc = false;
    void* field[2][3] = {{nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr}};
    unsigned* flags = nullptr;
    int row0 = 0, device = -1;
};

// ------------------------------------------------------------------------------------------------
// handle
// ------------------------------------------------------------------------------------------------
struct fdtd2d_sim {
    int dtype = 0, device = 0, batch = 1;
    int Rg = 0, C = 0, row_begin = 0, row_end = 0, halo = 0, row0 = 0, Rl = 0;
    bool has_top_nb = false, has_bot_nb = false;
    size_t esize = 4, pitch = 0, grid_elems = 0;
    void* field[2][3] = {{nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr}};
    void* ce = nullptr;
    void* ch = nullptr;
    void* mur = nullptr;
    int cur = 0;
    bool coeffs_set = false, mur_set = false;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    Options opt;
    // sources
    int n_src = 0, n_waves = 0, amp_steps = 0;
    Cell* d_src = nullptr;
    int* d_src_range = nullptr;
    double* d_amp = nullptr;
    // probes
    int n_probe = 0;
    Cell* d_probe = nullptr;
    int* d_probe_range = nullptr;
    void* d_trace = nullptr;
    long long trace_cap = 0;
    std::vector<int> probe_perm;  // sorted position -> caller's index
    long long step = 0, launches = 0, passes = 0;
    int variant = 0;
    cudaStream_t side_stream = nullptr;  // edge tiles run here, concurrently with the plain tiles / runs
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    PassPlan hybrid[FDTD2D_MAX_K + 1];
    std::vector<Cell> h_src, h_probe;  // host copies (sorted) for tile classification
    TmaMaps tma_maps[2];               // tensor maps with field set 0 / 1 as the pass input
    int tma_box_rows = 0;              // box height the maps were encoded for (0 = not built)
    int sm_count = 0;
    int open_pass_k = 0;  // > 0 between fdtd2d_pass_begin and fdtd2d_pass_end
    int ch_uniform = -1;  // dt/(mu*dx) the same in every cell? (-1 = not checked since the maps last changed)
    double ch_value = 0.0;
    int* d_flag = nullptr;
    // fdtd2d_set_materials_async runs the permeability check right behind the coefficient kernels on the copy stream and
    // parks the answer in pinned host memory: {flag, value bits}
    unsigned long long* h_check = nullptr;
    cudaEvent_t ev_check = nullptr;
    bool check_pending = false;
    int resident_ok = -1;  // cluster-resident kernel usable for this handle? (-1 = not decided yet)
    int resident_cluster = 0, resident_rpc = 0, resident_edge = 0, resident_cfg = 0;  // CTAs per grid, rows per middle / first CTA, kResCfgs index
    unsigned char* d_gray = nullptr;  // snapshot background (Rl x C per grid)
    unsigned char* d_rgb = nullptr;   // one rendered frame (Rl x C x 3)
    double* d_lut = nullptr;          // 256 x 3 colormap
    unsigned char* d_canvas = nullptr;  // structure canvas (Rl x C per grid), fdtd2d_canvas_*
    // y-slabs: flag block (common.cuh FLAG_*), the neighbours' buffers, passes stepped with a peer attached
    unsigned* d_slab_flags = nullptr;
    PeerLink peer[2];
    unsigned pass_seq = 0;
    unsigned settled_seq = 0;  // the newest state whose ghost rows are known to have arrived (peer_settle)
    long long fused_pairs = 0, fused_checked = 0;  // fused double passes launched / covered by the last look at the error flag
    // fdtd2d_*_async: copies run on a stream of their own, ordered against this handle's stepping work by two events
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_copy = nullptr;  // the last asynchronous copy
    cudaEvent_t ev_work = nullptr;  // the last work this handle put on its stream
    bool copy_pending = false;      // the next work on the stream has to wait for ev_copy first
};

static size_t round_up(size_t v, size_t m) { return (v + m - 1) / m * m; }

// Restores the caller's current device when an entry point returns (a multi-GPU host process must not find its
// device changed by a library call, least of all by a destroy that runs from a garbage collector).
struct DeviceGuard {
    int prev = -1, dev;
    cudaError_t err = cudaSuccess;
    explicit DeviceGuard(int d) : dev(d) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != dev) err = cudaSetDevice(dev);
    }
    ~DeviceGuard() {
        if (prev >= 0 && prev != dev) cudaSetDevice(prev);
    }
};
#define USE_DEVICE(s)               \
    DeviceGuard guard_((s)->device); \
    if (guard_.err != cudaSuccess) return fail(FDTD2D_ECUDA, "cudaSetDevice(%d) failed: %s", (s)->device, cudaGetErrorString(guard_.err))

// Work that reads or writes the handle's buffers is bracketed by these two: begin_work makes the stream wait for an
// asynchronous copy that is still in flight, mark_work remembers where this handle's work ends on a stream it may
// share with other handles (an asynchronous copy waits for that point, not for whatever else was queued since).
static int begin_work(fdtd2d_sim* s) {
    if (s->copy_pending) {
        CUDA_TRY(cudaStreamWaitEvent(s->stream, s->ev_copy, 0));
        s->copy_pending = false;
    }
    return 0;
}
static int mark_work(fdtd2d_sim* s) {
    if (s->ev_work) CUDA_TRY(cudaEventRecord(s->ev_work, s->stream));
    return 0;
}

static int sm_count(fdtd2d_sim* s) {
    if (!s->sm_count) cudaDeviceGetAttribute(&s->sm_count, cudaDevAttrMultiProcessorCount, s->device);
    return s->sm_count > 0 ? s->sm_count : 148;
}

static bool peer_mode(const fdtd2d_sim* s) { return s->peer[0].attached || s->peer[1].attached; }
static int own_first(const fdtd2d_sim* s) { return s->row_begin - s->row0; }  // local row of the first owned row
static int own_last(const fdtd2d_sim* s) { return s->row_end - s->row0; }     // one past the last owned row

static int plan_axis(int extent, int core_max, int quantum, int* core, int* tiles) {
    int c = core_max;
    while (c >= core_max / 2 && c > 0) {
        int t = (extent + c - 1) / c;
        int rem = extent - (t - 1) * c;
        if (t == 1 || rem >= MIN_LAST) {
            *core = c;
            *tiles = t;
            return 0;
        }
        c -= quantum;
    }
    return -1;
}

// The tile grid covers the OWNED rows of the handle (all rows unless it is a slab): a slab's ghost rows are only ever
// halo, so its first and last tile rows are as plain as any other.  cw_max > 0 caps the core width (fp64 wavefront:
// a tile column is two strips).
static int plan_tiles(const fdtd2d_sim* s, int k, int TH, TilePlan* tp, int col_quantum = 0, int cw_max = 0) {
    // columns are handled in groups: the 16-byte vector of the generic kernel, 4 cells per lane in the
    // register-resident kernels
    const int vn = col_quantum ? col_quantum : (int)(16 / s->esize);
    tp->k = k;
    tp->hx = (int)round_up((size_t)k, (size_t)vn);
    int cw = TILE_TW - 2 * tp->hx;
    if (cw_max > 0) cw = std::min(cw, cw_max);
    if (plan_axis(own_last(s) - own_first(s), TH - 2 * k, 1, &tp->CH, &tp->tiles_y) != 0 ||
        plan_axis(s->C, cw, vn, &tp->CW, &tp->tiles_x) != 0)
        return fail(FDTD2D_EINVAL, "cannot tile a %d x %d grid with k=%d", own_last(s) - own_first(s), s->C, k);
    return 0;
}

// ------------------------------------------------------------------------------------------------
// small kernels: coefficient maps and synthetic media
// ------------------------------------------------------------------------------------------------
// ce <- dt/(ce*dx), ch <- dt/(ch*dx) in place (the buffers hold eps and mu on entry), main.py:27,70,74.
template <typename T>
__global__ void coeff_from_materials_kernel(T* ce, T* ch, long long n, T dt, T dx) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        const T e = ce[i], m = ch[i];
        // padding cells hold 0: keep them 0 instead of dt/0 = inf
        ce[i] = (e == (T)0) ? (T)0 : div_rn(dt, mul_rn(e, dx));
        ch[i] = (m == (T)0) ? (T)0 : div_rn(dt, mul_rn(m, dx));
    }
}

// Mur coefficient from the corner cell's materials, main.py:30-31.
template <typename T> __device__ __forceinline__ T mur_from(T mu00, T eps00, T dt, T dx) {
    const T c = div_rn((T)1, sqrt_rn(mul_rn(mu00, eps00)));
    const T cdt = mul_rn(c, dt);
    return div_rn(sub_rn(cdt, dx), add_rn(cdt, dx));
}

// eps/mu of cell (0,0) of every grid are still in ce/ch when this runs (before the in-place transform).
template <typename T>
__global__ void mur_from_materials_kernel(const T* eps, const T* mu, long long grid_stride, int batch, T dt, T dx,
                                          T* mur) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < batch) mur[b] = mur_from(mu[b * grid_stride], eps[b * grid_stride], dt, dx);
}

// eps = eps0*(1 + span*u(seed, grid, global row, col)), mu = mu0, then the same transform as above.
template <typename T>
__global__ void random_materials_kernel(T* ce, T* ch, T* mur, int Rl, int C, int pitch, int row0, long long grid_stride,
                                        int batch, uint64_t seed, T span, T eps0, T mu0, T dt, T dx) {
    const long long per_grid = (long long)Rl * C;
    const long long n = per_grid * batch;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const T chv = div_rn(dt, mul_rn(mu0, dx));
    for (; i < n; i += stride) {
        const int b = (int)(i / per_grid);
        const long long r = i - (long long)b * per_grid;
        const int li = (int)(r / C), j = (int)(r - (long long)li * C);
        const T u = (T)hash_uniform(seed, (uint32_t)b, (uint32_t)(li + row0), (uint32_t)j);
        const T eps = mul_rn(eps0, add_rn((T)1, mul_rn(span, u)));
        const long long o = (long long)b * grid_stride + (long long)li * pitch + j;
        ce[o] = div_rn(dt, mul_rn(eps, dx));
        ch[o] = chv;
    }
    if (blockIdx.x == 0 && (int)threadIdx.x < batch && mur) {
        // every handle can form the coefficient of global cell (0,0) of its grids
        for (int b = threadIdx.x; b < batch; b += blockDim.x) {
            const T u = (T)hash_uniform(seed, (uint32_t)b, 0u, 0u);
            const T eps = mul_rn(eps0, add_rn((T)1, mul_rn(span, u)));
            mur[b] = mur_from(mu0, eps, dt, dx);
        }
    }
}

// Grayscale structure image -> materials (main.py:109-123): eps = (1 + (bp - 1) * (1 - g/255)) * eps0 in float64
// (the reference's dtype), then cast to the run dtype -- the cast the caller of material_init does for an fp32
// run; mu = mu0.  One byte per cell crosses PCIe instead of two float64 maps.
template <typename T>
__global__ void gray_materials_kernel(const unsigned char* gray, T* ce, T* ch, int Rl, int C, int pitch, long long grid_stride,
                                      int batch, double black_point, double eps0, double mu0) {
    const long long per_grid = (long long)Rl * C, n = per_grid * batch;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        const int b = (int)(i / per_grid);
        const long long r = i - (long long)b * per_grid;
        const int li = (int)(r / C), j = (int)(r - (long long)li * C);
        const double g = __ddiv_rn((double)gray[i], 255.0);
        const double factor = __dadd_rn(1.0, __dmul_rn(__dsub_rn(black_point, 1.0), __dsub_rn(1.0, g)));
        const long long o = (long long)b * grid_stride + (long long)li * pitch + j;
        ce[o] = (T)__dmul_rn(factor, eps0);
        ch[o] = (T)mu0;
    }
}

// Random two-phase medium of the dataset generator (diffusion_training.py:54-93): a uniform field u in [0,1),
// blurred with a 15 x 15 Gaussian (zero padding, float32, taps accumulated row-major with one rounding per
// multiply and per add), thresholded at 0.5 -> eps_hi / eps_lo; mu uniform.  u comes from the counter-based hash
// (hash_uniform), so a CPU checker rebuilds the same field.  One CTA per 32 x 32 output tile of one grid.
constexpr int BLOB_K = 15, BLOB_T = 32, BLOB_H = BLOB_T + BLOB_K - 1;
template <typename T>
__global__ void __launch_bounds__(256) blob_materials_kernel(T* ce, T* ch, int R, int C, int pitch, long long grid_stride,
                                                             uint64_t seed, const float* weights, T eps_lo, T eps_hi, T mu) {
    __shared__ float su[BLOB_H][BLOB_H + 1];
    __shared__ float sw[BLOB_K * BLOB_K];
    const int b = blockIdx.z, r0 = blockIdx.y * BLOB_T, c0 = blockIdx.x * BLOB_T;
    for (int i = threadIdx.x; i < BLOB_K * BLOB_K; i += blockDim.x) sw[i] = weights[(size_t)b * BLOB_K * BLOB_K + i];
    for (int i = threadIdx.x; i < BLOB_H * BLOB_H; i += blockDim.x) {
        const int y = i / BLOB_H, x = i - y * BLOB_H;
        const int gi = r0 + y - BLOB_K / 2, gj = c0 + x - BLOB_K / 2;
        su[y][x] = (gi >= 0 && gi < R && gj >= 0 && gj < C) ? (float)hash_uniform(seed, (uint32_t)b, (uint32_t)gi, (uint32_t)gj) : 0.0f;
    }
    __syncthreads();
    const int tx = threadIdx.x & 31, ty0 = threadIdx.x >> 5;
    for (int ty = ty0; ty < BLOB_T; ty += 8) {
        const int gi = r0 + ty, gj = c0 + tx;
        if (gi >= R || gj >= C) continue;
        float acc = 0.0f;
#pragma unroll
        for (int ky = 0; ky < BLOB_K; ++ky)
#pragma unroll
            for (int kx = 0; kx < BLOB_K; ++kx) acc = add_rn(acc, mul_rn(sw[ky * BLOB_K + kx], su[ty + ky][tx + kx]));
        const long long o = (long long)b * grid_stride + (long long)gi * pitch + gj;
        ce[o] = acc > 0.5f ? eps_hi : eps_lo;
        ch[o] = mu;
    }
}

// Field readout as an image (main.py:153-179): clip Ez to [vmin, vmax], normalise in the run dtype, look
// up a 256-entry colormap, alpha-blend (alpha = 0.7) over the grayscale permittivity background in
// float64 and truncate to uint8 -- the same operations, in the same order and precision, as the reference's
// numpy expression, so the frame leaves the GPU as 3 bytes per cell.
template <typename T>
__global__ void snapshot_kernel(const T* ez, const unsigned char* gray, const double* lut, unsigned char* rgb, int Rl,
                                int C, int pitch, T vmin, T vmax, T span) {
    const long long n = (long long)Rl * C;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const double alpha = 0.7, one_minus = 1 - 0.7;
    for (; i < n; i += stride) {
        const int r = (int)(i / C), c = (int)(i - (long long)r * C);
        T v = ez[(long long)r * pitch + c];
        v = v < vmin ? vmin : (v > vmax ? vmax : v);  // np.clip
        T x = mul_rn(div_rn(sub_rn(v, vmin), span), (T)256);
        if (x == (T)256) x = (T)255;
        int idx = (int)x;  // astype(int): truncation
        idx = idx < 0 ? 0 : (idx > 255 ? 255 : idx);
        const double bg = __dmul_rn(__ddiv_rn((double)gray[i], 255.0), one_minus);
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
            const double f = __dmul_rn(__dadd_rn(__dmul_rn(lut[idx * 3 + ch], alpha), bg), 255.0);
            rgb[i * 3 + ch] = (unsigned char)f;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// helpers
// ------------------------------------------------------------------------------------------------
static int hy_rows(const fdtd2d_sim* s) {
    // Hy has one row fewer than Ez when the handle holds the global last row (main.py:84)
    return (s->row0 + s->Rl == s->Rg) ? s->Rl - 1 : s->Rl;
}

constexpr int MAX_DEVICES = 64;  // function attributes are per device: remember where they were set

template <typename T, int TH> static int set_generic_attr(int dev) {
    static bool done_[MAX_DEVICES] = {};
    bool& done = done_[dev % MAX_DEVICES];
    if (!done) {
        CUDA_TRY(cudaFuncSetAttribute(tile_generic_kernel<T, TH, G_TW, generic_nt<T, TH>()>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)(6 * TH * G_TW * sizeof(T))));
        done = true;
    }
    return 0;
}

// ---- TMA tensor maps ---------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int encode_map(const fdtd2d_sim* s, void* base, int box_rows, CUtensorMap* out) {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult qr;
        CUDA_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qr));
        if (!sym || qr != cudaDriverEntryPointSuccess) return fail(FDTD2D_ECUDA, "cuTensorMapEncodeTiled not available");
        fn = reinterpret_cast<EncodeTiledFn>(sym);
    }
    // 2-D view: inner = padded row (pitch floats), outer = all rows of all grids of the batch
    const cuuint64_t dims[2] = {(cuuint64_t)s->pitch, (cuuint64_t)s->Rl * (cuuint64_t)s->batch};
    const cuuint64_t strides[1] = {(cuuint64_t)s->pitch * sizeof(float)};
    const cuuint32_t box[2] = {(cuuint32_t)TILE_TW, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(FDTD2D_ECUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return 0;
}

static int build_tma_maps(fdtd2d_sim* s, int box_rows) {
    if (s->tma_box_rows == box_rows) return 0;
    for (int h = 0; h < 2; ++h) {
        TmaMaps& m = s->tma_maps[h];
        if (int rc = encode_map(s, s->field[h][0], box_rows, &m.ez)) return rc;
        if (int rc = encode_map(s, s->field[h][1], box_rows, &m.hx)) return rc;
        if (int rc = encode_map(s, s->field[h][2], box_rows, &m.hy)) return rc;
        if (int rc = encode_map(s, s->ce, box_rows, &m.ce)) return rc;
        if (int rc = encode_map(s, s->ch, box_rows, &m.ch)) return rc;
    }
    s->tma_box_rows = box_rows;
    return 0;
}

template <int MR, int NW, bool PAIR> static int launch_tma_t(fdtd2d_sim* s, const PassParams<float>& p, int n_tiles) {
    static bool done_[MAX_DEVICES] = {};
    bool& done = done_[s->device % MAX_DEVICES];
    const size_t smem = (size_t)(5 * MR * NW * TILE_TW + (PAIR ? 4 : 2) * NW * TILE_TW) * sizeof(float);
    if (!done) {
        CUDA_TRY(cudaFuncSetAttribute(tile_tma_kernel<MR, NW, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        done = true;
    }
    if (int rc = build_tma_maps(s, MR * NW)) return rc;
    const int grid = std::min(n_tiles, sm_count(s));
    tile_tma_kernel<MR, NW, PAIR><<<grid, NW * 32, smem, s->stream>>>(s->tma_maps[s->cur], p, n_tiles);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

// Is dt/(mu*dx) one value in every cell of the local array?  (True for every material_init output: main.py:105,121.)
template <typename T> __global__ void uniform_check_kernel(const T* a, int rows, int cols, int pitch, int* differs) {
    const T v = a[0];
    const long long n = (long long)rows * cols;
    int bad = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(i / cols), c = (int)(i - (long long)r * cols);
        bad |= a[(long long)r * pitch + c] != v;
    }
    if (__syncthreads_or(bad) && threadIdx.x == 0) atomicOr(differs, 1);
}

// The check queued on `st`, the answer copied to the handle's pinned words; ev_check marks its arrival.
static int enqueue_ch_check(fdtd2d_sim* s, cudaStream_t st) {
    if (!s->d_flag) CUDA_TRY(cudaMalloc(&s->d_flag, sizeof(int)));
    if (!s->h_check) CUDA_TRY(cudaHostAlloc(&s->h_check, 2 * sizeof(unsigned long long), cudaHostAllocDefault));
    if (!s->ev_check) CUDA_TRY(cudaEventCreateWithFlags(&s->ev_check, cudaEventDisableTiming));
    CUDA_TRY(cudaMemsetAsync(s->d_flag, 0, sizeof(int), st));
    const int blocks = sm_count(s) * 8;
    if (s->dtype == FDTD2D_F32)
        uniform_check_kernel<float><<<blocks, 256, 0, st>>>((const float*)s->ch, s->batch * s->Rl, s->C, (int)s->pitch, s->d_flag);
    else
        uniform_check_kernel<double><<<blocks, 256, 0, st>>>((const double*)s->ch, s->batch * s->Rl, s->C, (int)s->pitch, s->d_flag);
    CUDA_TRY(cudaGetLastError());
    s->h_check[0] = 1, s->h_check[1] = 0;
    CUDA_TRY(cudaMemcpyAsync(&s->h_check[0], s->d_flag, sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(&s->h_check[1], s->ch, s->esize, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaEventRecord(s->ev_check, st));
    s->launches += 1;
    return 0;
}

static int check_ch_uniform(fdtd2d_sim* s) {
    if (s->ch_uniform >= 0) return 0;
    if (s->check_pending) {  // queued with the maps (fdtd2d_set_materials_async): usually long done, and no other handle's work
        CUDA_TRY(cudaEventSynchronize(s->ev_check));  // on a shared stream stands between the host and the answer
        s->check_pending = false;
        const bool differs = (int)(s->h_check[0] & 0xffffffffu) != 0;
        if (s->dtype == FDTD2D_F32) {
            float v;
            memcpy(&v, &s->h_check[1], sizeof v);
            s->ch_value = v;
        } else {
            double v;
            memcpy(&v, &s->h_check[1], sizeof v);
            s->ch_value = v;
        }
        s->ch_uniform = (differs || !s->opt.uniform_ch) ? 0 : 1;
        return 0;
    }
    if (!s->d_flag) CUDA_TRY(cudaMalloc(&s->d_flag, sizeof(int)));
    CUDA_TRY(cudaMemsetAsync(s->d_flag, 0, sizeof(int), s->stream));
    // the batch grids are back to back: rows = batch * Rl
    const int blocks = sm_count(s) * 8;
    int differs = 1;
    if (s->dtype == FDTD2D_F32) {
        float v = 0.0f;
        uniform_check_kernel<float><<<blocks, 256, 0, s->stream>>>((const float*)s->ch, s->batch * s->Rl, s->C, (int)s->pitch, s->d_flag);
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaMemcpyAsync(&v, s->ch, sizeof(float), cudaMemcpyDeviceToHost, s->stream));
        CUDA_TRY(cudaMemcpyAsync(&differs, s->d_flag, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
        CUDA_TRY(cudaStreamSynchronize(s->stream));
        s->ch_value = v;
    } else {
        double v = 0.0;
        uniform_check_kernel<double><<<blocks, 256, 0, s->stream>>>((const double*)s->ch, s->batch * s->Rl, s->C, (int)s->pitch, s->d_flag);
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaMemcpyAsync(&v, s->ch, sizeof(double), cudaMemcpyDeviceToHost, s->stream));
        CUDA_TRY(cudaMemcpyAsync(&differs, s->d_flag, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
        CUDA_TRY(cudaStreamSynchronize(s->stream));
        s->ch_value = v;
    }
    s->ch_uniform = (differs || !s->opt.uniform_ch) ? 0 : 1;
    s->launches += 1;
    return 0;
}

// ---- wavefront launches --------------------------------------------------------------------------
// One instantiation of the packed fp32 wavefront kernel: K levels, scalar or mapped dt/(mu*dx), P rows of prefetch, with /
// without the ring-strip and the slab-band forms of the run.
template <int K, bool UCH, int P, bool RING, int FLAVOUR>
static int launch_wave_x2_t(fdtd2d_sim* s, const PassParams<float>& p, const WaveTask* tasks, int n_tasks, int* ticket, int grid) {
    static bool done_[MAX_DEVICES] = {};
    bool& done = done_[s->device % MAX_DEVICES];
    const size_t smem = wave_smem_bytes(WAVE_NW, UCH);
    if (!done) CUDA_TRY(cudaFuncSetAttribute(strip_wave_x2_kernel<K, UCH, P, RING, FLAVOUR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    done = true;
    strip_wave_x2_kernel<K, UCH, P, RING, FLAVOUR><<<grid, WAVE_NW * 32, smem, s->stream>>>(p, tasks, n_tasks, ticket, UCH ? (float)s->ch_value : 0.0f,
                                                                                       0x8000000080000000ull);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

// the scalar form, used for fp64
template <int K, bool UCH, bool SLAB, bool RING>
static int launch_wave_f64_t(fdtd2d_sim* s, const PassParams<double>& p, const WaveTask* tasks, int n_tasks, int* ticket, int grid) {
    static bool done_[MAX_DEVICES] = {};
    bool& done = done_[s->device % MAX_DEVICES];
    const size_t smem = wave_smem_bytes(WAVE_NW, UCH);
    if (!done) CUDA_TRY(cudaFuncSetAttribute(strip_wave_kernel<double, K, UCH, WAVE_P, SLAB, RING>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    done = true;
    strip_wave_kernel<double, K, UCH, WAVE_P, SLAB, RING><<<grid, WAVE_NW * 32, smem, s->stream>>>(p, tasks, n_tasks, ticket, UCH ? s->ch_value : 0.0);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

// the staged wavefront: K = 12 as three warps of four levels (uniform permeability)
template <int NG, bool RING>
static int launch_stage_t(fdtd2d_sim* s, const PassParams<float>& p, const WaveTask* tasks, int n_tasks, int* ticket) {
    static bool done_[MAX_DEVICES] = {};
    bool& done = done_[s->device % MAX_DEVICES];
    const size_t smem = (size_t)NG * stage_group_bytes();
    if (!done) CUDA_TRY(cudaFuncSetAttribute(strip_stage_kernel<NG, RING>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    done = true;
    const int grid = std::min((n_tasks + NG - 1) / NG, sm_count(s));
    strip_stage_kernel<NG, RING><<<grid, NG * STAGE_S * 32, smem, s->stream>>>(p, tasks, n_tasks, ticket, (float)s->ch_value, 0x8000000080000000ull);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

// levels for which a wavefront instantiation exists
static bool wave_has_k(const fdtd2d_sim* s, int k) {
    if (s->dtype == FDTD2D_F64) return k == 4 || k == 8;
    return k == 8 || (k == 12 && !s->has_top_nb && !s->has_bot_nb);  // (12 levels: whole grids with uniform permeability)
}

// max_ctas > 0: the kernel takes at most that many SMs (the rest are left to the edge tiles of the pass)
static int launch_wave(fdtd2d_sim* s, const PassParams<float>& p, const WaveTask* tasks, int n_tasks, int* ticket, int k, bool ring, int max_ctas = 0) {
    if (int rc = check_ch_uniform(s)) return rc;
    const int grid = std::min((n_tasks + WAVE_NW - 1) / WAVE_NW, max_ctas > 0 ? std::min(max_ctas, sm_count(s)) : sm_count(s));
    const bool uch = s->ch_uniform == 1;  // uniform permeability: the map is not read at all (28 instead of 32 B per cell and pass)
    const bool slab = s->has_top_nb || s->has_bot_nb;
    if (k == 12) {  // (runs for k = 12 are only built when the permeability is uniform)
        if (!uch || slab) return fail(FDTD2D_EINVAL, "the 12-level wavefront kernel needs uniform permeability and a whole grid");
        if (s->opt.stage) {
            if (s->opt.stage == 4) return ring ? launch_stage_t<4, true>(s, p, tasks, n_tasks, ticket) : launch_stage_t<4, false>(s, p, tasks, n_tasks, ticket);
            return ring ? launch_stage_t<5, true>(s, p, tasks, n_tasks, ticket) : launch_stage_t<5, false>(s, p, tasks, n_tasks, ticket);
        }
        return launch_wave_x2_t<12, true, 2, false, 0>(s, p, tasks, n_tasks, ticket, grid);
    }
    if (k != 8) return fail(FDTD2D_EINVAL, "no fp32 wavefront kernel for k=%d", k);
    const int sel = (uch ? 4 : 0) | (ring ? 2 : 0) | (slab ? 1 : 0);
    switch (sel) {
        case 0: return launch_wave_x2_t<8, false, WAVE_P, false, 0>(s, p, tasks, n_tasks, ticket, grid);
        case 1: return launch_wave_x2_t<8, false, WAVE_P, false, 1>(s, p, tasks, n_tasks, ticket, grid);
        case 2: return launch_wave_x2_t<8, false, WAVE_P, true, 0>(s, p, tasks, n_tasks, ticket, grid);
        case 3: return launch_wave_x2_t<8, false, WAVE_P, true, 1>(s, p, tasks, n_tasks, ticket, grid);
        case 4: return launch_wave_x2_t<8, true, WAVE_P, false, 0>(s, p, tasks, n_tasks, ticket, grid);
        case 5: return launch_wave_x2_t<8, true, WAVE_P, false, 1>(s, p, tasks, n_tasks, ticket, grid);
        case 6: return launch_wave_x2_t<8, true, WAVE_P, true, 0>(s, p, tasks, n_tasks, ticket, grid);
        default: return launch_wave_x2_t<8, true, WAVE_P, true, 1>(s, p, tasks, n_tasks, ticket, grid);
    }
}

// the fused double pass: phase-0 and phase-1 runs of a k = 8 pass pair in one launch (whole fp32 grids)
static int launch_wave_fused(fdtd2d_sim* s, const PassParams<float>& p, const WaveTask* tasks, int n_tasks, int* ticket, bool ring) {
    if (int rc = check_ch_uniform(s)) return rc;
    const int grid = std::min((n_tasks + WAVE_NW - 1) / WAVE_NW, sm_count(s));
    const bool uch = s->ch_uniform == 1;
    if (uch) return ring ? launch_wave_x2_t<8, true, WAVE_P, true, 2>(s, p, tasks, n_tasks, ticket, grid) : launch_wave_x2_t<8, true, WAVE_P, false, 2>(s, p, tasks, n_tasks, ticket, grid);
    return ring ? launch_wave_x2_t<8, false, WAVE_P, true, 2>(s, p, tasks, n_tasks, ticket, grid) : launch_wave_x2_t<8, false, WAVE_P, false, 2>(s, p, tasks, n_tasks, ticket, grid);
}

template <int K, bool RING> static int launch_wave_f64_k(fdtd2d_sim* s, const PassParams<double>& p, const WaveTask* tasks, int n_tasks, int* ticket, int grid) {
    const bool uch = s->ch_uniform == 1, slab = s->has_top_nb || s->has_bot_nb;
    if (uch) return slab ? launch_wave_f64_t<K, true, true, RING>(s, p, tasks, n_tasks, ticket, grid) : launch_wave_f64_t<K, true, false, RING>(s, p, tasks, n_tasks, ticket, grid);
    return slab ? launch_wave_f64_t<K, false, true, RING>(s, p, tasks, n_tasks, ticket, grid) : launch_wave_f64_t<K, false, false, RING>(s, p, tasks, n_tasks, ticket, grid);
}

static int launch_wave(fdtd2d_sim* s, const PassParams<double>& p, const WaveTask* tasks, int n_tasks, int* ticket, int k, bool ring, int max_ctas = 0) {
    if (int rc = check_ch_uniform(s)) return rc;
    const int grid = std::min((n_tasks + WAVE_NW - 1) / WAVE_NW, max_ctas > 0 ? std::min(max_ctas, sm_count(s)) : sm_count(s));
    switch (k) {
        case 4: return launch_wave_f64_k<4, false>(s, p, tasks, n_tasks, ticket, grid);  // (ring strips are built for k = 8 only)
        case 8: return ring ? launch_wave_f64_k<8, true>(s, p, tasks, n_tasks, ticket, grid) : launch_wave_f64_k<8, false>(s, p, tasks, n_tasks, ticket, grid);
        default: return fail(FDTD2D_EINVAL, "no fp64 wavefront kernel for k=%d", k);
    }
}

template <typename T, int MR, int NW> static int launch_edge_tt(int dev, const PassParams<T>& p, int n_tiles, cudaStream_t st) {
    static bool done_[MAX_DEVICES] = {};
    bool& done = done_[dev % MAX_DEVICES];
    const size_t smem = (size_t)(2 * MR * NW * TILE_TW + 2 * NW * TILE_TW) * sizeof(T);
    if (!done) {
        CUDA_TRY(cudaFuncSetAttribute(tile_edge_kernel<T, MR, NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        done = true;
    }
    tile_edge_kernel<T, MR, NW><<<(unsigned)n_tiles, NW * 32, smem, st>>>(p);
    CUDA_TRY(cudaGetLastError());
    return 0;
}
static int launch_edge(int dev, const PassParams<float>& p, int n, cudaStream_t st) { return launch_edge_tt<float, F_MR, F_NW>(dev, p, n, st); }
static int launch_edge(int dev, const PassParams<double>& p, int n, cudaStream_t st) { return launch_edge_tt<double, D_MR, D_NW>(dev, p, n, st); }

template <int TH> static int launch_generic_list_t(int dev, const PassParams<float>& p, int n_tiles, cudaStream_t st) {
    if (int rc = set_generic_attr<float, TH>(dev)) return rc;
    constexpr int NT = generic_nt<float, TH>();
    tile_generic_kernel<float, TH, G_TW, NT><<<(unsigned)n_tiles, NT, 6 * TH * G_TW * sizeof(float), st>>>(p);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

// second_pass: the parameters of the SECOND pass of a fused pair (it reads what the first one writes, k steps later)
template <typename T> static void fill_params(const fdtd2d_sim* s, const TilePlan& tp, int phases, PassParams<T>* pp, const PassPlan* pl = nullptr,
                                              bool second_pass = false) {
    PassParams<T>& p = *pp;
    memset(&p, 0, sizeof p);
    const int cur = s->cur ^ (second_pass ? 1 : 0);
    for (int f = 0; f < 3; ++f) {
        p.in[f] = static_cast<const T*>(s->field[cur][f]);
        p.out[f] = static_cast<T*>(s->field[cur ^ 1][f]);
    }
    p.ce = static_cast<const T*>(s->ce);
    p.ch = static_cast<const T*>(s->ch);
    p.mur = static_cast<const T*>(s->mur);
    p.grid_stride = (long long)s->grid_elems;
    p.pitch = (int)s->pitch;
    p.Rg = s->Rg;
    p.C = s->C;
    p.row0 = s->row0;
    p.Rl = s->Rl;
    p.own_begin = s->row_begin;
    p.own_end = s->row_end;
    p.k = tp.k;
    p.hx = tp.hx;
    p.phases = phases;
    p.CH = tp.CH;
    p.CW = tp.CW;
    p.tiles_y = tp.tiles_y;
    p.tiles_x = tp.tiles_x;
    p.tile_list = nullptr;
    p.src = s->d_src;
    p.src_range = s->n_src ? s->d_src_range : nullptr;
    p.amp = s->d_amp;
    p.amp_steps = s->amp_steps;
    p.step0 = s->step + (second_pass ? tp.k : 0);
    p.probes = s->d_probe;
    p.probe_range = s->n_probe ? s->d_probe_range : nullptr;
    p.n_probe = s->n_probe;
    p.trace = static_cast<T*>(s->d_trace);
    p.trace_cap = s->trace_cap;
    // y-slabs: the tile grid starts at the first owned row; band rows; the neighbours' buffers
    p.org = own_first(s);
    p.store_lo = own_first(s);
    p.store_hi = own_last(s);
    p.band_lo[0] = own_first(s);
    p.band_hi[0] = own_first(s) + (s->has_top_nb ? s->halo : 0);
    p.band_lo[1] = own_last(s) - (s->has_bot_nb ? s->halo : 0);
    p.band_hi[1] = own_last(s);
    p.flags = s->d_slab_flags;
    p.seq = s->pass_seq;
    if (pl) p.fuse_flags = pl->d_fuse_flags, p.fuse_nblk = pl->fuse_nblk;
    for (int side = 0; side < 2; ++side) {
        const PeerLink& pe = s->peer[side];
        if (!pe.attached) continue;
        for (int f = 0; f < 3; ++f) p.peer_out[side][f] = static_cast<T*>(pe.field[s->cur ^ 1][f]);
        p.peer_shift[side] = (long long)(s->row0 - pe.row0) * (long long)s->pitch;
        // I am the bottom neighbour of the slab above me and the top neighbour of the slab below me
        p.peer_flag[side] = pe.flags + (side == 0 ? FLAG_IN_BOT : FLAG_IN_TOP);
        p.band_expected[side] = pl ? pl->band_expected[side] : 0;
    }
}

// Generic kernel over the whole tile grid (per-function passes, variant 1).
template <typename T> static int launch_generic_all(fdtd2d_sim* s, int k, int phases) {
    if (peer_mode(s)) return fail(FDTD2D_EINVAL, "the generic kernel (variant 1 / single phases) does not take part in the peer halo exchange");
    TilePlan tp;
    if (int rc = plan_tiles(s, k, G_TH, &tp)) return rc;
    PassParams<T> p;
    fill_params(s, tp, phases, &p);
    if (int rc = set_generic_attr<T, G_TH>(s->device)) return rc;
    const long long n_tiles = (long long)s->batch * tp.tiles_y * tp.tiles_x;
    if (n_tiles > 0x7fffffffLL) return fail(FDTD2D_EINVAL, "too many tiles");
    const size_t smem = 6 * G_TH * G_TW * sizeof(T);
    constexpr int NT = generic_nt<T, G_TH>();
    tile_generic_kernel<T, G_TH, G_TW, NT><<<(unsigned)n_tiles, NT, smem, s->stream>>>(p);
    CUDA_TRY(cudaGetLastError());
    s->launches += 1;
    return 0;
}

static void free_plans(fdtd2d_sim* s) {
    s->resident_ok = -1;
    for (PassPlan& pl : s->hybrid) {
        cudaFree(pl.d_edge);
        cudaFree(pl.d_fast);
        cudaFree(pl.d_wave);
        cudaFree(pl.d_ticket);
        cudaFree(pl.d_fused);
        cudaFree(pl.d_deferred);
        cudaFree(pl.d_fuse_flags);
        pl = PassPlan();
    }
}

// Run lengths of the wavefront kernel (host only, no CUDA; exported as fdtd2d_plan_wave_runs for the CPU tests).
// rows[i] is the height of stretch i (a vertical sequence of plain tiles of one strip), ring[i] != 0 marks a ring strip,
// whose rows cost about twice a plain row: its runs are half as long and its rows count double.  The result is the
// shortest plain run length L (>= 4k rows: a run spends 2k rows warming up) for which the stretches fall into at most
// m x `warps` runs, m = the smallest count of runs per warp that keeps a run under ~cap_rows rows; parts[i] runs of
// (nearly) equal height then cover stretch i.
// ring_cost > 0: a ring-strip row costs ring_cost percent of a plain row and a run's 2k warm-up rows are counted too, so a
// ring run of r rows takes as long as a plain run of L rows when (r + 2k) * ring_cost = (L + 2k) * 100.
static int plan_wave_runs(const std::vector<int>& rows, const std::vector<unsigned char>& ring, long long warps, long long cap_rows, int k,
                          std::vector<int>* parts, int ring_cost = 0) {
    long long total = 0;
    int longest = 1;
    for (size_t i = 0; i < rows.size(); ++i) total += (ring[i] ? 2LL : 1LL) * rows[i], longest = std::max(longest, rows[i]);
    const long long m = std::max<long long>(1, (total + warps * cap_rows - 1) / (warps * cap_rows));
    auto run_len = [&](size_t i, int len) {
        if (!ring[i]) return len;
        if (ring_cost <= 0) return std::max(4 * k, len / 2);
        return std::max(4 * k, (int)((long long)(len + 2 * k) * 100 / ring_cost) - 2 * k);
    };
    auto count_runs = [&](int len) {
        long long c = 0;
        for (size_t i = 0; i < rows.size(); ++i) c += (rows[i] + run_len(i, len) - 1) / run_len(i, len);
        return c;
    };
    int lo = std::min(longest, 4 * k), hi = std::max(longest, 8 * k);
    while (lo < hi) {
        const int mid = (lo + hi) / 2;
        if (count_runs(mid) <= m * warps) hi = mid; else lo = mid + 1;
    }
    parts->resize(rows.size());
    for (size_t i = 0; i < rows.size(); ++i) (*parts)[i] = (rows[i] + run_len(i, lo) - 1) / run_len(i, lo);
    return lo;
}

// How many SMs the wavefront kernel of a pass should leave to the pass's edge tiles (host arithmetic only; exported as
// fdtd2d_plan_edge_reserve for the CPU tests).  An edge tile is one CTA that fills an SM (512 threads, the whole register file
// next to a wavefront CTA's 8 x 32 x 232 registers), so an SM holds either.  Launched first, the edge tiles of a mid-size
// grid take half the SMs for their ~23 us and the wavefront CTAs of those SMs -- with one run per warp -- start that much
// later: 4096^2 fp32 measured 112 us per pass = 22.7 (edge tiles) + 89 (a plain run), with a work-conserving floor of 101.
// Instead the wavefront goes first on sms - r SMs, its runs cut for that many warps, and the edge tiles cycle through the r
// SMs left: r minimises max(edge rounds x tile time, rows per warp).  Units: rows of one wavefront warp (0.60 us at 4096^2);
// an edge tile takes ~38 of them (22.7 us; both scale with k).  Measured (profiles/r2_reserve_ab.txt, bit-equal fields):
// 4096^2 113.2 -> 108.2 us per pass (21 SMs), 8192^2 fp32 364 -> 351 (11), 16384^2 1317 -> 1295 (6), 8192^2 fp64 749 -> 739
// (6).  Below 2 % of the pass (65536^2: a reserved SM costs the DRAM-bound runs more than its edge tiles' 0.9 %) and above
// 25 % (small grids, where the wavefront is the minor part and the constants above were not measured) nothing is
// reserved.  wave_rows: rows of all stretches, ring strips weighted.
static int plan_edge_reserve(long long n_edge, long long wave_rows, int sms, int k) {
    constexpr long long EDGE_ROWS = 38;
    if (n_edge <= 0 || wave_rows <= 0 || sms < 8) return 0;
    const long long edge_work = n_edge * EDGE_ROWS * WAVE_NW;  // in rows of one warp, like wave_rows
    if (edge_work * 50 < edge_work + wave_rows || edge_work * 4 > edge_work + wave_rows) return 0;  // (2 % .. 25 % of the pass)
    int best = 0;
    long long best_t = -1;
    for (int r = 1; r <= sms / 2; ++r) {
        const long long w = (long long)(sms - r) * WAVE_NW;
        const long long t_edge = (n_edge + r - 1) / r * EDGE_ROWS, t_wave = (wave_rows + w - 1) / w + 2 * k;
        const long long t = std::max(t_edge, t_wave);
        if (best_t < 0 || t < best_t) best_t = t, best = r;
    }
    return best;
}

// Classify the tile grid of a k-step pass and build its task lists (cached per k until sources, probes, materials or
// options change).
//   * The tile grid (TH x 128 windows, core CH x CW) covers the owned rows.  A tile is PLAIN when nothing but interior
//     cells is within k cells of its core: no Mur ring, no array edge, no source in its window, no probe in its core.
//   * Wavefront mode (a k with a wavefront instantiation, enough plain tiles): vertical stretches of plain tiles become
//     runs of rows; the first / last tile column may ride along as ring strips; a slab's band rows are cut out as runs of
//     their own and put first.  Everything else is an edge tile.
//   * Tile mode: plain tiles whose whole window lies inside the local array and that hold no band row go to the
//     persistent TMA kernel (fp32), the rest to the edge kernel.
struct PlanLists {
    TilePlan tp;
    std::vector<unsigned char> kind;  // per tile: 0 edge, 1 wavefront (plain), 2 wavefront (ring strip), 3 TMA tile kernel
    std::vector<int> edge, fast;      // edge tiles (band tiles first), TMA tiles
    int n_edge_band = 0;
    std::vector<WaveTask> tasks;      // band runs first
    int n_wave_band = 0;
    bool wave_ring = false;
    int band_expected[2] = {0, 0};
    int reserve_sms = 0;              // SMs left to the edge tiles (plan_edge_reserve)
    std::vector<WaveTask> fused, deferred;  // fused double pass: stage-1 ticket order (both phases), stage-2 phase-1 pieces
    int fuse_nblk = 0;
};

// The planning itself: host arithmetic only (s->sm_count and s->ch_uniform are read as they are), so that the CPU
// tests can check it for every geometry (fdtd2d_plan_host).
static int plan_pass(const fdtd2d_sim* s, int k, PlanLists* pl) {
    const bool f64 = s->dtype == FDTD2D_F64;
    const int TH = f64 ? D_TH : F_TH;
    const bool wave_k = s->opt.wavefront && wave_has_k(s, k) && s->variant != 3;
    // fp64 strips are 64 columns wide (two columns per lane): a tile column is cut into two strips
    const int hxw = (int)round_up((size_t)k, 2), strip_core_max = 64 - 2 * hxw;
    TilePlan& tp = pl->tp;
    pl->wave_ring = false, pl->n_wave_band = 0, pl->n_edge_band = 0, pl->reserve_sms = 0;
    if (int rc = plan_tiles(s, k, TH, &tp, 4, (f64 && wave_k) ? 2 * strip_core_max : 0)) return rc;
    const int per_grid = tp.tiles_y * tp.tiles_x;
    const long long n_tiles = (long long)s->batch * per_grid;
    if (n_tiles > 0x7fffffffLL) return fail(FDTD2D_EINVAL, "too many tiles");
    const int org = own_first(s), own_hi = own_last(s);
    const bool slab = s->has_top_nb || s->has_bot_nb;
    const int band_lo[2] = {org, own_hi - (s->has_bot_nb ? s->halo : 0)};
    const int band_hi[2] = {org + (s->has_top_nb ? s->halo : 0), own_hi};
    auto row_lo = [&](int ty) { return org + ty * tp.CH; };
    auto row_hi = [&](int ty) { return std::min(org + (ty + 1) * tp.CH, own_hi); };
    auto tile_band = [&](int ty, int side) { return row_lo(ty) < band_hi[side] && row_hi(ty) > band_lo[side]; };

    std::vector<unsigned char> special((size_t)n_tiles, 0);
    auto mark = [&](int b, int trow, int col, int row_pad_lo, int row_pad_hi, int col_pad_lo, int col_pad_hi) {
        // trow = row relative to the tile grid's origin; marks the tiles whose window
        // [t*CH - pad_lo, t*CH + CH + pad_hi) x [t*CW - pad_lo, ...) contains the cell
        const int ty_lo = std::max(0, (trow - row_pad_hi) / tp.CH - 1);
        const int tx_lo = std::max(0, (col - col_pad_hi) / tp.CW - 1);
        for (int ty = ty_lo; ty < tp.tiles_y; ++ty) {
            const int r0 = ty * tp.CH - row_pad_lo, r1 = ty * tp.CH + tp.CH + row_pad_hi;
            if (trow < r0) break;
            if (trow >= r1) continue;
            for (int tx = tx_lo; tx < tp.tiles_x; ++tx) {
                const int c0 = tx * tp.CW - col_pad_lo, c1 = tx * tp.CW + tp.CW + col_pad_hi;
                if (col < c0) break;
                if (col >= c1) continue;
                special[(size_t)b * per_grid + (size_t)ty * tp.tiles_x + tx] = 1;
            }
        }
    };
    // a source anywhere in the haloed tile; a probe in the core
    for (const Cell& c : s->h_src)
        mark(c.grid, c.row - s->row0 - org, c.col, k, std::max(k, TH - k - tp.CH), tp.hx, TILE_TW - tp.hx - tp.CW);
    for (const Cell& c : s->h_probe) mark(c.grid, c.row - s->row0 - org, c.col, 0, 0, 0, 0);

    // Ring strips: the first / last 128 columns of the padded row.  A source / probe must be re-checked against them
    // (the right strip is not aligned with the tile grid).
    // (fp32: 128-column strips at k = 8, and at k = 12 on the staged kernel; fp64: 64-column strips at k = 8)
    const bool lr_ok = wave_k && (f64 ? k == 8 : (k == 8 || (k == 12 && s->opt.stage))) && s->opt.ring_strips && s->C >= 4 * TILE_TW;
    const int SW = f64 ? 64 : TILE_TW, NQc = f64 ? 2 : 4;  // strip width, columns per 16 bytes
    const int lr_x0[2] = {0, (s->C + NQc - 1) / NQc * NQc - SW};
    std::vector<unsigned char> lr_special((size_t)n_tiles, 0);
    if (lr_ok) {
        auto mark_lr = [&](const Cell& c, int row_pad) {
            for (int side = 0; side < 2; ++side) {
                // (the whole tile column is given up for a source / probe near it: conservative, and rare)
                const int tx = side ? tp.tiles_x - 1 : 0, lrow = c.row - s->row0;
                const int w0 = std::min(lr_x0[side], tx * tp.CW - tp.hx), w1 = std::max(lr_x0[side] + SW, side ? s->C : tp.CW + tp.hx);
                if (c.col < w0 || c.col >= w1) continue;
                for (int ty = 0; ty < tp.tiles_y; ++ty)
                    if (lrow >= row_lo(ty) - row_pad && lrow < row_hi(ty) + row_pad)
                        lr_special[(size_t)c.grid * per_grid + (size_t)ty * tp.tiles_x + tx] = 1;
            }
        };
        for (const Cell& c : s->h_src) mark_lr(c, k);
        for (const Cell& c : s->h_probe) mark_lr(c, 0);
    }

    // per tile row: are its rows plain for the wavefront (rows [lo - k, hi + k) exist and hold no top / bottom ring) and
    // for the tile kernels (the whole TH-row window as well, and a full core)?
    std::vector<unsigned char> rows_wave(tp.tiles_y), rows_tile(tp.tiles_y);
    for (int ty = 0; ty < tp.tiles_y; ++ty) {
        const int lo = row_lo(ty) - k, hi = row_hi(ty) + k, whi = row_lo(ty) - k + TH;
        rows_wave[ty] = lo >= 0 && hi <= s->Rl && lo + s->row0 >= RING && hi + s->row0 <= s->Rg - RING;
        rows_tile[ty] = rows_wave[ty] && row_hi(ty) - row_lo(ty) == tp.CH && whi <= s->Rl && whi + s->row0 <= s->Rg - RING;
    }
    auto cols_plain = [&](int tx) {
        const int lc0 = tx * tp.CW - tp.hx;
        return lc0 >= RING && lc0 + TILE_TW <= s->C - RING;
    };
    // wavefront mode?  Small grids: one run per warp gets too short against its 2k warm-up rows, and the persistent tile
    // kernel wins.  Measured on B200 with balanced runs (profiles/): 3000^2 696 vs 593, 2048^2 548 vs 491, 1536^2 364 vs
    // 390 Gcell/s (wavefront vs tiles) -> the wavefront takes over from ~0.4 plain fp32 tiles (48 x 112 cells) per warp
    long long n_plain_cells = 0;
    if (wave_k)
        for (int b = 0; b < s->batch; ++b)
            for (int ty = 0; ty < tp.tiles_y; ++ty)
                for (int tx = 0; tx < tp.tiles_x; ++tx)
                    if (rows_wave[ty] && cols_plain(tx) && !special[(size_t)b * per_grid + ty * tp.tiles_x + tx])
                        n_plain_cells += (long long)(row_hi(ty) - row_lo(ty)) * tp.CW;
    const long long warps = (long long)std::max(1, s->sm_count) * WAVE_NW, tile_cells = 48 * 112;
    const long long min_tiles = s->opt.wave_min_tiles >= 0 ? s->opt.wave_min_tiles : 2 * warps / 5;
    const long long ring_min_tiles = s->opt.ring_min_tiles >= 0 ? s->opt.ring_min_tiles : 2 * warps;
    bool use_wave = wave_k && n_plain_cells > 0 && n_plain_cells >= min_tiles * tile_cells;
    if (k == 12 && s->ch_uniform != 1) use_wave = false;  // (12 levels exist for uniform permeability only)
    // (on small grids the ring strips do not pay: 2048^2 539 with, 590 Gcell/s without)
    const bool use_lr = use_wave && lr_ok && n_plain_cells >= ring_min_tiles * tile_cells;

    // kind of every tile: 0 edge, 1 wavefront (plain), 2 wavefront (ring strip), 3 TMA tile kernel
    std::vector<unsigned char>& kind = pl->kind;
    kind.assign((size_t)n_tiles, 0);
    std::vector<int> edge_band, edge_rest;
    std::vector<int>& fast = pl->fast;
    fast.clear();
    pl->band_expected[0] = pl->band_expected[1] = 0;
    for (int b = 0; b < s->batch; ++b)
        for (int ty = 0; ty < tp.tiles_y; ++ty)
            for (int tx = 0; tx < tp.tiles_x; ++tx) {
                const int id = b * per_grid + ty * tp.tiles_x + tx;
                const bool band = tile_band(ty, 0) || tile_band(ty, 1);
                int kd = 0;
                if (!special[id]) {
                    if (cols_plain(tx)) {
                        if (use_wave && rows_wave[ty]) kd = 1;
                        else if (!use_wave && !f64 && rows_tile[ty] && !band && s->variant != 1) kd = 3;
                    } else if (use_lr && rows_wave[ty] && !lr_special[id] && (tx == 0 || tx == tp.tiles_x - 1)) {
                        kd = 2;
                    }
                }
                kind[id] = (unsigned char)kd;
                if (kd == 0) {
                    (band ? edge_band : edge_rest).push_back(id);
                    for (int side = 0; side < 2; ++side) pl->band_expected[side] += tile_band(ty, side) ? 1 : 0;
                } else if (kd == 3) {
                    fast.push_back(id);
                }
            }

    std::vector<WaveTask>& tasks = pl->tasks;
    tasks.clear();
    if (use_wave) {
        // stretches: maximal vertical sequences of tiles of one kind in one tile column, in rows; a slab's band rows are
        // cut out of them as tasks of their own
        std::vector<WaveTask> segs, band_tasks;
        auto add_stretch = [&](WaveTask t) {
            int cuts[4], nc = 0;
            cuts[nc++] = t.y0;
            for (int side = 0; side < 2; ++side) {
                const int c = side == 0 ? band_hi[0] : band_lo[1];
                if (band_hi[side] > band_lo[side] && c > cuts[nc - 1] && c < t.y1) cuts[nc++] = c;
            }
            cuts[nc++] = t.y1;
            for (int i = 0; i + 1 < nc; ++i) {
                WaveTask q = t;
                q.y0 = cuts[i], q.y1 = cuts[i + 1];
                q.band = 0;
                for (int side = 0; side < 2; ++side)
                    if (band_hi[side] > band_lo[side] && q.y0 >= band_lo[side] && q.y1 <= band_hi[side]) q.band = side + 1;
                (q.band ? band_tasks : segs).push_back(q);
            }
        };
        const int strips = f64 ? 2 : 1, sw = tp.CW / strips;  // strips per tile column, their core width
        for (int b = 0; b < s->batch; ++b)
            for (int tx = 0; tx < tp.tiles_x; ++tx) {
                int run = 0, run_kind = 0;
                for (int ty = 0; ty <= tp.tiles_y; ++ty) {
                    const int kd = ty < tp.tiles_y ? kind[b * per_grid + ty * tp.tiles_x + tx] : 0;
                    const bool wavey = kd == 1 || kd == 2;
                    if (wavey && (run == 0 || kd == run_kind)) {
                        ++run, run_kind = kd;
                        continue;
                    }
                    if (run) {
                        const int y0 = row_lo(ty - run), y1 = row_hi(ty - 1);
                        if (run_kind == 1) {
                            for (int q = 0; q < strips; ++q) {
                                WaveTask t;
                                t.b = b, t.y0 = y0, t.y1 = y1, t.side = 0, t.band = 0;
                                if (f64) {
                                    t.x0 = tx * tp.CW + q * sw - hxw, t.c0 = hxw, t.c1 = hxw + sw;
                                } else {
                                    t.x0 = tx * tp.CW - tp.hx, t.c0 = tp.hx, t.c1 = tp.hx + tp.CW;
                                }
                                add_stretch(t);
                            }
                        } else {
                            const int side = tx == 0 ? 0 : 1;
                            WaveTask t;
                            t.b = b, t.y0 = y0, t.y1 = y1, t.side = side + 1, t.band = 0, t.x0 = lr_x0[side];
                            // stored columns: the tile's core, in strip coordinates (whole 16-byte groups; the last group
                            // may reach into the pad columns, which keep their zeros)
                            if (!f64) {
                                t.c0 = side ? tx * tp.CW - t.x0 : 0;
                                t.c1 = side ? TILE_TW : tp.CW;
                                add_stretch(t);
                            } else if (side == 0) {
                                // fp64: the ring strip stores the first half of the tile column, a plain strip the second
                                t.c0 = 0, t.c1 = sw;
                                add_stretch(t);
                                WaveTask u = t;
                                u.side = 0, u.x0 = sw - hxw, u.c0 = hxw, u.c1 = hxw + sw;
                                add_stretch(u);
                            } else {
                                // fp64, right: the ring strip stores what it can of the (narrower) last tile column -- it needs
                                // hxw halo columns on its left -- and a plain strip the columns before that, if any
                                const int X = tx * tp.CW, first = std::max(X, t.x0 + hxw);
                                t.c0 = first - t.x0, t.c1 = SW;
                                add_stretch(t);
                                if (first > X) {
                                    WaveTask u = t;
                                    u.side = 0, u.x0 = X - hxw, u.c0 = hxw, u.c1 = hxw + (first - X);
                                    add_stretch(u);
                                }
                            }
                            pl->wave_ring = true;
                        }
                    }
                    run = wavey ? 1 : 0, run_kind = kd;
                }
            }
        // Non-band stretches are cut into RUNS of rows, one wavefront task each.  A run costs 2k warm-up rows, and the GPU
        // has W = SMs x 8 independent warps: the cut is chosen so that there are (at most) m x W runs of nearly the same
        // length -- every warp gets m of them -- with m as small as a cap of ~wave_run_rows rows per run allows.  (With
        // runs of whole tiles a 4096^2 grid gave 720 runs to 1184 warps.)
        std::vector<int> seg_rows(segs.size()), parts;
        std::vector<unsigned char> seg_ring(segs.size());
        for (size_t i = 0; i < segs.size(); ++i) seg_rows[i] = segs[i].y1 - segs[i].y0, seg_ring[i] = segs[i].side != 0;
        // whole grids: leave some SMs to the edge tiles and cut the runs for the rest (plan_edge_reserve)
        int reserve = 0, auto_ring_cost = 208;
        const long long n_edge_all = (long long)edge_band.size() + (long long)edge_rest.size();
        if (!slab && band_tasks.empty() && s->opt.edge_reserve != 0 && n_edge_all > 0) {
            if (s->opt.edge_reserve > 0) {
                reserve = std::min(s->opt.edge_reserve, std::max(1, s->sm_count) / 2);
            } else {
                long long wave_rows = 0;
                for (size_t i = 0; i < segs.size(); ++i) wave_rows += seg_ring[i] ? (long long)seg_rows[i] * 208 / 100 : seg_rows[i];
                reserve = plan_edge_reserve(n_edge_all, wave_rows, std::max(1, s->sm_count), k);
                // one run per warp (mid-size grids, edge tiles >= 5 % of the pass): the ring runs end the kernel, and 220
                // measured better than 208 there (4096^2: 108.2 against 109.5 us per pass)
                if (n_edge_all * 38 * WAVE_NW * 20 >= n_edge_all * 38 * WAVE_NW + wave_rows) auto_ring_cost = 220;
            }
        }
        pl->reserve_sms = reserve;
        const int ring_cost = s->opt.ring_cost > 0 ? s->opt.ring_cost : (reserve > 0 ? auto_ring_cost : 0);
        const long long run_warps = (long long)(std::max(1, s->sm_count) - reserve) * WAVE_NW;
        plan_wave_runs(seg_rows, seg_ring, run_warps, std::max(1, s->opt.wave_run_rows), k, &parts, ring_cost);
        for (size_t i = 0; i < segs.size(); ++i) {
            const WaveTask& g = segs[i];
            const int rows = seg_rows[i];
            for (int q = 0; q < parts[i]; ++q) {
                WaveTask t = g;
                t.y0 = g.y0 + (int)((long long)rows * q / parts[i]), t.y1 = g.y0 + (int)((long long)rows * (q + 1) / parts[i]);
                tasks.push_back(t);
            }
        }
        // ring-strip runs first (they are the heavier ones); then row band by row band: warps that work at the same time
        // then hold neighbouring strips of the same rows, so the halo columns they share are read from DRAM once and from
        // L2 the second time
        auto order = [](const WaveTask& a, const WaveTask& b) {
            if ((a.side != 0) != (b.side != 0)) return a.side != 0;
            if (a.b != b.b) return a.b < b.b;
            if (a.y0 != b.y0) return a.y0 < b.y0;
            return a.x0 < b.x0;
        };
        std::stable_sort(tasks.begin(), tasks.end(), order);
        std::stable_sort(band_tasks.begin(), band_tasks.end(), order);
        for (const WaveTask& t : band_tasks) pl->band_expected[t.band - 1] += 1;
        pl->n_wave_band = (int)band_tasks.size();
        tasks.insert(tasks.begin(), band_tasks.begin(), band_tasks.end());  // the band runs go first

        // ---- fused double pass (strip_wave.cuh "two passes per launch"): whole fp32 grids, k = 8, tile rows that are
        // whole 16-row blocks.  Phase 0 = every stretch, cut into runs at multiples of 16 rows; phase 1 = the same
        // stretches, but only the tile rows whose 3 x 3 tile neighbourhood (the tile columns the strip's window covers,
        // one tile row up and down) is produced by phase-0 RUNS are fused -- the rest are left to the second launch.
        const bool fuse_on = s->opt.fuse > 0 || (s->opt.fuse < 0 && n_plain_cells >= (long long)6000 * 6000);
        pl->fused.clear(), pl->deferred.clear();
        if (fuse_on && !f64 && k == 8 && !slab && org == 0 && tp.CH % (1 << FUSE_BLOCK_LOG2) == 0 && band_tasks.empty()) {
            const int B16 = 1 << FUSE_BLOCK_LOG2;
            pl->fuse_nblk = (s->Rl + B16 - 1) / B16;
            auto wave_tile = [&](int b, int ty, int tx) {
                if (ty < 0 || ty >= tp.tiles_y || tx < 0 || tx >= tp.tiles_x) return false;
                const int kd = kind[(size_t)b * per_grid + (size_t)ty * tp.tiles_x + tx];
                return kd == 1 || kd == 2;
            };
            std::vector<WaveTask> phase0, phase1;
            for (size_t i = 0; i < segs.size(); ++i) {
                WaveTask g = segs[i];
                g.tx = g.side == 0 ? (g.x0 + tp.hx) / tp.CW : (g.side == 1 ? 0 : tp.tiles_x - 1);
                g.txlo = std::max(0, g.x0 / tp.CW), g.txhi = std::min(tp.tiles_x - 1, (g.x0 + TILE_TW - 1) / tp.CW);
                const int rows = seg_rows[i], np = parts[i];
                auto cut = [&](int q) { return q >= np ? g.y1 : g.y0 + (int)((long long)rows * q / np) / B16 * B16; };
                // phase 0: the whole stretch
                for (int q = 0; q < np; ++q) {
                    WaveTask t = g;
                    t.phase = 0, t.y0 = cut(q), t.y1 = cut(q + 1);
                    if (t.y1 > t.y0) phase0.push_back(t);
                }
                // phase 1: maximal ranges of tile rows that can / cannot be fused.  Its runs are cut 16 rows ABOVE the cuts of
                // phase 0, so that a run reads nothing a later-starting phase-0 run of its strip produces (rows up to y1 + k
                // <= the phase-0 cut): together with the pairing below, a phase-1 run then only ever waits for phase-0 runs
                // that start with it or earlier.
                const int ta = g.y0 / tp.CH, tb = g.y1 / tp.CH;  // tile rows [ta, tb) (org = 0, stretches are whole tile rows)
                auto fusable = [&](int ty) {
                    for (int yy = ty - 1; yy <= ty + 1; ++yy)
                        for (int xx = g.txlo; xx <= g.txhi; ++xx)
                            if (!wave_tile(g.b, yy, xx)) return false;
                    return true;
                };
                for (int ty = ta; ty < tb;) {
                    const bool f = fusable(ty);
                    int te = ty + 1;
                    while (te < tb && fusable(te) == f) ++te;
                    const int r0 = ty * tp.CH, r1 = te * tp.CH;
                    int y = r0;
                    for (int q = 1; q <= np && y < r1; ++q) {
                        int c = q == np ? r1 : cut(q) - B16;
                        if (c <= y) continue;
                        c = std::min(c, r1);
                        if (r1 - c < 2 * k) c = r1;  // (no stub shorter than its own warm-up)
                        WaveTask t = g;
                        t.phase = 1, t.y0 = y, t.y1 = c;
                        (f ? phase1 : pl->deferred).push_back(t);
                        y = c;
                    }
                    if (y < r1) {
                        WaveTask t = g;
                        t.phase = 1, t.y0 = y, t.y1 = r1;
                        (f ? phase1 : pl->deferred).push_back(t);
                    }
                    ty = te;
                }
            }
            if (!phase1.empty()) {
                // One ticket order that PAIRS the two phases: the phase-1 run that starts 16 rows above a phase-0 cut comes
                // right after the phase-0 run that starts at that cut, so both start together and the second one trails
                // the first by 16..32 rows -- close enough for its reads to hit L2.  (Ordered one pass after the other, a
                // phase-1 run starts a whole round of runs later than its producers and finds their rows evicted: measured,
                // 15.5 GB of DRAM traffic per pair instead of 11.)  A phase-1 run waits only for phase-0 runs at most a few
                // tickets after its own, which the next free warp takes: no deadlock whatever the grid width.
                struct Keyed {
                    long long key;
                    WaveTask t;
                };
                std::vector<Keyed> all;
                auto key_of = [&](const WaveTask& t) {
                    const long long row = t.phase ? (t.y0 + B16) / B16 * B16 : t.y0;  // the phase-0 cut this run pairs with
                    return ((long long)t.b << 44) | (row << 24) | ((long long)(t.x0 + TILE_TW) << 1) | (long long)t.phase;
                };
                for (const WaveTask& t : phase0) all.push_back({key_of(t), t});
                for (const WaveTask& t : phase1) all.push_back({key_of(t), t});
                std::stable_sort(all.begin(), all.end(), [](const Keyed& a, const Keyed& b) { return a.key < b.key; });
                for (const Keyed& e : all) pl->fused.push_back(e.t);
                // Safety net: every phase-0 run a phase-1 run reads must have a ticket less than half the GPU's warps after
                // its own (it is then taken before the waiting runs could fill the machine).  With equal cuts in
                // neighbouring strips the distance is 1 or 2; strips cut differently (ring strips, stretches broken by
                // sources) can push it up to a row of strips.  If it ever fails, this grid is not fused.
                std::vector<std::vector<int>> by_col((size_t)s->batch * tp.tiles_x);
                for (size_t i = 0; i < pl->fused.size(); ++i)
                    if (!pl->fused[i].phase) by_col[(size_t)pl->fused[i].b * tp.tiles_x + pl->fused[i].tx].push_back((int)i);
                bool safe = true;
                for (size_t i = 0; i < pl->fused.size() && safe; ++i) {
                    const WaveTask& t = pl->fused[i];
                    if (!t.phase) continue;
                    for (int xx = t.txlo; xx <= t.txhi && safe; ++xx)
                        for (int a : by_col[(size_t)t.b * tp.tiles_x + xx]) {
                            const WaveTask& u = pl->fused[a];
                            if (u.y0 < t.y1 + k && u.y1 > t.y0 - k && (long long)a > (long long)i + warps / 2) safe = false;
                        }
                }
                if (!safe) pl->fused.clear(), pl->deferred.clear();
            } else {
                pl->deferred.clear();
            }
        }
    }
    if (!slab) pl->band_expected[0] = pl->band_expected[1] = 0;
    pl->n_edge_band = (int)edge_band.size();
    pl->edge = edge_band;
    pl->edge.insert(pl->edge.end(), edge_rest.begin(), edge_rest.end());
    return 0;
}

// Classify the tile grid of a k-step pass and put its task lists on the device (cached per k until sources, probes,
// materials, options or peer links change).
static int classify_tiles(fdtd2d_sim* s, int k, PassPlan* pl) {
    sm_count(s);
    if (s->opt.wavefront && wave_has_k(s, k) && (k == 12 || s->dtype == FDTD2D_F64))  // (these read the permeability check)
        if (int rc = check_ch_uniform(s)) return rc;
    PlanLists L;
    if (int rc = plan_pass(s, k, &L)) return rc;
    pl->tp = L.tp;
    pl->wave_ring = L.wave_ring;
    pl->reserve_sms = L.reserve_sms;
    pl->band_expected[0] = L.band_expected[0], pl->band_expected[1] = L.band_expected[1];
    pl->n_wave = (int)L.tasks.size(), pl->n_wave_band = L.n_wave_band;
    pl->n_edge = (int)L.edge.size(), pl->n_edge_band = L.n_edge_band;
    pl->n_fast = (int)L.fast.size();
    if (pl->n_wave) {
        CUDA_TRY(cudaMalloc(&pl->d_ticket, sizeof(int)));
        CUDA_TRY(cudaMemsetAsync(pl->d_ticket, 0, sizeof(int), s->stream));  // (the kernels put it back to zero themselves)
        CUDA_TRY(cudaMalloc(&pl->d_wave, sizeof(WaveTask) * L.tasks.size()));
        // on the handle's own stream: a plain cudaMemcpy goes through the legacy default stream, which a non-blocking
        // stream does not wait for -- with another handle keeping the GPU busy the kernel could read the list before it
        // had landed (seen as an illegal address with two handles in two host threads)
        CUDA_TRY(cudaMemcpyAsync(pl->d_wave, L.tasks.data(), sizeof(WaveTask) * L.tasks.size(), cudaMemcpyHostToDevice, s->stream));
    }
    pl->n_fused = (int)L.fused.size(), pl->n_deferred = (int)L.deferred.size(), pl->fuse_nblk = L.fuse_nblk;
    if (pl->n_fused) {
        CUDA_TRY(cudaMalloc(&pl->d_fused, sizeof(WaveTask) * L.fused.size()));
        CUDA_TRY(cudaMemcpyAsync(pl->d_fused, L.fused.data(), sizeof(WaveTask) * L.fused.size(), cudaMemcpyHostToDevice, s->stream));
        if (pl->n_deferred) {
            CUDA_TRY(cudaMalloc(&pl->d_deferred, sizeof(WaveTask) * L.deferred.size()));
            CUDA_TRY(cudaMemcpyAsync(pl->d_deferred, L.deferred.data(), sizeof(WaveTask) * L.deferred.size(), cudaMemcpyHostToDevice, s->stream));
        }
        pl->fuse_flag_bytes = sizeof(unsigned) * (size_t)s->batch * pl->tp.tiles_x * pl->fuse_nblk;
        CUDA_TRY(cudaMalloc(&pl->d_fuse_flags, pl->fuse_flag_bytes));
    }
    if (pl->n_edge) {
        CUDA_TRY(cudaMalloc(&pl->d_edge, sizeof(int) * L.edge.size()));
        CUDA_TRY(cudaMemcpyAsync(pl->d_edge, L.edge.data(), sizeof(int) * L.edge.size(), cudaMemcpyHostToDevice, s->stream));
    }
    if (pl->n_fast) {
        CUDA_TRY(cudaMalloc(&pl->d_fast, sizeof(int) * L.fast.size()));
        CUDA_TRY(cudaMemcpyAsync(pl->d_fast, L.fast.data(), sizeof(int) * L.fast.size(), cudaMemcpyHostToDevice, s->stream));
    }
    CUDA_TRY(cudaStreamSynchronize(s->stream));  // the host vectors die here
    const TilePlan& tp = pl->tp;
    if (s->opt.debug)
        fprintf(stderr, "[fdtd2d] plan k=%d: tiles %d x %d (core %d x %d), edge %d (band %d), tma %d, wave runs %d (band %d, ring %d), band tasks %d | %d, fused %d + deferred %d, SMs left to the edge tiles %d\n", k,
                tp.tiles_y, tp.tiles_x, tp.CH, tp.CW, pl->n_edge, pl->n_edge_band, pl->n_fast, pl->n_wave, pl->n_wave_band, (int)pl->wave_ring,
                pl->band_expected[0], pl->band_expected[1], pl->n_fused, pl->n_deferred, pl->reserve_sms);
    pl->valid = true;
    return 0;
}

// One pass of the tile / wavefront kernels: edge tiles on a side stream, concurrently with the runs / plain tiles.
// part: 0 = everything, 1 = only the tasks that hold band rows of a slab, 2 = everything else.
template <typename T> static int launch_hybrid_t(fdtd2d_sim* s, int k, int part) {
    PassPlan& pl = s->hybrid[k];
    if (pl.valid && k == 12 && s->ch_uniform < 0) {
        // the 12-level wavefront exists for uniform permeability only, and the maps have changed since the plan was made
        if (int rc = check_ch_uniform(s)) return rc;
        if ((pl.n_wave > 0) != (s->ch_uniform == 1)) {
            CUDA_TRY(cudaStreamSynchronize(s->stream));
            cudaFree(pl.d_edge), cudaFree(pl.d_fast), cudaFree(pl.d_wave), cudaFree(pl.d_ticket);
            cudaFree(pl.d_fused), cudaFree(pl.d_deferred), cudaFree(pl.d_fuse_flags);
            pl = PassPlan();
        }
    }
    if (!pl.valid)
        if (int rc = classify_tiles(s, k, &pl)) return rc;
    PassParams<T> p;
    fill_params(s, pl.tp, FDTD2D_PHASE_H | FDTD2D_PHASE_E | FDTD2D_PHASE_SRC, &p, &pl);
    const int e_off = part == 2 ? pl.n_edge_band : 0, w_off = part == 2 ? pl.n_wave_band : 0;
    int n_edge = part == 1 ? pl.n_edge_band : pl.n_edge - e_off;
    int n_wave = part == 1 ? pl.n_wave_band : pl.n_wave - w_off;
    int n_fast = part == 1 ? 0 : pl.n_fast;
    if (s->opt.measure_skip & 1) n_edge = 0;  // (timing of the parts of a pass; the results are wrong)
    if (s->opt.measure_skip & 2) n_wave = n_fast = 0;
    const bool both = n_edge > 0 && (n_wave > 0 || n_fast > 0) && !(s->opt.measure_skip & 4);
    cudaStream_t estream = s->stream;
    if (both) {
        if (!s->side_stream) {
            CUDA_TRY(cudaStreamCreateWithFlags(&s->side_stream, cudaStreamNonBlocking));
            CUDA_TRY(cudaEventCreateWithFlags(&s->ev_fork, cudaEventDisableTiming));
            CUDA_TRY(cudaEventCreateWithFlags(&s->ev_join, cudaEventDisableTiming));
        }
        CUDA_TRY(cudaEventRecord(s->ev_fork, s->stream));
        CUDA_TRY(cudaStreamWaitEvent(s->side_stream, s->ev_fork, 0));
        estream = s->side_stream;
    }
    // With SMs reserved for them (plan_edge_reserve) the edge tiles go out AFTER the wavefront kernel, which takes only its
    // share of the SMs: the edge CTAs then cycle through the SMs that are left instead of holding up half the wavefront CTAs.
    const bool wave_first = both && part == 0 && pl.reserve_sms > 0 && n_wave > 0;
    auto launch_edges = [&]() -> int {
        p.tile_list = pl.d_edge + e_off;
        int rc;
        if (s->variant == 3) {  // debugging aid: shared-memory generic kernel for the edge tiles (fp32)
            if constexpr (std::is_same<T, float>::value) {
                if (peer_mode(s)) return fail(FDTD2D_EINVAL, "variant 3 does not take part in the peer halo exchange");
                rc = launch_generic_list_t<F_TH>(s->device, p, n_edge, estream);
            } else {
                return fail(FDTD2D_EINVAL, "variant 3 is an fp32 debugging aid");
            }
        } else {
            rc = launch_edge(s->device, p, n_edge, estream);
        }
        p.tile_list = nullptr;
        if (rc) return rc;
        s->launches += 1;
        return 0;
    };
    if (n_edge && !wave_first)
        if (int rc = launch_edges()) return rc;
    p.tile_list = nullptr;
    if (n_wave) {
        const int max_ctas = wave_first ? sm_count(s) - pl.reserve_sms : 0;
        if (int rc = launch_wave(s, p, pl.d_wave + w_off, n_wave, pl.d_ticket, k, pl.wave_ring, max_ctas)) return rc;
        s->launches += 1;
    }
    if (n_edge && wave_first)
        if (int rc = launch_edges()) return rc;
    if (n_fast) {
        if constexpr (std::is_same<T, float>::value) {
            p.tile_list = pl.d_fast;
            const int rc = s->opt.tma_pair ? launch_tma_t<F_MR, F_NW, true>(s, p, n_fast) : launch_tma_t<F_MR, F_NW, false>(s, p, n_fast);
            if (rc) return rc;
            s->launches += 1;
        } else {
            return fail(FDTD2D_EINVAL, "internal: TMA tiles are fp32 only");
        }
    }
    if (both) {
        CUDA_TRY(cudaEventRecord(s->ev_join, s->side_stream));
        CUDA_TRY(cudaStreamWaitEvent(s->stream, s->ev_join, 0));
    }
    return 0;
}

static int launch_hybrid(fdtd2d_sim* s, int k, int part) {
    return s->dtype == FDTD2D_F64 ? launch_hybrid_t<double>(s, k, part) : launch_hybrid_t<float>(s, k, part);
}

// Can the next two k = 8 passes of this handle go out as one fused launch?  (Builds the k = 8 plan if need be.)
static int fused_available(fdtd2d_sim* s, bool* yes) {
    *yes = false;
    if (s->dtype != FDTD2D_F32 || s->has_top_nb || s->has_bot_nb || s->variant == 1 || s->variant == 3 || !s->opt.wavefront || s->opt.fuse == 0) return 0;
    PassPlan& pl = s->hybrid[8];
    if (!pl.valid)
        if (int rc = classify_tiles(s, 8, &pl)) return rc;
    *yes = pl.n_fused > 0;
    return 0;
}

// Two k = 8 passes: [edge tiles of pass 1 || all runs of pass 1 + the fusable runs of pass 2, in one ticket order], then
// [edge tiles of pass 2 || the pass-2 runs next to edge tiles].  The state ends where it started (A -> B -> A).
static int launch_fused_pair(fdtd2d_sim* s) {
    const int k = 8;
    PassPlan& pl = s->hybrid[k];
    const int all = FDTD2D_PHASE_H | FDTD2D_PHASE_E | FDTD2D_PHASE_SRC;
    if (!s->side_stream) {
        CUDA_TRY(cudaStreamCreateWithFlags(&s->side_stream, cudaStreamNonBlocking));
        CUDA_TRY(cudaEventCreateWithFlags(&s->ev_fork, cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&s->ev_join, cudaEventDisableTiming));
    }
    for (int stage = 0; stage < 2; ++stage) {
        PassParams<float> p;
        fill_params(s, pl.tp, all, &p, &pl, stage == 1);
        if (stage == 0) CUDA_TRY(cudaMemsetAsync(pl.d_fuse_flags, 0, pl.fuse_flag_bytes, s->stream));
        const int n_wave = stage == 0 ? pl.n_fused : pl.n_deferred;
        const bool both = pl.n_edge > 0 && n_wave > 0;
        cudaStream_t estream = s->stream;
        if (both) {
            CUDA_TRY(cudaEventRecord(s->ev_fork, s->stream));
            CUDA_TRY(cudaStreamWaitEvent(s->side_stream, s->ev_fork, 0));
            estream = s->side_stream;
        }
        if (pl.n_edge) {
            p.tile_list = pl.d_edge;
            if (int rc = launch_edge(s->device, p, pl.n_edge, estream)) return rc;
            s->launches += 1;
        }
        p.tile_list = nullptr;
        if (n_wave) {
            const int rc = stage == 0 ? launch_wave_fused(s, p, pl.d_fused, n_wave, pl.d_ticket, pl.wave_ring)
                                      : launch_wave(s, p, pl.d_deferred, n_wave, pl.d_ticket, k, pl.wave_ring);
            if (rc) return rc;
            s->launches += 1;
        }
        if (both) {
            CUDA_TRY(cudaEventRecord(s->ev_join, s->side_stream));
            CUDA_TRY(cudaStreamWaitEvent(s->stream, s->ev_join, 0));
        }
    }
    s->passes += 2;
    s->fused_pairs += 1;
    return 0;
}

// ---- cluster-resident path (grid_resident.cuh) ---------------------------------------------------
// Shapes of the cluster-resident kernels.  0..4: grid_resident.cuh, MR rows per thread x NW warps -> a CTA holds MR*NW rows x
// 256 columns (3 x 16 was the best of them on B200 for 256^2 grids); 5 (the default): the packed kernel of
// grid_resident_x2.cuh, 6 rows x 4 columns per thread, 8 row blocks x 2 column halves -- 686 against 649 Gcell/s on
// BASELINE configs[4], 563k against 435k steps/s on one 200^2 grid.  All are selectable (option resident_cfg) and covered
// by the parity tests; grids the packed kernel does not take (right ring astride column 128, more than 384 rows) use 0.
struct ResCfg {
    int MR, NW;
};
static const ResCfg kResCfgs[] = {{3, 16}, {4, 12}, {2, 16}, {4, 8}, {3, 12}, {RX_MR, RX_NW}};
constexpr int RES_CFG_X2 = 5;  // the packed kernel of grid_resident_x2.cuh (6 rows x 4 columns per thread, 8 row blocks x 2 column halves)
constexpr int N_RES_CFG = sizeof(kResCfgs) / sizeof(kResCfgs[0]);

static int resident_cfg(const fdtd2d_sim* s) {
    const int v = s->opt.resident_cfg;
    return (v >= 0 && v < N_RES_CFG) ? v : 0;
}

// Small fp32 grids that fit a thread-block cluster: whole-run residency instead of k-step tiles.
static bool resident_eligible(fdtd2d_sim* s) {
    if (s->resident_ok >= 0) return s->resident_ok != 0;
    s->resident_ok = 0;
    if (s->dtype != FDTD2D_F32 || s->has_top_nb || s->has_bot_nb) return false;
    int rcfg = resident_cfg(s);
    if (rcfg == RES_CFG_X2 && s->C > RX_HW && (((s->C - 6) >> 2) << 2) < RX_HW) rcfg = 0;  // right ring astride the two column halves
    if (rcfg == RES_CFG_X2 && s->Rg > 8 * RX_TH) rcfg = 0;
    const int mr = kResCfgs[rcfg].MR, band = rcfg == RES_CFG_X2 ? RX_TH : mr * kResCfgs[rcfg].NW;
    if (s->C < 16 || s->C > RES_TW || s->Rg < 16 || s->Rg > 8 * band) return false;
    if (!s->opt.resident) return false;
    // every source / probe cell may need a 4-cell slot in its CTA's slot frame
    std::vector<int> per_grid((size_t)s->batch, 0);
    for (const Cell& c : s->h_src) per_grid[c.grid] += 1;
    for (const Cell& c : s->h_probe) per_grid[c.grid] += 1;
    for (int v : per_grid)
        if (v > RES_MAX_SLOTS) return false;
    int n = (s->Rg + band - 1) / band;
    if (s->opt.resident_cluster >= n && s->opt.resident_cluster <= 8) n = s->opt.resident_cluster;  // tuning knob: more, thinner bands per grid
    if (rcfg == RES_CFG_X2) {
        // Bands of the packed kernel: first | (n-2) x rpc | last with rpc and last multiples of 6 (the six bottom ring rows are
        // ONE warp's rows; the six top ring rows always are) and the first band whatever is left (6..48).  One band holds both
        // rings only when it has twelve rows or more and a multiple of six.
        if (n == 1 && (s->Rg % mr != 0 || s->Rg < 2 * mr)) n = 2;
        int rpc = 0, edge = 0, last = 0;
        if (n == 1) {
            rpc = edge = last = s->Rg;
        } else {
            bool ok = false;
            // Warps w, w + 4, ... share a scheduler, i.e. row blocks 0 and 4: the two SLOW row blocks of a CTA -- the first
            // (ring rows or the CTA above) and the last (ring rows or the CTA below) -- must not be blocks 0 and 4
            // (measured on cfg5: 685 against 712 Gcell/s), so a band of 5 blocks is avoided where the grid leaves the choice.
            auto clash = [&](int rows) { return (rows + mr - 1) / mr == 5; };
            int best = 1 << 30;
            const int r6_lo = n > 2 ? std::max(mr, (s->Rg - 2 * band + (n - 2) * mr - 1) / ((n - 2) * mr) * mr) : mr;  // (the first and the last band hold at most 2 x 48 rows)
            for (int r6 = r6_lo; r6 <= band; r6 += mr) {
                const int rem = s->Rg - (n - 2) * r6;  // rows of the first and the last band together
                if (rem < 2 * mr) break;
                const int la_min = std::max(mr, (rem - band + mr - 1) / mr * mr), la_max = std::min(band, (rem - mr) / mr * mr);
                for (int la = la_min; la <= std::min(la_max, n > 2 ? r6 : la_max); la += mr) {  // (the kernel caps every band but the first at rpc rows)
                    // cost: clashes first, then thick middle bands, then a first and a last band of different size
                    // cost: clashes first; then a first / last CTA that is not full -- they hold the ring warps, the slowest of the
                    // cluster, and the fewer other warps share their SM the better (256 rows as 46 | 42 x 4 | 42: 694 Gcell/s, as
                    // 40 | 48 x 4 | 24: 712); then thin middle bands (200 rows as 38 | 42 x 3 | 36: 505 Gcell/s over 148 grids, as
                    // 32 | 48 x 3 | 24: 481); then a first and a last band of similar size
                    const int fi = rem - la, thick = std::max(fi, la), thin = std::min(fi, la);
                    auto blocks = [&](int rows) { return (rows + mr - 1) / mr; };
                    int cost = 10000 * ((n > 2 && clash(r6) ? n - 2 : 0) + (clash(la) ? 1 : 0) + (clash(fi) ? 1 : 0)) +
                               300 * (std::max(0, blocks(fi) - 7) + std::max(0, blocks(la) - 7)) + (n > 2 ? 100 * blocks(r6) : 0) + (thick - thin);
                    // tuning knobs: rows of the last band (resident_trim) and of the middle bands (resident_rows)
                    if (s->opt.resident_trim > 0 || s->opt.resident_rows > 0)
                        cost = (s->opt.resident_trim > 0 ? 100 * std::abs(la - s->opt.resident_trim / mr * mr) : 0) +
                               (s->opt.resident_rows > 0 ? 100 * std::abs(r6 - s->opt.resident_rows / mr * mr) : 0) + (thick - thin);
                    if (cost < best) best = cost, rpc = r6, last = la, edge = rem - la, ok = true;
                }
            }
            if (!ok) return false;
            if (n == 2) rpc = last;
        }
        s->resident_cfg = rcfg;
        s->resident_cluster = n;
        s->resident_rpc = rpc;
        s->resident_edge = edge;
        s->resident_ok = 1;
        return true;
    }
    // The first and last CTA of a cluster also run the top / bottom boundary pass: give them `trim` rows fewer
    // than the middle ones when the grid leaves room (rows: edge | (n-2) x rpc | what is left, at most edge).
    int trim = 4 * mr;
    if (s->opt.resident_trim >= 0) trim = s->opt.resident_trim / mr * mr;
    int rpc = 0, edge = 0, last = 0;
    for (;; trim -= mr) {
        if (n <= 2 || trim <= 0) {
            rpc = edge = mr * ((s->Rg + mr * n - 1) / (mr * n));
            last = s->Rg - (n - 1) * rpc;
            break;
        }
        // smallest rpc (multiple of mr) with 2 * (rpc - trim) + (n - 2) * rpc >= Rg
        rpc = mr * ((s->Rg + 2 * trim + mr * n - 1) / (mr * n));
        edge = rpc - trim;
        last = s->Rg - edge - (n - 2) * rpc;  // rows of the last CTA
        if (rpc <= band && edge >= 6 && last >= 6 && last <= rpc) break;
    }
    if (n == 1) last = s->Rg;
    if (rpc > band || edge < 6 || last < 6 || last > rpc) return false;  // first / last band hold the whole ring
    s->resident_cfg = rcfg;
    s->resident_cluster = n;
    s->resident_rpc = rpc;
    s->resident_edge = edge;
    s->resident_ok = 1;
    return true;
}

template <int MR, int NW> static int launch_resident_t(fdtd2d_sim* s, int n_steps) {
    static bool done_[MAX_DEVICES] = {};
    bool& done = done_[s->device % MAX_DEVICES];
    const size_t smem = resident_smem_floats(MR, NW) * sizeof(float);
    if (!done) {
        CUDA_TRY(cudaFuncSetAttribute(grid_resident_kernel<MR, NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        done = true;
    }
    TilePlan tp;
    tp.k = n_steps;
    tp.CH = s->resident_rpc;
    tp.CW = s->resident_edge;
    PassParams<float> p;
    fill_params(s, tp, FDTD2D_PHASE_H | FDTD2D_PHASE_E | FDTD2D_PHASE_SRC, &p);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(s->batch * s->resident_cluster));
    cfg.blockDim = dim3(NW * 32);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)s->resident_cluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (s->opt.debug) {
        int nc = -1;
        cudaOccupancyMaxActiveClusters(&nc, grid_resident_kernel<MR, NW>, &cfg);
        fprintf(stderr, "[fdtd2d] resident: %d grids x cluster %d (%d | %d rows per CTA, %d x %d warps), %zu B smem, max active clusters %d\n",
                s->batch, s->resident_cluster, s->resident_edge, s->resident_rpc, MR, NW, smem, nc);
    }
    CUDA_TRY(cudaLaunchKernelEx(&cfg, grid_resident_kernel<MR, NW>, p));
    s->launches += 1;
    s->passes += 1;
    s->cur ^= 1;
    return 0;
}

// the packed kernel (grid_resident_x2.cuh): uniform dt/(mu*dx) as an argument, or the map
static int launch_resident_x2(fdtd2d_sim* s, int n_steps, bool uch, int rl) {
    static bool done_[2][RX_MR][MAX_DEVICES] = {};
    bool& done = done_[uch ? 1 : 0][rl][s->device % MAX_DEVICES];
    const size_t smem = resident_x2_smem_floats() * sizeof(float);
    TilePlan tp;
    tp.k = n_steps;
    tp.CH = s->resident_rpc;
    tp.CW = s->resident_edge;
    PassParams<float> p;
    fill_params(s, tp, FDTD2D_PHASE_H | FDTD2D_PHASE_E | FDTD2D_PHASE_SRC, &p);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(s->batch * s->resident_cluster));
    cfg.blockDim = dim3(RX_NW * 32);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)s->resident_cluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int nc = -1;
    CUDA_TRY(resident_x2_launch(uch, rl, &cfg, p, (float)s->ch_value, !done, smem, s->opt.debug ? &nc : nullptr));
    done = true;
    if (s->opt.debug)
        fprintf(stderr, "[fdtd2d] resident x2: %d grids x cluster %d (%d | %d rows per CTA), uniform ch %d, %zu B smem, max active clusters %d\n",
                s->batch, s->resident_cluster, s->resident_edge, s->resident_rpc, (int)uch, smem, nc);
    s->launches += 1;
    s->passes += 1;
    s->cur ^= 1;
    return 0;
}

static int launch_resident(fdtd2d_sim* s, int n_steps) {
    if (s->resident_cfg == RES_CFG_X2) {
        if (int rc = check_ch_uniform(s)) return rc;
        // the row inside a row block at which the first band ends is a template parameter (see the kernel)
        const int rl = s->resident_cluster > 1 ? (s->resident_edge - 1) % RX_MR : RX_MR - 1;
        return launch_resident_x2(s, n_steps, s->ch_uniform == 1, rl);
    }
    switch (s->resident_cfg) {
        case 1: return launch_resident_t<4, 12>(s, n_steps);
        case 2: return launch_resident_t<2, 16>(s, n_steps);
        case 3: return launch_resident_t<4, 8>(s, n_steps);
        case 4: return launch_resident_t<3, 12>(s, n_steps);
        default: return launch_resident_t<3, 16>(s, n_steps);
    }
}

// ---- tiny grids (grid_small.cuh) ------------------------------------------------------------------------
static bool small_grid(const fdtd2d_sim* s) { return s->Rg < 2 * RING + 1 || s->C < 2 * RING + 1; }

template <typename T> static int launch_small_t(fdtd2d_sim* s, int n_steps, int phases) {
    TilePlan tp;
    tp.k = n_steps;
    PassParams<T> p;
    fill_params(s, tp, phases, &p);
    grid_small_kernel<T><<<s->batch, SMALL_NT, 0, s->stream>>>(p);  // in place: the current state stays current
    CUDA_TRY(cudaGetLastError());
    s->launches += 1;
    s->passes += 1;
    return 0;
}

static int launch_small(fdtd2d_sim* s, int n_steps, int phases = FDTD2D_PHASE_H | FDTD2D_PHASE_E | FDTD2D_PHASE_SRC) {
    return s->dtype == FDTD2D_F64 ? launch_small_t<double>(s, n_steps, phases) : launch_small_t<float>(s, n_steps, phases);
}

static bool uses_hybrid(const fdtd2d_sim* s, int phases) {
    const int all = FDTD2D_PHASE_H | FDTD2D_PHASE_E | FDTD2D_PHASE_SRC;
    return phases == all && s->variant != 1;
}

static int run_pass(fdtd2d_sim* s, int k, int phases) {
    int rc;
    if (uses_hybrid(s, phases))
        rc = launch_hybrid(s, k, 0);
    else
        rc = s->dtype == FDTD2D_F64 ? launch_generic_all<double>(s, k, phases) : launch_generic_all<float>(s, k, phases);
    if (rc) return rc;
    s->cur ^= 1;
    s->passes += 1;
    if (peer_mode(s)) s->pass_seq += 1;
    return 0;
}

// ---- y-slab peer links -------------------------------------------------------------------------------
// What a slab tells its neighbours (fdtd2d_peer_export / fdtd2d_peer_attach): geometry for the checks, the device
// pointers of its two field sets and its flag block -- used as they are inside one process -- and their CUDA IPC
// handles for a neighbour in another process.
struct PeerBlob {
    uint32_t magic;
    int32_t pid, device, dtype, Rg, C, row0, Rl, row_begin, row_end, halo, cur;
    uint64_t pitch;
    uint64_t ptr[7];            // field[0][0..2], field[1][0..2], flags
    cudaIpcMemHandle_t ipc[7];
};
static_assert(sizeof(PeerBlob) <= FDTD2D_PEER_BLOB_BYTES, "peer blob does not fit the size the header promises");
constexpr uint32_t PEER_MAGIC = 0x46443250u;  // "FD2P"

static void peer_close(fdtd2d_sim* s, int side) {
    PeerLink& pe = s->peer[side];
    if (pe.attached && pe.ipc) {
        for (int h = 0; h < 2; ++h)
            for (int f = 0; f < 3; ++f)
                if (pe.field[h][f]) cudaIpcCloseMemHandle(pe.field[h][f]);
        if (pe.flags) cudaIpcCloseMemHandle(pe.flags);
    }
    pe = PeerLink();
}

// Host wait for THIS handle's work: its own stream entirely; on a stream it was given (and may share with other handles,
// whose queued work is none of its business) up to the last work it put there; and its copies.
static int wait_own_work(fdtd2d_sim* s) {
    if (s->stream == s->own_stream)
        CUDA_TRY(cudaStreamSynchronize(s->stream));
    else
        CUDA_TRY(cudaEventSynchronize(s->ev_work));
    CUDA_TRY(cudaStreamSynchronize(s->copy_stream));
    return 0;
}

// Read the flag block; wait (on the host) until both attached neighbours have delivered the ghost rows of the current
// state, so that a download that follows sees them.  A timeout inside a kernel or here is an error.
static int peer_settle(fdtd2d_sim* s) {
    if (!peer_mode(s)) return 0;
    const auto t0 = std::chrono::steady_clock::now();
    for (;;) {
        unsigned f[FLAG_WORDS];
        CUDA_TRY(cudaMemcpyAsync(f, s->d_slab_flags, sizeof f, cudaMemcpyDeviceToHost, s->copy_stream));
        CUDA_TRY(cudaStreamSynchronize(s->copy_stream));
        if (f[FLAG_ERR]) return fail(FDTD2D_ESTATE, "halo wait timed out: the %s neighbour slab did not deliver its rows (are all slabs stepping the same passes?)",
                                     f[FLAG_ERR] == 1 ? "top" : "bottom");
        const bool top_ok = !s->peer[0].attached || (int)(f[FLAG_IN_TOP] - s->pass_seq) >= 0;
        const bool bot_ok = !s->peer[1].attached || (int)(f[FLAG_IN_BOT] - s->pass_seq) >= 0;
        if (top_ok && bot_ok) {
            s->settled_seq = s->pass_seq;
            return 0;
        }
        if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(20))
            return fail(FDTD2D_ESTATE, "halo rows of state %u have not arrived after 20 s (flags %u / %u)", s->pass_seq, f[FLAG_IN_TOP], f[FLAG_IN_BOT]);
        std::this_thread::sleep_for(std::chrono::microseconds(50));
    }
}

// 2-D copy between a dense host array (rows x width elements) and the padded device layout
static int copy2d(const fdtd2d_sim* s, void* dev, void* host, int rows, int width, bool to_device, cudaStream_t st) {
    const size_t wbytes = (size_t)width * s->esize;
    if ((size_t)width == s->pitch) {  // no padding on either side: one linear copy
        CUDA_TRY(cudaMemcpyAsync(to_device ? dev : host, to_device ? host : dev, wbytes * rows,
                                 to_device ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost, st));
        return 0;
    }
    if (to_device)
        CUDA_TRY(cudaMemcpy2DAsync(dev, s->pitch * s->esize, host, wbytes, wbytes, rows, cudaMemcpyHostToDevice,
                                   st));
    else
        CUDA_TRY(cudaMemcpy2DAsync(host, wbytes, dev, s->pitch * s->esize, wbytes, rows, cudaMemcpyDeviceToHost,
                                   st));
    return 0;
}

static int transfer_field(fdtd2d_sim* s, void* dev, void* host, int rows_host, int width, bool to_device, cudaStream_t st) {
    // per grid: rows_host x width on the host, Rl x pitch on the device
    if (rows_host == s->Rl && (size_t)s->batch * rows_host < (size_t)0x7fffffff)
        // the grids are back to back on both sides: the whole batch is one strided copy (1024 small grids would
        // otherwise be 1024 copies of 256 KB each)
        return copy2d(s, dev, host, s->batch * rows_host, width, to_device, st);
    if (s->batch > 1) {  // Hy: one row fewer per grid on the host than on the device -> one 3-D copy for the batch
        const size_t wbytes = (size_t)width * s->esize;
        cudaMemcpy3DParms c = {};
        const cudaPitchedPtr h = make_cudaPitchedPtr(host, wbytes, wbytes, (size_t)rows_host);
        const cudaPitchedPtr d = make_cudaPitchedPtr(dev, s->pitch * s->esize, s->pitch * s->esize, (size_t)s->Rl);
        c.srcPtr = to_device ? h : d;
        c.dstPtr = to_device ? d : h;
        c.extent = make_cudaExtent(wbytes, (size_t)rows_host, (size_t)s->batch);
        c.kind = to_device ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost;
        CUDA_TRY(cudaMemcpy3DAsync(&c, st));
        return 0;
    }
    for (int b = 0; b < s->batch; ++b) {
        char* d = static_cast<char*>(dev) + (size_t)b * s->grid_elems * s->esize;
        char* h = static_cast<char*>(host) + (size_t)b * rows_host * width * s->esize;
        if (int rc = copy2d(s, d, h, rows_host, width, to_device, st)) return rc;
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
extern "C" {

int fdtd2d_abi_version(void) { return FDTD2D_ABI_VERSION; }

const char* fdtd2d_last_error(void) { return g_err.c_str(); }

int fdtd2d_device_count(int* count) {
    REQUIRE(count, "count is null");
    *count = 0;
    CUDA_TRY(cudaGetDeviceCount(count));
    return 0;
}

static int create_impl(fdtd2d_sim** out, int Rg, int C, int row_begin, int row_end, int halo, int dtype, int device,
                       int batch) {
    REQUIRE(out, "out is null");
    *out = nullptr;
    REQUIRE(Rg >= SMALL_MIN && C >= SMALL_MIN, "rows and cols must be >= %d (got %d x %d): the reference's own boundary code indexes 6 cells deep", SMALL_MIN, Rg, C);
    REQUIRE(dtype == FDTD2D_F32 || dtype == FDTD2D_F64, "dtype must be FDTD2D_F32 or FDTD2D_F64");
    REQUIRE(batch >= 1, "batch must be >= 1");
    REQUIRE(0 <= row_begin && row_begin < row_end && row_end <= Rg, "bad slab rows [%d, %d) of %d", row_begin, row_end, Rg);
    REQUIRE(halo >= 0 && halo <= FDTD2D_MAX_K, "halo must be in [0, %d]", FDTD2D_MAX_K);
    const bool top_nb = row_begin > 0, bot_nb = row_end < Rg;
    if (top_nb || bot_nb) {
        REQUIRE(batch == 1, "slab handles must have batch = 1");
        REQUIRE(halo >= 1, "a slab with neighbours needs halo >= 1");
        REQUIRE(Rg >= 2 * RING + 1 && C >= 2 * RING + 1, "slabs need a grid of at least 11 x 11");
        REQUIRE(row_end - row_begin >= 2 * FDTD2D_MAX_K + 2 * RING, "slab of %d rows is too thin", row_end - row_begin);
    }
    int ndev = 0;
    CUDA_TRY(cudaGetDeviceCount(&ndev));
    if (ndev <= 0) return fail(FDTD2D_ECUDA, "no CUDA device (libfdtd2d has no CPU fallback)");
    REQUIRE(device >= 0 && device < ndev, "device %d out of range (%d devices)", device, ndev);
    DeviceGuard guard(device);
    if (guard.err != cudaSuccess) return fail(FDTD2D_ECUDA, "cudaSetDevice(%d) failed: %s", device, cudaGetErrorString(guard.err));

    const int row0_ = row_begin - (top_nb ? halo : 0);
    const int rl_ = (row_end + (bot_nb ? halo : 0)) - row0_;
    REQUIRE(row0_ >= 0 && row0_ + rl_ <= Rg, "halo reaches outside the grid");

    fdtd2d_sim* s = new (std::nothrow) fdtd2d_sim();
    if (!s) return fail(FDTD2D_ENOMEM, "host allocation failed");
    s->dtype = dtype;
    s->device = device;
    s->batch = batch;
    s->Rg = Rg;
    s->C = C;
    s->row_begin = row_begin;
    s->row_end = row_end;
    s->halo = halo;
    s->has_top_nb = top_nb;
    s->has_bot_nb = bot_nb;
    s->row0 = row0_;
    s->Rl = rl_;
    s->esize = dtype == FDTD2D_F32 ? 4 : 8;
    s->pitch = round_up((size_t)C, 128 / s->esize);  // rows start on 128-byte lines
    s->grid_elems = (size_t)s->Rl * s->pitch;
    s->opt = options_from_env();  // the environment is read here, once; fdtd2d_set_option changes this handle only
    const size_t bytes = s->grid_elems * s->esize * (size_t)batch;

    cudaError_t e = cudaStreamCreateWithFlags(&s->own_stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        delete s;
        return fail(FDTD2D_ECUDA, "cudaStreamCreate failed: %s", cudaGetErrorString(e));
    }
    s->stream = s->own_stream;
    if (cudaStreamCreateWithFlags(&s->copy_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&s->ev_copy, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&s->ev_work, cudaEventDisableTiming) != cudaSuccess) {
        fdtd2d_destroy(s);
        return fail(FDTD2D_ECUDA, "stream / event creation failed");
    }
    void** bufs[10] = {&s->field[0][0], &s->field[0][1], &s->field[0][2], &s->field[1][0], &s->field[1][1],
                       &s->field[1][2], &s->ce,          &s->ch,          &s->mur,         reinterpret_cast<void**>(&s->d_slab_flags)};
    for (int i = 0; i < 10; ++i) {
        const size_t nb = i == 8 ? (size_t)batch * s->esize : (i == 9 ? FLAG_WORDS * sizeof(unsigned) : bytes);
        e = cudaMalloc(bufs[i], nb);
        if (e == cudaSuccess) e = cudaMemsetAsync(*bufs[i], 0, nb, s->stream);
        if (e != cudaSuccess) {
            fdtd2d_destroy(s);
            return fail(e == cudaErrorMemoryAllocation ? FDTD2D_ENOMEM : FDTD2D_ECUDA,
                        "device allocation of %zu bytes failed: %s", nb, cudaGetErrorString(e));
        }
    }
    // (the flag block must read zero before a neighbour can possibly raise it)
    e = cudaStreamSynchronize(s->stream);
    if (e != cudaSuccess) {
        fdtd2d_destroy(s);
        return fail(FDTD2D_ECUDA, "device initialisation failed: %s", cudaGetErrorString(e));
    }
    *out = s;
    return 0;
}

int fdtd2d_create(fdtd2d_sim** out, int rows, int cols, int dtype, int device, int batch) {
    return create_impl(out, rows, cols, 0, rows, 0, dtype, device, batch);
}

int fdtd2d_create_slab(fdtd2d_sim** out, int global_rows, int cols, int row_begin, int row_end, int halo, int dtype,
                       int device) {
    return create_impl(out, global_rows, cols, row_begin, row_end, halo, dtype, device, 1);
}

int fdtd2d_destroy(fdtd2d_sim* s) {
    if (!s) return 0;
    DeviceGuard guard(s->device);
    if (s->stream) cudaStreamSynchronize(s->stream);
    peer_close(s, 0);
    peer_close(s, 1);
    for (int h = 0; h < 2; ++h)
        for (int f = 0; f < 3; ++f) cudaFree(s->field[h][f]);
    cudaFree(s->ce);
    cudaFree(s->ch);
    cudaFree(s->mur);
    cudaFree(s->d_src);
    cudaFree(s->d_src_range);
    cudaFree(s->d_amp);
    cudaFree(s->d_probe);
    cudaFree(s->d_probe_range);
    cudaFree(s->d_trace);
    cudaFree(s->d_gray);
    cudaFree(s->d_rgb);
    cudaFree(s->d_lut);
    cudaFree(s->d_canvas);
    cudaFree(s->d_flag);
    if (s->h_check) cudaFreeHost(s->h_check);
    if (s->ev_check) cudaEventDestroy(s->ev_check);
    cudaFree(s->d_slab_flags);
    free_plans(s);
    if (s->side_stream) cudaStreamDestroy(s->side_stream);
    if (s->copy_stream) cudaStreamDestroy(s->copy_stream);
    if (s->ev_fork) cudaEventDestroy(s->ev_fork);
    if (s->ev_join) cudaEventDestroy(s->ev_join);
    if (s->ev_copy) cudaEventDestroy(s->ev_copy);
    if (s->ev_work) cudaEventDestroy(s->ev_work);
    if (s->own_stream) cudaStreamDestroy(s->own_stream);
    delete s;
    return 0;
}

int fdtd2d_set_option(fdtd2d_sim* s, const char* key, int value) {
    REQUIRE(s && key, "null argument");
    const OptionKey* k = find_option(key);
    REQUIRE(k, "unknown option '%s'", key);
    if (s->opt.*(k->field) == value) return 0;
    USE_DEVICE(s);
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    s->opt.*(k->field) = value;
    s->ch_uniform = -1;  // (uniform_ch decides how the result of the check is used)
    free_plans(s);
    return 0;
}

int fdtd2d_get_option(const fdtd2d_sim* s, const char* key, int* value) {
    REQUIRE(s && key && value, "null argument");
    const OptionKey* k = find_option(key);
    REQUIRE(k, "unknown option '%s'", key);
    *value = s->opt.*(k->field);
    return 0;
}

int fdtd2d_set_stream(fdtd2d_sim* s, void* cuda_stream) {
    REQUIRE(s, "handle is null");
    USE_DEVICE(s);
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    s->stream = static_cast<cudaStream_t>(cuda_stream);
    return 0;
}

int fdtd2d_reset_stream(fdtd2d_sim* s) {
    REQUIRE(s, "handle is null");
    USE_DEVICE(s);
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    s->stream = s->own_stream;
    return 0;
}

int fdtd2d_get_stream(const fdtd2d_sim* s, void** cuda_stream) {
    REQUIRE(s && cuda_stream, "null argument");
    *cuda_stream = s->stream;
    return 0;
}

int fdtd2d_sync(fdtd2d_sim* s) {
    REQUIRE(s, "handle is null");
    USE_DEVICE(s);
    if (int rc = wait_own_work(s)) return rc;
    if (s->fused_pairs != s->fused_checked) {  // a phase-1 run that gave up waiting for its producers leaves a mark
        unsigned err = 0;
        CUDA_TRY(cudaMemcpyAsync(&err, s->d_slab_flags + FLAG_ERR, sizeof err, cudaMemcpyDeviceToHost, s->copy_stream));
        CUDA_TRY(cudaStreamSynchronize(s->copy_stream));
        s->fused_checked = s->fused_pairs;
        if (err == 3) return fail(FDTD2D_ESTATE, "a fused double pass timed out waiting for its first pass (internal error)");
    }
    return peer_settle(s);
}

int fdtd2d_geometry(const fdtd2d_sim* s, int* local_rows, int* cols, int* row0, int* global_rows, int* batch,
                    int* dtype, size_t* pitch_elems) {
    REQUIRE(s, "handle is null");
    if (local_rows) *local_rows = s->Rl;
    if (cols) *cols = s->C;
    if (row0) *row0 = s->row0;
    if (global_rows) *global_rows = s->Rg;
    if (batch) *batch = s->batch;
    if (dtype) *dtype = s->dtype;
    if (pitch_elems) *pitch_elems = s->pitch;
    return 0;
}

static int upload_state_on(fdtd2d_sim* s, const void* Ez, const void* Hx, const void* Hy, cudaStream_t st) {
    void** f = s->field[s->cur];
    if (int rc = transfer_field(s, f[0], const_cast<void*>(Ez), s->Rl, s->C, true, st)) return rc;
    if (int rc = transfer_field(s, f[1], const_cast<void*>(Hx), s->Rl, s->C - 1, true, st)) return rc;
    if (int rc = transfer_field(s, f[2], const_cast<void*>(Hy), hy_rows(s), s->C, true, st)) return rc;
    return 0;
}

int fdtd2d_upload_state(fdtd2d_sim* s, const void* Ez, const void* Hx, const void* Hy) {
    REQUIRE(s && Ez && Hx && Hy, "null argument");
    USE_DEVICE(s);
    if (int rc = begin_work(s)) return rc;
    if (int rc = upload_state_on(s, Ez, Hx, Hy, s->stream)) return rc;
    return mark_work(s);
}

int fdtd2d_download_state(fdtd2d_sim* s, void* Ez, void* Hx, void* Hy) {
    REQUIRE(s, "handle is null");
    USE_DEVICE(s);
    if (int rc = begin_work(s)) return rc;
    if (peer_mode(s)) {  // a slab's ghost rows are written by its neighbours
        if (int rc = wait_own_work(s)) return rc;
        if (int rc = peer_settle(s)) return rc;
    }
    void** f = s->field[s->cur];
    if (Ez)
        if (int rc = transfer_field(s, f[0], Ez, s->Rl, s->C, false, s->stream)) return rc;
    if (Hx)
        if (int rc = transfer_field(s, f[1], Hx, s->Rl, s->C - 1, false, s->stream)) return rc;
    if (Hy)
        if (int rc = transfer_field(s, f[2], Hy, hy_rows(s), s->C, false, s->stream)) return rc;
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    return 0;
}

// ---- asynchronous copies (pinned host memory): one host thread can keep several handles busy -------------------
// The copy runs on the handle's copy stream, ordered after the stepping work issued so far (a download reads the
// result of the last fdtd2d_step; an upload waits until the stepping kernels that still read the old state are done);
// stepping work issued AFTER the call is ordered behind the copy.  The host does not block; fdtd2d_copy_wait (or
// fdtd2d_sync) does.
static int copy_fork(fdtd2d_sim* s) {
    if (s->stream == s->own_stream)
        if (int rc = mark_work(s)) return rc;  // a private stream: its tail is exactly this handle's work
    CUDA_TRY(cudaStreamWaitEvent(s->copy_stream, s->ev_work, 0));
    return 0;
}
// (the stream does not wait here: the handle's NEXT work does, in begin_work -- work of other handles that share the
// stream and is queued in between has no reason to wait for this copy)
static int copy_join(fdtd2d_sim* s) {
    CUDA_TRY(cudaEventRecord(s->ev_copy, s->copy_stream));
    s->copy_pending = true;
    return 0;
}

int fdtd2d_upload_state_async(fdtd2d_sim* s, const void* Ez, const void* Hx, const void* Hy) {
    REQUIRE(s && Ez && Hx && Hy, "null argument");
    USE_DEVICE(s);
    if (int rc = copy_fork(s)) return rc;
    if (int rc = upload_state_on(s, Ez, Hx, Hy, s->copy_stream)) return rc;
    return copy_join(s);
}

int fdtd2d_download_state_async(fdtd2d_sim* s, void* Ez, void* Hx, void* Hy) {
    REQUIRE(s, "handle is null");
    USE_DEVICE(s);
    if (int rc = copy_fork(s)) return rc;
    void** f = s->field[s->cur];
    if (Ez)
        if (int rc = transfer_field(s, f[0], Ez, s->Rl, s->C, false, s->copy_stream)) return rc;
    if (Hx)
        if (int rc = transfer_field(s, f[1], Hx, s->Rl, s->C - 1, false, s->copy_stream)) return rc;
    if (Hy)
        if (int rc = transfer_field(s, f[2], Hy, hy_rows(s), s->C, false, s->copy_stream)) return rc;
    return copy_join(s);
}

int fdtd2d_copy_wait(fdtd2d_sim* s) {
    REQUIRE(s, "handle is null");
    USE_DEVICE(s);
    CUDA_TRY(cudaStreamSynchronize(s->copy_stream));
    return 0;
}

int fdtd2d_zero_state(fdtd2d_sim* s) {
    REQUIRE(s, "handle is null");
    USE_DEVICE(s);
    if (int rc = begin_work(s)) return rc;
    if (peer_mode(s) && s->settled_seq != s->pass_seq) {
        // the neighbours' last ghost rows must have landed before they are cleared (nothing to wait for when the handle
        // was synchronised after its last step: a job pipeline that re-uses the handle then queues this without blocking)
        if (int rc = wait_own_work(s)) return rc;
        if (int rc = peer_settle(s)) return rc;
    }
    const size_t bytes = s->grid_elems * s->esize * (size_t)s->batch;
    // With peer links only the current set is cleared: the other set's ghost rows belong to the neighbours, whose first
    // pass after their own zero_state may already be storing into them (every pass overwrites all owned rows of its
    // output set, so stale values there are never read).
    for (int h = 0; h < 2; ++h)
        if (!peer_mode(s) || h == s->cur)
            for (int f = 0; f < 3; ++f) CUDA_TRY(cudaMemsetAsync(s->field[h][f], 0, bytes, s->stream));
    s->step = 0;
    return mark_work(s);
}

// the maps changed: the permeability check starts over (the cached plans stay -- a job that uploads new maps for every run
// must not pay a cudaFree, which waits for the whole device; the one plan that depends on the check, the 12-level
// wavefront, is re-examined when it is next used: launch_hybrid_t)
static void materials_changed(fdtd2d_sim* s) {
    s->coeffs_set = true;
    s->ch_uniform = -1;
    s->check_pending = false;
}

int fdtd2d_set_mur_coef(fdtd2d_sim* s, const void* mur_coef) {
    REQUIRE(s && mur_coef, "null argument");
    USE_DEVICE(s);
    CUDA_TRY(cudaMemcpyAsync(s->mur, mur_coef, (size_t)s->batch * s->esize, cudaMemcpyHostToDevice, s->stream));
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    s->mur_set = true;
    return 0;
}

int fdtd2d_set_coeffs(fdtd2d_sim* s, const void* ce, const void* ch, const void* mur_coef) {
    REQUIRE(s && ce && ch, "null argument");
    USE_DEVICE(s);
    if (int rc = begin_work(s)) return rc;
    if (int rc = transfer_field(s, s->ce, const_cast<void*>(ce), s->Rl, s->C, true, s->stream)) return rc;
    if (int rc = transfer_field(s, s->ch, const_cast<void*>(ch), s->Rl, s->C, true, s->stream)) return rc;
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    materials_changed(s);
    if (mur_coef) return fdtd2d_set_mur_coef(s, mur_coef);
    return 0;
}

static int finish_materials(fdtd2d_sim* s, double dt, double dx, bool wait, bool on_copy_stream = false);

int fdtd2d_set_materials(fdtd2d_sim* s, const void* eps, const void* mu, double dt, double dx) {
    REQUIRE(s && eps && mu, "null argument");
    USE_DEVICE(s);
    if (int rc = begin_work(s)) return rc;
    // stage eps in ce and mu in ch, then transform in place on the device
    if (int rc = transfer_field(s, s->ce, const_cast<void*>(eps), s->Rl, s->C, true, s->stream)) return rc;
    if (int rc = transfer_field(s, s->ch, const_cast<void*>(mu), s->Rl, s->C, true, s->stream)) return rc;
    return finish_materials(s, dt, dx, true);
}

int fdtd2d_set_materials_async(fdtd2d_sim* s, const void* eps, const void* mu, double dt, double dx) {
    REQUIRE(s && eps && mu, "null argument");
    USE_DEVICE(s);
    if (int rc = copy_fork(s)) return rc;
    if (int rc = transfer_field(s, s->ce, const_cast<void*>(eps), s->Rl, s->C, true, s->copy_stream)) return rc;
    if (int rc = transfer_field(s, s->ch, const_cast<void*>(mu), s->Rl, s->C, true, s->copy_stream)) return rc;
    return finish_materials(s, dt, dx, false, true);
}

double fdtd2d_hash_uniform(uint64_t seed, uint32_t grid, uint32_t row, uint32_t col) {
    return hash_uniform(seed, grid, row, col);
}

int fdtd2d_set_materials_random(fdtd2d_sim* s, uint64_t seed, double span, double dt, double dx) {
    REQUIRE(s, "handle is null");
    USE_DEVICE(s);
    if (int rc = begin_work(s)) return rc;
    const double eps0 = 8.85418e-12, mu0 = 4 * 3.141592653589793 * 1e-7;  // main.py:100-101
    const long long n = (long long)s->Rl * s->C * s->batch;
    const int blocks = (int)std::min<long long>((n + 255) / 256, sm_count(s) * 16);
    if (s->dtype == FDTD2D_F32)
        random_materials_kernel<float><<<blocks, 256, 0, s->stream>>>(
            (float*)s->ce, (float*)s->ch, (float*)s->mur, s->Rl, s->C, (int)s->pitch, s->row0, (long long)s->grid_elems,
            s->batch, seed, (float)span, (float)eps0, (float)mu0, (float)dt, (float)dx);
    else
        random_materials_kernel<double><<<blocks, 256, 0, s->stream>>>(
            (double*)s->ce, (double*)s->ch, (double*)s->mur, s->Rl, s->C, (int)s->pitch, s->row0,
            (long long)s->grid_elems, s->batch, seed, span, eps0, mu0, dt, dx);
    CUDA_TRY(cudaGetLastError());
    s->launches += 1;
    materials_changed(s);
    s->mur_set = true;
    return mark_work(s);
}

// eps/mu are staged in ce/ch: form the Mur coefficient(s) and the coefficient maps in place (device-side tail of
// every fdtd2d_set_materials* entry point)
// on_copy_stream: the asynchronous form -- the kernels follow the uploads on the copy stream (the handle's next stepping
// work waits for them through ev_copy), so that on a compute stream shared with other handles nothing of this job queues
// behind their kernels; the permeability check rides along.
static int finish_materials(fdtd2d_sim* s, double dt, double dx, bool wait, bool on_copy_stream) {
    if (!on_copy_stream)
        if (int rc = begin_work(s)) return rc;  // (an asynchronous upload of eps / mu may still be in flight)
    cudaStream_t st = on_copy_stream ? s->copy_stream : s->stream;
    const long long n = (long long)s->grid_elems * s->batch;
    const int blocks = (int)std::min<long long>((n + 255) / 256, sm_count(s) * 16);
    const bool has_corner = s->row0 == 0;
    if (s->dtype == FDTD2D_F32) {
        if (has_corner)
            mur_from_materials_kernel<float><<<(s->batch + 127) / 128, 128, 0, st>>>(
                (const float*)s->ce, (const float*)s->ch, (long long)s->grid_elems, s->batch, (float)dt, (float)dx,
                (float*)s->mur);
        coeff_from_materials_kernel<float><<<blocks, 256, 0, st>>>((float*)s->ce, (float*)s->ch, n, (float)dt,
                                                                         (float)dx);
    } else {
        if (has_corner)
            mur_from_materials_kernel<double><<<(s->batch + 127) / 128, 128, 0, st>>>(
                (const double*)s->ce, (const double*)s->ch, (long long)s->grid_elems, s->batch, dt, dx, (double*)s->mur);
        coeff_from_materials_kernel<double><<<blocks, 256, 0, st>>>((double*)s->ce, (double*)s->ch, n, dt, dx);
    }
    CUDA_TRY(cudaGetLastError());
    if (wait) CUDA_TRY(cudaStreamSynchronize(st));
    s->launches += has_corner ? 2 : 1;
    materials_changed(s);
    if (has_corner) s->mur_set = true;
    if (on_copy_stream) {
        if (int rc = enqueue_ch_check(s, st)) return rc;
        s->check_pending = true;
        return copy_join(s);
    }
    return mark_work(s);
}

int fdtd2d_set_materials_gray(fdtd2d_sim* s, const unsigned char* gray, double black_point, double dt, double dx) {
    REQUIRE(s && gray, "null argument");
    USE_DEVICE(s);
    if (int rc = begin_work(s)) return rc;
    const double eps0 = 8.85418e-12, mu0 = 4 * 3.141592653589793 * 1e-7;  // main.py:100-101
    const size_t n = (size_t)s->Rl * s->C * s->batch;
    unsigned char* d_g = nullptr;
    CUDA_TRY(cudaMalloc(&d_g, n));
    cudaError_t e = cudaMemcpyAsync(d_g, gray, n, cudaMemcpyHostToDevice, s->stream);
    if (e != cudaSuccess) {
        cudaFree(d_g);
        return fail(FDTD2D_ECUDA, "gray upload failed: %s", cudaGetErrorString(e));
    }
    const int blocks = (int)std::min<size_t>((n + 255) / 256, (size_t)sm_count(s) * 16);
    if (s->dtype == FDTD2D_F32)
        gray_materials_kernel<float><<<blocks, 256, 0, s->stream>>>(d_g, (float*)s->ce, (float*)s->ch, s->Rl, s->C, (int)s->pitch,
                                                                   (long long)s->grid_elems, s->batch, black_point, eps0, mu0);
    else
        gray_materials_kernel<double><<<blocks, 256, 0, s->stream>>>(d_g, (double*)s->ce, (double*)s->ch, s->Rl, s->C,
                                                                    (int)s->pitch, (long long)s->grid_elems, s->batch, black_point,
                                                                    eps0, mu0);
    e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(s->stream);
    cudaFree(d_g);
    if (e != cudaSuccess) return fail(FDTD2D_ECUDA, "gray_materials_kernel failed: %s", cudaGetErrorString(e));
    s->launches += 1;
    return finish_materials(s, dt, dx, true);
}

int fdtd2d_generate_materials_blobs(fdtd2d_sim* s, uint64_t seed, const float* weights, double eps_lo, double eps_hi, double mu,
                                    double dt, double dx, void* eps_out) {
    REQUIRE(s && weights, "null argument");
    REQUIRE(!s->has_top_nb && !s->has_bot_nb, "the blob generator works on whole grids, not slabs");
    USE_DEVICE(s);
    if (int rc = begin_work(s)) return rc;
    float* d_w = nullptr;
    const size_t wbytes = sizeof(float) * BLOB_K * BLOB_K * (size_t)s->batch;
    CUDA_TRY(cudaMalloc(&d_w, wbytes));
    cudaError_t e = cudaMemcpyAsync(d_w, weights, wbytes, cudaMemcpyHostToDevice, s->stream);
    const dim3 grid((s->C + BLOB_T - 1) / BLOB_T, (s->Rg + BLOB_T - 1) / BLOB_T, s->batch);
    if (e == cudaSuccess) {
        if (s->dtype == FDTD2D_F32)
            blob_materials_kernel<float><<<grid, 256, 0, s->stream>>>((float*)s->ce, (float*)s->ch, s->Rg, s->C, (int)s->pitch,
                                                                     (long long)s->grid_elems, seed, d_w, (float)eps_lo,
                                                                     (float)eps_hi, (float)mu);
        else
            blob_materials_kernel<double><<<grid, 256, 0, s->stream>>>((double*)s->ce, (double*)s->ch, s->Rg, s->C, (int)s->pitch,
                                                                      (long long)s->grid_elems, seed, d_w, eps_lo, eps_hi, mu);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(s->stream);
    cudaFree(d_w);
    if (e != cudaSuccess) return fail(FDTD2D_ECUDA, "blob_materials_kernel failed: %s", cudaGetErrorString(e));
    s->launches += 1;
    if (eps_out) {  // the permittivity maps themselves are part of a dataset sample
        if (int rc = transfer_field(s, s->ce, eps_out, s->Rl, s->C, false, s->stream)) return rc;
        CUDA_TRY(cudaStreamSynchronize(s->stream));
    }
    return finish_materials(s, dt, dx, true);
}

int fdtd2d_download_coeffs(fdtd2d_sim* s, void* ce, void* ch, void* mur_coef) {
    REQUIRE(s, "handle is null");
    USE_DEVICE(s);
    if (int rc = begin_work(s)) return rc;
    if (ce)
        if (int rc = transfer_field(s, s->ce, ce, s->Rl, s->C, false, s->stream)) return rc;
    if (ch)
        if (int rc = transfer_field(s, s->ch, ch, s->Rl, s->C, false, s->stream)) return rc;
    if (mur_coef)
        CUDA_TRY(cudaMemcpyAsync(mur_coef, s->mur, (size_t)s->batch * s->esize, cudaMemcpyDeviceToHost, s->stream));
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    return 0;
}

// sort cells by grid; build [batch+1] ranges; reject out-of-range and duplicate cells
static int build_cells(const fdtd2d_sim* s, int n, const int32_t* grid, const int32_t* row, const int32_t* col,
                       const int32_t* wave, int n_waves, std::vector<Cell>* cells, std::vector<int>* range,
                       std::vector<int>* perm) {
    std::vector<int> order(n);
    for (int i = 0; i < n; ++i) order[i] = i;
    for (int i = 0; i < n; ++i) {
        const int g = grid ? grid[i] : 0;
        REQUIRE(g >= 0 && g < s->batch, "cell %d: grid %d out of range", i, g);
        REQUIRE(row[i] >= 0 && row[i] < s->Rg && col[i] >= 0 && col[i] < s->C, "cell %d: (%d, %d) outside %d x %d", i,
                row[i], col[i], s->Rg, s->C);
        if (wave) REQUIRE(wave[i] >= 0 && wave[i] < n_waves, "cell %d: waveform %d out of range", i, wave[i]);
    }
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
        const int ga = grid ? grid[a] : 0, gb = grid ? grid[b] : 0;
        if (ga != gb) return ga < gb;
        if (row[a] != row[b]) return row[a] < row[b];
        return col[a] < col[b];
    });
    cells->resize(n);
    range->assign(s->batch + 1, 0);
    for (int q = 0; q < n; ++q) {
        const int i = order[q];
        Cell c;
        c.grid = grid ? grid[i] : 0;
        c.row = row[i];
        c.col = col[i];
        c.wave = wave ? wave[i] : 0;
        (*cells)[q] = c;
        (*range)[c.grid + 1] += 1;
        if (wave && q > 0) {  // duplicate source cells would make the float64-add order ambiguous
            const Cell& pc = (*cells)[q - 1];
            REQUIRE(!(pc.grid == c.grid && pc.row == c.row && pc.col == c.col), "duplicate source cell (%d, %d) in grid %d",
                    c.row, c.col, c.grid);
        }
    }
    for (int b = 0; b < s->batch; ++b) (*range)[b + 1] += (*range)[b];
    if (perm) *perm = order;
    return 0;
}

int fdtd2d_set_sources(fdtd2d_sim* s, int n_cells, const int32_t* grid, const int32_t* row, const int32_t* col,
                       const int32_t* wave, int n_waves, int n_steps, const double* tables) {
    REQUIRE(s, "handle is null");
    USE_DEVICE(s);
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    cudaFree(s->d_src);
    cudaFree(s->d_src_range);
    cudaFree(s->d_amp);
    s->d_src = nullptr;
    s->d_src_range = nullptr;
    s->d_amp = nullptr;
    s->n_src = s->n_waves = s->amp_steps = 0;
    s->h_src.clear();
    free_plans(s);
    if (n_cells == 0) return 0;
    REQUIRE(n_cells > 0 && row && col && wave && tables && n_waves > 0 && n_steps > 0, "bad source arguments");
    std::vector<Cell> cells;
    std::vector<int> range;
    if (int rc = build_cells(s, n_cells, grid, row, col, wave, n_waves, &cells, &range, nullptr)) return rc;
    CUDA_TRY(cudaMalloc(&s->d_src, sizeof(Cell) * n_cells));
    CUDA_TRY(cudaMalloc(&s->d_src_range, sizeof(int) * range.size()));
    CUDA_TRY(cudaMalloc(&s->d_amp, sizeof(double) * (size_t)n_waves * n_steps));
    // (stream-ordered copies, then a wait: see the note at the wavefront task list)
    CUDA_TRY(cudaMemcpyAsync(s->d_src, cells.data(), sizeof(Cell) * n_cells, cudaMemcpyHostToDevice, s->stream));
    CUDA_TRY(cudaMemcpyAsync(s->d_src_range, range.data(), sizeof(int) * range.size(), cudaMemcpyHostToDevice, s->stream));
    CUDA_TRY(cudaMemcpyAsync(s->d_amp, tables, sizeof(double) * (size_t)n_waves * n_steps, cudaMemcpyHostToDevice, s->stream));
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    s->h_src = cells;
    s->n_src = n_cells;
    s->n_waves = n_waves;
    s->amp_steps = n_steps;
    return 0;
}

int fdtd2d_set_probes(fdtd2d_sim* s, int n_probes, const int32_t* grid, const int32_t* row, const int32_t* col,
                      int capacity_steps) {
    REQUIRE(s, "handle is null");
    USE_DEVICE(s);
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    cudaFree(s->d_probe);
    cudaFree(s->d_probe_range);
    cudaFree(s->d_trace);
    s->d_probe = nullptr;
    s->d_probe_range = nullptr;
    s->d_trace = nullptr;
    s->n_probe = 0;
    s->trace_cap = 0;
    s->probe_perm.clear();
    s->h_probe.clear();
    free_plans(s);
    if (n_probes == 0) return 0;
    REQUIRE(n_probes > 0 && row && col && capacity_steps > 0, "bad probe arguments");
    std::vector<Cell> cells;
    std::vector<int> range;
    if (int rc = build_cells(s, n_probes, grid, row, col, nullptr, 0, &cells, &range, &s->probe_perm)) return rc;
    const size_t tbytes = (size_t)capacity_steps * n_probes * s->esize;
    CUDA_TRY(cudaMalloc(&s->d_probe, sizeof(Cell) * n_probes));
    CUDA_TRY(cudaMalloc(&s->d_probe_range, sizeof(int) * range.size()));
    CUDA_TRY(cudaMalloc(&s->d_trace, tbytes));
    CUDA_TRY(cudaMemcpyAsync(s->d_probe, cells.data(), sizeof(Cell) * n_probes, cudaMemcpyHostToDevice, s->stream));
    CUDA_TRY(cudaMemcpyAsync(s->d_probe_range, range.data(), sizeof(int) * range.size(), cudaMemcpyHostToDevice, s->stream));
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    CUDA_TRY(cudaMemsetAsync(s->d_trace, 0, tbytes, s->stream));
    s->h_probe = cells;
    s->n_probe = n_probes;
    s->trace_cap = capacity_steps;
    return 0;
}

int fdtd2d_read_probes(fdtd2d_sim* s, void* out, int64_t first_step, int n_steps) {
    REQUIRE(s && out, "null argument");
    REQUIRE(s->n_probe > 0, "no probes set");
    REQUIRE(first_step >= 0 && n_steps >= 0 && first_step + n_steps <= s->trace_cap, "probe rows [%lld, %lld) outside capacity %lld",
            (long long)first_step, (long long)(first_step + n_steps), s->trace_cap);
    USE_DEVICE(s);
    const size_t row_bytes = (size_t)s->n_probe * s->esize;
    std::vector<char> tmp(row_bytes * (size_t)n_steps);
    if (int rc = copy_fork(s)) return rc;  // behind this handle's stepping work, not behind whatever else shares its stream
    CUDA_TRY(cudaMemcpyAsync(tmp.data(), static_cast<char*>(s->d_trace) + (size_t)first_step * row_bytes, tmp.size(),
                             cudaMemcpyDeviceToHost, s->copy_stream));
    CUDA_TRY(cudaStreamSynchronize(s->copy_stream));
    // device columns are in sorted order; give them back in the caller's order
    char* o = static_cast<char*>(out);
    for (int t = 0; t < n_steps; ++t)
        for (int q = 0; q < s->n_probe; ++q)
            memcpy(o + (size_t)t * row_bytes + (size_t)s->probe_perm[q] * s->esize,
                   tmp.data() + (size_t)t * row_bytes + (size_t)q * s->esize, s->esize);
    return 0;
}

int fdtd2d_step(fdtd2d_sim* s, int n_steps, int k_temporal) {
    REQUIRE(s, "handle is null");
    REQUIRE(n_steps >= 0, "n_steps must be >= 0");
    REQUIRE(k_temporal >= 0 && k_temporal <= FDTD2D_MAX_K, "k_temporal must be in [0, %d]", FDTD2D_MAX_K);
    if (!s->coeffs_set || !s->mur_set) return fail(FDTD2D_ESTATE, "coefficients / Mur coefficient not set");
    USE_DEVICE(s);
    if (int rc = begin_work(s)) return rc;
    if (n_steps > 0 && small_grid(s)) {
        // rows or cols below 11: the reference's boundary statements overlap, so they are executed one by one
        if (int rc = launch_small(s, n_steps)) return rc;
        s->step += n_steps;
        return mark_work(s);
    }
    if (n_steps > 0 && (s->variant == 0 || s->variant == 4) && resident_eligible(s)) {
        // the whole run in one launch, the grid resident on chip (k_temporal does not apply)
        if (int rc = launch_resident(s, n_steps)) return rc;
        s->step += n_steps;
        return mark_work(s);
    }
    if (s->variant == 4 && n_steps > 0) return fail(FDTD2D_EINVAL, "variant 4 (cluster-resident) needs fp32, 16..256 columns, 16..384 rows, no slabs");
    const bool slab = s->has_top_nb || s->has_bot_nb;
    int k = k_temporal;
    if (!k) {
        if (s->dtype == FDTD2D_F32) {
            k = 8;
            if (s->opt.auto_k12 && s->variant != 1 && !slab && n_steps >= 12) {
                // large grid with uniform permeability: the 12-level wavefront moves a third less DRAM traffic per step, but on
                // B200 it is bound by latency (2 warps per scheduler at 255 registers), not by DRAM: 1418 against 1564 Gcell/s
                // at 16384^2 -- so it is opt-in (k_temporal = 12, or this option for the automatic choice)
                PassPlan& pl = s->hybrid[12];
                if (!pl.valid)
                    if (int rc = classify_tiles(s, 12, &pl)) return rc;
                if (pl.d_wave) k = 12;
            }
        } else {
            // fp64: 8 levels where the wavefront kernel takes the grid or where the k = 8 tiles fit one wave of CTAs (small
            // grids are bound by launch and barrier latency: 200^2 runs 275k steps/s at k = 8, 217k at k = 4 -- and 12 steps per
            // launch where even those tiles fit one wave); otherwise 4
            k = s->opt.f64_k > 0 ? std::min(s->opt.f64_k, FDTD2D_MAX_K) : 8;
            if (s->opt.f64_k <= 0) {
                if (s->variant == 1) {
                    k = 4;
                } else {
                    PassPlan& pl = s->hybrid[8];
                    if (!pl.valid)
                        if (int rc = classify_tiles(s, 8, &pl)) return rc;
                    if (!pl.d_wave) {
                        if (pl.n_edge > sm_count(s)) {
                            k = 4;
                        } else if (!slab && n_steps >= FDTD2D_MAX_K) {  // still one wave of CTAs at 12 steps per launch? (200^2: 307k steps/s)
                            PassPlan& p12 = s->hybrid[FDTD2D_MAX_K];
                            if (!p12.valid)
                                if (int rc = classify_tiles(s, FDTD2D_MAX_K, &p12)) return rc;
                            if (!p12.d_wave && p12.n_edge <= sm_count(s)) k = FDTD2D_MAX_K;
                        }
                    }
                }
            }
        }
    }
    if (slab) {
        k = std::min(k, s->halo);
        // the host layer exchanges halos after every pass unless the peer links do it inside the kernels
        const bool linked = (!s->has_top_nb || s->peer[0].attached) && (!s->has_bot_nb || s->peer[1].attached);
        REQUIRE(linked || n_steps <= s->halo, "a slab handle without peer links can advance at most halo=%d steps between halo exchanges", s->halo);
    }
    int left = n_steps;
    bool fuse = false;
    if (k == 8 && left >= 16)
        if (int rc = fused_available(s, &fuse)) return rc;
    while (left > 0) {
        if (fuse && left >= 16) {  // two passes per launch, the second one fed from L2
            if (int rc = launch_fused_pair(s)) return rc;
            s->step += 16;
            left -= 16;
            continue;
        }
        const int kk = std::min(k, left);
        if (int rc = run_pass(s, kk, FDTD2D_PHASE_H | FDTD2D_PHASE_E | FDTD2D_PHASE_SRC)) return rc;
        s->step += kk;
        left -= kk;
    }
    return mark_work(s);
}

int fdtd2d_step_phases(fdtd2d_sim* s, int phases) {
    REQUIRE(s, "handle is null");
    REQUIRE(phases > 0 && phases < 8, "phases must be a non-empty FDTD2D_PHASE_* mask");
    if (!s->coeffs_set || !s->mur_set) return fail(FDTD2D_ESTATE, "coefficients / Mur coefficient not set");
    USE_DEVICE(s);
    if (int rc = begin_work(s)) return rc;
    if (small_grid(s)) {
        if (int rc = launch_small(s, 1, phases)) return rc;
    } else if (int rc = run_pass(s, 1, phases)) {
        return rc;
    }
    if (phases & FDTD2D_PHASE_SRC) s->step += 1;
    return mark_work(s);
}

int fdtd2d_get_step_index(const fdtd2d_sim* s, int64_t* step) {
    REQUIRE(s && step, "null argument");
    *step = s->step;
    return 0;
}

int fdtd2d_set_step_index(fdtd2d_sim* s, int64_t step) {
    REQUIRE(s && step >= 0, "bad argument");
    s->step = step;
    return 0;
}

int fdtd2d_source_steps(const fdtd2d_sim* s, int* n_cells, int* n_steps, int64_t* probe_capacity) {
    REQUIRE(s, "handle is null");
    if (n_cells) *n_cells = s->n_src;
    if (n_steps) *n_steps = s->amp_steps;
    if (probe_capacity) *probe_capacity = s->n_probe ? s->trace_cap : 0;
    return 0;
}

int fdtd2d_set_kernel_variant(fdtd2d_sim* s, int variant) {
    REQUIRE(s && variant >= 0 && variant <= 4, "bad argument");
    if (variant == s->variant) return 0;
    USE_DEVICE(s);
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    s->variant = variant;
    free_plans(s);
    return 0;
}

int fdtd2d_plan_wave_runs(int n_stretches, const int32_t* rows, const uint8_t* ring, int warps, int cap_rows, int k, int32_t* parts,
                          int32_t* run_rows) {
    REQUIRE(n_stretches >= 0 && warps > 0 && cap_rows > 0 && k > 0 && (n_stretches == 0 || (rows && ring && parts)), "bad argument");
    std::vector<int> r(rows, rows + n_stretches), pr;
    std::vector<unsigned char> g(ring, ring + n_stretches);
    for (int v : r) REQUIRE(v > 0, "a stretch has at least one row");
    const int len = plan_wave_runs(r, g, warps, cap_rows, k, &pr);
    for (int i = 0; i < n_stretches; ++i) parts[i] = pr[i];
    if (run_rows) *run_rows = len;
    return 0;
}

int fdtd2d_plan_edge_reserve(int64_t n_edge, int64_t wave_rows, int sm_count, int k) {
    if (k < 1 || k > FDTD2D_MAX_K) return 0;
    return plan_edge_reserve(n_edge, wave_rows, sm_count, k);
}

int fdtd2d_plan_resident(int rows, int cols, int cfg, int cluster, int32_t* out) {
    REQUIRE(out && rows > 0 && cols > 0, "bad argument");
    fdtd2d_sim s;  // geometry only: no device resources are created or touched
    s.dtype = FDTD2D_F32, s.batch = 1, s.Rg = rows, s.C = cols, s.row_begin = 0, s.row_end = rows;
    s.opt = Options();
    s.opt.resident_cfg = cfg, s.opt.resident_cluster = cluster;
    out[0] = out[1] = out[2] = out[3] = -1;
    if (resident_eligible(&s)) out[0] = s.resident_cfg, out[1] = s.resident_cluster, out[2] = s.resident_edge, out[3] = s.resident_rpc;
    return 0;
}

int fdtd2d_plan_host(const int32_t* geom, int n_src, const int32_t* src, int n_probe, const int32_t* probe, int32_t* plan,
                     int32_t* tile_kind, int cap_tiles, int32_t* tasks, int cap_tasks) {
    REQUIRE(geom && plan && (n_src == 0 || src) && (n_probe == 0 || probe) && n_src >= 0 && n_probe >= 0, "bad argument");
    fdtd2d_sim s;  // geometry only: no device resources are created or touched
    s.dtype = geom[0], s.batch = geom[1], s.Rg = geom[2], s.C = geom[3], s.row_begin = geom[4], s.row_end = geom[5], s.halo = geom[6];
    const int k = geom[7];
    s.sm_count = geom[8], s.variant = geom[9];
    s.opt = Options();
    s.opt.wave_min_tiles = geom[10], s.opt.ring_min_tiles = geom[11], s.opt.wavefront = geom[12], s.opt.ring_strips = geom[13];
    s.ch_uniform = geom[14];
    REQUIRE((s.dtype == FDTD2D_F32 || s.dtype == FDTD2D_F64) && s.batch >= 1 && s.Rg >= 2 * RING + 1 && s.C >= 2 * RING + 1 && k >= 1 && k <= FDTD2D_MAX_K &&
                0 <= s.row_begin && s.row_begin < s.row_end && s.row_end <= s.Rg && s.halo >= 0 && s.halo <= FDTD2D_MAX_K && s.sm_count > 0,
            "bad geometry");
    s.has_top_nb = s.row_begin > 0, s.has_bot_nb = s.row_end < s.Rg;
    REQUIRE(!(s.has_top_nb || s.has_bot_nb) || (s.halo >= k && s.batch == 1 && s.row_end - s.row_begin >= 2 * FDTD2D_MAX_K + 2 * RING), "bad slab");
    s.row0 = s.row_begin - (s.has_top_nb ? s.halo : 0);
    s.Rl = s.row_end + (s.has_bot_nb ? s.halo : 0) - s.row0;
    s.esize = s.dtype == FDTD2D_F32 ? 4 : 8;
    s.pitch = round_up((size_t)s.C, 128 / s.esize);
    for (int i = 0; i < n_src; ++i) s.h_src.push_back(Cell{src[3 * i], src[3 * i + 1], src[3 * i + 2], 0});
    for (int i = 0; i < n_probe; ++i) s.h_probe.push_back(Cell{probe[3 * i], probe[3 * i + 1], probe[3 * i + 2], 0});
    PlanLists L;
    if (int rc = plan_pass(&s, k, &L)) return rc;
    const int32_t v[16] = {L.tp.tiles_y, L.tp.tiles_x, L.tp.CH, L.tp.CW, L.tp.hx, s.row_begin - s.row0, s.Rl, (int32_t)L.edge.size(), L.n_edge_band,
                           (int32_t)L.fast.size(), (int32_t)L.tasks.size(), L.n_wave_band, L.wave_ring ? 1 : 0, L.band_expected[0], L.band_expected[1],
                           (int32_t)s.pitch};
    memcpy(plan, v, sizeof v);
    if (tile_kind) {
        REQUIRE((size_t)cap_tiles >= L.kind.size(), "tile_kind holds %d entries, %zu needed", cap_tiles, L.kind.size());
        for (size_t i = 0; i < L.kind.size(); ++i) tile_kind[i] = L.kind[i];
    }
    if (tasks) {
        REQUIRE((size_t)cap_tasks >= L.tasks.size(), "tasks holds %d entries, %zu needed", cap_tasks, L.tasks.size());
        for (size_t i = 0; i < L.tasks.size(); ++i) memcpy(tasks + 8 * i, &L.tasks[i], 8 * sizeof(int32_t));
    }
    return 0;
}

int fdtd2d_plan_host_fused(const int32_t* geom, int fuse, int n_src, const int32_t* src, int n_probe, const int32_t* probe, int32_t* counts,
                           int32_t* fused, int cap_fused, int32_t* deferred, int cap_deferred) {
    REQUIRE(geom && counts && (n_src == 0 || src) && (n_probe == 0 || probe) && n_src >= 0 && n_probe >= 0, "bad argument");
    fdtd2d_sim s;  // geometry only
    s.dtype = geom[0], s.batch = geom[1], s.Rg = geom[2], s.C = geom[3], s.row_begin = 0, s.row_end = geom[2], s.halo = 0;
    s.sm_count = geom[8], s.variant = geom[9];
    s.opt = Options();
    s.opt.wave_min_tiles = geom[10], s.opt.ring_min_tiles = geom[11], s.opt.wavefront = geom[12], s.opt.ring_strips = geom[13];
    s.opt.fuse = fuse;
    s.ch_uniform = geom[14];
    REQUIRE(s.dtype == FDTD2D_F32 && s.batch >= 1 && s.Rg >= 2 * RING + 1 && s.C >= 2 * RING + 1 && s.sm_count > 0, "bad geometry");
    s.row0 = 0, s.Rl = s.Rg, s.esize = 4;
    s.pitch = round_up((size_t)s.C, 32);
    for (int i = 0; i < n_src; ++i) s.h_src.push_back(Cell{src[3 * i], src[3 * i + 1], src[3 * i + 2], 0});
    for (int i = 0; i < n_probe; ++i) s.h_probe.push_back(Cell{probe[3 * i], probe[3 * i + 1], probe[3 * i + 2], 0});
    PlanLists L;
    if (int rc = plan_pass(&s, 8, &L)) return rc;
    counts[0] = (int32_t)L.fused.size(), counts[1] = (int32_t)L.deferred.size(), counts[2] = L.fuse_nblk, counts[3] = (int32_t)L.tasks.size();
    if (fused) {
        REQUIRE((size_t)cap_fused >= L.fused.size(), "fused holds %d entries, %zu needed", cap_fused, L.fused.size());
        if (!L.fused.empty()) memcpy(fused, L.fused.data(), sizeof(WaveTask) * L.fused.size());
    }
    if (deferred) {
        REQUIRE((size_t)cap_deferred >= L.deferred.size(), "deferred holds %d entries, %zu needed", cap_deferred, L.deferred.size());
        if (!L.deferred.empty()) memcpy(deferred, L.deferred.data(), sizeof(WaveTask) * L.deferred.size());
    }
    return 0;
}

int fdtd2d_plan_info(fdtd2d_sim* s, int k, int32_t* info, int n_info) {
    REQUIRE(s && info && n_info >= 0 && k >= 1 && k <= FDTD2D_MAX_K, "bad argument");
    if (!s->coeffs_set) return fail(FDTD2D_ESTATE, "coefficients not set");
    USE_DEVICE(s);
    PassPlan& pl = s->hybrid[k];
    if (!pl.valid)
        if (int rc = classify_tiles(s, k, &pl)) return rc;
    const int32_t v[FDTD2D_PLAN_INFO_WORDS] = {pl.tp.tiles_y, pl.tp.tiles_x, pl.tp.CH, pl.tp.CW, pl.n_edge, pl.n_edge_band, pl.n_fast,
                                              pl.n_wave, pl.n_wave_band, pl.wave_ring ? 1 : 0, pl.band_expected[0], pl.band_expected[1],
                                              pl.reserve_sms};
    for (int i = 0; i < n_info && i < FDTD2D_PLAN_INFO_WORDS; ++i) info[i] = v[i];
    return 0;
}

int fdtd2d_pass_count(const fdtd2d_sim* s, int64_t* passes) {
    REQUIRE(s && passes, "null argument");
    *passes = s->passes;
    return 0;
}

int fdtd2d_launch_count(const fdtd2d_sim* s, int64_t* launches) {
    REQUIRE(s && launches, "null argument");
    *launches = s->launches;
    return 0;
}

static int halo_block_impl(fdtd2d_sim* s, int state, int field, int side, void** send_ptr, void** recv_ptr, size_t* nbytes) {
    REQUIRE(s && field >= 0 && field < 3 && (side == 0 || side == 1), "bad argument");
    const bool has = side == 0 ? s->has_top_nb : s->has_bot_nb;
    REQUIRE(has, "no neighbour on that side");
    char* base = static_cast<char*>(s->field[state][field]);
    const size_t row_bytes = s->pitch * s->esize;
    const int h = s->halo;
    // local rows: [0,h) top ghosts | owned | [Rl-h, Rl) bottom ghosts
    const int send_row = side == 0 ? own_first(s) : own_last(s) - h;
    const int recv_row = side == 0 ? 0 : own_last(s);
    if (send_ptr) *send_ptr = base + (size_t)send_row * row_bytes;
    if (recv_ptr) *recv_ptr = base + (size_t)recv_row * row_bytes;
    if (nbytes) *nbytes = (size_t)h * row_bytes;
    return 0;
}

int fdtd2d_halo_block(fdtd2d_sim* s, int field, int side, void** send_ptr, void** recv_ptr, size_t* nbytes) {
    REQUIRE(s, "handle is null");
    return halo_block_impl(s, s->cur, field, side, send_ptr, recv_ptr, nbytes);
}

int fdtd2d_halo_block_next(fdtd2d_sim* s, int field, int side, void** send_ptr, void** recv_ptr, size_t* nbytes) {
    REQUIRE(s, "handle is null");
    return halo_block_impl(s, s->cur ^ 1, field, side, send_ptr, recv_ptr, nbytes);
}

int fdtd2d_pass_begin(fdtd2d_sim* s, int k) {
    REQUIRE(s, "handle is null");
    REQUIRE(k >= 1 && k <= FDTD2D_MAX_K && k <= std::max(1, s->halo), "k must be in [1, min(halo, %d)]", FDTD2D_MAX_K);
    REQUIRE(s->open_pass_k == 0, "a pass is already open");
    if (!s->coeffs_set || !s->mur_set) return fail(FDTD2D_ESTATE, "coefficients / Mur coefficient not set");
    USE_DEVICE(s);
    if (int rc = begin_work(s)) return rc;
    const int all = FDTD2D_PHASE_H | FDTD2D_PHASE_E | FDTD2D_PHASE_SRC;
    int rc;
    if (uses_hybrid(s, all))
        rc = launch_hybrid(s, k, 1);
    else  // no band split for the generic kernel: do the whole pass now
        rc = s->dtype == FDTD2D_F64 ? launch_generic_all<double>(s, k, all) : launch_generic_all<float>(s, k, all);
    if (rc) return rc;
    s->open_pass_k = k;
    return 0;
}

int fdtd2d_pass_end(fdtd2d_sim* s) {
    REQUIRE(s, "handle is null");
    REQUIRE(s->open_pass_k > 0, "no open pass");
    USE_DEVICE(s);
    const int k = s->open_pass_k;
    if (uses_hybrid(s, FDTD2D_PHASE_H | FDTD2D_PHASE_E | FDTD2D_PHASE_SRC))
        if (int rc = launch_hybrid(s, k, 2)) return rc;
    s->open_pass_k = 0;
    s->cur ^= 1;
    s->passes += 1;
    if (peer_mode(s)) s->pass_seq += 1;
    s->step += k;
    return mark_work(s);
}

// ---- peer links: the halo exchange inside the stepping kernels (NVLink peer stores + flags) ----------------------
int fdtd2d_peer_export(fdtd2d_sim* s, void* blob) {
    REQUIRE(s && blob, "null argument");
    REQUIRE(s->has_top_nb || s->has_bot_nb, "not a slab handle");
    USE_DEVICE(s);
    PeerBlob b;
    memset(&b, 0, sizeof b);
    b.magic = PEER_MAGIC;
    b.pid = (int32_t)getpid();
    b.device = s->device, b.dtype = s->dtype, b.Rg = s->Rg, b.C = s->C, b.row0 = s->row0, b.Rl = s->Rl;
    b.row_begin = s->row_begin, b.row_end = s->row_end, b.halo = s->halo, b.cur = s->cur;
    b.pitch = s->pitch;
    void* ptrs[7] = {s->field[0][0], s->field[0][1], s->field[0][2], s->field[1][0], s->field[1][1], s->field[1][2], s->d_slab_flags};
    for (int i = 0; i < 7; ++i) {
        b.ptr[i] = reinterpret_cast<uint64_t>(ptrs[i]);
        CUDA_TRY(cudaIpcGetMemHandle(&b.ipc[i], ptrs[i]));
    }
    memset(blob, 0, FDTD2D_PEER_BLOB_BYTES);
    memcpy(blob, &b, sizeof b);
    return 0;
}

int fdtd2d_peer_attach(fdtd2d_sim* s, int side, const void* blob) {
    REQUIRE(s && blob && (side == 0 || side == 1), "bad argument");
    REQUIRE(side == 0 ? s->has_top_nb : s->has_bot_nb, "no neighbour on that side");
    REQUIRE(s->variant != 1, "the generic kernel (variant 1) does not take part in the peer halo exchange");
    PeerBlob b;
    memcpy(&b, blob, sizeof b);
    REQUIRE(b.magic == PEER_MAGIC, "not a peer blob");
    REQUIRE(b.dtype == s->dtype && b.Rg == s->Rg && b.C == s->C && b.pitch == s->pitch && b.halo == s->halo,
            "the neighbour is a slab of a different grid (%d x %d, dtype %d, halo %d)", b.Rg, b.C, b.dtype, b.halo);
    REQUIRE(side == 0 ? b.row_end == s->row_begin : b.row_begin == s->row_end, "the neighbour's rows [%d, %d) do not touch mine [%d, %d) on that side",
            b.row_begin, b.row_end, s->row_begin, s->row_end);
    REQUIRE(b.cur == s->cur, "the neighbour has stepped a different number of passes: attach before stepping");
    USE_DEVICE(s);
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    peer_close(s, side);
    PeerLink pe;
    pe.row0 = b.row0;
    pe.device = b.device;
    void* ptrs[7];
    if (b.pid == (int32_t)getpid()) {  // same process: the pointers are valid as they are
        if (b.device != s->device) {
            int can = 0;
            CUDA_TRY(cudaDeviceCanAccessPeer(&can, s->device, b.device));
            REQUIRE(can, "device %d cannot access device %d", s->device, b.device);
            const cudaError_t e = cudaDeviceEnablePeerAccess(b.device, 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled)
                cudaGetLastError();
            else
                CUDA_TRY(e);
        }
        for (int i = 0; i < 7; ++i) ptrs[i] = reinterpret_cast<void*>(b.ptr[i]);
    } else {
        for (int i = 0; i < 7; ++i) {
            const cudaError_t e = cudaIpcOpenMemHandle(&ptrs[i], b.ipc[i], cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) {
                for (int j = 0; j < i; ++j) cudaIpcCloseMemHandle(ptrs[j]);
                return fail(FDTD2D_ECUDA, "cudaIpcOpenMemHandle failed: %s", cudaGetErrorString(e));
            }
        }
        pe.ipc = true;
    }
    for (int i = 0; i < 6; ++i) pe.field[i / 3][i % 3] = ptrs[i];
    pe.flags = static_cast<unsigned*>(ptrs[6]);
    pe.attached = true;
    s->peer[side] = pe;
    free_plans(s);
    return 0;
}

int fdtd2d_peer_detach(fdtd2d_sim* s) {
    REQUIRE(s, "handle is null");
    USE_DEVICE(s);
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    peer_close(s, 0);
    peer_close(s, 1);
    return 0;
}

int fdtd2d_peer_status(fdtd2d_sim* s, uint32_t* flags_out) {
    REQUIRE(s && flags_out, "null argument");
    USE_DEVICE(s);
    flags_out[0] = (s->peer[0].attached ? 1u : 0u) | (s->peer[1].attached ? 2u : 0u);
    flags_out[1] = s->pass_seq;
    unsigned f[FLAG_WORDS] = {0};
    if (s->d_slab_flags) {
        CUDA_TRY(cudaMemcpyAsync(f, s->d_slab_flags, sizeof f, cudaMemcpyDeviceToHost, s->copy_stream));
        CUDA_TRY(cudaStreamSynchronize(s->copy_stream));
    }
    flags_out[2] = f[FLAG_IN_TOP], flags_out[3] = f[FLAG_IN_BOT], flags_out[4] = f[FLAG_ERR];
    return 0;
}

// ---- structure drawing on the device (structure.cuh): replaces RegionDrawer (region_drawer.py:5-87) + the image -> eps
// mapping of material_init (main.py:109-123) ------------------------------------------------------------------
static Canvas canvas_of(const fdtd2d_sim* s) {
    Canvas cv;
    cv.px = s->d_canvas, cv.Rl = s->Rl, cv.C = s->C, cv.row0 = s->row0, cv.grid_stride = (long long)s->Rl * s->C;
    return cv;
}
static int canvas_blocks(fdtd2d_sim* s, long long n) { return (int)std::max<long long>(1, std::min<long long>((n + 255) / 256, (long long)sm_count(s) * 16)); }

int fdtd2d_canvas_clear(fdtd2d_sim* s, int value) {
    REQUIRE(s && value >= 0 && value <= 255, "bad argument");
    USE_DEVICE(s);
    const long long n = (long long)s->Rl * s->C * s->batch;
    if (!s->d_canvas) CUDA_TRY(cudaMalloc(&s->d_canvas, (size_t)n));
    canvas_fill_kernel<<<canvas_blocks(s, n), 256, 0, s->stream>>>(s->d_canvas, n, (unsigned char)value);
    CUDA_TRY(cudaGetLastError());
    s->launches += 1;
    return 0;
}

int fdtd2d_canvas_rect(fdtd2d_sim* s, int grid, int x0, int y0, int x1, int y1, int value) {
    REQUIRE(s && grid >= 0 && grid < s->batch && value >= 0 && value <= 255, "bad argument");
    if (!s->d_canvas) return fail(FDTD2D_ESTATE, "no canvas: call fdtd2d_canvas_clear first");
    USE_DEVICE(s);
    // clip to the canvas of this handle (y in global rows)
    x0 = std::max(x0, 0), x1 = std::min(x1, s->C - 1), y0 = std::max(y0, s->row0), y1 = std::min(y1, s->row0 + s->Rl - 1);
    if (x0 > x1 || y0 > y1) return 0;
    const long long n = (long long)(x1 - x0 + 1) * (y1 - y0 + 1);
    canvas_rect_kernel<<<canvas_blocks(s, n), 256, 0, s->stream>>>(canvas_of(s), grid, x0, y0, x1, y1, (unsigned char)value);
    CUDA_TRY(cudaGetLastError());
    s->launches += 1;
    return 0;
}

int fdtd2d_canvas_ellipse(fdtd2d_sim* s, int grid, int x0, int y0, int x1, int y1, int width, int value) {
    REQUIRE(s && grid >= 0 && grid < s->batch && value >= 0 && value <= 255, "bad argument");
    if (!s->d_canvas) return fail(FDTD2D_ESTATE, "no canvas: call fdtd2d_canvas_clear first");
    if (x1 < x0 || y1 < y0 || (x1 - x0) + (y1 - y0) < 1) return 0;  // (PIL draws nothing for an inverted or a one-cell box)
    USE_DEVICE(s);
    const int a = x1 - x0, b = y1 - y0;
    int ix0 = 0, iy0 = 0, ix1 = -1, iy1 = -1;  // inner box of a ring (empty: filled ellipse)
    if (width > 0 && a - 2 * width >= 0 && b - 2 * width >= 0 && a + b - 4 * width >= 1) ix0 = x0 + width, iy0 = y0 + width, ix1 = x1 - width, iy1 = y1 - width;
    const int rows = b / 2 + 1, irows = ix1 >= ix0 ? (iy1 - iy0) / 2 + 1 : 0;
    int* d_half = nullptr;
    CUDA_TRY(cudaMalloc(&d_half, sizeof(int) * (size_t)(rows + irows + 1)));
    ellipse_walk_kernel<<<1, 32, 0, s->stream>>>(a, b, d_half);
    if (irows) ellipse_walk_kernel<<<1, 32, 0, s->stream>>>(ix1 - ix0, iy1 - iy0, d_half + rows);
    const long long n = (long long)(a + 1) * (b + 1);
    canvas_ellipse_kernel<<<canvas_blocks(s, n), 256, 0, s->stream>>>(canvas_of(s), grid, x0, y0, x1, y1, d_half, ix0, iy0, ix1, iy1, d_half + rows,
                                                                   (unsigned char)value);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(s->stream);
    cudaFree(d_half);
    if (e != cudaSuccess) return fail(FDTD2D_ECUDA, "ellipse kernels failed: %s", cudaGetErrorString(e));
    s->launches += irows ? 3 : 2;
    return 0;
}

int fdtd2d_canvas_segment(fdtd2d_sim* s, int grid, double x0, double y0, double x1, double y1, double width, int value) {
    REQUIRE(s && grid >= 0 && grid < s->batch && value >= 0 && value <= 255 && width >= 0, "bad argument");
    if (!s->d_canvas) return fail(FDTD2D_ESTATE, "no canvas: call fdtd2d_canvas_clear first");
    USE_DEVICE(s);
    const double hw = 0.5 * width + 1.0;
    const int bx0 = std::max(0, (int)floor(std::min(x0, x1) - hw)), bx1 = std::min(s->C - 1, (int)ceil(std::max(x0, x1) + hw));
    const int by0 = std::max(s->row0, (int)floor(std::min(y0, y1) - hw)), by1 = std::min(s->row0 + s->Rl - 1, (int)ceil(std::max(y0, y1) + hw));
    if (bx0 > bx1 || by0 > by1) return 0;
    const long long n = (long long)(bx1 - bx0 + 1) * (by1 - by0 + 1);
    canvas_segment_kernel<<<canvas_blocks(s, n), 256, 0, s->stream>>>(canvas_of(s), grid, x0, y0, x1, y1, width, bx0, by0, bx1, by1, (unsigned char)value);
    CUDA_TRY(cudaGetLastError());
    s->launches += 1;
    return 0;
}

int fdtd2d_canvas_download(fdtd2d_sim* s, unsigned char* gray) {
    REQUIRE(s && gray, "null argument");
    if (!s->d_canvas) return fail(FDTD2D_ESTATE, "no canvas: call fdtd2d_canvas_clear first");
    USE_DEVICE(s);
    CUDA_TRY(cudaMemcpyAsync(gray, s->d_canvas, (size_t)s->Rl * s->C * s->batch, cudaMemcpyDeviceToHost, s->stream));
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    return 0;
}

int fdtd2d_canvas_apply(fdtd2d_sim* s, double black_point, double dt, double dx) {
    REQUIRE(s, "handle is null");
    if (!s->d_canvas) return fail(FDTD2D_ESTATE, "no canvas: call fdtd2d_canvas_clear first");
    USE_DEVICE(s);
    if (int rc = begin_work(s)) return rc;
    const double eps0 = 8.85418e-12, mu0 = 4 * 3.141592653589793 * 1e-7;  // main.py:100-101
    const long long n = (long long)s->Rl * s->C * s->batch;
    if (s->dtype == FDTD2D_F32)
        gray_materials_kernel<float><<<canvas_blocks(s, n), 256, 0, s->stream>>>(s->d_canvas, (float*)s->ce, (float*)s->ch, s->Rl, s->C, (int)s->pitch,
                                                                               (long long)s->grid_elems, s->batch, black_point, eps0, mu0);
    else
        gray_materials_kernel<double><<<canvas_blocks(s, n), 256, 0, s->stream>>>(s->d_canvas, (double*)s->ce, (double*)s->ch, s->Rl, s->C, (int)s->pitch,
                                                                                (long long)s->grid_elems, s->batch, black_point, eps0, mu0);
    CUDA_TRY(cudaGetLastError());
    s->launches += 1;
    return finish_materials(s, dt, dx, true);
}

int fdtd2d_set_snapshot_background(fdtd2d_sim* s, const unsigned char* gray, const double* lut) {
    REQUIRE(s && gray && lut, "null argument");
    USE_DEVICE(s);
    const size_t n = (size_t)s->Rl * s->C * s->batch;
    if (!s->d_gray) {
        CUDA_TRY(cudaMalloc(&s->d_gray, n));
        CUDA_TRY(cudaMalloc(&s->d_rgb, (size_t)s->Rl * s->C * 3));
        CUDA_TRY(cudaMalloc(&s->d_lut, sizeof(double) * 768));
    }
    CUDA_TRY(cudaMemcpyAsync(s->d_gray, gray, n, cudaMemcpyHostToDevice, s->stream));
    CUDA_TRY(cudaMemcpyAsync(s->d_lut, lut, sizeof(double) * 768, cudaMemcpyHostToDevice, s->stream));
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    return 0;
}

int fdtd2d_render_snapshot(fdtd2d_sim* s, int grid, double vmin, double vmax, unsigned char* out_rgb) {
    REQUIRE(s && out_rgb && grid >= 0 && grid < s->batch, "bad argument");
    if (!s->d_gray) return fail(FDTD2D_ESTATE, "snapshot background not set");
    USE_DEVICE(s);
    if (int rc = begin_work(s)) return rc;
    const long long n = (long long)s->Rl * s->C;
    const int blocks = (int)std::min<long long>((n + 255) / 256, sm_count(s) * 8);
    const unsigned char* gray = s->d_gray + (size_t)grid * n;
    if (s->dtype == FDTD2D_F32) {
        const float* ez = static_cast<const float*>(s->field[s->cur][0]) + (size_t)grid * s->grid_elems;
        // python-float bounds are weak scalars: the whole normalisation runs in float32 (NEP 50)
        snapshot_kernel<float><<<blocks, 256, 0, s->stream>>>(ez, gray, s->d_lut, s->d_rgb, s->Rl, s->C, (int)s->pitch, (float)vmin,
                                                             (float)vmax, (float)(vmax - vmin));
    } else {
        const double* ez = static_cast<const double*>(s->field[s->cur][0]) + (size_t)grid * s->grid_elems;
        snapshot_kernel<double><<<blocks, 256, 0, s->stream>>>(ez, gray, s->d_lut, s->d_rgb, s->Rl, s->C, (int)s->pitch, vmin, vmax,
                                                              vmax - vmin);
    }
    CUDA_TRY(cudaGetLastError());
    s->launches += 1;
    CUDA_TRY(cudaMemcpyAsync(out_rgb, s->d_rgb, (size_t)n * 3, cudaMemcpyDeviceToHost, s->stream));
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    return 0;
}

int fdtd2d_device_field(fdtd2d_sim* s, int field, void** ptr) {
    REQUIRE(s && ptr && field >= 0 && field < 5, "bad argument");
    *ptr = field < 3 ? s->field[s->cur][field] : (field == 3 ? s->ce : s->ch);
    return 0;
}

}  // extern "C"
