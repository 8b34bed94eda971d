// Shared definitions for the fdtd2d CUDA kernels (sm_100a).
//
// Arithmetic discipline: every floating-point operation is a single
// IEEE round-to-nearest op in the run dtype, in the reference's order and parenthesisation
// (python-src/main.py:21-27,34-51,54-61,69-74).  We use the explicit *_rn intrinsics, which the
// compiler never contracts into FMAs, and additionally build with -fmad=false.  Denormals are kept
// (nvcc default -ftz=false).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace fdtd2d {

constexpr int RING = 5;  // depth of the Mur boundary (main.py:33,38,43,48)

__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub_rn(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float div_rn(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ float sqrt_rn(float a) { return __fsqrt_rn(a); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double sub_rn(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double div_rn(double a, double b) { return __ddiv_rn(a, b); }
__device__ __forceinline__ double sqrt_rn(double a) { return __dsqrt_rn(a); }

// Source add of fdtd.py:34: the reference adds a float64 array into Ez in place, i.e. the sum is
// formed in float64 and then cast to Ez's dtype (one extra rounding when Ez is float32).
__device__ __forceinline__ float add_source(float ez, double amp) {
    return __double2float_rn(__dadd_rn((double)ez, amp));
}
__device__ __forceinline__ double add_source(double ez, double amp) { return __dadd_rn(ez, amp); }

template <typename T> struct Vec;
template <> struct Vec<float> {
    using type = float4;
    static constexpr int N = 4;
};
template <> struct Vec<double> {
    using type = double2;
    static constexpr int N = 2;
};

__device__ __forceinline__ void unpack4(const float4& v, float* a) { a[0] = v.x, a[1] = v.y, a[2] = v.z, a[3] = v.w; }

// A source or probe cell: grid index in the batch, GLOBAL row, column, waveform index (sources only).
struct Cell {
    int32_t grid, row, col, wave;
};

// Everything one pass (k leapfrog steps per HBM round trip) needs.  Plain data, passed by value.
template <typename T> struct PassParams {
    const T* in[3];  // Ez, Hx, Hy of the current state
    T* out[3];       // the other half of the ping-pong pair
    const T* ce;     // dt/(eps*dx), (Rl, pitch) per grid
    const T* ch;     // dt/(mu*dx)
    const T* mur;    // Mur coefficient, one per grid
    long long grid_stride;  // elements between consecutive grids of the batch
    int pitch;              // elements per padded row
    int Rg, C;              // global rows, columns
    int row0;               // global index of local row 0
    int Rl;                 // local rows (ghost rows included)
    int own_begin, own_end; // global rows owned by this handle (probes are recorded by the owner)
    int k;                  // leapfrog steps in this pass
    int hx;                 // column halo = k rounded up to the vector width
    int phases;             // FDTD2D_PHASE_* bits
    int CH, CW;             // core tile size
    int tiles_y, tiles_x;   // tiles per grid
    const int* tile_list;   // optional explicit tile ids (else blockIdx.x is the tile id)
    // sources (sorted by grid; src_range[b]..src_range[b+1])
    const Cell* src;
    const int* src_range;
    const double* amp;  // [n_waves][amp_steps]
    int amp_steps;
    long long step0;  // step index of the first leapfrog step of this pass
    // probes (sorted by grid)
    const Cell* probes;
    const int* probe_range;
    int n_probe;
    T* trace;  // [trace_cap][n_probe]
    long long trace_cap;
    // ---- y-slabs ----------------------------------------------------------------------
    int org;                 // local row where the tile grid starts (= first owned row: ghost rows are halo, never core)
    int store_lo, store_hi;  // local rows this handle stores: its owned rows
    int band_lo[2], band_hi[2];  // local rows next to the top / bottom neighbour whose results the neighbour needs (its ghost rows)
    // peer mode: the tasks that produce band rows store them into the neighbour's ghost rows as well (NVLink peer stores)
    // and the last one to finish raises the neighbour's flag; the tasks that read ghost rows wait for their own flag.
    T* peer_out[2][3];        // the state set the neighbour's pass is writing (mapped peer memory); null = no peer protocol
    long long peer_shift[2];  // element offset: my local offset + shift = the same cell in the neighbour's arrays
    unsigned* flags;          // this handle's flag block (FLAG_* below); the IN words are written by the neighbours
    unsigned* peer_flag[2];   // the word of the neighbour's block that I raise (its IN word for my side)
    int band_expected[2];     // band tasks (wavefront runs + edge tiles) per side in this pass
    unsigned seq;             // sequence number of the state this pass reads (passes since the handle was created)
    // fused double pass (strip_wave.cuh): one flag per (grid, tile column, 16-row block) of what phase 0 has stored
    unsigned* fuse_flags;
    int fuse_nblk;
};

// Flag block of a slab handle: 8 words of device memory.  IN_TOP / IN_BOT = sequence number of the newest state whose
// ghost rows the top / bottom neighbour has delivered; CNT_* = band tasks of the running pass that are done (reset by
// the host before every pass); ERR != 0: a wait for a neighbour timed out (reported by fdtd2d_sync).
enum { FLAG_IN_TOP = 0, FLAG_IN_BOT = 1, FLAG_CNT_TOP = 2, FLAG_CNT_BOT = 3, FLAG_ERR = 4, FLAG_WORDS = 8 };
constexpr unsigned long long HALO_WAIT_NS = 2000000000ull;  // give up on a neighbour after 2 s instead of hanging the GPU

__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// the neighbour's arrays of side 0 or 1 (selects, not indexed parameter loads: a run-time index would move the parameter
// block onto the stack)
template <typename T> __device__ __forceinline__ T* peer_field(const PassParams<T>& p, int side, int f) {
    return side ? p.peer_out[1][f] : p.peer_out[0][f];
}
template <typename T> __device__ __forceinline__ long long peer_shift(const PassParams<T>& p, int side) {
    return side ? p.peer_shift[1] : p.peer_shift[0];
}

// Which band a local row belongs to: 0 top, 1 bottom, -1 none.
template <typename T> __device__ __forceinline__ int band_of_row(const PassParams<T>& p, int row) {
    if (row >= p.band_lo[0] && row < p.band_hi[0]) return 0;
    if (row >= p.band_lo[1] && row < p.band_hi[1]) return 1;
    return -1;
}

// A task that reads ghost rows of `side` calls this (every thread, before its first load): the neighbour's band tasks of
// the previous pass have stored those rows once its flag has reached the sequence number of the state this pass reads.
// Every thread does its own acquire load, so its later loads are ordered behind it without relying on a barrier.
// (parameter arrays are read through selects: a run-time index would move the parameter block onto the stack)
template <typename T> __device__ __forceinline__ void band_wait(const PassParams<T>& p, int side) {
    if (!(side ? p.peer_out[1][0] : p.peer_out[0][0])) return;
    const unsigned* f = p.flags + FLAG_IN_TOP + side;
    if ((int)(ld_acquire_sys(f) - p.seq) >= 0) return;
    const unsigned long long t0 = globaltimer_ns();
    while ((int)(ld_acquire_sys(f) - p.seq) < 0) {
        if (ld_acquire_sys(p.flags + FLAG_ERR)) return;  // somebody already gave up: do not wait 2 s per task
        __nanosleep(256);
        if (globaltimer_ns() - t0 > HALO_WAIT_NS) {
            atomicExch(p.flags + FLAG_ERR, 1u + (unsigned)side);
            return;
        }
    }
}

// One thread of a band task calls this after ALL threads of the task have issued their mirrored stores, executed
// __threadfence_system() and synchronised (warp or CTA barrier): counts the task; the last task of the side publishes.
template <typename T> __device__ __forceinline__ void band_done(const PassParams<T>& p, int side) {
    if (!(side ? p.peer_out[1][0] : p.peer_out[0][0])) return;
    __threadfence_system();  // cumulative: covers the stores of the threads this one has synchronised with
    // (atomicInc wraps to zero at the last task: the counter is ready for the next pass without a memset)
    const unsigned last = (unsigned)(side ? p.band_expected[1] : p.band_expected[0]) - 1u;
    if (atomicInc(p.flags + FLAG_CNT_TOP + side, last) == last) {
        __threadfence_system();
        st_release_sys(side ? p.peer_flag[1] : p.peer_flag[0], p.seq + 1u);
    }
}

// Counter-based uniform in [0,1) with 24 random bits (exact in fp32): splitmix64 finaliser over
// (seed, grid, row, col).  Same code runs on host (fdtd2d_hash_uniform) and device.
__host__ __device__ __forceinline__ double hash_uniform(uint64_t seed, uint32_t grid, uint32_t row, uint32_t col) {
    uint64_t key = ((uint64_t)grid << 40) ^ ((uint64_t)row << 20) ^ (uint64_t)col;
    uint64_t z = key + seed * 0x9E3779B97F4A7C15ull + 0x632BE59BD9B4E019ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z = z ^ (z >> 31);
    return (double)(z >> 40) * (1.0 / 16777216.0);
}

}  // namespace fdtd2d
