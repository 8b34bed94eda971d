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
};

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
