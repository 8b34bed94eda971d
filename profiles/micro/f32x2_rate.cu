// Issue rate of the packed fp32 instructions of sm_100a (FADD2 / FFMA2) against scalar FADD / FMUL.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false -O3 -o f32x2_rate f32x2_rate.cu ; run on one B200.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) { uint64_t c; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(c) : "l"(a), "l"(b)); return c; }
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b, uint64_t nz) { uint64_t c; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(c) : "l"(a), "l"(b), "l"(nz)); return c; }
constexpr int CH = 8, IT = 2048;
// mode 0: scalar FADD, 1: scalar FMUL+FADD alternating, 2: FADD2, 3: FFMA2 (as mul), 4: FFMA2+FADD2 alternating,
// 5: FADD2 + one integer LOP3 each (does the second cycle of a packed instruction leave the issue port free?),
// 6: FADD2 + one shared-memory load each, 7: FADD2 + one shuffle each, 8: scalar FADD + one LOP3 each
template <int MODE> __global__ void rate(float* out, float seed, uint64_t nz, long long* cyc) {
    float x[2 * CH];
    uint64_t y[CH];
    uint32_t z[CH];
    __shared__ uint32_t sm[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = i & 3;
    __syncthreads();
    const uint32_t zk = (uint32_t)nz | 0x1234567u;
    for (int i = 0; i < CH; ++i) z[i] = threadIdx.x * 7 + i;
    for (int i = 0; i < 2 * CH; ++i) x[i] = seed * (threadIdx.x + i);
    for (int i = 0; i < CH; ++i) y[i] = ((uint64_t)__float_as_uint(x[2 * i]) << 32) | __float_as_uint(x[2 * i + 1]);
    const uint64_t c2 = ((uint64_t)__float_as_uint(seed) << 32) | __float_as_uint(seed);
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < IT; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            if (MODE == 0) {
#pragma unroll
                for (int i = 0; i < 2 * CH; ++i) x[i] = __fadd_rn(x[i], seed);
            } else if (MODE == 1) {
#pragma unroll
                for (int i = 0; i < 2 * CH; ++i) x[i] = (r & 1) ? __fmul_rn(x[i], seed) : __fadd_rn(x[i], seed);
            } else if (MODE == 2) {
#pragma unroll
                for (int i = 0; i < CH; ++i) y[i] = add2(y[i], c2);
            } else if (MODE == 3) {
#pragma unroll
                for (int i = 0; i < CH; ++i) y[i] = mul2(y[i], c2, nz);
            } else if (MODE == 4) {
#pragma unroll
                for (int i = 0; i < CH; ++i) y[i] = (r & 1) ? mul2(y[i], c2, nz) : add2(y[i], c2);
            } else if (MODE == 5) {
#pragma unroll
                for (int i = 0; i < CH; ++i) { y[i] = add2(y[i], c2); z[i] = (z[i] ^ zk) + (z[i] >> 1 & 0x55u); }
            } else if (MODE == 9) {
#pragma unroll
                for (int i = 0; i < CH; ++i) { y[i] = add2(y[i], c2); asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(z[i]) : "r"(zk), "r"(z[(i + 1) % CH])); }
            } else if (MODE == 10) {
#pragma unroll
                for (int i = 0; i < CH; ++i) { y[i] = add2(y[i], c2); if (i & 1) z[i] = __shfl_down_sync(0xffffffffu, z[i], 1); }
            } else if (MODE == 6) {
#pragma unroll
                for (int i = 0; i < CH; ++i) { y[i] = add2(y[i], c2); z[i] += sm[(threadIdx.x + 32 * i + z[i]) & 1023]; }
            } else if (MODE == 7) {
#pragma unroll
                for (int i = 0; i < CH; ++i) { y[i] = add2(y[i], c2); z[i] = __shfl_down_sync(0xffffffffu, z[i], 1); }
            } else {
#pragma unroll
                for (int i = 0; i < CH; ++i) { x[i] = __fadd_rn(x[i], seed); z[i] = (z[i] ^ zk) & (z[i] + 0x55u); }
            }
        }
    }
    long long t1 = clock64();
    float s = 0;
    for (int i = 0; i < 2 * CH; ++i) s += x[i];
    for (int i = 0; i < CH; ++i) s += __uint_as_float((uint32_t)y[i]) + __uint_as_float((uint32_t)(y[i] >> 32)) + (float)z[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int MODE> void run(const char* name, int warps, float* out, long long* cyc) {
    rate<MODE><<<148, warps * 32>>>(out, 1.0000001f, 0x8000000080000000ull, cyc);
    rate<MODE><<<148, warps * 32>>>(out, 1.0000001f, 0x8000000080000000ull, cyc);
    cudaDeviceSynchronize();
    long long c;
    cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    const double per_thread = (double)IT * 4 * ((MODE <= 1) ? 2 * CH : CH);  // FP instructions only
    const double warp_instr_per_smsp = per_thread * warps / 4.0;
    printf("%-28s warps/SM=%2d  cycles=%lld  warp-instr/clk/SMSP=%.3f  fp32 lane-ops/clk/SM=%.1f\n", name, warps, c, warp_instr_per_smsp / c,
           warp_instr_per_smsp / c * 4 * 32 * ((MODE <= 1 || MODE == 8) ? 1 : 2));
}
int main() {
    float* out; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
    for (int warps : {4, 8, 16}) {
        run<0>("FADD", warps, out, cyc);
        run<1>("FADD/FMUL", warps, out, cyc);
        run<2>("FADD2", warps, out, cyc);
        run<3>("FFMA2 (mul, -0 addend)", warps, out, cyc);
        run<4>("FADD2/FFMA2", warps, out, cyc);
        run<5>("FADD2 + 2 int ops each", warps, out, cyc);
        run<6>("FADD2 + LDS + IADD each", warps, out, cyc);
        run<7>("FADD2 + SHFL each", warps, out, cyc);
        run<8>("FADD + 2 int ops each", warps, out, cyc);
        run<9>("FADD2 + 1 LOP3 each", warps, out, cyc);
        run<10>("2 FADD2 + 1 SHFL", warps, out, cyc);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
