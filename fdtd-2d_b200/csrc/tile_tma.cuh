// Persistent TMA-fed register-resident tile kernel (fp32, plain tiles): the sm_100a flagship path.
//
// Same arithmetic and same tile grid as tile_fast.cuh (interior updates of python-src/main.py:69-74 and
// :21-27, k leapfrog steps per HBM round trip, results bit-identical), restructured around Blackwell's
// asynchronous copy engine:
//   * one persistent CTA per SM (grid = #SMs) walks the plain-tile list with stride gridDim.x;
//   * while the CTA advances tile t for k steps out of REGISTERS, the TMA unit
//     (cp.async.bulk.tensor.2d -> shared memory, completion on an mbarrier) prefetches the five
//     TH x 128 boxes (Ez, Hx, Hy, ce, ch) of tile t + gridDim.x into a 160 KB staging area, so DRAM
//     latency is never exposed to the compute warps and no LSU instructions are spent on loads;
//   * at the tile switch the threads pull their MR x 4 cells of all five arrays from the staging area
//     into registers (the coefficient maps live in registers too: the compute loop touches shared
//     memory only for the one-row exchanges between warps), release the stage, and an elected thread
//     re-arms the mbarrier and issues the next five TMA loads;
//   * results go straight from registers to HBM with 128-bit stores.
// Tile: TH = MR*NW rows x 128 columns; warp w owns rows [w*MR, (w+1)*MR), lane l columns [4l, 4l+4).
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace fdtd2d {

struct TmaMaps {
    CUtensorMap ez, hx, hy, ce, ch;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// (the bound is on TIME -- a spin count trips under a debugger, compute-sanitizer or heavy preemption -- and generous: a
// wait of 10 s is a protocol bug, and a trap is better than a GPU that has to be reset)
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    unsigned long long t0 = 0;
    for (;;) {
        asm volatile(
            "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (ok) return;
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        if (t0 == 0) t0 = t;
        if (t - t0 > 10000000000ull) __trap();
    }
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int x, int y, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(x), "r"(y), "r"(smem_u32(bar))
        : "memory");
}

// PAIR = true replaces the two CTA-wide barriers per leapfrog step by point-to-point mbarriers between
// neighbouring warps (warp w only ever needs one row from warp w+1 and one from warp w-1), with the
// exchanged rows double-buffered by step parity, so the 16 warps of the CTA drift apart and fill each
// other's pipeline bubbles instead of draining the SM twice per step.
template <int MR, int NW, bool PAIR>
__global__ void __launch_bounds__(NW * 32, 1)
    tile_tma_kernel(const __grid_constant__ TmaMaps maps, const PassParams<float> p, const int n_tiles) {
    constexpr int TW = 128, TH = MR * NW, N = TH * TW, XB = PAIR ? 2 : 1;
    constexpr uint32_t STAGE_BYTES = 5u * N * sizeof(float);
    extern __shared__ __align__(128) unsigned char smem_tma[];
    float* stage = reinterpret_cast<float*>(smem_tma);  // [5][TH][TW]: Ez, Hx, Hy, ce, ch of the NEXT tile
    float* sEz = stage + 5 * N;                         // [XB][NW][TW] first Ez row of every warp
    float* sHx = sEz + XB * NW * TW;                    // [XB][NW][TW] last Hx row of every warp
    __shared__ __align__(8) uint64_t full_bar;
    __shared__ __align__(8) uint64_t bar_e[NW], bar_h[NW];  // PAIR: "row of warp w is published"

    const int tid = threadIdx.x, w = tid >> 5, l = tid & 31;
    const int li0 = w * MR, lj = 4 * l;
    const int k = p.k;
    const int per_grid = p.tiles_y * p.tiles_x;

    auto issue = [&](int t) {  // elected thread: arm the barrier and launch the five box loads of tile t
        const int tile = p.tile_list[t];
        const int b = tile / per_grid, rem = tile - b * per_grid;
        const int ty = rem / p.tiles_x, tx = rem - ty * p.tiles_x;
        const int x = tx * p.CW - p.hx, y = b * p.Rl + p.org + ty * p.CH - k;
        mbar_expect_tx(&full_bar, STAGE_BYTES);
        tma_load_2d(stage + 0 * N, &maps.ez, x, y, &full_bar);
        tma_load_2d(stage + 1 * N, &maps.hx, x, y, &full_bar);
        tma_load_2d(stage + 2 * N, &maps.hy, x, y, &full_bar);
        tma_load_2d(stage + 3 * N, &maps.ce, x, y, &full_bar);
        tma_load_2d(stage + 4 * N, &maps.ch, x, y, &full_bar);
    };

    if (tid == 0) {
        mbar_init(&full_bar, 1);
        if (PAIR)
            for (int i = 0; i < NW; ++i) {
                mbar_init(&bar_e[i], 1);
                mbar_init(&bar_h[i], 1);
            }
        fence_mbar_init();
    }
    uint32_t g = 0;  // leapfrog steps done by this CTA so far (phase counter of the pair barriers)
    __syncthreads();
    int t = blockIdx.x;
    if (tid == 0 && t < n_tiles) issue(t);
    uint32_t parity = 0;
    const int wb = (w + 1 < NW ? w + 1 : NW - 1) * TW + lj;
    const int wa = (w > 0 ? w - 1 : 0) * TW + lj;

    for (; t < n_tiles; t += gridDim.x) {
        const int tile = p.tile_list[t];
        const int b = tile / per_grid, rem = tile - b * per_grid;
        const int ty = rem / p.tiles_x, tx = rem - ty * p.tiles_x;
        const int lr0 = p.org + ty * p.CH - k, lc0 = tx * p.CW - p.hx;

        // ---- tile switch: staging area -> registers ------------------------------------------
        mbar_wait(&full_bar, parity);
        parity ^= 1;
        float e[MR][4], hx[MR][4], hy[MR][4], ce[MR][4], ch[MR][4];
#pragma unroll
        for (int r = 0; r < MR; ++r) {
            const int so = (li0 + r) * TW + lj;
            const float4 a = *reinterpret_cast<const float4*>(stage + 0 * N + so);
            const float4 bx = *reinterpret_cast<const float4*>(stage + 1 * N + so);
            const float4 by = *reinterpret_cast<const float4*>(stage + 2 * N + so);
            const float4 c1 = *reinterpret_cast<const float4*>(stage + 3 * N + so);
            const float4 c2 = *reinterpret_cast<const float4*>(stage + 4 * N + so);
            e[r][0] = a.x, e[r][1] = a.y, e[r][2] = a.z, e[r][3] = a.w;
            hx[r][0] = bx.x, hx[r][1] = bx.y, hx[r][2] = bx.z, hx[r][3] = bx.w;
            hy[r][0] = by.x, hy[r][1] = by.y, hy[r][2] = by.z, hy[r][3] = by.w;
            ce[r][0] = c1.x, ce[r][1] = c1.y, ce[r][2] = c1.z, ce[r][3] = c1.w;
            ch[r][0] = c2.x, ch[r][1] = c2.y, ch[r][2] = c2.z, ch[r][3] = c2.w;
        }
        __syncthreads();  // every thread has drained the stage
        if (tid == 0 && t + (int)gridDim.x < n_tiles) {
            fence_proxy_async();  // order the generic-proxy reads above before the async-proxy writes
            issue(t + gridDim.x);
        }

        // ---- k leapfrog steps out of registers ---------------------------------------------
        for (int s = 0; s < k; ++s, ++g) {
            const int xo = PAIR ? (int)(g & 1u) * NW * TW : 0;
            *reinterpret_cast<float4*>(sEz + xo + w * TW + lj) = make_float4(e[0][0], e[0][1], e[0][2], e[0][3]);
            if (PAIR) {
                __syncwarp();
                if (l == 0) mbar_arrive(&bar_e[w]);
                if (w + 1 < NW) mbar_wait(&bar_e[w + 1], g & 1u);
            } else {
                __syncthreads();
            }
            const float4 eb = *reinterpret_cast<const float4*>(sEz + xo + wb);
            const float below[4] = {eb.x, eb.y, eb.z, eb.w};
#pragma unroll
            for (int r = 0; r < MR; ++r) {  // H half-step, main.py:69-74
                const float right3 = __shfl_down_sync(0xffffffffu, e[r][0], 1);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float down = (r + 1 < MR) ? e[r + 1 < MR ? r + 1 : r][q] : below[q];
                    const float right = (q < 3) ? e[r][q < 3 ? q + 1 : 3] : right3;
                    hx[r][q] = sub_rn(hx[r][q], mul_rn(ch[r][q], sub_rn(down, e[r][q])));
                    hy[r][q] = add_rn(hy[r][q], mul_rn(ch[r][q], sub_rn(right, e[r][q])));
                }
            }
            *reinterpret_cast<float4*>(sHx + xo + w * TW + lj) =
                make_float4(hx[MR - 1][0], hx[MR - 1][1], hx[MR - 1][2], hx[MR - 1][3]);
            if (PAIR) {
                __syncwarp();
                if (l == 0) mbar_arrive(&bar_h[w]);
                if (w > 0) mbar_wait(&bar_h[w - 1], g & 1u);
            } else {
                __syncthreads();
            }
            const float4 ha = *reinterpret_cast<const float4*>(sHx + xo + wa);
            const float above[4] = {ha.x, ha.y, ha.z, ha.w};
#pragma unroll
            for (int r = 0; r < MR; ++r) {  // Ez update, main.py:21-27
                const float left0 = __shfl_up_sync(0xffffffffu, hy[r][3], 1);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float up = (r > 0) ? hx[r > 0 ? r - 1 : 0][q] : above[q];
                    const float left = (q > 0) ? hy[r][q > 0 ? q - 1 : 0] : left0;
                    const float curl = sub_rn(sub_rn(hy[r][q], left), sub_rn(hx[r][q], up));
                    e[r][q] = add_rn(e[r][q], mul_rn(curl, ce[r][q]));
                }
            }
        }

        // ---- store the core straight from registers -------------------------------------------
        if (lj >= p.hx && lj < p.hx + p.CW) {
            const long long base = (long long)b * p.grid_stride + (long long)(lr0 + li0) * p.pitch + (lc0 + lj);
#pragma unroll
            for (int r = 0; r < MR; ++r) {
                const int li = li0 + r;
                if (li >= k && li < k + p.CH) {
                    const long long o = base + (long long)r * p.pitch;
                    *reinterpret_cast<float4*>(p.out[0] + o) = make_float4(e[r][0], e[r][1], e[r][2], e[r][3]);
                    *reinterpret_cast<float4*>(p.out[1] + o) = make_float4(hx[r][0], hx[r][1], hx[r][2], hx[r][3]);
                    *reinterpret_cast<float4*>(p.out[2] + o) = make_float4(hy[r][0], hy[r][1], hy[r][2], hy[r][3]);
                }
            }
        }
    }
}

}  // namespace fdtd2d
