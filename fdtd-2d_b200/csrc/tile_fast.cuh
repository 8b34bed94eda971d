// Register-resident fast tile kernel (fp32): k leapfrog steps per HBM round trip for PLAIN tiles.
//
// A plain tile is one whose haloed extent contains no Mur-ring cell, no source cell, no probe (in its
// core) and lies entirely inside the local array, so every cell takes the unconditional interior
// updates of python-src/main.py:69-74 (H) and :21-27 (Ez); the host (api.cu: classify_tiles) routes every
// other tile to tile_generic_kernel.  Same tile grid, same arithmetic, same results bit for bit.
//
// Layout of the work: the tile is TH = MR*NW rows by TW = 128 columns.  Warp w owns rows
// [w*MR, (w+1)*MR); lane l owns columns [4l, 4l+4).  Each thread keeps its MR x 4 cells of Ez, Hx, Hy in
// REGISTERS for all k steps (HBM is touched once per field per pass, 128-bit coalesced).  Neighbours:
//   column j+1 / j-1 across the lane boundary -> warp shuffles,
//   row i+1 (Ez) / i-1 (Hx) across the warp boundary -> one 512-byte row per warp through shared memory.
// The coefficient maps are parked in shared memory (each thread only ever reads the slots it wrote, so
// they need no barrier) and re-read with one LDS.128 per 4 cells per half-step.
// Per cell-update: 11 FP ops + ~1.3 other instructions (vs ~88 in the generic kernel).
// Cells next to a tile edge read clamped/garbage neighbours; they go stale one cell per step and are
// never stored (halo = k rows, hx = round_up(k, 4) columns).
#pragma once
#include "common.cuh"

namespace fdtd2d {

constexpr int FAST_TW = 128;

template <int MR, int NW, int MINB>
__global__ void __launch_bounds__(NW * 32, MINB) tile_fast_kernel(const PassParams<float> p) {
    constexpr int TW = FAST_TW, TH = MR * NW;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* sCe = reinterpret_cast<float*>(smem_raw);  // [TH][TW]
    float* sCh = sCe + TH * TW;                       // [TH][TW]
    float* sEz = sCh + TH * TW;                       // [NW][TW] first Ez row of every warp
    float* sHx = sEz + NW * TW;                       // [NW][TW] last Hx row of every warp

    const int tile = p.tile_list[blockIdx.x];
    const int per_grid = p.tiles_y * p.tiles_x;
    const int b = tile / per_grid;
    const int rem = tile - b * per_grid;
    const int ty = rem / p.tiles_x;
    const int tx = rem - ty * p.tiles_x;
    const int k = p.k;
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    const int li0 = w * MR, lj = 4 * l;
    const int lr0 = ty * p.CH - k, lc0 = tx * p.CW - p.hx;
    const long long base = (long long)b * p.grid_stride + (long long)(lr0 + li0) * p.pitch + (lc0 + lj);

    float e[MR][4], hx[MR][4], hy[MR][4];
    {
        float4 ve[MR], vx[MR], vy[MR], vce[MR], vch[MR];
#pragma unroll
        for (int r = 0; r < MR; ++r) {
            const long long o = base + (long long)r * p.pitch;
            ve[r] = __ldg(reinterpret_cast<const float4*>(p.in[0] + o));
            vx[r] = __ldg(reinterpret_cast<const float4*>(p.in[1] + o));
            vy[r] = __ldg(reinterpret_cast<const float4*>(p.in[2] + o));
            vce[r] = __ldg(reinterpret_cast<const float4*>(p.ce + o));
            vch[r] = __ldg(reinterpret_cast<const float4*>(p.ch + o));
        }
#pragma unroll
        for (int r = 0; r < MR; ++r) {
            e[r][0] = ve[r].x, e[r][1] = ve[r].y, e[r][2] = ve[r].z, e[r][3] = ve[r].w;
            hx[r][0] = vx[r].x, hx[r][1] = vx[r].y, hx[r][2] = vx[r].z, hx[r][3] = vx[r].w;
            hy[r][0] = vy[r].x, hy[r][1] = vy[r].y, hy[r][2] = vy[r].z, hy[r][3] = vy[r].w;
            *reinterpret_cast<float4*>(sCe + (li0 + r) * TW + lj) = vce[r];
            *reinterpret_cast<float4*>(sCh + (li0 + r) * TW + lj) = vch[r];
        }
    }
    const int wb = (w + 1 < NW ? w + 1 : NW - 1) * TW + lj;  // warp below (clamped: garbage at the tile edge)
    const int wa = (w > 0 ? w - 1 : 0) * TW + lj;            // warp above

    for (int s = 0; s < k; ++s) {
        // ---- H half-step (main.py:69-74) ----------------------------------------------------
        *reinterpret_cast<float4*>(sEz + w * TW + lj) = make_float4(e[0][0], e[0][1], e[0][2], e[0][3]);
        __syncthreads();
        const float4 eb = *reinterpret_cast<const float4*>(sEz + wb);
        const float below[4] = {eb.x, eb.y, eb.z, eb.w};
#pragma unroll
        for (int r = 0; r < MR; ++r) {
            const float4 c4 = *reinterpret_cast<const float4*>(sCh + (li0 + r) * TW + lj);
            const float c[4] = {c4.x, c4.y, c4.z, c4.w};
            const float right3 = __shfl_down_sync(0xffffffffu, e[r][0], 1);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float down = (r + 1 < MR) ? e[r + 1][q] : below[q];
                const float right = (q < 3) ? e[r][q + 1 < 4 ? q + 1 : 3] : right3;
                hx[r][q] = sub_rn(hx[r][q], mul_rn(c[q], sub_rn(down, e[r][q])));
                hy[r][q] = add_rn(hy[r][q], mul_rn(c[q], sub_rn(right, e[r][q])));
            }
        }
        // ---- Ez update (main.py:21-27) ------------------------------------------------------
        *reinterpret_cast<float4*>(sHx + w * TW + lj) =
            make_float4(hx[MR - 1][0], hx[MR - 1][1], hx[MR - 1][2], hx[MR - 1][3]);
        __syncthreads();
        const float4 ha = *reinterpret_cast<const float4*>(sHx + wa);
        const float above[4] = {ha.x, ha.y, ha.z, ha.w};
#pragma unroll
        for (int r = 0; r < MR; ++r) {
            const float4 c4 = *reinterpret_cast<const float4*>(sCe + (li0 + r) * TW + lj);
            const float c[4] = {c4.x, c4.y, c4.z, c4.w};
            const float left0 = __shfl_up_sync(0xffffffffu, hy[r][3], 1);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float up = (r > 0) ? hx[r > 0 ? r - 1 : 0][q] : above[q];
                const float left = (q > 0) ? hy[r][q > 0 ? q - 1 : 0] : left0;
                const float curl = sub_rn(sub_rn(hy[r][q], left), sub_rn(hx[r][q], up));
                e[r][q] = add_rn(e[r][q], mul_rn(curl, c[q]));
            }
        }
    }

    // ---- store the core -------------------------------------------------------------------
    if (lj >= p.hx && lj < p.hx + p.CW) {
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

}  // namespace fdtd2d
