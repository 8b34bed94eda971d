// Register-resident tile kernel for the tiles the plain kernels cannot take (fp32): tiles that reach the
// Mur ring, the array edge or a slab's ghost rows, or that hold a source cell or a probe.
//
// The bulk of such a tile is still ordinary interior cells, so the H half-step (main.py:69-74) and the
// interior Ez update (main.py:21-27) run exactly as in tile_fast.cuh -- MR x 4 cells per thread in
// registers, shuffles across lanes, one row per warp through shared memory -- with per-row / per-column
// masks that keep cells outside the reference's index ranges unchanged (rows 0..R-2 / cols 0..C-2 for H,
// 1..R-2 / 1..C-2 for Ez).  Only when the tile actually touches a ring, a source or a probe does a step
// take the detour the reference's boundary code needs: the pre-step field (S0) and the post-interior
// field (S1) are parked in shared memory, the shared ring stages (ring_ops.cuh: Mur left/right,
// top/bottom, corner means, source add, probes) run on them, and the threads pull the finished field
// back into registers.  Bit-identical to tile_generic_kernel, ~3-4x faster per tile.
#pragma once
#include "common.cuh"
#include "ring_ops.cuh"

namespace fdtd2d {

template <int MR, int NW>
__global__ void __launch_bounds__(NW * 32, 1) tile_edge_kernel(const PassParams<float> p) {
    constexpr int TW = 128, TH = MR * NW, N = TH * TW, NT = NW * 32;
    extern __shared__ __align__(16) unsigned char smem_edge[];
    float* s0 = reinterpret_cast<float*>(smem_edge);  // [TH][TW] Ez before the step
    float* s1 = s0 + N;                               // [TH][TW] Ez after the interior update
    float* sEz = s1 + N;                              // [NW][TW] first Ez row of every warp
    float* sHx = sEz + NW * TW;                       // [NW][TW] last Hx row of every warp

    const int tid = threadIdx.x, w = tid >> 5, l = tid & 31;
    const int tile = p.tile_list ? p.tile_list[blockIdx.x] : (int)blockIdx.x;
    const int per_grid = p.tiles_y * p.tiles_x;
    const int b = tile / per_grid, rem = tile - b * per_grid;
    const int ty = rem / p.tiles_x, tx = rem - ty * p.tiles_x;
    const int k = p.k;
    const int li0 = w * MR, lj = 4 * l;
    const int lr0 = ty * p.CH - k, lc0 = tx * p.CW - p.hx;
    const int gr0 = lr0 + p.row0;
    const int Rg = p.Rg, C = p.C;
    const long long base = (long long)b * p.grid_stride + (long long)(lr0 + li0) * p.pitch + (lc0 + lj);

    // ---- load (zero outside the local array) ------------------------------------------------
    float e[MR][4], hx[MR][4], hy[MR][4], ce[MR][4], ch[MR][4];
    const bool col_in = (lc0 + lj >= 0) && (lc0 + lj < p.pitch);
#pragma unroll
    for (int r = 0; r < MR; ++r) {
        const int row = lr0 + li0 + r;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f), bx = a, by = a, c1 = a, c2 = a;
        if (col_in && row >= 0 && row < p.Rl) {
            const long long o = base + (long long)r * p.pitch;
            a = __ldg(reinterpret_cast<const float4*>(p.in[0] + o));
            bx = __ldg(reinterpret_cast<const float4*>(p.in[1] + o));
            by = __ldg(reinterpret_cast<const float4*>(p.in[2] + o));
            c1 = __ldg(reinterpret_cast<const float4*>(p.ce + o));
            c2 = __ldg(reinterpret_cast<const float4*>(p.ch + o));
        }
        e[r][0] = a.x, e[r][1] = a.y, e[r][2] = a.z, e[r][3] = a.w;
        hx[r][0] = bx.x, hx[r][1] = bx.y, hx[r][2] = bx.z, hx[r][3] = bx.w;
        hy[r][0] = by.x, hy[r][1] = by.y, hy[r][2] = by.z, hy[r][3] = by.w;
        ce[r][0] = c1.x, ce[r][1] = c1.y, ce[r][2] = c1.z, ce[r][3] = c1.w;
        ch[r][0] = c2.x, ch[r][1] = c2.y, ch[r][2] = c2.z, ch[r][3] = c2.w;
    }
    // index-range masks of the reference's slices (main.py:70,74 and :27)
    bool hrow[MR], erow[MR], hcol[4], ecol[4];
#pragma unroll
    for (int r = 0; r < MR; ++r) {
        const int gi = gr0 + li0 + r;
        hrow[r] = gi >= 0 && gi <= Rg - 2;
        erow[r] = gi >= 1 && gi <= Rg - 2;
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int gj = lc0 + lj + q;
        hcol[q] = gj >= 0 && gj <= C - 2;
        ecol[q] = gj >= 1 && gj <= C - 2;
    }

    TileCtx<float> tc;
    tc.gr0 = gr0, tc.lc0 = lc0, tc.Rg = Rg, tc.C = C, tc.k = k;
    tc.coef = p.mur[b];
    tc.touchL = lc0 < RING;
    tc.touchR = lc0 + TW > C - RING;
    tc.touchT = gr0 < RING;
    tc.touchB = gr0 + TH > Rg - RING;
    tc.src_lo = p.src_range ? p.src_range[b] : 0;
    tc.src_hi = p.src_range ? p.src_range[b + 1] : 0;
    tc.prb_lo = p.probe_range ? p.probe_range[b] : 0;
    tc.prb_hi = p.probe_range ? p.probe_range[b + 1] : 0;
    // does any source / probe of this grid fall inside this tile?
    int mine = 0;
    for (int q = tc.src_lo + tid; q < tc.src_hi; q += NT) {
        const Cell c = p.src[q];
        mine |= (c.row >= gr0 && c.row < gr0 + TH && c.col >= lc0 && c.col < lc0 + TW);
    }
    for (int q = tc.prb_lo + tid; q < tc.prb_hi; q += NT) {
        const Cell c = p.probes[q];
        mine |= (c.row >= gr0 && c.row < gr0 + TH && c.col >= lc0 && c.col < lc0 + TW);
    }
    const bool staged = __syncthreads_or(mine) || tc.touchL || tc.touchR || tc.touchT || tc.touchB;

    const int wb = (w + 1 < NW ? w + 1 : NW - 1) * TW + lj;
    const int wa = (w > 0 ? w - 1 : 0) * TW + lj;

    for (int s = 0; s < k; ++s) {
        // ---- H half-step ----------------------------------------------------------------------
        *reinterpret_cast<float4*>(sEz + w * TW + lj) = make_float4(e[0][0], e[0][1], e[0][2], e[0][3]);
        __syncthreads();
        const float4 eb = *reinterpret_cast<const float4*>(sEz + wb);
        const float below[4] = {eb.x, eb.y, eb.z, eb.w};
#pragma unroll
        for (int r = 0; r < MR; ++r) {
            const float right3 = __shfl_down_sync(0xffffffffu, e[r][0], 1);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float down = (r + 1 < MR) ? e[r + 1 < MR ? r + 1 : r][q] : below[q];
                const float right = (q < 3) ? e[r][q < 3 ? q + 1 : 3] : right3;
                const float nx = sub_rn(hx[r][q], mul_rn(ch[r][q], sub_rn(down, e[r][q])));
                const float ny = add_rn(hy[r][q], mul_rn(ch[r][q], sub_rn(right, e[r][q])));
                const bool ok = hrow[r] && hcol[q];
                hx[r][q] = ok ? nx : hx[r][q];
                hy[r][q] = ok ? ny : hy[r][q];
            }
        }
        // ---- interior Ez update (S1) ----------------------------------------------------------
        *reinterpret_cast<float4*>(sHx + w * TW + lj) =
            make_float4(hx[MR - 1][0], hx[MR - 1][1], hx[MR - 1][2], hx[MR - 1][3]);
        if (staged) {
#pragma unroll
            for (int r = 0; r < MR; ++r)
                *reinterpret_cast<float4*>(s0 + (li0 + r) * TW + lj) = make_float4(e[r][0], e[r][1], e[r][2], e[r][3]);
        }
        __syncthreads();
        const float4 ha = *reinterpret_cast<const float4*>(sHx + wa);
        const float above[4] = {ha.x, ha.y, ha.z, ha.w};
#pragma unroll
        for (int r = 0; r < MR; ++r) {
            const float left0 = __shfl_up_sync(0xffffffffu, hy[r][3], 1);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float up = (r > 0) ? hx[r > 0 ? r - 1 : 0][q] : above[q];
                const float left = (q > 0) ? hy[r][q > 0 ? q - 1 : 0] : left0;
                const float curl = sub_rn(sub_rn(hy[r][q], left), sub_rn(hx[r][q], up));
                const float nv = add_rn(e[r][q], mul_rn(curl, ce[r][q]));
                e[r][q] = (erow[r] && ecol[q]) ? nv : e[r][q];
            }
        }
        // ---- boundary stages through shared memory (S2, S3, S4, source, probes) -------------------
        if (staged) {
#pragma unroll
            for (int r = 0; r < MR; ++r)
                *reinterpret_cast<float4*>(s1 + (li0 + r) * TW + lj) = make_float4(e[r][0], e[r][1], e[r][2], e[r][3]);
            __syncthreads();
            ring_stages<float, TH, TW, NT>(s0, s1, tc, tid);
            source_and_probes<float, TH, TW, NT>(s1, p, tc, p.step0 + s, tid);
#pragma unroll
            for (int r = 0; r < MR; ++r) {
                const float4 v = *reinterpret_cast<const float4*>(s1 + (li0 + r) * TW + lj);
                e[r][0] = v.x, e[r][1] = v.y, e[r][2] = v.z, e[r][3] = v.w;
            }
        }
    }

    // ---- store the core (inside the local array) ----------------------------------------------
    if (lj >= p.hx && lj < p.hx + p.CW && lc0 + lj < p.pitch) {
#pragma unroll
        for (int r = 0; r < MR; ++r) {
            const int li = li0 + r;
            if (li >= k && li < k + p.CH && lr0 + li < p.Rl) {
                const long long o = base + (long long)r * p.pitch;
                *reinterpret_cast<float4*>(p.out[0] + o) = make_float4(e[r][0], e[r][1], e[r][2], e[r][3]);
                *reinterpret_cast<float4*>(p.out[1] + o) = make_float4(hx[r][0], hx[r][1], hx[r][2], hx[r][3]);
                *reinterpret_cast<float4*>(p.out[2] + o) = make_float4(hy[r][0], hy[r][1], hy[r][2], hy[r][3]);
            }
        }
    }
}

}  // namespace fdtd2d
