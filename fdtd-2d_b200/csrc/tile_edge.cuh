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

// four consecutive cells of one row <-> registers (one 128-bit access in fp32, two in fp64)
__device__ __forceinline__ void load4(const float* p, float* a) { unpack4(*reinterpret_cast<const float4*>(p), a); }
__device__ __forceinline__ void load4(const double* p, double* a) {
    const double2 u = *reinterpret_cast<const double2*>(p), v = *reinterpret_cast<const double2*>(p + 2);
    a[0] = u.x, a[1] = u.y, a[2] = v.x, a[3] = v.y;
}
__device__ __forceinline__ void ldg4(const float* p, float* a) { unpack4(__ldg(reinterpret_cast<const float4*>(p)), a); }
__device__ __forceinline__ void ldg4(const double* p, double* a) {
    const double2 u = __ldg(reinterpret_cast<const double2*>(p)), v = __ldg(reinterpret_cast<const double2*>(p + 2));
    a[0] = u.x, a[1] = u.y, a[2] = v.x, a[3] = v.y;
}
// (L2 only: a slab's ghost rows are written by the neighbour GPU while this kernel may already be running)
__device__ __forceinline__ void ldcg4(const float* p, float* a) { unpack4(__ldcg(reinterpret_cast<const float4*>(p)), a); }
__device__ __forceinline__ void ldcg4(const double* p, double* a) {
    const double2 u = __ldcg(reinterpret_cast<const double2*>(p)), v = __ldcg(reinterpret_cast<const double2*>(p + 2));
    a[0] = u.x, a[1] = u.y, a[2] = v.x, a[3] = v.y;
}
__device__ __forceinline__ void store4(float* p, const float* a) {
    *reinterpret_cast<float4*>(p) = make_float4(a[0], a[1], a[2], a[3]);
}
__device__ __forceinline__ void store4(double* p, const double* a) {
    *reinterpret_cast<double2*>(p) = make_double2(a[0], a[1]);
    *reinterpret_cast<double2*>(p + 2) = make_double2(a[2], a[3]);
}

template <typename T, int MR, int NW>
__global__ void __launch_bounds__(NW * 32, 1) tile_edge_kernel(const PassParams<T> p) {
    constexpr int TW = 128, TH = MR * NW, N = TH * TW, NT = NW * 32;
    extern __shared__ __align__(16) unsigned char smem_edge[];
    T* s0 = reinterpret_cast<T*>(smem_edge);  // [TH][TW] Ez before the step
    T* s1 = s0 + N;                           // [TH][TW] Ez after the interior update
    T* sEz = s1 + N;                          // [NW][TW] first Ez row of every warp
    T* sHx = sEz + NW * TW;                   // [NW][TW] last Hx row of every warp

    const int tid = threadIdx.x, w = tid >> 5, l = tid & 31;
    const int tile = p.tile_list ? p.tile_list[blockIdx.x] : (int)blockIdx.x;
    const int per_grid = p.tiles_y * p.tiles_x;
    const int b = tile / per_grid, rem = tile - b * per_grid;
    const int ty = rem / p.tiles_x, tx = rem - ty * p.tiles_x;
    const int k = p.k;
    const int li0 = w * MR, lj = 4 * l;
    const int lr0 = p.org + ty * p.CH - k, lc0 = tx * p.CW - p.hx;
    const int gr0 = lr0 + p.row0;
    const int Rg = p.Rg, C = p.C;
    const long long base = (long long)b * p.grid_stride + (long long)(lr0 + li0) * p.pitch + (lc0 + lj);
    // y-slabs: a tile whose core holds band rows reads the ghost rows of that side -- wait for the neighbour -- and
    // mirrors those rows into the neighbour's ghost rows when it stores (common.cuh)
    const int core_lo = lr0 + k, core_hi = min(core_lo + p.CH, p.store_hi);
    const bool band0 = core_lo < p.band_hi[0] && core_hi > p.band_lo[0];
    const bool band1 = core_lo < p.band_hi[1] && core_hi > p.band_lo[1];
    if (band0) band_wait(p, 0);
    if (band1) band_wait(p, 1);

    // ---- load (zero outside the local array) ------------------------------------------------
    T e[MR][4], hx[MR][4], hy[MR][4], ce[MR][4], ch[MR][4];
    const bool col_in = (lc0 + lj >= 0) && (lc0 + lj < p.pitch);
#pragma unroll
    for (int r = 0; r < MR; ++r) {
        const int row = lr0 + li0 + r;
#pragma unroll
        for (int q = 0; q < 4; ++q) e[r][q] = hx[r][q] = hy[r][q] = ce[r][q] = ch[r][q] = (T)0;
        if (col_in && row >= 0 && row < p.Rl) {
            const long long o = base + (long long)r * p.pitch;
            ldcg4(p.in[0] + o, e[r]);
            ldcg4(p.in[1] + o, hx[r]);
            ldcg4(p.in[2] + o, hy[r]);
            ldg4(p.ce + o, ce[r]);
            ldg4(p.ch + o, ch[r]);
        }
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

    TileCtx<T> tc;
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
    // does any source / probe of this grid fall inside this tile, and in which warps' rows?
    __shared__ unsigned cell_warps;
    if (tid == 0) cell_warps = 0u;
    __syncthreads();
    for (int q = tc.src_lo + tid; q < tc.src_hi; q += NT) {
        const Cell c = p.src[q];
        if (c.row >= gr0 && c.row < gr0 + TH && c.col >= lc0 && c.col < lc0 + TW) atomicOr(&cell_warps, 1u << ((c.row - gr0) / MR));
    }
    for (int q = tc.prb_lo + tid; q < tc.prb_hi; q += NT) {
        const Cell c = p.probes[q];
        if (c.row >= gr0 && c.row < gr0 + TH && c.col >= lc0 && c.col < lc0 + TW) atomicOr(&cell_warps, 1u << ((c.row - gr0) / MR));
    }
    __syncthreads();
    // Sources / probes in the tile: the warps that hold their rows park the field after the interior update (s1), the
    // source is added and the probes are read there.  Top / bottom ring: the warps that hold ring rows (global rows 0..5,
    // Rg-6..Rg-1) park the field before (s0) and after (s1) the update for S3 / S4.  Left / right ring: nothing is parked,
    // S2 runs in registers below.
    // (fp64 keeps S2 in shared memory with every row parked: its kernel sits at the 128-register limit of a 512-thread
    // CTA and the register form cost it a third of its speed)
    constexpr bool S2REG = sizeof(T) == 4;
    const bool full = cell_warps != 0u;
    const bool tb = tc.touchT || tc.touchB;
    const bool ring_smem = S2REG ? tb : (tb || tc.touchL || tc.touchR);  // ring stages that run in shared memory
    const bool staged = full || ring_smem;
    const int gw0 = gr0 + li0, gw1 = gw0 + MR - 1;  // global rows of this warp
    const bool park0 = S2REG ? (tb && ((gw0 <= RING && gw1 >= 0) || (gw1 >= Rg - 1 - RING && gw0 <= Rg - 1))) : ring_smem;
    const bool park = park0 || ((cell_warps >> w) & 1u);
    const bool lr = S2REG && (tc.touchL || tc.touchR);
    // Mur left / right (main.py:33-41) on registers: in the reference's order every column reads its inward neighbour
    // before that one is overwritten, so for rows 1..Rg-2
    //   Ez[i, q]     = S0[i, q+1]   + coef * (S1[i, q+1]   - S0[i, q])      q = 0..4
    //   Ez[i, C-1-q] = S0[i, C-2-q] + coef * (S1[i, C-2-q] - S0[i, C-1-q])
    // with S0 the field before the step and S1 the field after the interior update: both are in this thread or one
    // lane away.  (Left and right do not interact for C >= 11, the smallest grid the library accepts.)
    bool lcol[4], rcol[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int gj = lc0 + lj + q;
        lcol[q] = gj >= 0 && gj < RING;
        rcol[q] = gj >= C - RING && gj <= C - 1;
    }
    const T coef = tc.coef;

    const int wb = (w + 1 < NW ? w + 1 : NW - 1) * TW + lj;
    const int wa = (w > 0 ? w - 1 : 0) * TW + lj;

    for (int s = 0; s < k; ++s) {
        // ---- H half-step ----------------------------------------------------------------------
        store4(sEz + w * TW + lj, e[0]);
        __syncthreads();
        T below[4];
        load4(sEz + wb, below);
#pragma unroll
        for (int r = 0; r < MR; ++r) {
            const T right3 = __shfl_down_sync(0xffffffffu, e[r][0], 1);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const T down = (r + 1 < MR) ? e[r + 1 < MR ? r + 1 : r][q] : below[q];
                const T right = (q < 3) ? e[r][q < 3 ? q + 1 : 3] : right3;
                const T nx = sub_rn(hx[r][q], mul_rn(ch[r][q], sub_rn(down, e[r][q])));
                const T ny = add_rn(hy[r][q], mul_rn(ch[r][q], sub_rn(right, e[r][q])));
                const bool ok = hrow[r] && hcol[q];
                hx[r][q] = ok ? nx : hx[r][q];
                hy[r][q] = ok ? ny : hy[r][q];
            }
        }
        // ---- interior Ez update (S1), Mur left/right (S2) -------------------------------------------
        store4(sHx + w * TW + lj, hx[MR - 1]);
        if (park0) {
#pragma unroll
            for (int r = 0; r < MR; ++r) store4(s0 + (li0 + r) * TW + lj, e[r]);
        }
        __syncthreads();
        T above[4];
        load4(sHx + wa, above);
#pragma unroll
        for (int r = 0; r < MR; ++r) {
            const T left0 = __shfl_up_sync(0xffffffffu, hy[r][3], 1);
            T nv[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const T up = (r > 0) ? hx[r > 0 ? r - 1 : 0][q] : above[q];
                const T left = (q > 0) ? hy[r][q > 0 ? q - 1 : 0] : left0;
                const T curl = sub_rn(sub_rn(hy[r][q], left), sub_rn(hx[r][q], up));
                const T v = add_rn(e[r][q], mul_rn(curl, ce[r][q]));
                nv[q] = (erow[r] && ecol[q]) ? v : e[r][q];
            }
            if (lr) {
                const T s0r = __shfl_down_sync(0xffffffffu, e[r][0], 1), s1r = __shfl_down_sync(0xffffffffu, nv[0], 1);
                const T s0l = __shfl_up_sync(0xffffffffu, e[r][3], 1), s1l = __shfl_up_sync(0xffffffffu, nv[3], 1);
                T out[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const T a0 = (q < 3) ? e[r][q < 3 ? q + 1 : 3] : s0r, a1 = (q < 3) ? nv[q < 3 ? q + 1 : 3] : s1r;
                    const T b0 = (q > 0) ? e[r][q > 0 ? q - 1 : 0] : s0l, b1 = (q > 0) ? nv[q > 0 ? q - 1 : 0] : s1l;
                    const T ml = add_rn(a0, mul_rn(coef, sub_rn(a1, e[r][q])));
                    const T mr = add_rn(b0, mul_rn(coef, sub_rn(b1, e[r][q])));
                    out[q] = (erow[r] && lcol[q]) ? ml : ((erow[r] && rcol[q]) ? mr : nv[q]);
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) e[r][q] = out[q];
            } else {
#pragma unroll
                for (int q = 0; q < 4; ++q) e[r][q] = nv[q];
            }
        }
        // ---- the other boundary stages through shared memory (S3, S4, source, probes) ----------------
        if (staged) {
            if (park) {
#pragma unroll
                for (int r = 0; r < MR; ++r) store4(s1 + (li0 + r) * TW + lj, e[r]);
            }
            __syncthreads();
            if (ring_smem) ring_stages<T, TH, TW, NT, S2REG>(s0, s1, tc, tid);
            if (full) source_and_probes<T, TH, TW, NT>(s1, p, tc, p.step0 + s, tid);
            if (park) {
#pragma unroll
                for (int r = 0; r < MR; ++r) load4(s1 + (li0 + r) * TW + lj, e[r]);
            }
        }
    }

    // ---- store the core (owned rows only; band rows also go to the neighbour slab) --------------------------
    if (lj >= p.hx && lj < p.hx + p.CW && lc0 + lj < p.pitch) {
#pragma unroll
        for (int r = 0; r < MR; ++r) {
            const int li = li0 + r, row = lr0 + li;
            if (li >= k && li < k + p.CH && row < p.store_hi) {
                const long long o = base + (long long)r * p.pitch;
                store4(p.out[0] + o, e[r]);
                store4(p.out[1] + o, hx[r]);
                store4(p.out[2] + o, hy[r]);
                const int bs = band_of_row(p, row);
                if (bs >= 0) {
                    T* const q0 = peer_field(p, bs, 0);
                    if (q0) {
                        const long long po = o + peer_shift(p, bs);
                        store4(q0 + po, e[r]);
                        store4(peer_field(p, bs, 1) + po, hx[r]);
                        store4(peer_field(p, bs, 2) + po, hy[r]);
                    }
                }
            }
        }
    }
    if (band0 || band1) {
        __threadfence_system();
        __syncthreads();
        if (tid == 0) {
            if (band0) band_done(p, 0);
            if (band1) band_done(p, 1);
        }
    }
}

}  // namespace fdtd2d
