// Boundary stages shared by the tile kernels: they operate on a TH x TW tile of Ez held in shared memory
// as two buffers, s0 = Ez at the start of the step (the reference's Ez_prev, main.py:18) and s1 = Ez after
// the interior update (S1), and turn s1 into the end-of-step field:
//   S2  Mur left/right   python-src/main.py:33-41
//   S3  Mur top/bottom   python-src/main.py:43-51
//   S4  corner means     python-src/main.py:54-61
//   source add           python-src/fdtd.py:34 (+ main.py:182-195), probes sampled after it.
// Every function is called by all threads of the CTA (they contain __syncthreads()).
#pragma once
#include "common.cuh"

namespace fdtd2d {

template <typename T> struct TileCtx {
    int gr0, lc0;  // global row / column held in shared row 0 / column 0
    int Rg, C, k;
    T coef;        // Mur coefficient of this grid
    bool touchL, touchR, touchT, touchB;  // the tile (halo included) reaches that ring
    int src_lo, src_hi, prb_lo, prb_hi;   // this grid's ranges in the sorted source / probe lists
};

// S2_DONE: the caller has already applied S2 to nxt (tile_edge_kernel does it in registers).
template <typename T, int TH, int TW, int NT, bool S2_DONE = false>
__device__ __forceinline__ void ring_stages(const T* cur, T* nxt, const TileCtx<T>& tc, const int tid) {
    const int gr0 = tc.gr0, lc0 = tc.lc0, Rg = tc.Rg, C = tc.C;
    const T coef = tc.coef;
    const bool touchL = tc.touchL, touchR = tc.touchR, touchT = tc.touchT, touchB = tc.touchB;
                // ---- S2: Mur left/right, main.py:33-41. One thread per (row, side) runs the
                // reference's five column updates in the reference's order (outermost first), so every
                // read of the inward neighbour sees the value S1 left there. -----------------------
                if (!S2_DONE && (touchL || touchR)) {
                    for (int w = tid; w < 2 * TH; w += NT) {
                        const int side = w / TH, li = w - side * TH;
                        const int gi = gr0 + li;
                        if (gi < 1 || gi > Rg - 2) continue;
                        T* n1 = nxt + li * TW;
                        const T* s0 = cur + li * TW;
    #pragma unroll
                        for (int q = 0; q < RING; ++q) {
                            const int gj = side ? C - 1 - q : q;
                            const int lj = gj - lc0;
                            const int ln = side ? lj - 1 : lj + 1;
                            if (lj < 0 || lj >= TW || ln < 0 || ln >= TW) continue;
                            n1[lj] = add_rn(s0[ln], mul_rn(coef, sub_rn(n1[ln], s0[lj])));
                        }
                    }
                    __syncthreads();
                }
                // ---- S3: Mur top/bottom, main.py:43-51. One thread per (column, side). -------------
                if (touchT || touchB) {
                    for (int w = tid; w < 2 * TW; w += NT) {
                        const int side = w / TW, lj = w - side * TW;
                        const int gj = lc0 + lj;
                        if (gj < 1 || gj > C - 2) continue;
    #pragma unroll
                        for (int q = 0; q < RING; ++q) {
                            const int gi = side ? Rg - 1 - q : q;
                            const int li = gi - gr0;
                            const int ln = side ? li - 1 : li + 1;
                            if (li < 0 || li >= TH || ln < 0 || ln >= TH) continue;
                            const int o = li * TW + lj, on = ln * TW + lj;
                            nxt[o] = add_rn(cur[on], mul_rn(coef, sub_rn(nxt[on], cur[o])));
                        }
                    }
                    __syncthreads();
                }
                // ---- S4: 5x5 corner means, main.py:54-61.  In the reference's order every cell reads its right / lower
                // (mirrored: inward) neighbours before they are overwritten, so all 100 means are functions of the
                // post-S3 field: one thread per cell reads, the CTA synchronises, then the cells are written. ---------
                if ((touchL || touchR) && (touchT || touchB)) {
                    T val = (T)0;
                    int dst = -1;
                    if (tid < 4 * RING * RING) {
                        const int corner = tid / (RING * RING), a = (tid / RING) % RING, c = tid % RING;
                        const bool top = corner < 2, left = (corner & 1) == 0;
                        const int gi = top ? a : Rg - 1 - a, li = gi - gr0, lin = top ? li + 1 : li - 1;
                        const int gj = left ? c : C - 1 - c, lj = gj - lc0, ljn = left ? lj + 1 : lj - 1;
                        if (li >= 0 && li < TH && lin >= 0 && lin < TH && lj >= 0 && lj < TW && ljn >= 0 && ljn < TW) {
                            val = mul_rn(add_rn(nxt[li * TW + ljn], nxt[lin * TW + lj]), (T)0.5);  // == sum / 2 exactly
                            dst = li * TW + lj;
                        }
                    }
                    __syncthreads();
                    if (dst >= 0) nxt[dst] = val;
                    __syncthreads();
                }
}

template <typename T, int TH, int TW, int NT>
__device__ __forceinline__ void source_and_probes(T* cur, const PassParams<T>& p, const TileCtx<T>& tc,
                                                  const long long step, const int tid) {
    const int gr0 = tc.gr0, lc0 = tc.lc0, k = tc.k;
    // ---- source add, fdtd.py:34 ---------------------------------------------------------------
    if (tc.src_hi > tc.src_lo) {
        if (step < p.amp_steps) {
            for (int q = tc.src_lo + tid; q < tc.src_hi; q += NT) {
                const Cell sc = p.src[q];
                const int li = sc.row - gr0, lj = sc.col - lc0;
                if (li >= 0 && li < TH && lj >= 0 && lj < TW) {
                    const double a = p.amp[(long long)sc.wave * p.amp_steps + step];
                    cur[li * TW + lj] = add_source(cur[li * TW + lj], a);
                }
            }
        }
        __syncthreads();
    }
    // ---- probes: recorded by the tile whose core holds the cell, on the owning slab -----------
    if (tc.prb_hi > tc.prb_lo && step < p.trace_cap) {
        for (int q = tc.prb_lo + tid; q < tc.prb_hi; q += NT) {
            const Cell pc = p.probes[q];
            const int li = pc.row - gr0, lj = pc.col - lc0;
            if (pc.row >= p.own_begin && pc.row < p.own_end && li >= k && li < k + p.CH && lj >= p.hx && lj < p.hx + p.CW)
                p.trace[step * p.n_probe + q] = cur[li * TW + lj];
        }
    }
}

}  // namespace fdtd2d
