// Generic shared-memory tile kernel: k leapfrog steps per HBM round trip on one haloed tile.
//
// One CTA loads a TH x TW tile (core + halo) of Ez, Hx, Hy, ce, ch into shared memory, advances it k
// steps entirely on chip (overlapped/"trapezoid" temporal blocking: cells near a tile edge that is
// interior to the domain go stale one cell per step and are discarded), and stores the core to the
// other half of the ping-pong state.  It implements EVERYTHING the reference step does, for every
// tile position:
//   H half-step                     python-src/main.py:66-76
//   Ez interior update              python-src/main.py:18-27
//   Mur ABC, 5 px, L/R then T/B     python-src/main.py:29-51  (stages S2, S3 of SURVEY Appendix A)
//   5x5 corner means                python-src/main.py:53-61  (stage S4)
//   point/line source add           python-src/fdtd.py:34     (float64 add, cast to run dtype)
//   probe sampling                  (field readout after each step)
// It is the path for edge tiles, tiles holding sources/probes, fp64, and the per-function entry
// points (phases mask); plain interior fp32 tiles are taken by the register-resident fast kernel.
//
// Validity argument (why halo = k suffices): a cell update reads radius-1 neighbours in the interior,
// so staleness entering from an interior-side tile edge moves one cell per step.  Ring cells read
// only themselves and cells further INWARD (S2: same row; S3: same column; S4: both), so staleness
// can only run faster than that while travelling OUTWARD inside the 6-cell ring zone.  The tiling
// chosen by the host (api.cu: plan_tiles) guarantees that no tile has an interior-side edge within
// k+6 cells of a ring zone that lies in its core, so that case never arises.
#pragma once
#include "common.cuh"

namespace fdtd2d {

template <typename T> __device__ __forceinline__ typename Vec<T>::type zero_vec();
template <> __device__ __forceinline__ float4 zero_vec<float>() { return make_float4(0.f, 0.f, 0.f, 0.f); }
template <> __device__ __forceinline__ double2 zero_vec<double>() { return make_double2(0.0, 0.0); }

template <typename T, int TH, int TW, int NT>
__global__ void __launch_bounds__(NT) tile_generic_kernel(const PassParams<T> p) {
    constexpr int N = TH * TW;
    constexpr int VN = Vec<T>::N;
    constexpr int TWV = TW / VN;
    using V = typename Vec<T>::type;
    static_assert(TW % VN == 0, "tile width must be a multiple of the vector width");

    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* sEa = reinterpret_cast<T*>(smem_raw);
    T* sEb = sEa + N;
    T* sHx = sEb + N;
    T* sHy = sHx + N;
    T* sCe = sHy + N;
    T* sCh = sCe + N;

    const int tid = threadIdx.x;
    const int tile = p.tile_list ? p.tile_list[blockIdx.x] : (int)blockIdx.x;
    const int per_grid = p.tiles_y * p.tiles_x;
    const int b = tile / per_grid;
    const int rem = tile - b * per_grid;
    const int ty = rem / p.tiles_x;
    const int tx = rem - ty * p.tiles_x;
    const int k = p.k;
    const int lr0 = ty * p.CH - k;      // local row held in shared row 0
    const int lc0 = tx * p.CW - p.hx;   // column held in shared column 0
    const int gr0 = lr0 + p.row0;       // its global row
    const long long gbase = (long long)b * p.grid_stride;
    const int Rg = p.Rg, C = p.C;

    // ---- load tile (+halo), zero outside the local array --------------------------------------
    for (int v = tid; v < N / VN; v += NT) {
        const int li = v / TWV, lv = v - li * TWV;
        const int r = lr0 + li, c = lc0 + lv * VN;
        V e = zero_vec<T>(), hx = e, hy = e, ce = e, ch = e;
        if (r >= 0 && r < p.Rl && c >= 0 && c < p.pitch) {
            const long long off = gbase + (long long)r * p.pitch + c;
            e = *reinterpret_cast<const V*>(p.in[0] + off);
            hx = *reinterpret_cast<const V*>(p.in[1] + off);
            hy = *reinterpret_cast<const V*>(p.in[2] + off);
            ce = *reinterpret_cast<const V*>(p.ce + off);
            ch = *reinterpret_cast<const V*>(p.ch + off);
        }
        const int so = li * TW + lv * VN;
        *reinterpret_cast<V*>(sEa + so) = e;
        *reinterpret_cast<V*>(sHx + so) = hx;
        *reinterpret_cast<V*>(sHy + so) = hy;
        *reinterpret_cast<V*>(sCe + so) = ce;
        *reinterpret_cast<V*>(sCh + so) = ch;
    }
    __syncthreads();

    const T coef = p.mur[b];
    const bool touchL = lc0 < RING;
    const bool touchR = lc0 + TW > C - RING;
    const bool touchT = gr0 < RING;
    const bool touchB = gr0 + TH > Rg - RING;
    const int src_lo = p.src_range ? p.src_range[b] : 0;
    const int src_hi = p.src_range ? p.src_range[b + 1] : 0;
    const int prb_lo = p.probe_range ? p.probe_range[b] : 0;
    const int prb_hi = p.probe_range ? p.probe_range[b + 1] : 0;

    T* cur = sEa;  // Ez at the start of the step (S0)
    T* nxt = sEb;  // Ez being built (S1..S4)

    for (int s = 0; s < k; ++s) {
        // ---- H half-step, main.py:69-74 : rows 0..R-2, cols 0..C-2 -----------------------------
        if (p.phases & 1) {
            for (int idx = tid; idx < N; idx += NT) {
                const int li = idx / TW, lj = idx - li * TW;
                const int gi = gr0 + li, gj = lc0 + lj;
                if (li < TH - 1 && lj < TW - 1 && gi >= 0 && gi <= Rg - 2 && gj >= 0 && gj <= C - 2) {
                    const T e = cur[idx];
                    const T c = sCh[idx];
                    sHx[idx] = sub_rn(sHx[idx], mul_rn(c, sub_rn(cur[idx + TW], e)));
                    sHy[idx] = add_rn(sHy[idx], mul_rn(c, sub_rn(cur[idx + 1], e)));
                }
            }
            __syncthreads();
        }
        if (p.phases & 2) {
            // ---- S1: interior Ez update, main.py:21-27 : rows 1..R-2, cols 1..C-2 ---------------
            for (int idx = tid; idx < N; idx += NT) {
                const int li = idx / TW, lj = idx - li * TW;
                const int gi = gr0 + li, gj = lc0 + lj;
                T v = cur[idx];
                if (li >= 1 && lj >= 1 && gi >= 1 && gi <= Rg - 2 && gj >= 1 && gj <= C - 2) {
                    const T dhy = sub_rn(sHy[idx], sHy[idx - 1]);
                    const T dhx = sub_rn(sHx[idx], sHx[idx - TW]);
                    v = add_rn(v, mul_rn(sub_rn(dhy, dhx), sCe[idx]));
                }
                nxt[idx] = v;
            }
            __syncthreads();
            // ---- S2: Mur left/right, main.py:33-41. One thread per (row, side) runs the
            // reference's five column updates in the reference's order (outermost first), so every
            // read of the inward neighbour sees the value S1 left there. -----------------------
            if (touchL || touchR) {
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
            // ---- S4: 5x5 corner means, main.py:54-61. One thread per corner, reference order. ---
            if ((touchL || touchR) && (touchT || touchB)) {
                if (tid < 4) {
                    const bool top = tid < 2, left = (tid & 1) == 0;
                    for (int a = 0; a < RING; ++a) {
                        const int gi = top ? a : Rg - 1 - a;
                        const int li = gi - gr0, lin = top ? li + 1 : li - 1;
                        if (li < 0 || li >= TH || lin < 0 || lin >= TH) continue;
                        for (int c = 0; c < RING; ++c) {
                            const int gj = left ? c : C - 1 - c;
                            const int lj = gj - lc0, ljn = left ? lj + 1 : lj - 1;
                            if (lj < 0 || lj >= TW || ljn < 0 || ljn >= TW) continue;
                            const T sum = add_rn(nxt[li * TW + ljn], nxt[lin * TW + lj]);
                            nxt[li * TW + lj] = mul_rn(sum, (T)0.5);  // == sum / 2 exactly
                        }
                    }
                }
                __syncthreads();
            }
            T* t = cur;
            cur = nxt;
            nxt = t;
        }
        if (p.phases & 4) {
            const long long step = p.step0 + s;
            // ---- source add, fdtd.py:34 -------------------------------------------------------
            if (src_hi > src_lo) {
                if (step < p.amp_steps) {
                    for (int q = src_lo + tid; q < src_hi; q += NT) {
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
            // ---- probes: recorded by the tile whose core holds the cell, on the owning slab ---
            if (prb_hi > prb_lo && step < p.trace_cap) {
                for (int q = prb_lo + tid; q < prb_hi; q += NT) {
                    const Cell pc = p.probes[q];
                    const int li = pc.row - gr0, lj = pc.col - lc0;
                    if (pc.row >= p.own_begin && pc.row < p.own_end && li >= k && li < k + p.CH && lj >= p.hx &&
                        lj < p.hx + p.CW)
                        p.trace[step * p.n_probe + q] = cur[li * TW + lj];
                }
            }
        }
    }

    // ---- store the core ---------------------------------------------------------------------
    const int CWV = p.CW / VN;
    const int ncore = p.CH * CWV;
    for (int v = tid; v < ncore; v += NT) {
        const int ci = v / CWV, cv = v - ci * CWV;
        const int li = k + ci, lj = p.hx + cv * VN;
        const int r = lr0 + li, c = lc0 + lj;
        if (r < p.Rl && c < p.pitch) {
            const long long off = gbase + (long long)r * p.pitch + c;
            const int so = li * TW + lj;
            *reinterpret_cast<V*>(p.out[0] + off) = *reinterpret_cast<const V*>(cur + so);
            *reinterpret_cast<V*>(p.out[1] + off) = *reinterpret_cast<const V*>(sHx + so);
            *reinterpret_cast<V*>(p.out[2] + off) = *reinterpret_cast<const V*>(sHy + so);
        }
    }
}

}  // namespace fdtd2d
