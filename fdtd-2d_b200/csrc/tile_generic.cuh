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
#include "ring_ops.cuh"

namespace fdtd2d {

template <typename T> __device__ __forceinline__ typename Vec<T>::type zero_vec();
template <> __device__ __forceinline__ float4 zero_vec<float>() { return make_float4(0.f, 0.f, 0.f, 0.f); }
template <> __device__ __forceinline__ double2 zero_vec<double>() { return make_double2(0.0, 0.0); }

__device__ __forceinline__ void unpack(const float4& v, float* a) { a[0] = v.x, a[1] = v.y, a[2] = v.z, a[3] = v.w; }
__device__ __forceinline__ void unpack(const double2& v, double* a) { a[0] = v.x, a[1] = v.y; }
__device__ __forceinline__ float4 pack(const float* a) { return make_float4(a[0], a[1], a[2], a[3]); }
__device__ __forceinline__ double2 pack(const double* a) { return make_double2(a[0], a[1]); }

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
    const int lr0 = p.org + ty * p.CH - k;  // local row held in shared row 0
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

    T* cur = sEa;  // Ez at the start of the step (S0)
    T* nxt = sEb;  // Ez being built (S1..S4)

    for (int s = 0; s < k; ++s) {
        // ---- H half-step, main.py:69-74 : rows 0..R-2, cols 0..C-2 -----------------------------
        // One vector (VN cells of one row) per thread and iteration; row predicates are uniform per
        // vector, column predicates select between the updated and the old value per element.
        if (p.phases & 1) {
            for (int v = tid; v < N / VN; v += NT) {
                const int li = v / TWV, lj0 = (v - li * TWV) * VN;
                const int gi = gr0 + li;
                if (li >= TH - 1 || gi < 0 || gi > Rg - 2) continue;
                const int o = li * TW + lj0;
                T e[VN + 1], dn[VN], c[VN], hx[VN], hy[VN];
                unpack(*reinterpret_cast<const V*>(cur + o), e);
                unpack(*reinterpret_cast<const V*>(cur + o + TW), dn);
                unpack(*reinterpret_cast<const V*>(sCh + o), c);
                unpack(*reinterpret_cast<const V*>(sHx + o), hx);
                unpack(*reinterpret_cast<const V*>(sHy + o), hy);
                e[VN] = (lj0 + VN < TW) ? cur[o + VN] : e[VN - 1];
#pragma unroll
                for (int q = 0; q < VN; ++q) {
                    const int lj = lj0 + q, gj = lc0 + lj;
                    const bool ok = (lj < TW - 1) & (gj >= 0) & (gj <= C - 2);
                    const T nx = sub_rn(hx[q], mul_rn(c[q], sub_rn(dn[q], e[q])));
                    const T ny = add_rn(hy[q], mul_rn(c[q], sub_rn(e[q + 1], e[q])));
                    hx[q] = ok ? nx : hx[q];
                    hy[q] = ok ? ny : hy[q];
                }
                *reinterpret_cast<V*>(sHx + o) = pack(hx);
                *reinterpret_cast<V*>(sHy + o) = pack(hy);
            }
            __syncthreads();
        }
        if (p.phases & 2) {
            // ---- S1: interior Ez update, main.py:21-27 : rows 1..R-2, cols 1..C-2 ---------------
            for (int v = tid; v < N / VN; v += NT) {
                const int li = v / TWV, lj0 = (v - li * TWV) * VN;
                const int gi = gr0 + li;
                const int o = li * TW + lj0;
                T e[VN];
                unpack(*reinterpret_cast<const V*>(cur + o), e);
                if (li >= 1 && gi >= 1 && gi <= Rg - 2) {
                    T hy[VN + 1], hx[VN], up[VN], c[VN];
                    unpack(*reinterpret_cast<const V*>(sHy + o), hy + 1);
                    unpack(*reinterpret_cast<const V*>(sHx + o), hx);
                    unpack(*reinterpret_cast<const V*>(sHx + o - TW), up);
                    unpack(*reinterpret_cast<const V*>(sCe + o), c);
                    hy[0] = (lj0 > 0) ? sHy[o - 1] : hy[1];
#pragma unroll
                    for (int q = 0; q < VN; ++q) {
                        const int lj = lj0 + q, gj = lc0 + lj;
                        const bool ok = (lj >= 1) & (gj >= 1) & (gj <= C - 2);
                        const T curl = sub_rn(sub_rn(hy[q + 1], hy[q]), sub_rn(hx[q], up[q]));
                        const T nv = add_rn(e[q], mul_rn(curl, c[q]));
                        e[q] = ok ? nv : e[q];
                    }
                }
                *reinterpret_cast<V*>(nxt + o) = pack(e);
            }
            __syncthreads();
            ring_stages<T, TH, TW, NT>(cur, nxt, tc, tid);
            T* t = cur;
            cur = nxt;
            nxt = t;
        }
        if (p.phases & 4) {
            source_and_probes<T, TH, TW, NT>(cur, p, tc, p.step0 + s, tid);
        }
    }

    // ---- store the core ---------------------------------------------------------------------
    const int CWV = p.CW / VN;
    const int ncore = p.CH * CWV;
    for (int v = tid; v < ncore; v += NT) {
        const int ci = v / CWV, cv = v - ci * CWV;
        const int li = k + ci, lj = p.hx + cv * VN;
        const int r = lr0 + li, c = lc0 + lj;
        if (r < p.store_hi && c < p.pitch) {
            const long long off = gbase + (long long)r * p.pitch + c;
            const int so = li * TW + lj;
            *reinterpret_cast<V*>(p.out[0] + off) = *reinterpret_cast<const V*>(cur + so);
            *reinterpret_cast<V*>(p.out[1] + off) = *reinterpret_cast<const V*>(sHx + so);
            *reinterpret_cast<V*>(p.out[2] + off) = *reinterpret_cast<const V*>(sHy + so);
        }
    }
}

}  // namespace fdtd2d
