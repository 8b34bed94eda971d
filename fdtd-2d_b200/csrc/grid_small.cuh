// Grids with fewer than 11 rows or columns (6 is the smallest size on which the reference's own boundary code indexes
// inside the arrays): the five-deep Mur strips of opposite sides overlap, so the staged dataflow of the other kernels
//  no longer equals the reference, whose statements then read what earlier statements of the SAME
// step wrote.  This kernel therefore executes python-src/main.py:12-76 statement by statement, in the reference's
// order: each numpy statement is one data-parallel sweep of the CTA over its slice, with a CTA barrier between
// statements; the 100 corner assignments (main.py:54-61) are scalar statements and run on one thread, in order.
// One CTA per grid of the batch, the whole fdtd2d_step call in one launch, fields in place in global memory (such a
// grid fits L1/L2 many times over); the other half of the ping-pong state serves as Ez_prev (main.py:18).
#pragma once
#include "common.cuh"

namespace fdtd2d {

constexpr int SMALL_MIN = 6;  // main.py:34-61 reads columns / rows 0..5 and -1..-6
constexpr int SMALL_NT = 256;

// in/out: p.in = the state (updated in place), p.out[0] = scratch for Ez_prev.  p.k = leapfrog steps.
template <typename T>
__global__ void __launch_bounds__(SMALL_NT) grid_small_kernel(const PassParams<T> p) {
    const int b = blockIdx.x, tid = threadIdx.x;
    const int R = p.Rg, C = p.C, P = p.pitch;
    const long long g = (long long)b * p.grid_stride;
    T* Ez = const_cast<T*>(p.in[0]) + g;
    T* Hx = const_cast<T*>(p.in[1]) + g;
    T* Hy = const_cast<T*>(p.in[2]) + g;
    T* S0 = p.out[0] + g;
    const T* ce = p.ce + g;
    const T* ch = p.ch + g;
    const T coef = p.mur[b];
    const int s_lo = p.src_range ? p.src_range[b] : 0, s_hi = p.src_range ? p.src_range[b + 1] : 0;
    const int p_lo = p.probe_range ? p.probe_range[b] : 0, p_hi = p.probe_range ? p.probe_range[b + 1] : 0;
    for (int s = 0; s < p.k; ++s) {
        const long long step = p.step0 + s;
        if (p.phases & 1) {  // main.py:66-76: rows 0..R-2, columns 0..C-2
            for (int c = tid; c < (R - 1) * (C - 1); c += SMALL_NT) {
                const int i = c / (C - 1), j = c - i * (C - 1), o = i * P + j;
                const T e = Ez[o];
                Hx[o] = sub_rn(Hx[o], mul_rn(ch[o], sub_rn(Ez[o + P], e)));
                Hy[o] = add_rn(Hy[o], mul_rn(ch[o], sub_rn(Ez[o + 1], e)));
            }
            __syncthreads();
        }
        if (p.phases & 2) {
            for (int c = tid; c < R * C; c += SMALL_NT) {  // main.py:18
                const int i = c / C, j = c - i * C;
                S0[i * P + j] = Ez[i * P + j];
            }
            __syncthreads();
            for (int c = tid; c < (R - 2) * (C - 2); c += SMALL_NT) {  // main.py:21-27
                const int i = 1 + c / (C - 2), j = 1 + c % (C - 2), o = i * P + j;
                const T curl = sub_rn(sub_rn(Hy[o], Hy[o - 1]), sub_rn(Hx[o], Hx[o - P]));
                Ez[o] = add_rn(Ez[o], mul_rn(curl, ce[o]));
            }
            __syncthreads();
            // main.py:33-51: twenty statements, each over rows 1..R-2 (left / right) or columns 1..C-2 (top / bottom)
            for (int q = 0; q < 2 * RING; ++q) {
                const int k = q % RING, col = q < RING ? k : C - 1 - k, nb = q < RING ? col + 1 : col - 1;
                for (int i = 1 + tid; i < R - 1; i += SMALL_NT)
                    Ez[i * P + col] = add_rn(S0[i * P + nb], mul_rn(coef, sub_rn(Ez[i * P + nb], S0[i * P + col])));
                __syncthreads();
            }
            for (int q = 0; q < 2 * RING; ++q) {
                const int k = q % RING, row = q < RING ? k : R - 1 - k, nb = q < RING ? row + 1 : row - 1;
                for (int j = 1 + tid; j < C - 1; j += SMALL_NT)
                    Ez[row * P + j] = add_rn(S0[nb * P + j], mul_rn(coef, sub_rn(Ez[nb * P + j], S0[row * P + j])));
                __syncthreads();
            }
            if (tid == 0) {  // main.py:54-61, in order
                for (int i = 0; i < RING; ++i)
                    for (int j = 0; j < RING; ++j) {
                        const int bi = R - 1 - i, rj = C - 1 - j;
                        Ez[i * P + j] = mul_rn(add_rn(Ez[i * P + j + 1], Ez[(i + 1) * P + j]), (T)0.5);     // == sum / 2 exactly
                        Ez[i * P + rj] = mul_rn(add_rn(Ez[i * P + rj - 1], Ez[(i + 1) * P + rj]), (T)0.5);
                        Ez[bi * P + j] = mul_rn(add_rn(Ez[(bi - 1) * P + j], Ez[bi * P + j + 1]), (T)0.5);
                        Ez[bi * P + rj] = mul_rn(add_rn(Ez[(bi - 1) * P + rj], Ez[bi * P + rj - 1]), (T)0.5);
                    }
            }
            __syncthreads();
        }
        if (p.phases & 4) {
            if (step < p.amp_steps)
                for (int q = s_lo + tid; q < s_hi; q += SMALL_NT) {  // fdtd.py:34
                    const Cell sc = p.src[q];
                    const int o = sc.row * P + sc.col;
                    Ez[o] = add_source(Ez[o], p.amp[(long long)sc.wave * p.amp_steps + step]);
                }
            __syncthreads();
            if (step < p.trace_cap)
                for (int q = p_lo + tid; q < p_hi; q += SMALL_NT) {
                    const Cell pc = p.probes[q];
                    p.trace[step * p.n_probe + q] = Ez[pc.row * P + pc.col];
                }
        }
    }
}

}  // namespace fdtd2d
