/*
 * Plain-C restatement of the reference's FDTD leapfrog path.
 * TEST INFRASTRUCTURE ONLY -- never linked into, loaded by, or called from the
 * product (fdtd-2d_b200/).  Used by tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py.
 *
 * Follows, statement by statement and in the reference's order:
 *   H half-step                 python-src/main.py:66-76
 *   Ez interior update          python-src/main.py:18-27
 *   Mur ABC left/right/top/bot  python-src/main.py:29-51  (in-place, sequential)
 *   corner means                python-src/main.py:53-61  (in-place, sequential)
 *   point source add            python-src/fdtd.py:34 + main.py:182-187
 *                               (float64 add on one cell, cast back to the run type)
 *   loop order                  python-src/fdtd.py:30-34
 * Arrays use the reference's shapes: Ez[R][C], Hx[R][C-1], Hy[R-1][C], C-contiguous.
 * Coefficient maps ce = dt/(eps*dx), ch = dt/(mu*dx) and the Mur coefficient are
 * formed by the caller with numpy exactly as main.py:27,30-31,70,74 forms them, so
 * the only arithmetic here is sub/mul/add (one IEEE rounding each).
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math [-fopenmp]  (see c_oracle.py).
 * -ffp-contract=off is mandatory: a fused multiply-add changes the bits.
 *
 * Parity pin: checked bit-for-bit against the .npz files under tests/golden (outputs of the real
 * reference) by tests/test_oracle_golden.py.
 */
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define RING 5

#define DEFINE_ORACLE(T, SFX)                                                                          \
    /* main.py:66-76 */                                                                                \
    void fdtd_oracle_update_h_##SFX(const T* Ez, T* Hx, T* Hy, const T* ch, int R, int C) {            \
        _Pragma("omp parallel for schedule(static)")                                                   \
        for (int i = 0; i < R - 1; ++i) {                                                              \
            const T* e0 = Ez + (size_t)i * C;                                                          \
            const T* e1 = e0 + C;                                                                      \
            const T* k = ch + (size_t)i * C;                                                           \
            T* hx = Hx + (size_t)i * (C - 1);                                                          \
            T* hy = Hy + (size_t)i * C;                                                                \
            for (int j = 0; j < C - 1; ++j) {                                                          \
                T dy = e1[j] - e0[j];                                                                  \
                T dxv = e0[j + 1] - e0[j];                                                             \
                T py = k[j] * dy;                                                                      \
                T px = k[j] * dxv;                                                                     \
                hx[j] = hx[j] - py;                                                                    \
                hy[j] = hy[j] + px;                                                                    \
            }                                                                                          \
        }                                                                                              \
    }                                                                                                  \
    /* main.py:12-63; prev is caller-provided scratch of R*C elements (Ez_prev) */                     \
    void fdtd_oracle_update_e_##SFX(T* Ez, const T* Hx, const T* Hy, const T* ce, T coef, int R,       \
                                    int C, T* prev) {                                                  \
        memcpy(prev, Ez, (size_t)R* C * sizeof(T)); /* main.py:18 */                                   \
        _Pragma("omp parallel for schedule(static)")                                                   \
        for (int i = 1; i < R - 1; ++i) { /* main.py:21-27 */                                          \
            T* e = Ez + (size_t)i * C;                                                                 \
            const T* k = ce + (size_t)i * C;                                                           \
            const T* hy = Hy + (size_t)i * C;                                                          \
            const T* hx = Hx + (size_t)i * (C - 1);                                                    \
            const T* hxm = hx - (C - 1);                                                               \
            for (int j = 1; j < C - 1; ++j) {                                                          \
                T dhy = hy[j] - hy[j - 1];                                                             \
                T dhx = hx[j] - hxm[j];                                                                \
                T curl = dhy - dhx;                                                                    \
                T inc = curl * k[j];                                                                   \
                e[j] = e[j] + inc;                                                                     \
            }                                                                                          \
        }                                                                                              \
        /* main.py:33-35 left, 38-41 right */                                                          \
        for (int k = 0; k < RING; ++k)                                                                 \
            for (int i = 1; i < R - 1; ++i) {                                                          \
                size_t o = (size_t)i * C;                                                              \
                T d = Ez[o + k + 1] - prev[o + k];                                                     \
                T m = coef * d;                                                                        \
                Ez[o + k] = prev[o + k + 1] + m;                                                       \
            }                                                                                          \
        for (int k = 0; k < RING; ++k)                                                                 \
            for (int i = 1; i < R - 1; ++i) {                                                          \
                size_t o = (size_t)i * C;                                                              \
                int c1 = C - 1 - k, c2 = C - 2 - k;                                                    \
                T d = Ez[o + c2] - prev[o + c1];                                                       \
                T m = coef * d;                                                                        \
                Ez[o + c1] = prev[o + c2] + m;                                                         \
            }                                                                                          \
        /* main.py:43-45 top, 48-51 bottom */                                                          \
        for (int k = 0; k < RING; ++k)                                                                 \
            for (int j = 1; j < C - 1; ++j) {                                                          \
                size_t a = (size_t)k * C + j, b = (size_t)(k + 1) * C + j;                             \
                T d = Ez[b] - prev[a];                                                                 \
                T m = coef * d;                                                                        \
                Ez[a] = prev[b] + m;                                                                   \
            }                                                                                          \
        for (int k = 0; k < RING; ++k)                                                                 \
            for (int j = 1; j < C - 1; ++j) {                                                          \
                size_t a = (size_t)(R - 1 - k) * C + j, b = (size_t)(R - 2 - k) * C + j;               \
                T d = Ez[b] - prev[a];                                                                 \
                T m = coef * d;                                                                        \
                Ez[a] = prev[b] + m;                                                                   \
            }                                                                                          \
        /* main.py:54-61 */                                                                            \
        for (int a = 0; a < RING; ++a)                                                                 \
            for (int b = 0; b < RING; ++b) {                                                           \
                size_t ra = (size_t)a * C, ra1 = (size_t)(a + 1) * C;                                  \
                size_t rb = (size_t)(R - 1 - a) * C, rb1 = (size_t)(R - 2 - a) * C;                    \
                T s;                                                                                   \
                s = Ez[ra + b + 1] + Ez[ra1 + b];                                                      \
                Ez[ra + b] = s / (T)2;                                                                 \
                s = Ez[ra + C - 2 - b] + Ez[ra1 + C - 1 - b];                                          \
                Ez[ra + C - 1 - b] = s / (T)2;                                                         \
                s = Ez[rb1 + b] + Ez[rb + b + 1];                                                      \
                Ez[rb + b] = s / (T)2;                                                                 \
                s = Ez[rb1 + C - 1 - b] + Ez[rb + C - 2 - b];                                          \
                Ez[rb + C - 1 - b] = s / (T)2;                                                         \
            }                                                                                          \
    }                                                                                                  \
    /* fdtd.py:30-34. amp[n] is the float64 source value of step step0+n (may be NULL).              \
     * n_src source cells share amp. trace (may be NULL) is [nsteps][n_probe], sampled after the add. \
     * Returns 0, or -1 if scratch allocation fails. */                                                \
    int fdtd_oracle_run_##SFX(T* Ez, T* Hx, T* Hy, const T* ce, const T* ch, T coef, int R, int C,     \
                              int nsteps, const double* amp, int n_src, const int* src_rc,             \
                              int n_probe, const int* probe_rc, T* trace) {                            \
        T* prev = (T*)malloc((size_t)R * C * sizeof(T));                                               \
        if (!prev) return -1;                                                                          \
        for (int n = 0; n < nsteps; ++n) {                                                             \
            fdtd_oracle_update_h_##SFX(Ez, Hx, Hy, ch, R, C);                                          \
            fdtd_oracle_update_e_##SFX(Ez, Hx, Hy, ce, coef, R, C, prev);                              \
            if (amp)                                                                                   \
                for (int s = 0; s < n_src; ++s) {                                                      \
                    size_t o = (size_t)src_rc[2 * s] * C + src_rc[2 * s + 1];                          \
                    double v = (double)Ez[o] + amp[n];                                                 \
                    Ez[o] = (T)v;                                                                      \
                }                                                                                      \
            if (trace)                                                                                 \
                for (int p = 0; p < n_probe; ++p)                                                      \
                    trace[(size_t)n * n_probe + p] =                                                   \
                        Ez[(size_t)probe_rc[2 * p] * C + probe_rc[2 * p + 1]];                         \
        }                                                                                              \
        free(prev);                                                                                    \
        return 0;                                                                                      \
    }

DEFINE_ORACLE(float, f32)
DEFINE_ORACLE(double, f64)

int fdtd_oracle_abi_version(void) { return 1; }
