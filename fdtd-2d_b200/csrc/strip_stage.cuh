// Staged wavefront (fp32, uniform permeability): the row-streaming wavefront of strip_wave.cuh with its K time levels
// split over S WARPS that pass rows to one another through shared memory -- K = S x L levels per HBM round trip at the
// register cost of L.
//
// Why: at K = 8 the one-warp kernel moves 28 B per cell and pass at 94 % of the copy bandwidth; more levels per pass is the
// only way to fewer DRAM bytes per step, but one warp cannot hold more: 8 levels take 232 registers, the 12-level instance
// sits at 255 with two warps per scheduler and is bound by latency.  Here a GROUP of S = 3 warps works on one run of one
// strip: warp t holds the window of levels tL .. (t+1)L - 1 (L = 4: 2 x 5 rows x 3 fields x 4 columns = 120 registers), so
// 12 levels cost 28 B per cell and 12 steps (a third less DRAM per step than K = 8) with 12 .. 15 warps per SM instead of 8.
//   warp 0   prefetches level-0 rows from HBM with cp.async (fields into a 4-row ring, dt/(eps*dx) into a 32-row ring that
//            all three warps read), steps them through levels 0..L-1 and puts what leaves level L-1 into hand-off ring 0;
//   warp t   takes its arriving rows from hand-off ring t-1, steps them through its L levels, and puts the result into
//            hand-off ring t -- or, the last warp, stores it to HBM straight from registers.
// A hand-off ring has 4 slots of 3 x 512 B and two mbarriers per slot ("full" / "empty", 32 arrivals each: every lane
// arrives after its own 16-byte stores / loads, and lane l only ever reads what lane l of the producer wrote).  No
// __syncthreads; the three warps drift a few rows apart and fill each other's pipeline bubbles.  Rows that do not exist
// (above the first level-0 row of the run) are not handed on: every warp starts from a zeroed window, exactly like the
// levels of the one-warp kernel, so the results are bit-identical to it (same operations, same order).
// The per-level arithmetic is that of wave_run_x2 (packed FADD2 / FFMA2, the left / right Mur ring riding along in the LR
// form); runs come from the same task lists (k = 12: column halo 12, core 104).
#pragma once
#include "strip_wave.cuh"

namespace fdtd2d {

constexpr int STAGE_L = 4;    // levels per warp
constexpr int STAGE_S = 3;    // warps per group: K = 12
constexpr int STAGE_D = 4;    // slots of a hand-off ring
constexpr int STAGE_NCR = 32; // rows of the coefficient ring: >= P + 1 + (S-1)(L + D) + L + 1 = 25
constexpr int STAGE_P = 3;    // rows prefetched ahead by warp 0

// shared memory of one group: field ring, coefficient ring, S-1 hand-off rings, their mbarriers, the task slot
__host__ __device__ constexpr size_t stage_group_bytes() {
    return (size_t)(WAVE_NF * 3 + STAGE_NCR + (STAGE_S - 1) * STAGE_D * 3) * WAVE_ROW_BYTES + (STAGE_S - 1) * STAGE_D * 2 * 8 + 64;
}

__device__ __forceinline__ void mbar_arrive_cta(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ bool mbar_try_cta(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cta(uint32_t bar, uint32_t parity) {
    if (mbar_try_cta(bar, parity)) return;  // (the common case costs one instruction; the clock is only read when waiting)
    const unsigned long long t0 = globaltimer_ns();
    while (!mbar_try_cta(bar, parity))
        if (globaltimer_ns() - t0 > 4000000000ull) __trap();  // a protocol bug must not hang the GPU
}
__device__ __forceinline__ void group_sync(int g) { asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "n"(STAGE_S * 32) : "memory"); }

// One level of the wavefront on one row (the body of wave_run_x2's level loop): st = the row stored at this level, ar = the
// row arriving at it, hx_above = Hx of the row above at the next level; out = the stored row one level on.
template <bool LR>
__device__ __forceinline__ void stage_level(const u64 (&st)[3][2], const u64 (&ar)[3][2], const u64 (&hx_above)[2], u64 (&out)[3][2], const u64 (&ce)[2],
                                            const u64 chu, const u64 negzero, const int side, const float coef, const uint32_t (&ringm)[4],
                                            const uint32_t (&hoffm)[4], const uint32_t (&padm)[4]) {
    constexpr unsigned FULL = 0xffffffffu;
    auto bsel = [](uint32_t m, float a, float b) { return __uint_as_float((__float_as_uint(a) & m) | (__float_as_uint(b) & ~m)); };
    const u64 e0 = st[0][0], e1 = st[0][1];
    // H half-step of the stored row (main.py:69-74)
    const float right3 = __shfl_down_sync(FULL, lo2(e0), 1);
    const u64 dx0 = pack2(sub_rn(hi2(e0), lo2(e0)), sub_rn(lo2(e1), hi2(e0)));
    const u64 dx1 = pack2(sub_rn(hi2(e1), lo2(e1)), sub_rn(right3, hi2(e1)));
    out[1][0] = sub2(st[1][0], mul2(chu, sub2(ar[0][0], e0), negzero));
    out[1][1] = sub2(st[1][1], mul2(chu, sub2(ar[0][1], e1), negzero));
    const u64 y0 = add2(st[2][0], mul2(chu, dx0, negzero));
    const u64 y1 = add2(st[2][1], mul2(chu, dx1, negzero));
    out[2][0] = y0;
    out[2][1] = y1;
    // its Ez update (main.py:21-27); Hx of the row above is one level up
    const float left0 = __shfl_up_sync(FULL, hi2(y1), 1);
    const u64 dy0 = pack2(sub_rn(lo2(y0), left0), sub_rn(hi2(y0), lo2(y0)));
    const u64 dy1 = pack2(sub_rn(lo2(y1), hi2(y0)), sub_rn(hi2(y1), lo2(y1)));
    const u64 curl0 = sub2(dy0, sub2(out[1][0], hx_above[0]));
    const u64 curl1 = sub2(dy1, sub2(out[1][1], hx_above[1]));
    out[0][0] = add2(e0, mul2(curl0, ce[0], negzero));
    out[0][1] = add2(e1, mul2(curl1, ce[1], negzero));
    if (LR && side != 0) {  // the left / right Mur ring (main.py:33-41), see wave_run_x2
        const float s0[4] = {lo2(e0), hi2(e0), lo2(e1), hi2(e1)};
        const float s1[4] = {lo2(out[0][0]), hi2(out[0][0]), lo2(out[0][1]), hi2(out[0][1])};
        float o[4];
        if (side == 1) {
            const float s1r = __shfl_down_sync(FULL, s1[0], 1);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float a0 = q < 3 ? s0[q < 3 ? q + 1 : 3] : right3, a1 = q < 3 ? s1[q < 3 ? q + 1 : 3] : s1r;
                o[q] = bsel(ringm[q], add_rn(a0, mul_rn(coef, sub_rn(a1, s0[q]))), s1[q]);
            }
        } else {
            const float s0l = __shfl_up_sync(FULL, s0[3], 1), s1l = __shfl_up_sync(FULL, s1[3], 1);
            const float hxo[4] = {lo2(st[1][0]), hi2(st[1][0]), lo2(st[1][1]), hi2(st[1][1])};
            const float hyo[4] = {lo2(st[2][0]), hi2(st[2][0]), lo2(st[2][1]), hi2(st[2][1])};
            const float hxn[4] = {lo2(out[1][0]), hi2(out[1][0]), lo2(out[1][1]), hi2(out[1][1])};
            const float hyn[4] = {lo2(y0), hi2(y0), lo2(y1), hi2(y1)};
            float hx2[4], hy2[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float b0 = q > 0 ? s0[q > 0 ? q - 1 : 0] : s0l, b1 = q > 0 ? s1[q > 0 ? q - 1 : 0] : s1l;
                o[q] = bsel(padm[q], s0[q], bsel(ringm[q], add_rn(b0, mul_rn(coef, sub_rn(b1, s0[q]))), s1[q]));
                hx2[q] = bsel(hoffm[q], hxo[q], hxn[q]);
                hy2[q] = bsel(hoffm[q], hyo[q], hyn[q]);
            }
            out[1][0] = pack2(hx2[0], hx2[1]), out[1][1] = pack2(hx2[2], hx2[3]);
            out[2][0] = pack2(hy2[0], hy2[1]), out[2][1] = pack2(hy2[2], hy2[3]);
        }
        out[0][0] = pack2(o[0], o[1]), out[0][1] = pack2(o[2], o[3]);
    }
}

// One run, seen by the warp of stage T of its group.  n0 = level-0 rows of the run.  `sent` / `recv` count the rows this
// warp has put into / taken from its hand-off rings since the kernel started (slot and barrier phase follow from them).
template <int T, bool LR>
__device__ __forceinline__ void stage_run(const PassParams<float>& p, const WaveTask& tk, float* fring, float* cring, float* hin, float* hout,
                                          const uint32_t bar_in, const uint32_t bar_out, unsigned& recv, unsigned& sent, const int l, const u64 chu,
                                          const u64 negzero) {
    constexpr int L = STAGE_L, S = STAGE_S, K = L * S, D = STAGE_D, TW = WAVE_TW, NF = WAVE_NF, NC = STAGE_NCR, P = STAGE_P;
    constexpr bool FIRST = T == 0, LAST = T == S - 1;
    const bool core = 4 * l >= tk.c0 && 4 * l < tk.c1;
    uint32_t ringm[4] = {0, 0, 0, 0}, hoffm[4] = {0, 0, 0, 0}, padm[4] = {0, 0, 0, 0};
    float coef = 0.0f;
    if (LR) {
        coef = p.mur[tk.b];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int gj = tk.x0 + 4 * l + q;
            auto lt = [](int a, int b) {  // a < b ? ~0 : 0
                uint32_t m;
                asm("shr.s32 %0, %1, 31;" : "=r"(m) : "r"(a - b));
                return m;
            };
            ringm[q] = tk.side == 1 ? lt(gj, RING) : (lt(p.C - RING - 1, gj) & lt(gj, p.C));
            hoffm[q] = lt(p.C - 2, gj);
            padm[q] = lt(p.C - 1, gj);
        }
    }
    const int n = tk.y1 - tk.y0 + 2 * K - T * L;  // rows arriving at this stage: the first one is row y0 - K for every stage
    long long of = (long long)tk.b * p.grid_stride + tk.x0 + 4 * l + (long long)(tk.y0 - K) * p.pitch;  // FIRST: next row to fetch
    auto fetch = [&](int fs, int cs) {
        cp_async16(fring + (fs * 3 + 0) * TW, p.in[0] + of);
        cp_async16(fring + (fs * 3 + 1) * TW, p.in[1] + of);
        cp_async16(fring + (fs * 3 + 2) * TW, p.in[2] + of);
        cp_async16(cring + cs * TW, p.ce + of);
    };
    u64 X[L + 1][3][2], Y[L + 1][3][2];
#pragma unroll
    for (int s = 0; s <= L; ++s)
#pragma unroll
        for (int f = 0; f < 3; ++f) X[s][f][0] = X[s][f][1] = Y[s][f][0] = Y[s][f][1] = 0ull;
    int fs = 0, cs = 1;  // FIRST: ring slots of the next row to fetch; coefficient slot of row (y0 - K - 1 + i) is i & (NC - 1)
    if (FIRST) {
#pragma unroll
        for (int d = 0; d < P; ++d) {
            fetch(fs, cs);
            of += p.pitch;
            cp_async_commit();
            fs = (fs + 1) & (NF - 1);
            cs = (cs + 1) & (NC - 1);
        }
    }
    int fr = 0;
    auto iter = [&](u64 (&ST)[L + 1][3][2], u64 (&AR)[L + 1][3][2], const int j) {
        // ---- the arriving row of this stage's first level ----
        if (FIRST) {
            if (j + P < n) fetch(fs, cs);
            of += p.pitch;
            cp_async_commit();
            fs = (fs + 1) & (NF - 1);
            cs = (cs + 1) & (NC - 1);
            cp_async_wait<P>();
            load22(fring + (fr * 3 + 0) * TW, AR[0][0]);
            load22(fring + (fr * 3 + 1) * TW, AR[0][1]);
            load22(fring + (fr * 3 + 2) * TW, AR[0][2]);
            fr = (fr + 1) & (NF - 1);
        } else {
            const uint32_t slot = recv & (D - 1);
            mbar_wait_cta(bar_in + slot * 16, (recv / D) & 1u);  // "full"
            const float* h = hin + slot * 3 * TW;
            load22(h + 0 * TW, AR[0][0]);
            load22(h + 1 * TW, AR[0][1]);
            load22(h + 2 * TW, AR[0][2]);
            mbar_arrive_cta(bar_in + slot * 16 + 8);  // "empty"
            ++recv;
        }
        // ---- L levels ----
#pragma unroll
        for (int s = 0; s < L; ++s) {
            u64 ce[2];
            // stored row of level s: row (y0 - K + j - s - 1) -> coefficient slot (j - s)
            load22(cring + ((j - s) & (NC - 1)) * TW, ce);
            stage_level<LR>(ST[s], AR[s], ST[s + 1][1], AR[s + 1], ce, chu, negzero, tk.side, coef, ringm, hoffm, padm);
        }
        // ---- what leaves the last level: row (y0 - K + j - L), L levels on ----
        if (LAST) {
            if (core && j >= K + L) {  // rows y0 .. y1 - 1
                const long long o = (long long)tk.b * p.grid_stride + tk.x0 + 4 * l + (long long)(tk.y0 - K + j - L) * p.pitch;
                store22(p.out[0] + o, AR[L][0]);
                store22(p.out[1] + o, AR[L][1]);
                store22(p.out[2] + o, AR[L][2]);
            }
        } else if (j >= L) {  // rows from y0 - K on exist; the ones before are never handed on
            const uint32_t slot = sent & (D - 1);
            mbar_wait_cta(bar_out + slot * 16 + 8, ((sent / D) & 1u) ^ 1u);  // "empty" (free at the start)
            float* h = hout + slot * 3 * TW;
            store22(h + 0 * TW, AR[L][0]);
            store22(h + 1 * TW, AR[L][1]);
            store22(h + 2 * TW, AR[L][2]);
            mbar_arrive_cta(bar_out + slot * 16);  // "full"
            ++sent;
        }
    };
    int j = 0;
#pragma unroll 1
    for (; j + 1 < n; j += 2) {
        iter(X, Y, j);
        iter(Y, X, j + 1);
    }
    if (j < n) iter(X, Y, j);
    if (FIRST) cp_async_wait<0>();
}

// CTA = NG groups of S warps.  Shared memory per group: see stage_group_bytes().
template <int NG, bool RING>
__global__ void __launch_bounds__(NG * STAGE_S * 32, 1) strip_stage_kernel(const PassParams<float> p, const WaveTask* tasks, const int n_tasks, int* ticket,
                                                                         const float ch_uniform, const u64 negzero) {
    constexpr int S = STAGE_S, D = STAGE_D, TW = WAVE_TW, NF = WAVE_NF, NC = STAGE_NCR;
    extern __shared__ __align__(16) unsigned char smem_stage[];
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31, g = w / S, t = w - g * S;
    unsigned char* gbase = smem_stage + (size_t)g * stage_group_bytes();
    float* fring = reinterpret_cast<float*>(gbase) + 4 * l;                       // [NF][3][TW]
    float* cring = fring + NF * 3 * TW;                                           // [NC][TW]
    float* hand = cring + NC * TW;                                                // [S-1][D][3][TW]
    uint64_t* bars = reinterpret_cast<uint64_t*>(gbase + (size_t)(NF * 3 + NC + (S - 1) * D * 3) * WAVE_ROW_BYTES);  // [S-1][D][full, empty]
    int* gtask = reinterpret_cast<int*>(bars + (S - 1) * D * 2);
    if (t == 0 && l == 0)
        for (int i = 0; i < (S - 1) * D * 2; ++i) mbar_init(bars + i, 32);
    if (t == 0 && l == 0) fence_mbar_init();
    group_sync(g);
    const u64 chu = pack2(ch_uniform, ch_uniform);
    float* hin = t > 0 ? hand + (size_t)(t - 1) * D * 3 * TW : hand;
    float* hout = t < S - 1 ? hand + (size_t)t * D * 3 * TW : hand;
    const uint32_t bar_in = smem_u32(bars + (t > 0 ? t - 1 : 0) * D * 2), bar_out = smem_u32(bars + (t < S - 1 ? t : 0) * D * 2);
    unsigned recv = 0, sent = 0;
    for (;;) {
        if (t == 0 && l == 0) *gtask = atomicAdd(ticket, 1);
        if (t == 0) {  // row y0 - K - 1 does not exist: its coefficient slot (0) reads as zeros
            const float z[4] = {0.0f, 0.0f, 0.0f, 0.0f};
            store4(cring, z);
        }
        group_sync(g);
        const int ti = *gtask;
        if (ti >= n_tasks) {  // (the last ticket drawn by the launch resets the counter, as in strip_wave.cuh)
            if (t == 0 && l == 0 && ti == n_tasks + (int)gridDim.x * NG - 1) atomicExch(ticket, 0);
            break;
        }
        const WaveTask tk = tasks[ti];
        const bool lr = RING && tk.side != 0;
        if (t == 0) {
            if (lr) stage_run<0, true>(p, tk, fring, cring, hin, hout, bar_in, bar_out, recv, sent, l, chu, negzero);
            else stage_run<0, false>(p, tk, fring, cring, hin, hout, bar_in, bar_out, recv, sent, l, chu, negzero);
        } else if (t == 1) {
            if (lr) stage_run<1, true>(p, tk, fring, cring, hin, hout, bar_in, bar_out, recv, sent, l, chu, negzero);
            else stage_run<1, false>(p, tk, fring, cring, hin, hout, bar_in, bar_out, recv, sent, l, chu, negzero);
        } else {
            if (lr) stage_run<2, true>(p, tk, fring, cring, hin, hout, bar_in, bar_out, recv, sent, l, chu, negzero);
            else stage_run<2, false>(p, tk, fring, cring, hin, hout, bar_in, bar_out, recv, sent, l, chu, negzero);
        }
        group_sync(g);  // the rings and the task slot are reused by the next run
    }
}

}  // namespace fdtd2d
