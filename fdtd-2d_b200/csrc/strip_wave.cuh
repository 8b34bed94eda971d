// Row-streaming wavefront kernels: K leapfrog steps per HBM round trip with redundancy only in the COLUMN halo.
// Two forms: strip_wave_x2_kernel (fp32, sm_100a's two-wide fp32 instructions; in its LR form the left / right Mur ring
// rides along) and strip_wave_kernel (scalar arithmetic, used for fp64: two columns per lane, 64-column strips).
//
// The overlapped tiles of tile_tma.cuh recompute a halo of K rows above and below every 64-row tile (core 48 x 112 of
// 64 x 128: 34 % of the arithmetic is thrown away at K = 8).  Here one WARP owns a strip of 32 x 16 bytes of columns
// (fp32: 128 columns, core 112 at K = 8) and marches down a long run of rows, carrying all K time levels of a sliding
// window of rows in its REGISTERS:
//   level 0 row i arrives from HBM; for s = 0..K-1 the row stored at level s (row i-s-1) and the arriving one (row i-s)
//   give level s+1 of row i-s-1 (H half-step main.py:69-74, then the interior Ez update main.py:21-27, which takes
//   Hx of the row above from the row stored at level s+1); that result is the arriving row of the next level; what
//   leaves level K-1 is row i-K advanced K steps and goes straight to HBM.
// So every row is loaded once, stepped K times and stored once; the only recomputation is the halo columns on each
// side of the strip (12.5 % at K = 8 in fp32) and K warm-up rows per run of rows (< 4 %).  A warp never talks to another
// warp: no __syncthreads, no shared-memory exchange of field rows; column neighbours come from two shuffles per row and
// level.  Rows are prefetched three iterations ahead with cp.async (16 B per lane and array) into a small per-warp ring;
// the coefficient rows stay in their ring until the last level has used them (K rows later).  Runs of rows are handed
// out to the warps dynamically (one atomic per run).
// Cells next to the strip's edge columns and above the first / below the last row of the run read neighbours that are
// missing; they go stale one cell per level exactly as in the overlapped tiles and are never stored.
// Arithmetic and results are bit-identical to the other kernels (same operations, same order).
//
// y-slabs (SLAB instantiations): a run whose rows are a slab's BAND (the `halo` owned rows next to a neighbour slab) first
// waits until that neighbour has delivered the ghost rows it reads, stores its result rows into the neighbour's ghost
// rows as well (peer stores over NVLink) and, when it is the last band task of its side, raises the neighbour's flag
// (common.cuh: band_wait / band_done).  Band runs come first in the task list, so the exchange is over long before the
// pass is.
//
// Two passes per launch (FUSE instantiations): at K = 8 the kernel moves 28 B per cell and pass at 94 % of the copy bandwidth;
// the way past that wall is to let the SECOND pass of a pair read what the first one wrote while it is still in the 126 MB
// L2.  One launch holds the runs of pass p (phase 0: set A -> set B) and of pass p + 1 (phase 1: set B -> set A) in one
// ticket order in which every phase-1 run comes after the phase-0 runs it reads.  A phase-0 run publishes a flag per
// 16-row block of its tile column as soon as the block is in memory (fence + store); a phase-1 run, before it prefetches
// the first row of a block, waits for the flags of that block in the (up to three) tile columns its window covers.  So it
// trails its producers by 16..32 rows: its 16 B per cell come from L2, and DRAM sees 16 B read + 12 B written by phase 0
// and 12 B written by phase 1 per 16 steps -- 2.5 B per cell-step instead of 3.5.  Only runs whose whole window is
// produced by phase-0 RUNS are fused; the pieces next to edge tiles (top / bottom ring, sources, probes) are left to a
// short second launch, after the phase-0 edge tiles are done (api.cu).  Waiting is safe: a run only ever waits for runs
// with smaller ticket numbers, which are running or done, and those never wait.
#pragma once
#include "common.cuh"
#include "tile_edge.cuh"

namespace fdtd2d {

constexpr int WAVE_NW = 8;      // warps per CTA: 2 per scheduler, each may use up to 255 registers for the K-level window
constexpr int WAVE_P = 3;       // rows prefetched ahead at K = 8 (6 was measured: no gain); K = 12 uses 2 so the ring still fits
constexpr int WAVE_NF = 4;      // rows of the field ring (a power of two > P)
constexpr int WAVE_NC = 16;     // rows of the coefficient ring (a power of two >= K + P + 2)
constexpr int WAVE_ROW_BYTES = 512;  // one ring row of one array: 16 bytes per lane (128 fp32 / 64 fp64 columns)
constexpr int WAVE_TW = 128;    // fp32 strip width (columns per warp)

// one run of rows of one strip: grid b, the strip's first column is x0, rows [y0, y1) are stored, of the strip's columns
// [c0, c1) (strip coordinates; whole 16-byte groups).
struct WaveTask {
    int32_t b, x0, y0, y1;
    int32_t c0, c1;
    int32_t side;  // 0 plain strip, 1 holds the left Mur ring (columns 0..4), 2 the right one
    int32_t band;  // 0, or 1 / 2: the rows are the band next to the top / bottom neighbour slab
    // fused double pass (see "two passes per launch" below)
    int32_t phase = 0;           // 0: first pass of the launch (producer), 1: second pass (consumer)
    int32_t tx = 0;              // tile column whose cells this run stores (index of its progress flags)
    int32_t txlo = 0, txhi = 0;  // tile columns its 128-column window reads: the producers a phase-1 run waits for
};
constexpr int WAVE_TASK_WORDS = 12;
constexpr int FUSE_BLOCK_LOG2 = 4;  // progress is published per block of 16 rows

// per warp: NF rows of the three fields + NC rows of dt/(eps*dx) (+ NC rows of dt/(mu*dx) unless that is a scalar)
__host__ __device__ constexpr size_t wave_smem_bytes(int warps, bool no_ch_ring) {
    return (size_t)warps * (WAVE_NF * 3 + WAVE_NC * (no_ch_ring ? 1 : 2)) * WAVE_ROW_BYTES;
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// 16 bytes <-> registers
__device__ __forceinline__ void ld16(const float* p, float (&a)[4]) { unpack4(*reinterpret_cast<const float4*>(p), a); }
__device__ __forceinline__ void ld16(const double* p, double (&a)[2]) {
    const double2 v = *reinterpret_cast<const double2*>(p);
    a[0] = v.x, a[1] = v.y;
}
__device__ __forceinline__ void st16(float* p, const float (&a)[4]) { *reinterpret_cast<float4*>(p) = make_float4(a[0], a[1], a[2], a[3]); }
__device__ __forceinline__ void st16(double* p, const double (&a)[2]) { *reinterpret_cast<double2*>(p) = make_double2(a[0], a[1]); }

// ---- scalar form (fp64; also valid for fp32) ---------------------------------------------------------------------------
// One run of rows of one strip, K levels deep.  UCH: dt/(mu*dx) is the same in every cell (true for every material_init
// output, main.py:105,121): it comes as a kernel argument and the map is neither fetched nor kept in the ring.
// The window: two register sets X, Y that swap roles every iteration so that no row is ever moved.  In an iteration
// ST[s] (s < K) is the row stored at level s (row i-s-1 when row i arrives), ST[K][1] the Hx of the last row that left;
// AR[0] receives the arriving level-0 row and AR[s+1] the result of level s, i.e. the row arriving at level s+1.  After
// the iteration the arrived rows ARE the stored rows: the sets swap.
// LR: the strip holds the left (tk.side == 1) or right (2) Mur ring, which rides along exactly as in wave_run_x2 below: at
// every level Ez[i, q] = S0[i, q+1] + coef * (S1[i, q+1] - S0[i, q]) on the five ring columns (mirrored on the right), and
// the right strip keeps H beyond column C-2 and the pad columns as they were.
template <typename T, int K, bool UCH, int P, bool BAND, bool LR = false>
__device__ __forceinline__ void wave_run_scalar(const PassParams<T>& p, const WaveTask& tk, T* fring, T* cring, const int l, const T ch_uniform) {
    constexpr int NQ = 16 / (int)sizeof(T), TW = 32 * NQ, NF = WAVE_NF, NC = WAVE_NC, CS = UCH ? 1 : 2;
    constexpr unsigned FULL = 0xffffffffu;
    const bool core = NQ * l >= tk.c0 && NQ * l < tk.c1;
    const int side = tk.band - 1;
    if (BAND) band_wait(p, side);
    bool ringc[NQ], hoffc[NQ], padc[NQ];
    T coef = (T)0;
#pragma unroll
    for (int q = 0; q < NQ; ++q) ringc[q] = hoffc[q] = padc[q] = false;
    if (LR) {
        coef = p.mur[tk.b];
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            const int gj = tk.x0 + NQ * l + q;
            ringc[q] = tk.side == 1 ? gj < RING : (gj >= p.C - RING && gj < p.C);
            hoffc[q] = gj > p.C - 2;  // H is updated in columns 0..C-2 (main.py:70,74)
            padc[q] = gj > p.C - 1;
        }
    }
    // j counts the level-0 rows of the run, `of` is the element offset of the next row to fetch and moves one row per
    // iteration; the row that leaves level K-1 in iteration j is K rows behind the arriving one, i.e. P + 1 + K rows behind `of`
    const int n = tk.y1 - tk.y0 + 2 * K;  // level-0 rows [y0 - K, y1 + K)
    long long of = (long long)tk.b * p.grid_stride + tk.x0 + NQ * l + (long long)(tk.y0 - K) * p.pitch;
    auto fetch = [&](int fs, int cs) {
        cp_async16(fring + (fs * 3 + 0) * TW, p.in[0] + of);
        cp_async16(fring + (fs * 3 + 1) * TW, p.in[1] + of);
        cp_async16(fring + (fs * 3 + 2) * TW, p.in[2] + of);
        cp_async16(cring + (cs * CS + 0) * TW, p.ce + of);
        if (!UCH) cp_async16(cring + (cs * CS + 1) * TW, p.ch + of);
    };
    T X[K + 1][3][NQ], Y[K + 1][3][NQ];
#pragma unroll
    for (int s = 0; s <= K; ++s)
#pragma unroll
        for (int f = 0; f < 3; ++f)
#pragma unroll
            for (int q = 0; q < NQ; ++q) X[s][f][q] = Y[s][f][q] = (T)0;
    // Row y0-K-1 (the row "stored" at level 0 before the first arrival) does not exist as data; its coefficient slot
    // (slot 0) is read by level 0 in the first iteration: give it zeros.
    int fs = 0, cs = 1;
    {
        T z[NQ];
#pragma unroll
        for (int q = 0; q < NQ; ++q) z[q] = (T)0;
        st16(cring + 0 * TW, z);
        if (!UCH) st16(cring + 1 * TW, z);
    }
#pragma unroll
    for (int d = 0; d < P; ++d) {
        fetch(fs, cs);
        of += p.pitch;
        cp_async_commit();
        fs = (fs + 1) & (NF - 1);
        cs = (cs + 1) & (NC - 1);
    }
    int fr = 0, cr = 0;
    auto iter = [&](T (&ST)[K + 1][3][NQ], T (&AR)[K + 1][3][NQ], const int j) {
        if (j + P < n) fetch(fs, cs);
        of += p.pitch;
        cp_async_commit();  // (an empty group keeps the wait count uniform at the end of the run)
        fs = (fs + 1) & (NF - 1);
        cs = (cs + 1) & (NC - 1);
        cp_async_wait<P>();  // the arriving row has landed (every lane reads back only the 16 bytes it copied itself)
        ld16(fring + (fr * 3 + 0) * TW, AR[0][0]);
        ld16(fring + (fr * 3 + 1) * TW, AR[0][1]);
        ld16(fring + (fr * 3 + 2) * TW, AR[0][2]);
        fr = (fr + 1) & (NF - 1);
#pragma unroll
        for (int s = 0; s < K; ++s) {
            T ce[NQ], ch[NQ];
            const int c = (cr - s) & (NC - 1);  // coefficient slot of row i-s-1
            ld16(cring + (c * CS + 0) * TW, ce);
            if (UCH) {
#pragma unroll
                for (int q = 0; q < NQ; ++q) ch[q] = ch_uniform;
            } else {
                ld16(cring + (c * CS + 1) * TW, ch);
            }
            const T right_nb = __shfl_down_sync(FULL, ST[s][0][0], 1);
#pragma unroll
            for (int q = 0; q < NQ; ++q) {  // H half-step of the stored row (main.py:69-74)
                const T right = q < NQ - 1 ? ST[s][0][q < NQ - 1 ? q + 1 : q] : right_nb;
                AR[s + 1][1][q] = sub_rn(ST[s][1][q], mul_rn(ch[q], sub_rn(AR[s][0][q], ST[s][0][q])));
                AR[s + 1][2][q] = add_rn(ST[s][2][q], mul_rn(ch[q], sub_rn(right, ST[s][0][q])));
            }
            const T left_nb = __shfl_up_sync(FULL, AR[s + 1][2][NQ - 1], 1);
#pragma unroll
            for (int q = 0; q < NQ; ++q) {  // its Ez update (main.py:21-27); Hx of the row above is one level up
                const T left = q > 0 ? AR[s + 1][2][q > 0 ? q - 1 : 0] : left_nb;
                const T curl = sub_rn(sub_rn(AR[s + 1][2][q], left), sub_rn(AR[s + 1][1][q], ST[s + 1][1][q]));
                AR[s + 1][0][q] = add_rn(ST[s][0][q], mul_rn(curl, ce[q]));
            }
            if (LR && tk.side != 0) {  // the left / right Mur ring (main.py:33-41)
                T out[NQ];
                if (tk.side == 1) {
                    const T s1r = __shfl_down_sync(FULL, AR[s + 1][0][0], 1);
#pragma unroll
                    for (int q = 0; q < NQ; ++q) {
                        const T a0 = q < NQ - 1 ? ST[s][0][q < NQ - 1 ? q + 1 : q] : right_nb;
                        const T a1 = q < NQ - 1 ? AR[s + 1][0][q < NQ - 1 ? q + 1 : q] : s1r;
                        const T m = add_rn(a0, mul_rn(coef, sub_rn(a1, ST[s][0][q])));
                        out[q] = ringc[q] ? m : AR[s + 1][0][q];
                    }
                } else {
                    const T s0l = __shfl_up_sync(FULL, ST[s][0][NQ - 1], 1), s1l = __shfl_up_sync(FULL, AR[s + 1][0][NQ - 1], 1);
#pragma unroll
                    for (int q = 0; q < NQ; ++q) {
                        const T b0 = q > 0 ? ST[s][0][q > 0 ? q - 1 : 0] : s0l, b1 = q > 0 ? AR[s + 1][0][q > 0 ? q - 1 : 0] : s1l;
                        const T m = add_rn(b0, mul_rn(coef, sub_rn(b1, ST[s][0][q])));
                        out[q] = padc[q] ? ST[s][0][q] : (ringc[q] ? m : AR[s + 1][0][q]);
                        AR[s + 1][1][q] = hoffc[q] ? ST[s][1][q] : AR[s + 1][1][q];
                        AR[s + 1][2][q] = hoffc[q] ? ST[s][2][q] : AR[s + 1][2][q];
                    }
                }
#pragma unroll
                for (int q = 0; q < NQ; ++q) AR[s + 1][0][q] = out[q];
            }
        }
        cr = (cr + 1) & (NC - 1);
        if (core && j >= 2 * K) {  // row y0 + (j - 2K) < y1 has left level K-1, K steps on
            const long long o = of - (long long)(P + 1 + K) * p.pitch;
            st16(p.out[0] + o, AR[K][0]);
            st16(p.out[1] + o, AR[K][1]);
            st16(p.out[2] + o, AR[K][2]);
            if (BAND) {
                T* const q0 = peer_field(p, side, 0);
                if (q0) {
                    const long long po = o + peer_shift(p, side);
                    st16(q0 + po, AR[K][0]);
                    st16(peer_field(p, side, 1) + po, AR[K][1]);
                    st16(peer_field(p, side, 2) + po, AR[K][2]);
                }
            }
        }
    };
    int j = 0;
#pragma unroll 1
    for (; j + 1 < n; j += 2) {
        iter(X, Y, j);
        iter(Y, X, j + 1);
    }
    if (j < n) iter(X, Y, j);
    cp_async_wait<0>();
    if (BAND) {
        __threadfence_system();
        __syncwarp();
        if (l == 0) band_done(p, side);
    }
    __syncwarp();  // the ring is reused by the next run
}

template <typename T, int K, bool UCH, int P, bool SLAB, bool RING = false>
__global__ void __launch_bounds__(WAVE_NW * 32, 1) strip_wave_kernel(const PassParams<T> p, const WaveTask* tasks, const int n_tasks, int* ticket, const T ch_uniform) {
    constexpr int NQ = 16 / (int)sizeof(T), TW = 32 * NQ, NF = WAVE_NF, NC = WAVE_NC, CS = UCH ? 1 : 2;
    static_assert((NF & (NF - 1)) == 0 && NF > P && (NC & (NC - 1)) == 0 && NC >= K + P + 2, "ring sizes");
    extern __shared__ __align__(16) unsigned char smem_wave[];
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    T* fring = reinterpret_cast<T*>(smem_wave) + (size_t)w * (NF * 3 + NC * CS) * TW + NQ * l;  // [NF][3][TW], my 16 bytes
    T* cring = fring + NF * 3 * TW;                                                           // [NC][CS][TW]
    for (;;) {  // runs are handed out dynamically: a warp takes the next one as soon as it is done
        int t = 0;
        if (l == 0) t = atomicAdd(ticket, 1);
        t = __shfl_sync(0xffffffffu, t, 0);
        if (t >= n_tasks) {
            // every warp of the launch draws exactly one ticket past the end: the last of those puts the counter back to
            // zero for the next launch (no memset per pass -- a memset can queue behind a host copy on a copy engine)
            if (l == 0 && t == n_tasks + (int)gridDim.x * WAVE_NW - 1) atomicExch(ticket, 0);
            break;
        }
        const WaveTask tk = tasks[t];
        if (RING && tk.side != 0) {
            if (SLAB && tk.band)
                wave_run_scalar<T, K, UCH, P, true, true>(p, tk, fring, cring, l, ch_uniform);
            else
                wave_run_scalar<T, K, UCH, P, false, true>(p, tk, fring, cring, l, ch_uniform);
        } else if (SLAB && tk.band) {
            wave_run_scalar<T, K, UCH, P, true>(p, tk, fring, cring, l, ch_uniform);
        } else {
            wave_run_scalar<T, K, UCH, P, false>(p, tk, fring, cring, l, ch_uniform);
        }
    }
}

// ---- packed form (fp32): the same wavefront with Blackwell's two-wide fp32 instructions ---------------------------------
// sm_100a has add/sub/fma.rn.f32x2 (SASS FADD2 / FFMA2): one instruction, two IEEE round-to-nearest results on an aligned
// register pair.  They run at half the issue rate of FADD (measured, profiles/micro/f32x2_rate.cu: 0.50 against 0.96
// warp-instructions per clock and scheduler), i.e. the same 128 lane-operations per clock and SM, but they take half the
// ISSUE SLOTS -- and the scalar form is bound by issue (79 %), its FP pipe only 63 % busy.  A lane's four columns
// are two pairs (c, c+1), (c+2, c+3), exactly as the 16-byte loads deliver them; only the two column differences, whose
// operands straddle the pairs, stay scalar (their results land in a pair directly).
// A level is 18 packed + 8 scalar instructions + 2 shuffles instead of 44 + 2 (437 -> 362 instructions per row at K = 8).
// Bit-exactness: ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into one FFMA2 even with --fmad false (scalar code is not
// touched), so the product is written as fma.rn.f32x2(a, b, -0) with the -0 pair coming in as a kernel argument the
// compiler cannot see through: rn(a*b + -0) = rn(a*b) for every input incl. signed zeros, and an FFMA2 that already
// has an addend cannot absorb the add that follows.
__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned ld_relaxed_gpu(const unsigned* p) {
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// A phase-1 run waits until block `blk` of the tile columns [txlo, txlo + ntx) of grid b has been published by phase 0.
// Every lane polls (lane l the flag of column txlo + min(l, ntx - 1)) with RELAXED loads -- an acquire load costs an
// invalidation of the whole L1 (CCTL.IVALL) per poll, which stalled the SM's other warps (measured: 24 M of them per
// launch, 13 % of all stall samples) -- and does one acquire load once the flag is up.  Gives up after 2 s and reports
// through the flag block instead of hanging the GPU.
__device__ __forceinline__ void fuse_wait(const PassParams<float>& p, const WaveTask& tk, const int blk, const int l) {
    const int ntx = tk.txhi - tk.txlo + 1;
    const unsigned* f = p.fuse_flags + ((long long)tk.b * p.tiles_x + tk.txlo + (l < ntx ? l : ntx - 1)) * p.fuse_nblk + blk;
    if (!__all_sync(0xffffffffu, ld_relaxed_gpu(f) != 0u)) {
        const unsigned long long t0 = globaltimer_ns();
        for (;;) {
            __nanosleep(400);
            if (__all_sync(0xffffffffu, ld_relaxed_gpu(f) != 0u)) break;
            if (globaltimer_ns() - t0 > HALO_WAIT_NS) {
                if (l == 0) atomicExch(p.flags + FLAG_ERR, 3u);
                return;
            }
        }
    }
    (void)ld_acquire_gpu(f);  // orders the prefetches that follow behind the flag
}

using u64 = unsigned long long;
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 c; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(c) : "l"(a), "l"(b)); return c; }
__device__ __forceinline__ u64 sub2(u64 a, u64 b) { u64 c; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(c) : "l"(a), "l"(b)); return c; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b, u64 negzero) { u64 c; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(c) : "l"(a), "l"(b), "l"(negzero)); return c; }
__device__ __forceinline__ u64 pack2(float lo, float hi) { u64 c; asm("mov.b64 %0, {%1, %2};" : "=l"(c) : "f"(lo), "f"(hi)); return c; }
__device__ __forceinline__ float lo2(u64 a) { return __uint_as_float((uint32_t)a); }
__device__ __forceinline__ float hi2(u64 a) { return __uint_as_float((uint32_t)(a >> 32)); }
__device__ __forceinline__ void load22(const float* p, u64* a) { const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(p); a[0] = v.x, a[1] = v.y; }
__device__ __forceinline__ void store22(float* p, const u64* a) { *reinterpret_cast<ulonglong2*>(p) = make_ulonglong2(a[0], a[1]); }

// One run of rows of one strip, K levels deep (the body of strip_wave_x2_kernel).
// LR = true: the strip holds the left or right Mur ring (main.py:33-41) over rows that are plain (no top / bottom ring,
// no source, no probe in reach).  The ring rides along the wavefront: at every level the five ring columns are set from
// the row's own values before (S0) and after (S1) the interior update,
//   Ez[i, q] = S0[i, q+1] + coef * (S1[i, q+1] - S0[i, q])   (left; mirrored on the right),
// which is the reference's column-by-column loop with every read resolved (each column reads its inward neighbour
// before that one is overwritten).  The reference's slice bounds (H: columns 0..C-2, Ez: 1..C-2) are imposed by selects
// on the right strip, which also covers the pad columns >= C (kept as loaded: zero).
// MODE: 0 plain; 1 the rows are a slab's band; 2 / 3 phase 0 / phase 1 of a fused double pass (see the top of the file;
// phase 1 reads what the launch's PassParams call `out` and writes what they call `in`).
template <int K, bool UCH, int P, bool LR, int MODE>
__device__ __forceinline__ void wave_run_x2(const PassParams<float>& p, const WaveTask& tk, float* fring, float* cring, const int l, const u64 chu, const u64 negzero) {
    constexpr int TW = WAVE_TW, NF = WAVE_NF, NC = WAVE_NC, CS = UCH ? 1 : 2;
    constexpr unsigned FULL = 0xffffffffu;
    constexpr bool BAND = MODE == 1, PUB = MODE == 2, SUB = MODE == 3;
    const bool core = 4 * l >= tk.c0 && 4 * l < tk.c1;
    const int bside = tk.band - 1;
    if (BAND) band_wait(p, bside);
    // (compile-time selects: the pointers stay constant-bank operands)
    auto IN = [&](int f) -> const float* { return SUB ? p.out[f] : p.in[f]; };
    auto OUT = [&](int f) -> float* { return SUB ? const_cast<float*>(p.in[f]) : p.out[f]; };
    {
        // ring strips: which of my four columns are ring columns / beyond the reference's slices, as all-ones masks for
        // bitwise selects (one LOP3 each; predicates would have to be recomputed at every use)
        uint32_t ringm[4] = {0, 0, 0, 0}, hoffm[4] = {0, 0, 0, 0}, padm[4] = {0, 0, 0, 0};
        float coef = 0.0f;
        if (LR) {
            coef = p.mur[tk.b];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int gj = tk.x0 + 4 * l + q;
                // (sign bits smeared by an arithmetic shift: written as a comparison the compiler turns the masks
                // back into predicates)
                auto lt = [](int a, int b) {  // a < b ? ~0 : 0
                    uint32_t m;
                    asm("shr.s32 %0, %1, 31;" : "=r"(m) : "r"(a - b));
                    return m;
                };
                ringm[q] = tk.side == 1 ? lt(gj, RING) : (lt(p.C - RING - 1, gj) & lt(gj, p.C));
                hoffm[q] = lt(p.C - 2, gj);  // H is updated in columns 0..C-2 (main.py:70,74)
                padm[q] = lt(p.C - 1, gj);
            }
        }
        auto bsel = [](uint32_t m, float a, float b) {  // m ? a : b
            return __uint_as_float((__float_as_uint(a) & m) | (__float_as_uint(b) & ~m));
        };
        // Loop state is kept small (the window takes 13 x 12 registers at K = 12): j counts the level-0 rows of the run,
        // `of` is the element offset of the next row to fetch and moves one row per iteration; the row that leaves
        // level K-1 in iteration j is K rows behind the arriving one, i.e. P + 1 + K rows behind `of`.
        const int n = tk.y1 - tk.y0 + 2 * K;  // level-0 rows [y0 - K, y1 + K)
        long long of = (long long)tk.b * p.grid_stride + tk.x0 + 4 * l + (long long)(tk.y0 - K) * p.pitch;
        auto fetch = [&](int fs, int cs, int jf) {  // jf = index of the fetched row among the run's level-0 rows
            if (SUB) {  // entering a new 16-row block of what phase 0 produces: wait for it
                const int row = tk.y0 - K + jf;
                if (jf == 0 || (row & ((1 << FUSE_BLOCK_LOG2) - 1)) == 0) fuse_wait(p, tk, row >> FUSE_BLOCK_LOG2, l);
            }
            const long long o = of;
            cp_async16(fring + (fs * 3 + 0) * TW, IN(0) + o);
            cp_async16(fring + (fs * 3 + 1) * TW, IN(1) + o);
            cp_async16(fring + (fs * 3 + 2) * TW, IN(2) + o);
            cp_async16(cring + (cs * CS + 0) * TW, p.ce + o);
            if (!UCH) cp_async16(cring + (cs * CS + 1) * TW, p.ch + o);
        };
        // the window of wave_run_scalar, every row as two column pairs
        u64 X[K + 1][3][2], Y[K + 1][3][2];
#pragma unroll
        for (int s = 0; s <= K; ++s)
#pragma unroll
            for (int f = 0; f < 3; ++f) X[s][f][0] = X[s][f][1] = Y[s][f][0] = Y[s][f][1] = 0ull;
        int fs = 0, cs = 1;
        {
            const float z[4] = {0.0f, 0.0f, 0.0f, 0.0f};
            store4(cring + 0 * TW, z);
            if (!UCH) store4(cring + 1 * TW, z);
        }
#pragma unroll
        for (int d = 0; d < P; ++d) {
            fetch(fs, cs, d);
            of += p.pitch;
            cp_async_commit();
            fs = (fs + 1) & (NF - 1);
            cs = (cs + 1) & (NC - 1);
        }
        int fr = 0, cr = 0;
        auto iter = [&](u64 (&ST)[K + 1][3][2], u64 (&AR)[K + 1][3][2], const int j) {
            if (j + P < n) fetch(fs, cs, j + P);
            of += p.pitch;
            cp_async_commit();
            fs = (fs + 1) & (NF - 1);
            cs = (cs + 1) & (NC - 1);
            cp_async_wait<P>();
            load22(fring + (fr * 3 + 0) * TW, AR[0][0]);
            load22(fring + (fr * 3 + 1) * TW, AR[0][1]);
            load22(fring + (fr * 3 + 2) * TW, AR[0][2]);
            fr = (fr + 1) & (NF - 1);
#pragma unroll
            for (int s = 0; s < K; ++s) {
                u64 ce[2], ch[2];
                const int c = (cr - s) & (NC - 1);  // coefficient slot of row i-s-1
                load22(cring + (c * CS + 0) * TW, ce);
                if (UCH)
                    ch[0] = ch[1] = chu;
                else
                    load22(cring + (c * CS + 1) * TW, ch);
                const u64 e0 = ST[s][0][0], e1 = ST[s][0][1];
                // H half-step of the stored row (main.py:69-74).  The column differences straddle the pairs, so they
                // are four scalar subtractions written straight into a pair; the rest is two-wide.
                const float right3 = __shfl_down_sync(FULL, lo2(e0), 1);
                const u64 dx0 = pack2(sub_rn(hi2(e0), lo2(e0)), sub_rn(lo2(e1), hi2(e0)));
                const u64 dx1 = pack2(sub_rn(hi2(e1), lo2(e1)), sub_rn(right3, hi2(e1)));
                AR[s + 1][1][0] = sub2(ST[s][1][0], mul2(ch[0], sub2(AR[s][0][0], e0), negzero));
                AR[s + 1][1][1] = sub2(ST[s][1][1], mul2(ch[1], sub2(AR[s][0][1], e1), negzero));
                const u64 y0 = add2(ST[s][2][0], mul2(ch[0], dx0, negzero));
                const u64 y1 = add2(ST[s][2][1], mul2(ch[1], dx1, negzero));
                AR[s + 1][2][0] = y0;
                AR[s + 1][2][1] = y1;
                // its Ez update (main.py:21-27); Hx of the row above is one level up
                const float left0 = __shfl_up_sync(FULL, hi2(y1), 1);
                const u64 dy0 = pack2(sub_rn(lo2(y0), left0), sub_rn(hi2(y0), lo2(y0)));
                const u64 dy1 = pack2(sub_rn(lo2(y1), hi2(y0)), sub_rn(hi2(y1), lo2(y1)));
                const u64 curl0 = sub2(dy0, sub2(AR[s + 1][1][0], ST[s + 1][1][0]));
                const u64 curl1 = sub2(dy1, sub2(AR[s + 1][1][1], ST[s + 1][1][1]));
                AR[s + 1][0][0] = add2(e0, mul2(curl0, ce[0], negzero));
                AR[s + 1][0][1] = add2(e1, mul2(curl1, ce[1], negzero));
                if (LR && tk.side != 0) {
                    const float s0[4] = {lo2(e0), hi2(e0), lo2(e1), hi2(e1)};
                    const float s1[4] = {lo2(AR[s + 1][0][0]), hi2(AR[s + 1][0][0]), lo2(AR[s + 1][0][1]), hi2(AR[s + 1][0][1])};
                    float out[4];
                    if (tk.side == 1) {
                        const float s1r = __shfl_down_sync(FULL, s1[0], 1);
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const float a0 = q < 3 ? s0[q < 3 ? q + 1 : 3] : right3, a1 = q < 3 ? s1[q < 3 ? q + 1 : 3] : s1r;
                            const float m = add_rn(a0, mul_rn(coef, sub_rn(a1, s0[q])));
                            out[q] = bsel(ringm[q], m, s1[q]);
                        }
                    } else {
                        const float s0l = __shfl_up_sync(FULL, s0[3], 1), s1l = __shfl_up_sync(FULL, s1[3], 1);
                        const float hxo[4] = {lo2(ST[s][1][0]), hi2(ST[s][1][0]), lo2(ST[s][1][1]), hi2(ST[s][1][1])};
                        const float hyo[4] = {lo2(ST[s][2][0]), hi2(ST[s][2][0]), lo2(ST[s][2][1]), hi2(ST[s][2][1])};
                        const float hxn[4] = {lo2(AR[s + 1][1][0]), hi2(AR[s + 1][1][0]), lo2(AR[s + 1][1][1]), hi2(AR[s + 1][1][1])};
                        const float hyn[4] = {lo2(y0), hi2(y0), lo2(y1), hi2(y1)};
                        float hx2[4], hy2[4];
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const float b0 = q > 0 ? s0[q > 0 ? q - 1 : 0] : s0l, b1 = q > 0 ? s1[q > 0 ? q - 1 : 0] : s1l;
                            const float m = add_rn(b0, mul_rn(coef, sub_rn(b1, s0[q])));
                            out[q] = bsel(padm[q], s0[q], bsel(ringm[q], m, s1[q]));
                            hx2[q] = bsel(hoffm[q], hxo[q], hxn[q]);
                            hy2[q] = bsel(hoffm[q], hyo[q], hyn[q]);
                        }
                        AR[s + 1][1][0] = pack2(hx2[0], hx2[1]), AR[s + 1][1][1] = pack2(hx2[2], hx2[3]);
                        AR[s + 1][2][0] = pack2(hy2[0], hy2[1]), AR[s + 1][2][1] = pack2(hy2[2], hy2[3]);
                    }
                    AR[s + 1][0][0] = pack2(out[0], out[1]), AR[s + 1][0][1] = pack2(out[2], out[3]);
                }
            }
            cr = (cr + 1) & (NC - 1);
            if (core && j >= 2 * K) {  // row y0 + (j - 2K) < y1 has left level K-1, K steps on
                const long long o = of - (long long)(P + 1 + K) * p.pitch;
                store22(OUT(0) + o, AR[K][0]);
                store22(OUT(1) + o, AR[K][1]);
                store22(OUT(2) + o, AR[K][2]);
                if (BAND) {  // the same rows into the neighbour slab's ghost rows
                    float* const q0 = peer_field(p, bside, 0);
                    if (q0) {
                        const long long po = o + peer_shift(p, bside);
                        store22(q0 + po, AR[K][0]);
                        store22(peer_field(p, bside, 1) + po, AR[K][1]);
                        store22(peer_field(p, bside, 2) + po, AR[K][2]);
                    }
                }
            }
            if (PUB && j >= 2 * K) {  // the last row of a 16-row block is in flight: publish the block (runs are cut at blocks)
                const int row = tk.y0 + j - 2 * K;
                if (((row + 1) & ((1 << FUSE_BLOCK_LOG2) - 1)) == 0) {
                    __syncwarp();  // the other lanes' stores happen before lane 0's fence, which is cumulative
                    if (l == 0) {
                        __threadfence();
                        *(volatile unsigned*)(p.fuse_flags + ((long long)tk.b * p.tiles_x + tk.tx) * p.fuse_nblk + (row >> FUSE_BLOCK_LOG2)) = 1u;
                    }
                }
            }
        };
        int j = 0;
#pragma unroll 1
        for (; j + 1 < n; j += 2) {
            iter(X, Y, j);
            iter(Y, X, j + 1);
        }
        if (j < n) iter(X, Y, j);
        cp_async_wait<0>();
        if (BAND) {
            __threadfence_system();
            __syncwarp();
            if (l == 0) band_done(p, bside);
        }
        __syncwarp();  // the ring is reused by the next run
    }
}

// The kernel: warps take runs from one ticket; with RING the ring-strip runs go through the LR instantiation of the run,
// with SLAB the band runs through the BAND one, everything else through the plain one -- separate loops in one kernel,
// so the plain strips pay nothing for the ring / band code and one launch balances everything.
// FLAVOUR: 0 whole grid, one pass; 1 y-slab (band runs); 2 fused double pass (phase-0 and phase-1 runs).
template <int K, bool UCH, int P, bool RING, int FLAVOUR>
__global__ void __launch_bounds__(WAVE_NW * 32, 1) strip_wave_x2_kernel(const PassParams<float> p, const WaveTask* tasks, const int n_tasks, int* ticket, const float ch_uniform, const u64 negzero) {
    constexpr int TW = WAVE_TW, NF = WAVE_NF, NC = WAVE_NC, CS = UCH ? 1 : 2;  // CS: coefficient maps kept in the ring
    static_assert((NF & (NF - 1)) == 0 && NF > P && (NC & (NC - 1)) == 0 && NC >= K + P + 2, "ring sizes");
    extern __shared__ __align__(16) unsigned char smem_wave[];
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    float* fring = reinterpret_cast<float*>(smem_wave) + (size_t)w * (NF * 3 + NC * CS) * TW + 4 * l;  // [NF][3][TW], my 4 columns
    float* cring = fring + NF * 3 * TW;                                                               // [NC][CS][TW]
    const u64 chu = pack2(ch_uniform, ch_uniform);
    for (;;) {
        int t = 0;
        if (l == 0) t = atomicAdd(ticket, 1);
        t = __shfl_sync(0xffffffffu, t, 0);
        if (t >= n_tasks) {
            // every warp of the launch draws exactly one ticket past the end: the last of those puts the counter back to
            // zero for the next launch (no memset per pass -- a memset can queue behind a host copy on a copy engine)
            if (l == 0 && t == n_tasks + (int)gridDim.x * WAVE_NW - 1) atomicExch(ticket, 0);
            break;
        }
        const WaveTask tk = tasks[t];
        if (FLAVOUR == 2) {
            if (RING && tk.side != 0) {
                if (tk.phase)
                    wave_run_x2<K, UCH, P, true, 3>(p, tk, fring, cring, l, chu, negzero);
                else
                    wave_run_x2<K, UCH, P, true, 2>(p, tk, fring, cring, l, chu, negzero);
            } else {
                if (tk.phase)
                    wave_run_x2<K, UCH, P, false, 3>(p, tk, fring, cring, l, chu, negzero);
                else
                    wave_run_x2<K, UCH, P, false, 2>(p, tk, fring, cring, l, chu, negzero);
            }
        } else if (RING && tk.side != 0) {
            if (FLAVOUR == 1 && tk.band)
                wave_run_x2<K, UCH, P, true, 1>(p, tk, fring, cring, l, chu, negzero);
            else
                wave_run_x2<K, UCH, P, true, 0>(p, tk, fring, cring, l, chu, negzero);
        } else {
            if (FLAVOUR == 1 && tk.band)
                wave_run_x2<K, UCH, P, false, 1>(p, tk, fring, cring, l, chu, negzero);
            else
                wave_run_x2<K, UCH, P, false, 0>(p, tk, fring, cring, l, chu, negzero);
        }
    }
}

}  // namespace fdtd2d
