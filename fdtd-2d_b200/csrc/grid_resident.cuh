// Cluster-resident kernel (fp32): one thread-block CLUSTER per grid; the whole grid stays on chip for ALL
// n leapfrog steps of a call -- HBM is touched once to load the five arrays and once to store the three
// fields, whatever n is.  This is the path for the small independent grids of the batched mode
// (BASELINE configs[4]: 1024 x 256^2) and for the reference demo grid (fdtd.py:14-19, 200^2), where
// overlapped tiling has nothing to amortise: every tile of such a grid touches the Mur ring.
//
// Work layout.  A cluster of n <= 8 CTAs splits the rows of one grid into n bands (multiples of MR rows, at most
// MR*NW rows, at most 256 columns; the first and last band are shorter because those CTAs also run the top /
// bottom boundary pass).  Inside a CTA warp w owns rows [w*MR, (w+1)*MR) of the band and lane l owns columns
// [4l, 4l+4) and [128+4l, 128+4l+4): MR x 8 cells of Ez, Hx, Hy per thread live in REGISTERS from the first step
// to the last.  The coefficient maps dt/(eps*dx), dt/(mu*dx) live in shared memory (each thread re-reads only the
// slots it wrote).  Neighbours:
//   columns across lanes -> warp shuffles (rotating, so column 127 <-> 128 is one more select);
//   rows across warps    -> every warp publishes its first and last Ez row in shared memory (double-buffered by
//     step parity) and keeps a REDUNDANT copy of the Hx row just above its rows, which it advances itself from the
//     Ez row it receives.  So only Ez is exchanged, once per step, and a step has ONE CTA-wide barrier;
//   rows across the CTAs of a cluster -> DISTRIBUTED SHARED MEMORY: the band's first / last Ez row goes straight
//     from registers into the neighbour CTA's shared memory with st.async (...mbarrier::complete_tx::bytes): the
//     stores themselves signal the receiver's mbarrier, so there is no fence (a release.cluster arrive costs a
//     MEMBAR.GPU) and no cluster-wide barrier inside the time loop.  Only the one warp that needs the row arms
//     the barrier (arrive.expect_tx) and waits, after all its other rows.  A CTA can never run more than one step
//     ahead of a neighbour, so two buffers per direction suffice.
// Index ranges of the reference's slices (main.py:70,74 rows 0..R-2 / cols 0..C-2 for H, :27 rows 1..R-2 /
// cols 1..C-2 for Ez) are imposed by zeroing the on-chip copy of the coefficient at the excluded cells:
// x -/+ 0*(..) leaves x unchanged (for finite fields), so the inner loops carry no masks.  Every cell the Ez
// mask excludes is overwritten by the boundary stages below, exactly as in the reference.
//
// Boundary stages (main.py:29-61) run on small shared-memory frames holding only the ring: the outermost columns
// of every row (frame LR), rows 0..5 (frame T, first CTA) and R-6..R-1 (frame B, last CTA), each as S0 (Ez before
// the step, the reference's Ez_prev) and S1 (after the interior update).
//   S2 (Mur left/right): by the warp that owns the row, one lane per ring cell, in place, no CTA barrier;
//   S3 + S4 (Mur top/bottom, 5x5 corner means -- a Jacobi sweep, every read is of a not-yet-processed cell): one
//     pass of all threads of the first / last CTA over the six rows, each cell evaluated from S0 and S2;
// then the owners pull the finished ring back into registers.  Source cells and probes that are not in a ring
// frame get a 4-cell slot: the owner thread parks the group there, adds the float64 source amplitude
// (fdtd.py:34) and reloads it; probes are sampled from the frames after the next step's barrier.
// Results are bit-identical to the tile kernels and to the reference (tests/test_gpu_resident.py).
#pragma once
#include "common.cuh"
#include "tile_edge.cuh"
#include "tile_tma.cuh"

#ifndef FDTD2D_RES_DIAG
#define FDTD2D_RES_DIAG 0  // MEASUREMENT AID (wrong results): 1 = no top / bottom pass, 2 = no ring work at all
#endif

namespace fdtd2d {

constexpr int RES_TW = 256;         // columns per CTA (two 128-column halves per warp row)
constexpr int RES_MAX_SLOTS = 255;  // 4-cell slots for source / probe cells outside the ring frames (per CTA)
constexpr int RES_MAX_CELLS = 256;  // probes per CTA; sources inside ring frames per CTA
constexpr int RES_LW = 8, RES_RW = 12, RES_ZW = RES_LW + RES_RW;  // left | right ring frame columns of one row

// shared-memory floats of one CTA (see the carve-up at the top of the kernel)
__host__ __device__ constexpr size_t resident_smem_floats(int MR, int RES_NW) {
    return (size_t)2 * MR * RES_NW * RES_TW                   // coefficient maps
           + RES_TW                                           // dt/(mu*dx) of the row above the band
           + 4 * RES_NW * RES_TW                              // Ez rows exchanged between warps: 2 parities x (first, last)
           + 4 * RES_TW                                       // Ez rows from the neighbour CTAs: 2 parities x (below, above)
           + 2 * (MR * RES_NW * RES_ZW + 12 * RES_TW)         // S0 and S1 frames: LR, T, B
           + 12 * RES_TW                                      // T2, B2: finished top / bottom rows
           + 2 * 256 * 4                                      // slot frame + per-slot source waveforms
           + 2 * RES_MAX_CELLS * 2                            // probe list, ring-source list
           + MR * RES_NW * (RES_TW / 4) / 4;                  // slot table (bytes)
}

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t cluster_nctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cta address -> shared::cluster address of the same variable in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t map_to_rank(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
// 16 bytes straight into another CTA's shared memory through the async proxy; the store itself signals the
// destination CTA's mbarrier (complete_tx), so no fence and no separate arrive are needed.
__device__ __forceinline__ void st_async4(uint32_t remote_addr, const float* a, uint32_t remote_bar) {
    asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(remote_addr),
                 "f"(a[0]), "f"(a[1]), "f"(a[2]), "f"(a[3]), "r"(remote_bar)
                 : "memory");
}

// A value the compiler must keep in a register: it cannot see through the empty asm, so it does not re-derive
// thread-constant shared-memory offsets from %tid in every step of the time loop (that was ~1/4 of the loop).
__device__ __forceinline__ int keep(int v) {
    asm volatile("" : "+r"(v));
    return v;
}

// p.k = number of leapfrog steps of this launch; p.CW / p.CH = rows of the first / of every other CTA of a cluster
// (multiples of MR); gridDim.x = batch * cluster size.
// MR rows per thread, NW warps per CTA: a CTA holds a band of at most MR * NW rows.
template <int MR, int NW>
__global__ void __launch_bounds__(NW * 32, 1) grid_resident_kernel(const PassParams<float> p) {
    static_assert(MR >= 1 && MR <= 4, "the S2 lane mapping (one lane per ring cell, 8 lanes per row) covers at most 4 rows per warp");
    static_assert(resident_smem_floats(MR, NW) * 4 + 256 <= 232448, "band does not fit shared memory");
    static_assert(NW * 32 >= RES_MAX_CELLS, "one thread per probe / ring source");
    constexpr int TW = RES_TW, TH = MR * NW, NT = NW * 32, LW = RES_LW, ZW = RES_ZW;
    constexpr unsigned FULL = 0xffffffffu;
    constexpr uint32_t ROW_BYTES = TW * sizeof(float);  // one exchanged row: every lane sends 2 x 16 bytes
    // ring frames, addressed as float offsets from F: [LR0 | T0 | B0] = S0, the same again = S1, then T2, B2, slots
    constexpr int oLR = 0, oT = TH * ZW, oB = oT + 6 * TW, DELTA = oB + 6 * TW;
    constexpr int oT2 = 2 * DELTA, oB2 = oT2 + 6 * TW, oSlot = oB2 + 6 * TW;
    extern __shared__ __align__(16) unsigned char smem_res[];
    float* sCe = reinterpret_cast<float*>(smem_res);  // [TH][TW] dt/(eps*dx), zero where Ez is not updated
    float* sCh = sCe + TH * TW;                       // [TH][TW] dt/(mu*dx), zero where H is not updated
    float* sChA = sCh + TH * TW;                      // [TW] dt/(mu*dx) of the row above the band (zero for the first CTA)
    float* sEx = sChA + TW;                           // [2 parities][first | last][NW][TW] Ez rows published by the warps
    float* rEz = sEx + 4 * NW * TW;                   // [2 parities][from below | from above][TW] written by the neighbour CTAs
    float* F = rEz + 4 * TW;                          // ring frames (see the offsets above)
    int* slotW = reinterpret_cast<int*>(F + oSlot + 256 * 4);  // [256][4] waveform of a slot cell's source or -1
    int* plist = slotW + 256 * 4;                     // [RES_MAX_CELLS][2] probes of this band: frame offset, trace column
    int* rlist = plist + RES_MAX_CELLS * 2;           // [RES_MAX_CELLS][2] sources inside ring frames: offset, wave*amp_steps
    unsigned char* slot_tbl = reinterpret_cast<unsigned char*>(rlist + RES_MAX_CELLS * 2);  // [TH][TW/4] slot or 0xFF
    __shared__ __align__(8) uint64_t barB[2], barA[2];  // "the row from below / above of this parity has landed"
    __shared__ int s_counts[2];  // probes, ring sources of this band

    const int tid = threadIdx.x, w = tid >> 5, l = tid & 31;
    const int crank = (int)cluster_ctarank(), csize = (int)cluster_nctarank();
    const int b = blockIdx.x / csize;
    const int R = p.Rg, C = p.C, n_steps = p.k;
    // bands: p.CW rows for the first CTA, p.CH for the others (the last one takes what is left).  The first and
    // last CTA also run the top / bottom boundary pass, so the host gives them fewer rows.
    const int row_lo = crank == 0 ? 0 : p.CW + (crank - 1) * p.CH;
    const int nrows = crank == 0 ? min(p.CW, R) : min(p.CH, R - row_lo);
    const bool isTop = crank == 0, isBot = crank == csize - 1;
    const bool has_above = !isTop, has_below = !isBot;
    const int wl = (nrows - 1) / MR;  // the warp that holds the band's last row
    const int li0 = w * MR;
    const int cR0 = ((C - 6) >> 2) << 2;  // first column of the right ring frame (4-aligned, >= 8 for C >= 16)
    const int cg[2] = {4 * l, 128 + 4 * l};
    const float coef = p.mur[b];
    const long long gbase = (long long)b * p.grid_stride;

    // ---- load: fields -> registers, coefficient maps -> shared memory (masked) ---------------------
    float e[MR][2][4], hx[MR][2][4], hy[MR][2][4];
#pragma unroll
    for (int r = 0; r < MR; ++r) {
        const int lr = li0 + r, gi = row_lo + lr;
        const bool hrow = gi <= R - 2, erow = gi >= 1 && gi <= R - 2;
#pragma unroll
        for (int g = 0; g < 2; ++g) {
            float ce4[4], ch4[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) e[r][g][q] = hx[r][g][q] = hy[r][g][q] = ce4[q] = ch4[q] = 0.0f;
            if (lr < nrows && cg[g] < p.pitch) {
                const long long o = gbase + (long long)gi * p.pitch + cg[g];
                ldg4(p.in[0] + o, e[r][g]);
                ldg4(p.in[1] + o, hx[r][g]);
                ldg4(p.in[2] + o, hy[r][g]);
                ldg4(p.ce + o, ce4);
                ldg4(p.ch + o, ch4);
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int gj = cg[g] + q;
                if (!(hrow && gj <= C - 2)) ch4[q] = 0.0f;             // main.py:70,74: rows 0..R-2, cols 0..C-2
                if (!(erow && gj >= 1 && gj <= C - 2)) ce4[q] = 0.0f;  // main.py:27: rows 1..R-2, cols 1..C-2
            }
            store4(sCe + lr * TW + cg[g], ce4);
            store4(sCh + lr * TW + cg[g], ch4);
        }
    }

    // Redundant copy of the Hx row just above my rows (owned by the warp / CTA above): every warp advances it
    // itself from the Ez row it receives, so the Ez update needs no second exchange and no second barrier.
    float hxa[2][4];
    {
        const int ga = row_lo + li0 - 1;  // global row above my first row
#pragma unroll
        for (int g = 0; g < 2; ++g) {
            float cha[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) hxa[g][q] = cha[q] = 0.0f;
            if (ga >= 0 && li0 < nrows && cg[g] < p.pitch) {
                const long long o = gbase + (long long)ga * p.pitch + cg[g];
                ldg4(p.in[1] + o, hxa[g]);
                if (w == 0) ldg4(p.ch + o, cha);
            }
            if (w == 0) {
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (cg[g] + q > C - 2) cha[q] = 0.0f;
                store4(sChA + cg[g], cha);
            }
        }
    }
    const int qA = keep(w == 0 ? TH * TW + cg[0] : (li0 - 1) * TW + cg[0]);  // dt/(mu*dx) of that row, offset from sCh
    const int qC = keep(li0 * TW + cg[0]);                                   // my first row in sCe / sCh
    const int qX = keep(w * TW + cg[0]);                                     // my slot in a row-exchange block

    // ---- set-up: barriers; where the source / probe cells of this band live in shared memory --------
    for (int i = tid; i < TH * (TW / 4) / 4; i += NT) reinterpret_cast<unsigned*>(slot_tbl)[i] = 0xffffffffu;
    for (int i = tid; i < 256 * 4; i += NT) slotW[i] = -1;
    __syncthreads();
    // offset (from F) of the finished value of cell (gi, gj) of this band once a step's boundary stages are done
    auto ring_off = [&](int gi, int gj) -> int {
        const int lr = gi - row_lo;
        if (isTop && gi <= 5) return oT2 + gi * TW + gj;
        if (isBot && gi >= R - 6) return oB2 + (gi - (R - 6)) * TW + gj;
        if (gj < LW) return DELTA + oLR + lr * ZW + gj;
        if (gj >= cR0) return DELTA + oLR + lr * ZW + LW + (gj - cR0);
        return -1;  // not in a ring frame: the cell gets a slot
    };
    if (tid == 0) {
        mbar_init(&barB[0], 1), mbar_init(&barB[1], 1), mbar_init(&barA[0], 1), mbar_init(&barA[1], 1);
        fence_mbar_init();
        int n_slot = 0, n_prb = 0, n_rsrc = 0;
        auto slot_of = [&](int lr, int col) -> int {
            unsigned char& t = slot_tbl[lr * (TW / 4) + (col >> 2)];
            if (t == 0xFF) {
                if (n_slot >= RES_MAX_SLOTS) __trap();  // the host checks eligibility; never overrun the frame
                t = (unsigned char)n_slot++;
            }
            return t;
        };
        const int s_lo = p.src_range ? p.src_range[b] : 0, s_hi = p.src_range ? p.src_range[b + 1] : 0;
        for (int q = s_lo; q < s_hi; ++q) {
            const Cell c = p.src[q];
            if (c.row < row_lo || c.row >= row_lo + nrows) continue;
            const int off = ring_off(c.row, c.col);
            if (off >= 0) {
                if (n_rsrc >= RES_MAX_CELLS) __trap();
                rlist[2 * n_rsrc] = off, rlist[2 * n_rsrc + 1] = c.wave * p.amp_steps, ++n_rsrc;
            } else {
                slotW[slot_of(c.row - row_lo, c.col) * 4 + (c.col & 3)] = c.wave;
            }
        }
        const int p_lo = p.probe_range ? p.probe_range[b] : 0, p_hi = p.probe_range ? p.probe_range[b + 1] : 0;
        for (int q = p_lo; q < p_hi; ++q) {
            const Cell c = p.probes[q];
            if (c.row < row_lo || c.row >= row_lo + nrows) continue;
            int off = ring_off(c.row, c.col);
            if (off < 0) off = oSlot + slot_of(c.row - row_lo, c.col) * 4 + (c.col & 3);
            if (n_prb >= RES_MAX_CELLS) __trap();
            plist[2 * n_prb] = off, plist[2 * n_prb + 1] = q, ++n_prb;
        }
        s_counts[0] = n_prb, s_counts[1] = n_rsrc;
    }
    __syncthreads();
    const int n_prb = s_counts[0], n_rsrc = s_counts[1];
    unsigned spmask = 0;  // bit r*2+g: my group (r, g) holds a source or probe cell and has a slot
#pragma unroll
    for (int r = 0; r < MR; ++r)
#pragma unroll
        for (int g = 0; g < 2; ++g)
            if (slot_tbl[(li0 + r) * (TW / 4) + (cg[g] >> 2)] != 0xFF) spmask |= 1u << (r * 2 + g);
    // my groups' columns inside the LR frame of a row (-1: not a ring group)
    const int zo[2] = {cg[0] < LW ? cg[0] : (cg[0] >= cR0 && cg[0] < cR0 + RES_RW ? LW + cg[0] - cR0 : -1),
                       cg[1] >= cR0 && cg[1] < cR0 + RES_RW ? LW + cg[1] - cR0 : -1};
    const bool z0 = zo[0] >= 0, z1 = zo[1] >= 0;
    const int zq0 = keep(oLR + li0 * ZW + zo[0]);  // S0 slot of my group 0 in the LR frame of my first row (offset from F)
    const int zq1 = keep(oLR + li0 * ZW + zo[1]);  // (S1 is DELTA further; never dereferenced when !z0 / !z1)
    float* const zp0 = F + zq0;
    float* const zp1 = F + zq1;
    // do my rows reach the top / bottom ring rows?
    const bool warp_tb = (isTop && row_lo + li0 <= 5) || (isBot && row_lo + li0 + MR - 1 >= R - 6 && row_lo + li0 < R);
    // S2 (Mur left/right) of my warp's rows: lane -> (row l>>3, ring cell l&7).  Rows of the top / bottom ring
    // are updated in their T / B frame, the others in the LR frame; S0 is always DELTA before S1.
    const int s2row = li0 + (l >> 3), s2k = l & 7, s2gi = row_lo + s2row;
    const bool s2act = (l >> 3) < MR && s2k < RING && s2row < nrows && s2gi >= 1 && s2gi <= R - 2;
    int s2l, s2r;  // S1 of cell (row, k) / (row, C-1-k), as offsets from F
    if (isTop && s2gi <= 5) {
        s2l = DELTA + oT + s2gi * TW + s2k, s2r = DELTA + oT + s2gi * TW + (C - 1 - s2k);
    } else if (isBot && s2gi >= R - 6) {
        s2l = DELTA + oB + (s2gi - (R - 6)) * TW + s2k, s2r = DELTA + oB + (s2gi - (R - 6)) * TW + (C - 1 - s2k);
    } else {
        s2l = DELTA + oLR + s2row * ZW + s2k, s2r = DELTA + oLR + s2row * ZW + LW + (C - 1 - s2k - cR0);
    }
    float* const s2L = F + keep(s2l);
    float* const s2R = F + keep(s2r);
    // registers -> S0 / S1 frames (delta = 0 / DELTA)
    auto park = [&](int delta) {
        if (FDTD2D_RES_DIAG >= 2) return;
#pragma unroll
        for (int r = 0; r < MR; ++r) {
            if (z0) store4(zp0 + delta + r * ZW, e[r][0]);
            if (z1) store4(zp1 + delta + r * ZW, e[r][1]);
        }
        if (warp_tb) {
#pragma unroll
            for (int r = 0; r < MR; ++r) {
                const int gi = row_lo + li0 + r;
                if (isTop && gi <= 5) {
                    store4(F + delta + oT + gi * TW + cg[0], e[r][0]);
                    store4(F + delta + oT + gi * TW + cg[1], e[r][1]);
                } else if (isBot && gi >= R - 6 && gi < R) {
                    store4(F + delta + oB + (gi - (R - 6)) * TW + cg[0], e[r][0]);
                    store4(F + delta + oB + (gi - (R - 6)) * TW + cg[1], e[r][1]);
                }
            }
        }
    };
    // Finished value of every cell of the 6 top (or bottom) rows from their frames: S0 and S2 (S1 with the Mur
    // left/right update already applied in place by the rows' owners); S3 and S4 of SURVEY Appendix A evaluated
    // per cell.  `i` is the depth from the edge (0 = edge row), `top` picks the frame.
    auto tb_pass = [&](const bool top) {
        const float* A0 = F + (top ? oT : oB);
        const float* A1 = A0 + DELTA;
        float* A2 = F + (top ? oT2 : oB2);
        auto fr = [&](int i) { return (top ? i : 5 - i) * TW; };
        auto S3v = [&](int i, int j) -> float {  // main.py:43-51, rows 0..4 from the edge, columns 1..C-2
            const float s2 = A1[fr(i) + j];
            if (i <= 4 && j >= 1 && j <= C - 2) {
                const float a = A0[fr(i + 1) + j], b2 = A1[fr(i + 1) + j], c0 = A0[fr(i) + j];
                return add_rn(a, mul_rn(coef, sub_rn(b2, c0)));
            }
            return s2;
        };
        static_assert((6 * TW) % NT == 0, "the six rows split evenly over the threads");
        if constexpr (NT % TW == 0) {  // a thread keeps its column: the corner test is hoisted out of the cell loop
            const int j = tid & (TW - 1);
            const bool corner_col = j <= 4 || (j >= C - 5 && j <= C - 1);
            const int jn = j <= 4 ? j + 1 : j - 1;
#pragma unroll
            for (int t = 0; t < 6 * TW / NT; ++t) {
                const int i = tid / TW + t * (NT / TW);
                float v = S3v(i, j);
                if (i <= 4 && corner_col)  // main.py:54-61: every read is of a not-yet-processed cell
                    v = mul_rn(add_rn(S3v(i, jn), S3v(i + 1, j)), 0.5f);  // == sum / 2 exactly
                A2[fr(i) + j] = v;
            }
        } else {
#pragma unroll
            for (int t = 0; t < 6 * TW / NT; ++t) {
                const int c = tid + t * NT, i = c / TW, j = c & (TW - 1);
                float v = S3v(i, j);
                if (i <= 4 && (j <= 4 || (j >= C - 5 && j <= C - 1)))
                    v = mul_rn(add_rn(S3v(i, j <= 4 ? j + 1 : j - 1), S3v(i + 1, j)), 0.5f);
                A2[fr(i) + j] = v;
            }
        }
    };
    auto sample_probes = [&](long long step) {
        if (tid < n_prb && step < p.trace_cap) p.trace[step * p.n_probe + plist[2 * tid + 1]] = F[plist[2 * tid]];
    };
    // addresses in the neighbours' shared memory
    const uint32_t up_rank = has_above ? crank - 1 : crank, dn_rank = has_below ? crank + 1 : crank;
    const uint32_t up_rEz = map_to_rank(smem_u32(rEz), up_rank), up_barB = map_to_rank(smem_u32(&barB[0]), up_rank);
    const uint32_t dn_rEz = map_to_rank(smem_u32(rEz), dn_rank), dn_barA = map_to_rank(smem_u32(&barA[0]), dn_rank);
    const bool edge_dn = (w == wl) && has_below;  // my last row needs the Ez row of the CTA below
    const bool edge_up = (w == 0) && has_above;   // my first row needs the Hx row of the CTA above
    const bool cta_tb = FDTD2D_RES_DIAG >= 1 ? false : (isTop || isBot);
    const bool warp_on = li0 < nrows;  // warps past the end of the band only help with the top / bottom pass
    {  // rows nobody publishes are still read as a neighbour row by the last active warp: keep them finite
        const float z[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
        for (int i = 0; i < 4; ++i) store4(sEx + (i * NW + w) * TW + cg[0], z), store4(sEx + (i * NW + w) * TW + cg[1], z);
    }
    // one row of the H half-step (main.py:69-74); dn = the Ez row below it
    auto h_row = [&](const int r, const float (&dn)[2][4]) {
        float c[2][4];
        load4(sCh + qC + r * TW, c[0]);
        load4(sCh + qC + r * TW + 128, c[1]);
        const float ra = __shfl_sync(FULL, e[r][0][0], (l + 1) & 31);
        const float rb = __shfl_sync(FULL, e[r][1][0], (l + 1) & 31);
        const float right3[2] = {l == 31 ? rb : ra, rb};
#pragma unroll
        for (int g = 0; g < 2; ++g)
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float right = (q < 3) ? e[r][g][q < 3 ? q + 1 : 3] : right3[g];
                hx[r][g][q] = sub_rn(hx[r][g][q], mul_rn(c[g][q], sub_rn(dn[g][q], e[r][g][q])));
                hy[r][g][q] = add_rn(hy[r][g][q], mul_rn(c[g][q], sub_rn(right, e[r][g][q])));
            }
    };
    // one row of the interior Ez update (main.py:21-27); up = the Hx row above it
    auto e_row = [&](const int r, const float (&up)[2][4]) {
        float c[2][4];
        load4(sCe + qC + r * TW, c[0]);
        load4(sCe + qC + r * TW + 128, c[1]);
        const float la = __shfl_sync(FULL, hy[r][0][3], (l + 31) & 31);
        const float lb = __shfl_sync(FULL, hy[r][1][3], (l + 31) & 31);
        const float left0[2] = {la, l == 0 ? la : lb};
#pragma unroll
        for (int g = 0; g < 2; ++g)
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float left = (q > 0) ? hy[r][g][q > 0 ? q - 1 : 0] : left0[g];
                const float curl = sub_rn(sub_rn(hy[r][g][q], left), sub_rn(hx[r][g][q], up[g][q]));
                e[r][g][q] = add_rn(e[r][g][q], mul_rn(curl, c[g][q]));
            }
    };
    cluster_sync_all();  // every CTA's barriers are initialised before anyone signals them

#ifdef FDTD2D_RES_TIMING
    long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tprev = clock64();
#define RES_STAMP(i) { const long long tn = clock64(); tacc[i] += tn - tprev; tprev = tn; }
#else
#define RES_STAMP(i)
#endif
#pragma unroll 1
    for (int s = 0; s < n_steps; ++s) {
        const int par = s & 1;
        const uint32_t ph = (uint32_t)(s >> 1) & 1u;
        const long long step = p.step0 + s;
        float* const xF = sEx + par * (2 * NW * TW) + qX;  // my slot (group 0) for the first row; the last row is NW rows further
        if (warp_on) {
            // ---- publish my first and last Ez rows; the band's first / last row also go to the neighbour CTAs ----
            store4(xF, e[0][0]);
            store4(xF + 128, e[0][1]);
            store4(xF + NW * TW, e[MR - 1][0]);
            store4(xF + NW * TW + 128, e[MR - 1][1]);
            if (edge_up) {  // -> "row from below" of the CTA above
                st_async4(up_rEz + (uint32_t)(par * 2 * TW + cg[0]) * 4u, e[0][0], up_barB + 8u * par);
                st_async4(up_rEz + (uint32_t)(par * 2 * TW + cg[1]) * 4u, e[0][1], up_barB + 8u * par);
            }
            if (edge_dn) {  // -> "row from above" of the CTA below
                st_async4(dn_rEz + (uint32_t)((par * 2 + 1) * TW + cg[0]) * 4u, e[MR - 1][0], dn_barA + 8u * par);
                st_async4(dn_rEz + (uint32_t)((par * 2 + 1) * TW + cg[1]) * 4u, e[MR - 1][1], dn_barA + 8u * par);
            }
            park(0);  // S0: Ez is not changed by the H half-step
        }
        RES_STAMP(0)
        __syncthreads();  // the only CTA-wide barrier of a step away from the top / bottom ring
        RES_STAMP(1)
        if (s > 0) sample_probes(step - 1);  // the previous step's frames are intact until the next park
        if (warp_on) {
            // ---- H half-step; the rows that need a neighbour CTA's Ez row come last ---------------------------
#pragma unroll
            for (int r = 0; r + 1 < MR; ++r) h_row(r, e[r + 1 < MR ? r + 1 : r]);
            RES_STAMP(2)
            {
                float dn[2][4];
                const float* belowp = xF + (w + 1 < NW ? TW : 0);  // first row of the warp below
                if (edge_dn) {
                    if (l == 0) mbar_expect_tx(&barB[par], ROW_BYTES);
                    mbar_wait(&barB[par], ph);
                    belowp = rEz + par * 2 * TW + cg[0];
                }
                load4(belowp, dn[0]);
                load4(belowp + 128, dn[1]);
                h_row(MR - 1, dn);
            }
            {   // my copy of the Hx row above my rows (main.py:69-70 for that row)
                float up[2][4], c[2][4];
                const float* abovep = xF + NW * TW - (w > 0 ? TW : 0);  // last row of the warp above
                if (edge_up) {
                    if (l == 0) mbar_expect_tx(&barA[par], ROW_BYTES);
                    mbar_wait(&barA[par], ph);
                    abovep = rEz + (par * 2 + 1) * TW + cg[0];
                }
                load4(abovep, up[0]);
                load4(abovep + 128, up[1]);
                load4(sCh + qA, c[0]);
                load4(sCh + qA + 128, c[1]);
#pragma unroll
                for (int g = 0; g < 2; ++g)
#pragma unroll
                    for (int q = 0; q < 4; ++q) hxa[g][q] = sub_rn(hxa[g][q], mul_rn(c[g][q], sub_rn(e[0][g][q], up[g][q])));
            }
            RES_STAMP(3)
            // ---- interior Ez update (no barrier: every Hx row it reads is in my registers) ---------------------
#pragma unroll
            for (int r = MR - 1; r > 0; --r) e_row(r, hx[r > 0 ? r - 1 : 0]);
            e_row(0, hxa);
            RES_STAMP(4)
            // ---- S1 -> ring frames; source / probe cells outside the frames go through their slot -------------
            park(DELTA);
            if (spmask) {
#pragma unroll
                for (int r = 0; r < MR; ++r)
#pragma unroll
                    for (int g = 0; g < 2; ++g)
                        if (spmask >> (r * 2 + g) & 1u) {
                            const int sl = slot_tbl[(li0 + r) * (TW / 4) + (cg[g] >> 2)];
                            float* f = F + oSlot + sl * 4;
                            store4(f, e[r][g]);
                            if (step < p.amp_steps)
                                for (int q = 0; q < 4; ++q) {  // source add (fdtd.py:34): float64 sum, then cast
                                    const int wv = slotW[sl * 4 + q];
                                    if (wv >= 0) f[q] = add_source(f[q], p.amp[(long long)wv * p.amp_steps + step]);
                                }
                            load4(f, e[r][g]);
                        }
            }
            __syncwarp();
            // ---- S2: Mur left/right (main.py:33-41) of my warp's own rows, one lane per ring cell; all cells are
            // read before any is written (the reference's k order reads column k+1 before overwriting it) ------
            float vl = 0.0f, vr = 0.0f;
            if (s2act && FDTD2D_RES_DIAG < 2) {
                vl = add_rn(s2L[1 - DELTA], mul_rn(coef, sub_rn(s2L[1], s2L[-DELTA])));
                vr = add_rn(s2R[-1 - DELTA], mul_rn(coef, sub_rn(s2R[-1], s2R[-DELTA])));
            }
            __syncwarp();
            if (s2act && FDTD2D_RES_DIAG < 2) *s2L = vl, *s2R = vr;
        }
        RES_STAMP(5)
        // ---- top / bottom rows: S3, S4 in one pass over their frames (all threads of the CTA) ----------------
        if (cta_tb) {
            __syncthreads();
            RES_STAMP(6)
            if (isTop) tb_pass(true);
            if (isBot) tb_pass(false);
        }
        if (n_rsrc) {  // sources inside a ring frame are added once the frame is finished
            __syncthreads();
            if (tid < n_rsrc && step < p.amp_steps) {
                float* f = F + rlist[2 * tid];
                *f = add_source(*f, p.amp[(long long)rlist[2 * tid + 1] + step]);
            }
        }
        if (cta_tb || n_rsrc)
            __syncthreads();
        else
            __syncwarp();
        RES_STAMP(7)
        // ---- finished ring -> registers ---------------------------------------------------------------------
        if (warp_on && FDTD2D_RES_DIAG < 2) {
#pragma unroll
            for (int r = 0; r < MR; ++r) {
                if (z0) load4(zp0 + DELTA + r * ZW, e[r][0]);
                if (z1) load4(zp1 + DELTA + r * ZW, e[r][1]);
            }
            if (warp_tb) {
#pragma unroll
                for (int r = 0; r < MR; ++r) {
                    const int gi = row_lo + li0 + r;
                    if (isTop && gi <= 5) {
                        load4(F + oT2 + gi * TW + cg[0], e[r][0]);
                        load4(F + oT2 + gi * TW + cg[1], e[r][1]);
                    } else if (isBot && gi >= R - 6 && gi < R) {
                        load4(F + oB2 + (gi - (R - 6)) * TW + cg[0], e[r][0]);
                        load4(F + oB2 + (gi - (R - 6)) * TW + cg[1], e[r][1]);
                    }
                }
            }
        }
    }
#ifdef FDTD2D_RES_TIMING
    if (l == 0 && b == 0)
        for (int i = 0; i < 8; ++i) p.trace[(crank * NW + w) * 8 + i] = (float)tacc[i] / (float)n_steps;
    if (false)
#endif
    if (n_steps > 0 && n_prb) {
        __syncthreads();
        sample_probes(p.step0 + n_steps - 1);
    }

    // ---- store the fields -----------------------------------------------------------------------------
#pragma unroll
    for (int r = 0; r < MR; ++r) {
        const int lr = li0 + r, gi = row_lo + lr;
#pragma unroll
        for (int g = 0; g < 2; ++g)
            if (lr < nrows && cg[g] < p.pitch) {
                const long long o = gbase + (long long)gi * p.pitch + cg[g];
                store4(p.out[0] + o, e[r][g]);
                store4(p.out[1] + o, hx[r][g]);
                store4(p.out[2] + o, hy[r][g]);
            }
    }
    cluster_sync_all();  // no CTA may exit while a neighbour can still write into its shared memory
}

}  // namespace fdtd2d
