// Cluster-resident kernel, packed form (fp32): the successor of grid_resident.cuh for the small independent grids of the
// batched mode (BASELINE configs[4]: 1024 x 256^2) and the reference demo grid (fdtd.py:14-19).  Same idea -- one
// thread-block CLUSTER holds one grid for ALL n leapfrog steps of a call, HBM is touched once to load and once to store --
// rebuilt around what limited the first kernel (profiles/r1_resident_cfg5_full.summary.txt and the FDTD2D_RES_DIAG
// timings): 550 instructions per warp and step of which 268 were arithmetic, 22 shared-memory transactions of 512 bytes
// per warp and step (the shared-memory pipe as busy as the FP pipe), and three CTA-wide barriers per step in the first /
// last CTA of a cluster for the top / bottom ring.
//
//   * A warp owns SIX rows x 128 columns (6 x 4 cells per thread in registers), 16 warps per CTA as 8 row blocks x 2
//     column halves: a 48 x 256 band like before, but a warp exchanges one 512-byte row per three of its rows instead of
//     two 1024-byte rows per three, and the arithmetic is sm_100a's two-wide fp32 (add/sub/fma.rn.f32x2 on the column
//     pairs a 16-byte access delivers; strip_wave.cuh explains why the product is an fma with a -0 addend).
//     (A first version with 6 x 8 cells per thread and 8 warps was bit-exact and SLOWER than the old kernel: with two
//     warps per scheduler every serial stretch of a step runs at a fifth of an instruction per clock.)
//   * dt/(mu*dx) is a kernel argument when it is uniform (every material_init output): no map on chip, no loads.  H beyond
//     column C-2 (main.py:70,74) is then simply put back from the input when the fields are stored -- those values
//     never change and only feed Ez cells whose own update is masked.
//   * the top / bottom Mur ring needs NO CTA-wide barrier and no full-row frames: the six rows next to the edge belong to
//     ONE row block (the first of the first CTA; the last active one of the last CTA, whose band is a multiple of six
//     rows), so main.py:43-51 is elementwise between a warp's own registers -- it rides along the Ez update, which walks
//     those rows TOWARDS the edge so that Ez_prev of the next row is still in its register -- and only the 5 x 5 corners
//     (main.py:53-61) go through the warp's ring frame, one lane per corner cell.
// A step has ONE CTA-wide barrier in every CTA.
//
// Work layout.  A cluster of n <= 8 CTAs splits the rows into bands: rows [0, CW) for the first CTA, CH rows for the
// others (multiples of 6; the last band is what is left and the host makes it a multiple of 6 as well; CW is arbitrary).
// Warp w = 8 * wc + wr owns rows [6 wr, 6 wr + 6) of the band and columns [128 wc + 4 l, + 4) per lane l, as two register
// PAIRS per row and field.  The first band may end inside a row block: those warps park the row they receive from the CTA
// below in the registers of their first unused row (a ghost row: zero coefficients keep it inert), so the row loops stay
// static.
// Neighbours: columns inside a warp -> shuffles; across the two halves -> the boundary Ez column of every row block goes
// through shared memory with the rows, and the right half keeps a REDUNDANT copy of Hy column 127 (advanced locally), as
// every warp keeps a redundant copy of the Hx row above its rows: only Ez is exchanged, once per step.  Rows across warps
// -> first / last Ez row of every warp in shared memory (double buffered by step parity); rows across CTAs -> DSMEM
// st.async with mbarrier complete_tx.  Index ranges of the reference's slices are imposed by zero coefficients (maps in
// shared memory), a row count, and the restore at the end (uniform dt/(mu*dx)).
// Ring: every row block keeps the outermost columns of its rows in a small frame (S0 = Ez before the step, S1 = after the
// interior update); S2 (Mur left/right, main.py:33-41) one lane per ring cell in place; the finished frame is pulled
// back into registers.  Sources / probes outside the frames go through 4-cell slots exactly as in grid_resident.cuh;
// probes are sampled by the warp that owns the cell.  Grids whose right ring straddles column 128 (129..133 columns)
// stay with grid_resident.cuh.
// Results are bit-identical to every other kernel and to the reference (tests/test_gpu_resident.py).
#pragma once
#include <type_traits>

#include "grid_resident.cuh"
#include "strip_wave.cuh"

namespace fdtd2d {

#ifndef FDTD2D_RX_NR
#define FDTD2D_RX_NR 8  // row blocks per CTA (8: 48-row bands, 512 threads at 128 registers; 6: 36-row bands, 384 threads at 168)
#endif
constexpr int RX_MR = 6, RX_NR = FDTD2D_RX_NR, RX_NC = 2, RX_NW = RX_NR * RX_NC, RX_TH = RX_MR * RX_NR, RX_HW = 128;

__host__ __device__ constexpr size_t resident_x2_smem_floats() {
    return (size_t)2 * RX_TH * RES_TW      // coefficient maps
           + RES_TW                        // dt/(mu*dx) of the row above the band
           + 4 * RX_NR * RES_TW            // Ez rows exchanged between row blocks: 2 parities x (first, last)
           + 4 * RES_TW                    // Ez rows from the neighbour CTAs: 2 parities x (below, above)
           + 2 * RX_NR * 2 * 8             // Ez columns 127 | 128 of every row block: 2 parities x block x side x 6 (8) rows
           + 2 * RX_TH * RES_ZW            // S0 and S1 ring frames
           + 2 * 256 * 4                   // slot frame + per-slot source waveforms
           + RES_MAX_CELLS * 3             // probe list
           + RES_MAX_CELLS * 3             // ring-source list
           + RX_TH * (RES_TW / 4) / 4;     // slot table (bytes)
}

__device__ __forceinline__ void st_async22(uint32_t remote_addr, const u64* a, uint32_t remote_bar) {
    asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.v2.b64 [%0], {%1, %2}, [%3];" ::"r"(remote_addr), "l"(a[0]),
                 "l"(a[1]), "r"(remote_bar)
                 : "memory");
}

// UCH: dt/(mu*dx) is `ch_uniform` in every cell.  RL: the last real row of the row block in which the FIRST band ends
// ((CW - 1) % 6; 5 = the band ends with a block) -- a compile-time constant, because a run-time row index into the register
// arrays made ptxas keep a second copy of every Ez row (24 registers and 22 moves per step in EVERY warp).
// p.k = steps of this launch; p.CW / p.CH = rows of the first / of every other CTA of a cluster; gridDim.x = batch * cluster size.
template <bool UCH, int RL>
__global__ void __launch_bounds__(RX_NW * 32, 1) grid_resident_x2_kernel(const PassParams<float> p, const float ch_uniform, const u64 negzero) {
    constexpr int MR = RX_MR, NR = RX_NR, NW = RX_NW, TW = RES_TW, HW = RX_HW, TH = RX_TH, NT = NW * 32, LW = RES_LW, ZW = RES_ZW;
    static_assert(resident_x2_smem_floats() * 4 + 256 <= 232448, "band does not fit shared memory");
    static_assert(MR * RING <= 32 && MR >= RING + 1, "S2: one lane per ring cell of the warp's rows; the ring rows fit one warp");
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int oLR = 0, DELTA = TH * ZW, oSlot = 2 * DELTA;  // ring frames as float offsets from F: S0 | S1 | slots
    extern __shared__ __align__(16) unsigned char smem_res[];
    float* sCe = reinterpret_cast<float*>(smem_res);  // [TH][TW] dt/(eps*dx), zero where Ez is not updated
    float* sCh = sCe + TH * TW;                       // [TH][TW] dt/(mu*dx), zero where H is not updated (unused when UCH)
    float* sChA = sCh + TH * TW;                      // [TW] dt/(mu*dx) of the row above the band
    float* sEx = sChA + TW;                           // [2 parities][first | last][NR][TW]
    float* rEz = sEx + 4 * NR * TW;                   // [2 parities][from below | from above][TW]
    float* sCx = rEz + 4 * TW;                        // [2 parities][NR][column 127 | column 128][8]
    float* F = sCx + 2 * NR * 2 * 8;                  // ring frames
    int* slotW = reinterpret_cast<int*>(F + oSlot + 256 * 4);
    int* plist = slotW + 256 * 4;                     // [RES_MAX_CELLS][3] probes SORTED BY OWNER: frame offset, trace column, owner warp
    int* rlist = plist + RES_MAX_CELLS * 3;           // [RES_MAX_CELLS][3] sources inside ring frames: offset, wave*amp_steps, owner warp
    unsigned char* slot_tbl = reinterpret_cast<unsigned char*>(rlist + RES_MAX_CELLS * 3);  // [TH][TW/4] slot or 0xFF
    __shared__ __align__(8) uint64_t barB[2], barA[2];
    __shared__ int s_counts[2];
    __shared__ int s_pbeg[NW + 1];  // probes of warp w: plist[s_pbeg[w] .. s_pbeg[w + 1])

    const int tid = threadIdx.x, w = tid >> 5, l = tid & 31;
    const int wr = w % NR, wc = w / NR;  // (warps w, w+4, w+8, w+12 share a scheduler: two row blocks x both halves)
    const int crank = (int)cluster_ctarank(), csize = (int)cluster_nctarank();
    const int b = blockIdx.x / csize;
    const int R = p.Rg, C = p.C, n_steps = p.k;
    const int row_lo = crank == 0 ? 0 : p.CW + (crank - 1) * p.CH;
    const int nrows = crank == 0 ? min(p.CW, R) : min(p.CH, R - row_lo);
    const bool isTop = crank == 0, isBot = crank == csize - 1;
    const bool has_above = !isTop, has_below = !isBot;
    const int wl = (nrows - 1) / MR;  // the row block that holds the band's last row
    const int li0 = wr * MR;
    const int nact = max(0, min(MR, nrows - li0));                        // my real rows
    const int hrows = max(0, min(MR, min(nrows, R - 1 - row_lo) - li0));  // ... of which H is updated (main.py:70,74: rows 0..R-2)
    const int cR0 = ((C - 6) >> 2) << 2;                                  // first column of the right ring frame
    const int wcR = cR0 >= HW ? 1 : 0;                                    // the half that holds the right ring (the host excludes a straddle)
    const bool two_halves = p.pitch > HW;
    const int cg = HW * wc + 4 * l;
    const bool col_on = cg < p.pitch;
    const float coef = p.mur[b];
    const u64 coef2 = pack2(coef, coef);
    const u64 chu2 = pack2(ch_uniform, ch_uniform);
    const long long gbase = (long long)b * p.grid_stride;
    const uint32_t ROW_BYTES = (two_halves ? 2u : 1u) * HW * (uint32_t)sizeof(float);  // what a neighbour CTA sends per row

    // ---- load: fields -> register pairs, coefficient maps -> shared memory (masked) ------------------
    u64 e[MR][2], hx[MR][2], hy[MR][2];
#pragma unroll
    for (int r = 0; r < MR; ++r) {
        const int lr = li0 + r, gi = row_lo + lr;
        const bool hrow = gi <= R - 2, erow = gi >= 1 && gi <= R - 2;
        float e4[4], x4[4], y4[4], ce4[4], ch4[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) e4[q] = x4[q] = y4[q] = ce4[q] = ch4[q] = 0.0f;
        if (lr < nrows && col_on) {
            const long long o = gbase + (long long)gi * p.pitch + cg;
            ldg4(p.in[0] + o, e4);
            ldg4(p.in[1] + o, x4);
            ldg4(p.in[2] + o, y4);
            ldg4(p.ce + o, ce4);
            if (!UCH) ldg4(p.ch + o, ch4);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int gj = cg + q;
            if (!(hrow && gj <= C - 2)) ch4[q] = 0.0f;
            if (!(erow && gj >= 1 && gj <= C - 2)) ce4[q] = 0.0f;
        }
        e[r][0] = pack2(e4[0], e4[1]), e[r][1] = pack2(e4[2], e4[3]);
        hx[r][0] = pack2(x4[0], x4[1]), hx[r][1] = pack2(x4[2], x4[3]);
        hy[r][0] = pack2(y4[0], y4[1]), hy[r][1] = pack2(y4[2], y4[3]);
        store4(sCe + lr * TW + cg, ce4);
        if (!UCH) store4(sCh + lr * TW + cg, ch4);
    }
    const bool warp_on = li0 < nrows && HW * wc < p.pitch;
    // redundant copy of the Hx row just above my rows
    u64 hxa[2];
    const bool has_up_row = row_lo + li0 - 1 >= 0 && warp_on;
    {
        const int ga = row_lo + li0 - 1;
        float x4[4], cha[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) x4[q] = cha[q] = 0.0f;
        if (has_up_row && col_on) {
            const long long o = gbase + (long long)ga * p.pitch + cg;
            ldg4(p.in[1] + o, x4);
            if (!UCH && wr == 0) ldg4(p.ch + o, cha);
        }
        hxa[0] = pack2(x4[0], x4[1]), hxa[1] = pack2(x4[2], x4[3]);
        if (!UCH && wr == 0) {
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (cg + q > C - 2) cha[q] = 0.0f;
            store4(sChA + cg, cha);
        }
    }
    // the right half's redundant copy of Hy column 127 of its rows (lane 0 uses it), and that column's dt/(mu*dx)
    float hyL[MR], chL[MR];
#pragma unroll
    for (int r = 0; r < MR; ++r) {
        const int lr = li0 + r, gi = row_lo + lr;
        hyL[r] = chL[r] = 0.0f;
        if (wc == 1 && warp_on && lr < nrows && gi <= R - 2) {  // (Hy has R-1 rows)
            const long long o = gbase + (long long)gi * p.pitch + (HW - 1);
            hyL[r] = __ldg(p.in[2] + o);
            if (!UCH) chL[r] = __ldg(p.ch + o);
        }
    }
    const int qA = keep(wr == 0 ? TH * TW + cg : (li0 - 1) * TW + cg);
    const int qC = keep(li0 * TW + cg);
    const int qX = keep(wr * TW + cg);

    // ---- set-up: barriers; where the source / probe cells of this band live in shared memory --------
    for (int i = tid; i < TH * (TW / 4) / 4; i += NT) reinterpret_cast<unsigned*>(slot_tbl)[i] = 0xffffffffu;
    for (int i = tid; i < 256 * 4; i += NT) slotW[i] = -1;
    for (int i = tid; i < 2 * NR * 2 * 8; i += NT) sCx[i] = 0.0f;
    __syncthreads();
    // offset (from F) of the finished value of cell (gi, gj) of this band once a step's boundary stages are done
    auto ring_off = [&](int gi, int gj) -> int {
        const int lr = gi - row_lo;
        if (gj < LW) return DELTA + oLR + lr * ZW + gj;
        if (gj >= cR0) return DELTA + oLR + lr * ZW + LW + (gj - cR0);
        return -1;  // not in a ring frame: the cell gets a slot
    };
    auto owner_of = [&](int gi, int gj) -> int {  // the warp that finishes cell (gi, gj)
        const int half = gj < LW ? 0 : (gj >= cR0 ? wcR : gj / HW);
        return half * NR + (gi - row_lo) / MR;
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
                rlist[3 * n_rsrc] = off, rlist[3 * n_rsrc + 1] = c.wave * p.amp_steps, rlist[3 * n_rsrc + 2] = owner_of(c.row, c.col), ++n_rsrc;
            } else {
                slotW[slot_of(c.row - row_lo, c.col) * 4 + (c.col & 3)] = c.wave;
            }
        }
        const int p_lo = p.probe_range ? p.probe_range[b] : 0, p_hi = p.probe_range ? p.probe_range[b + 1] : 0;
        for (int i = 0; i <= NW; ++i) s_pbeg[i] = 0;
        for (int q = p_lo; q < p_hi; ++q) {  // counting sort by owner: every warp samples its own probes
            const Cell c = p.probes[q];
            if (c.row < row_lo || c.row >= row_lo + nrows) continue;
            s_pbeg[owner_of(c.row, c.col) + 1] += 1, ++n_prb;
        }
        if (n_prb > RES_MAX_CELLS) __trap();
        for (int i = 0; i < NW; ++i) s_pbeg[i + 1] += s_pbeg[i];
        for (int q = p_lo; q < p_hi; ++q) {
            const Cell c = p.probes[q];
            if (c.row < row_lo || c.row >= row_lo + nrows) continue;
            int off = ring_off(c.row, c.col);
            if (off < 0) off = oSlot + slot_of(c.row - row_lo, c.col) * 4 + (c.col & 3);
            const int own = owner_of(c.row, c.col), at = s_pbeg[own]++;
            plist[3 * at] = off, plist[3 * at + 1] = q, plist[3 * at + 2] = own;
        }
        for (int i = NW; i > 0; --i) s_pbeg[i] = s_pbeg[i - 1];  // (the fill moved every start to its end)
        s_pbeg[0] = 0;
        s_counts[0] = n_prb, s_counts[1] = n_rsrc;
    }
    __syncthreads();
    const int n_rsrc = s_counts[1];
    const int pbeg = s_pbeg[w], pn = s_pbeg[w + 1] - pbeg;
    unsigned spmask = 0;  // bit r: my group of row r holds a source or probe cell and has a slot
#pragma unroll
    for (int r = 0; r < MR; ++r)
        if (slot_tbl[(li0 + r) * (TW / 4) + (cg >> 2)] != 0xFF) spmask |= 1u << r;
    // my group's columns inside the ring frame of a row (-1: not a ring group)
    const int zo = cg < LW ? cg : (cg >= cR0 && cg < cR0 + RES_RW ? LW + cg - cR0 : -1);
    const bool zg = zo >= 0;
    float* const zp = F + keep(oLR + li0 * ZW + zo);  // S0 slot of my group in the frame of my first row (S1 is DELTA further)
    // 1: my six rows are the top ring rows 0..5, 2: the bottom ring rows R-6..R-1 (the host makes the last band a multiple
    // of six rows, and a band that holds both has at least twelve), 0: neither
    const int tbmode = FDTD2D_RES_DIAG >= 1 ? 0 : (isTop && wr == 0) ? 1 : (isBot && li0 + MR == nrows) ? 2 : 0;
    // S2 (Mur left/right) of my rows: lane -> (row l / 5, ring cell l % 5); the left cells by the left half, the right
    // cells by the half that holds them
    const int s2row = li0 + l / RING, s2k = l % RING, s2gi = row_lo + s2row;
    const bool s2act = l < MR * RING && s2row < nrows && s2gi >= 1 && s2gi <= R - 2;
    const bool s2l_on = s2act && wc == 0, s2r_on = s2act && wc == wcR;
    float* const s2L = F + keep(DELTA + oLR + s2row * ZW + s2k);                       // S1 of cell (row, k)
    float* const s2R = F + keep(DELTA + oLR + s2row * ZW + LW + (C - 1 - s2k - cR0));  // S1 of cell (row, C-1-k)
    // the corners (main.py:53-61), by the warps that hold the six rows: lane -> (depth d = l / 5 from the edge, j = l % 5
    // from the side); frame row of depth d is d (top) or 5 - d (bottom)
    const bool cact = tbmode != 0 && l < RING * RING;
    const bool cl_on = cact && wc == 0, cr_on = cact && wc == wcR;
    const int cd = l / RING, cj = l % RING;
    float* const cF = F + oLR + li0 * ZW;  // S0 frame of my first row

    auto park = [&](int delta) {
        if (FDTD2D_RES_DIAG >= 2) return;
        if (zg) {
#pragma unroll
            for (int r = 0; r < MR; ++r) store22(zp + delta + r * ZW, e[r]);
        }
    };
    // the probes of my cells, from the finished frames / slots
    auto sample_probes = [&](long long step) {
        __syncwarp();
        if (step < p.trace_cap)
            for (int i = l; i < pn; i += 32) p.trace[step * p.n_probe + plist[3 * (pbeg + i) + 1]] = F[plist[3 * (pbeg + i)]];
        __syncwarp();
    };
    // (the addresses in the neighbours' shared memory are formed where they are used: two instructions in two warps
    // instead of four registers in all of them)
    const bool edge_dn = (wr == wl) && has_below;  // my last real row needs the Ez row of the CTA below
    const bool edge_up = (wr == 0) && has_above;   // my first row needs the Hx row of the CTA above
    const bool ghost_dn = RL < MR - 1 && edge_dn && isTop;  // the row from below is parked in the registers of row RL + 1
    const bool xb = wc == 0 ? l == 31 : l == 0;    // my lane is next to the other column half
    const int xo = keep(wc == 0 ? 8 : 0);          // where the OTHER half's boundary column sits in a block's column buffer
    if (wc == 0) {  // rows nobody publishes are still read as a neighbour row: keep them finite
        const float z[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
        for (int i = 0; i < 4; ++i) store4(sEx + (i * NR + wr) * TW + cg, z), store4(sEx + (i * NR + wr) * TW + cg + HW, z);
    }
    // one row of the H half-step (main.py:69-74); dn = the Ez row below it, xe = Ez of this row in the other half's boundary
    // column (the right neighbour of lane 31 of the left half)
    auto h_row = [&](const int r, const u64 (&dn)[2], const float xe) {
        u64 c[2];
        if (UCH)
            c[0] = chu2, c[1] = chu2;
        else
            load22(sCh + qC + r * TW, c);
        const float ra = __shfl_down_sync(FULL, lo2(e[r][0]), 1);
        const float right3 = (xb && wc == 0) ? xe : ra;
        const u64 e0 = e[r][0], e1 = e[r][1];
        const u64 dx0 = pack2(sub_rn(hi2(e0), lo2(e0)), sub_rn(lo2(e1), hi2(e0)));
        const u64 dx1 = pack2(sub_rn(hi2(e1), lo2(e1)), sub_rn(right3, hi2(e1)));
        hx[r][0] = sub2(hx[r][0], mul2(c[0], sub2(dn[0], e0), negzero));
        hx[r][1] = sub2(hx[r][1], mul2(c[1], sub2(dn[1], e1), negzero));
        hy[r][0] = add2(hy[r][0], mul2(c[0], dx0, negzero));
        hy[r][1] = add2(hy[r][1], mul2(c[1], dx1, negzero));
        // the copy of Hy column 127 (main.py:73-74 for that column; lane 0 of the right half holds column 128; every other
        // lane computes something nobody reads -- cheaper than a branch)
        hyL[r] = add_rn(hyL[r], mul_rn(UCH ? ch_uniform : chL[r], sub_rn(lo2(e0), xe)));
    };
    // the interior Ez update of one row (main.py:21-27) into t; up = the Hx row above it
    auto e_row = [&](const int r, const u64 (&up)[2], u64 (&t)[2]) {
        u64 c[2];
        load22(sCe + qC + r * TW, c);
        const float la = __shfl_up_sync(FULL, hi2(hy[r][1]), 1);
        const float left0 = (xb && wc == 1) ? hyL[r] : la;
        const u64 y0 = hy[r][0], y1 = hy[r][1];
        const u64 dy0 = pack2(sub_rn(lo2(y0), left0), sub_rn(hi2(y0), lo2(y0)));
        const u64 dy1 = pack2(sub_rn(lo2(y1), hi2(y0)), sub_rn(hi2(y1), lo2(y1)));
        const u64 curl0 = sub2(dy0, sub2(hx[r][0], up[0]));
        const u64 curl1 = sub2(dy1, sub2(hx[r][1], up[1]));
        t[0] = add2(e[r][0], mul2(curl0, c[0], negzero));
        t[1] = add2(e[r][1], mul2(curl1, c[1], negzero));
    };
    // Ez update of the warps that hold the top (TOP) or bottom ring rows.  main.py:43-51 for row r (r = depth from the edge,
    // 0..4) is S0[r+1] + coef * (S1[r+1] - S0[r]): the rows are walked AWAY from the edge with the interior update of the next
    // row formed one row ahead (T), so row r's result goes straight into its registers while row r + 1 still holds Ez_prev.
    // (On the ring columns the reference uses S2 instead of S1: those cells are corner cells and are overwritten through the
    // frame.)  S1 of the ring groups is parked on the way (S2 and the corners need it); the edge row itself has no
    // interior update (its S1 is never read).
    auto e_rows_tb = [&](auto top_c) {
        constexpr bool TOP = decltype(top_c)::value;
        u64 t[2];
#pragma unroll
        for (int d = 0; d + 1 < MR; ++d) {
            const int r = TOP ? d : MR - 1 - d;          // the row that gets its final value
            const int rx = TOP ? d + 1 : MR - 2 - d;     // the next row from the edge: its interior update is formed now
            if (rx == 0)
                e_row(rx, hxa, t);
            else
                e_row(rx, hx[rx > 0 ? rx - 1 : 0], t);
            if (zg) store22(zp + DELTA + rx * ZW, t);
            e[r][0] = add2(e[rx][0], mul2(coef2, sub2(t[0], e[r][0]), negzero));
            e[r][1] = add2(e[rx][1], mul2(coef2, sub2(t[1], e[r][1]), negzero));
        }
        e[TOP ? MR - 1 : 0][0] = t[0], e[TOP ? MR - 1 : 0][1] = t[1];  // the sixth row: interior update only
    };
    cluster_sync_all();  // every CTA's barriers are initialised before anyone signals them

#ifdef FDTD2D_RES_TIMING
    long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tprev = clock64();
#define RX_STAMP(i) { const long long tn = clock64(); tacc[i] += tn - tprev; tprev = tn; }
#else
#define RX_STAMP(i)
#endif
    // The time loop, specialised by what the warp is: ROLE bit 0 = its first row needs the CTA above (UP), bit 1 = its last row
    // needs the CTA below (DN), bit 2 = ... and that row arrives as a ghost row (GHOST), bits 3-4 = it holds the top (1) /
    // bottom (2) ring rows, bit 5 = none of the listed combinations: every flag is read at run time (GENERIC).  The general
    // loop spent 320 of its 480 instructions per warp and step on flags, branches and re-derived addresses; a warp takes
    // ~12 cycles per instruction of its own chain, and the step is as long as the slowest warp's chain.
    auto run = [&](auto role_c) {
        constexpr unsigned ROLE = decltype(role_c)::value;
        constexpr bool GEN = (ROLE & 32u) != 0;
        const bool up = GEN ? edge_up : (ROLE & 1u) != 0;
        const bool dn = GEN ? edge_dn : (ROLE & 2u) != 0;
        const bool ghost = GEN ? ghost_dn : (ROLE & 4u) != 0;
        const int tb = GEN ? tbmode : (int)((ROLE >> 3) & 3u);
        const int hr = GEN ? hrows : (ghost ? RL + 1 : (tb == 2 ? MR - 1 : MR));  // rows whose H is updated
        const bool uprow = GEN ? has_up_row : tb != 1;                            // there is a row above my rows
        // the corner lanes' frame offsets (from cF) and index tests do not change from step to step: side 0 = left, 1 = right;
        // value A = after-S3 value of (depth d, column next to mine towards the middle), value B = of (depth d + 1, my column)
        int cA0[2] = {0, 0}, cA1[2] = {0, 0}, cB0[2] = {0, 0}, cB1[2] = {0, 0}, cW[2] = {0, 0};
        bool cAs3[2] = {false, false}, cBs3[2] = {false, false}, con[2] = {false, false};
        if (tb) {
            auto fro = [&](int d) { return (tb == 1 ? d : MR - 1 - d) * ZW; };
#pragma unroll
            for (int sd = 0; sd < 2; ++sd) {
                const int gmine = sd == 0 ? cj : C - 1 - cj, gnext = sd == 0 ? cj + 1 : C - 2 - cj;
                const int fmine = sd == 0 ? cj : LW + gmine - cR0, fnext = sd == 0 ? cj + 1 : LW + gnext - cR0;
                con[sd] = sd == 0 ? cl_on : cr_on;
                cA0[sd] = fro(cd) + fnext, cA1[sd] = fro(cd + 1) + fnext;                      // rows d, d + 1 in the next column
                cB0[sd] = fro(cd + 1) + fmine, cB1[sd] = fro(cd + 2 < MR ? cd + 2 : MR - 1) + fmine;  // rows d + 1, d + 2 in my column
                cAs3[sd] = gnext >= 1 && gnext <= C - 2;                                      // (d <= 4 always)
                cBs3[sd] = cd + 1 <= 4 && gmine >= 1 && gmine <= C - 2;
                cW[sd] = DELTA + fro(cd) + fmine;
            }
        }
#pragma unroll 1
        for (int s = 0; s < n_steps; ++s) {
            const int par = s & 1;
            const uint32_t ph = (uint32_t)(s >> 1) & 1u;
            float* const xF = sEx + par * (2 * NR * TW) + qX;
            float* const xC = sCx + (par * NR + wr) * 16;  // my row block's boundary columns: [0..7] column 127, [8..15] column 128
            // ---- publish my first and last Ez rows; the band's first / last row also go to the neighbour CTAs ----
            store22(xF, e[0]);
            store22(xF + NR * TW, e[MR - 1]);
            if (xb) {  // my Ez column next to the other half
#pragma unroll
                for (int r = 0; r < MR; ++r) xC[8 - xo + r] = wc == 0 ? hi2(e[r][1]) : lo2(e[r][0]);
            }
            if (up) st_async22(map_to_rank(smem_u32(rEz + par * 2 * TW + cg), crank - 1), e[0], map_to_rank(smem_u32(&barB[par]), crank - 1));
            if (dn) {
                const uint32_t a0 = map_to_rank(smem_u32(rEz + (par * 2 + 1) * TW + cg), crank + 1);
                const uint32_t bar = map_to_rank(smem_u32(&barA[par]), crank + 1);
                if (!ghost)
                    st_async22(a0, e[MR - 1], bar);
                else
                    st_async22(a0, e[RL], bar);
            }
            park(0);  // S0: Ez is not changed by the H half-step
            RX_STAMP(0)
            __syncthreads();  // the only CTA-wide barrier of a step
            RX_STAMP(1)
            // ---- H half-step; the row that needs a neighbour CTA's Ez row comes last ---------------------------
            const float* const xe = xC + xo;  // Ez of my rows in the other half's boundary column
            if (ghost) {  // (the last row block of the first CTA) the row from below becomes the Ez of my first unused row
                if (tid == (wl << 5)) mbar_expect_tx(&barB[par], ROW_BYTES);
                mbar_wait(&barB[par], ph);
                load22(rEz + par * 2 * TW + cg, e[RL + 1 < MR ? RL + 1 : MR - 1]);
            }
#pragma unroll
            for (int r = 0; r + 1 < MR; ++r)
                if (r < hr) h_row(r, e[r + 1 < MR ? r + 1 : r], xe[r]);
            RX_STAMP(2)
            if (MR - 1 < hr) {
                u64 dnr[2];
                const float* belowp = xF + (wr + 1 < NR ? TW : 0);  // first row of the row block below
                if (dn) {
                    if (tid == (wl << 5)) mbar_expect_tx(&barB[par], ROW_BYTES);
                    mbar_wait(&barB[par], ph);
                    belowp = rEz + par * 2 * TW + cg;
                }
                load22(belowp, dnr);
                h_row(MR - 1, dnr, xe[MR - 1]);
            }
            if (uprow) {  // my copy of the Hx row above my rows (main.py:69-70 for that row)
                u64 upr[2], c[2];
                const float* abovep = xF + NR * TW - (wr > 0 ? TW : 0);  // last row of the row block above
                if (up) {
                    if (tid == 0) mbar_expect_tx(&barA[par], ROW_BYTES);
                    mbar_wait(&barA[par], ph);
                    abovep = rEz + (par * 2 + 1) * TW + cg;
                }
                load22(abovep, upr);
                if (UCH)
                    c[0] = chu2, c[1] = chu2;
                else
                    load22(sCh + qA, c);
                hxa[0] = sub2(hxa[0], mul2(c[0], sub2(e[0][0], upr[0]), negzero));
                hxa[1] = sub2(hxa[1], mul2(c[1], sub2(e[0][1], upr[1]), negzero));
            }
            RX_STAMP(3)
            // ---- interior Ez update (no barrier: every Hx row it reads is in my registers); S1 -> ring frame ------
            if (tb == 1) {
                e_rows_tb(std::true_type{});
            } else if (tb == 2) {
                e_rows_tb(std::false_type{});
            } else {
#pragma unroll
                for (int r = MR - 1; r > 0; --r) e_row(r, hx[r > 0 ? r - 1 : 0], e[r]);
                e_row(0, hxa, e[0]);
                park(DELTA);
            }
            RX_STAMP(4)
            // ---- source / probe cells outside the frames: the source add in registers (fdtd.py:34: float64 sum, then cast),
            // the result also goes to the cell's slot, where the probes are sampled.  (Fetching the amplitudes a step ahead
            // with cp.async was measured: slower -- sixteen steps share a cache line.)
            if (spmask) {
                const long long step = p.step0 + s;
#pragma unroll
                for (int r = 0; r < MR; ++r)
                    if (spmask >> r & 1u) {
                        const int sl = slot_tbl[(li0 + r) * (TW / 4) + (cg >> 2)];
                        if (step < p.amp_steps) {
                            const int4 wv = *reinterpret_cast<const int4*>(slotW + sl * 4);
                            float v[4] = {lo2(e[r][0]), hi2(e[r][0]), lo2(e[r][1]), hi2(e[r][1])};
                            if (wv.x >= 0) v[0] = add_source(v[0], p.amp[(long long)wv.x * p.amp_steps + step]);
                            if (wv.y >= 0) v[1] = add_source(v[1], p.amp[(long long)wv.y * p.amp_steps + step]);
                            if (wv.z >= 0) v[2] = add_source(v[2], p.amp[(long long)wv.z * p.amp_steps + step]);
                            if (wv.w >= 0) v[3] = add_source(v[3], p.amp[(long long)wv.w * p.amp_steps + step]);
                            e[r][0] = pack2(v[0], v[1]), e[r][1] = pack2(v[2], v[3]);
                        }
                        store22(F + oSlot + sl * 4, e[r]);
                    }
            }
            __syncwarp();
            // ---- S2: Mur left/right (main.py:33-41) of my warp's own rows, one lane per ring cell; all cells are
            // read before any is written (the reference's k order reads column k+1 before overwriting it) ------
            float vl = 0.0f, vr = 0.0f;
            if (FDTD2D_RES_DIAG < 2) {
                if (s2l_on) vl = add_rn(s2L[1 - DELTA], mul_rn(coef, sub_rn(s2L[1], s2L[-DELTA])));
                if (s2r_on) vr = add_rn(s2R[-1 - DELTA], mul_rn(coef, sub_rn(s2R[-1], s2R[-DELTA])));
                __syncwarp();
                if (s2l_on) *s2L = vl;
                if (s2r_on) *s2R = vr;
            }
            RX_STAMP(5)
            if (tb) {
                // ---- corners: Jacobi over values after S3 (every read is of a not-yet-processed cell) --------------
                __syncwarp();
                float cv[2] = {0.0f, 0.0f};
#pragma unroll
                for (int sd = 0; sd < 2; ++sd)
                    if (con[sd]) {
                        // value after main.py:43-51 of a frame cell: S0[d+1] + coef * (S2[d+1] - S0[d]) where that applies, else S2
                        const float a = cAs3[sd] ? add_rn(cF[cA1[sd]], mul_rn(coef, sub_rn(cF[DELTA + cA1[sd]], cF[cA0[sd]]))) : cF[DELTA + cA0[sd]];
                        const float bq = cBs3[sd] ? add_rn(cF[cB1[sd]], mul_rn(coef, sub_rn(cF[DELTA + cB1[sd]], cF[cB0[sd]]))) : cF[DELTA + cB0[sd]];
                        cv[sd] = mul_rn(add_rn(a, bq), 0.5f);  // == sum / 2 exactly
                    }
                __syncwarp();
                // the rows S3 finished in registers -> frame (columns of the ring groups that are not corner cells)
                if (zg) {
#pragma unroll
                    for (int r = 0; r < MR; ++r)
                        if (tb == 1 ? r < MR - 1 : r > 0) store22(zp + DELTA + r * ZW, e[r]);
                }
                __syncwarp();
                if (con[0]) cF[cW[0]] = cv[0];
                if (con[1]) cF[cW[1]] = cv[1];
            }
            RX_STAMP(6)
            if (n_rsrc) {  // sources inside a ring frame are added once the frame is finished (by the cells' own warp)
                __syncwarp();
                const long long step = p.step0 + s;
                for (int i = l; i < n_rsrc; i += 32)
                    if (rlist[3 * i + 2] == w && step < p.amp_steps) {
                        float* f = F + rlist[3 * i];
                        *f = add_source(*f, p.amp[(long long)rlist[3 * i + 1] + step]);
                    }
            }
            __syncwarp();
            // ---- finished ring -> registers -------------------------------------------------------------------
            if (FDTD2D_RES_DIAG < 2 && zg) {
#pragma unroll
                for (int r = 0; r < MR; ++r) load22(zp + DELTA + r * ZW, e[r]);
            }
            if (pn) sample_probes(p.step0 + s);
            RX_STAMP(7)
        }
    };
    if (!warp_on) {  // warps without rows only keep the barrier count
#pragma unroll 1
        for (int s = 0; s < n_steps; ++s) __syncthreads();
    } else {
        // the combinations that occur in a cluster of full bands; anything else (a band that is both first and last, ...) is GENERIC
        unsigned role = (edge_up ? 1u : 0u) | (edge_dn ? 2u : 0u) | (ghost_dn ? 4u : 0u) | ((unsigned)tbmode << 3);
        const int hr_role = ghost_dn ? RL + 1 : (tbmode == 2 ? MR - 1 : MR);
        if (hrows != hr_role || has_up_row != (tbmode != 1) || FDTD2D_RES_DIAG) role = 32u;
        switch (role) {
            case 0u: run(std::integral_constant<unsigned, 0u>{}); break;
            case 1u: run(std::integral_constant<unsigned, 1u>{}); break;
            case 2u: run(std::integral_constant<unsigned, 2u>{}); break;
            case 3u: run(std::integral_constant<unsigned, 3u>{}); break;
            case 6u: run(std::integral_constant<unsigned, 6u>{}); break;
            case 8u: run(std::integral_constant<unsigned, 8u>{}); break;
            case 10u: run(std::integral_constant<unsigned, 10u>{}); break;
            case 16u: run(std::integral_constant<unsigned, 16u>{}); break;
            case 17u: run(std::integral_constant<unsigned, 17u>{}); break;
            default: run(std::integral_constant<unsigned, 32u>{}); break;
        }
    }
#ifdef FDTD2D_RES_TIMING
    if (l == 0 && b == 0)
        for (int i = 0; i < 8; ++i) p.trace[(crank * NW + w) * 8 + i] = (float)tacc[i] / (float)n_steps;
#endif

    // ---- store the fields -----------------------------------------------------------------------------
#pragma unroll
    for (int r = 0; r < MR; ++r) {
        const int lr = li0 + r, gi = row_lo + lr;
        if (lr < nrows && col_on) {
            const long long o = gbase + (long long)gi * p.pitch + cg;
            if (UCH && cg + 3 > C - 2) {  // H beyond column C-2 is never updated (main.py:70,74): put the input back
                float x4[4], y4[4];
                ldg4(p.in[1] + o, x4);
                ldg4(p.in[2] + o, y4);
                float hx4[4] = {lo2(hx[r][0]), hi2(hx[r][0]), lo2(hx[r][1]), hi2(hx[r][1])};
                float hy4[4] = {lo2(hy[r][0]), hi2(hy[r][0]), lo2(hy[r][1]), hi2(hy[r][1])};
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (cg + q > C - 2) hx4[q] = x4[q], hy4[q] = y4[q];
                hx[r][0] = pack2(hx4[0], hx4[1]), hx[r][1] = pack2(hx4[2], hx4[3]);
                hy[r][0] = pack2(hy4[0], hy4[1]), hy[r][1] = pack2(hy4[2], hy4[3]);
            }
            store22(p.out[0] + o, e[r]);
            store22(p.out[1] + o, hx[r]);
            store22(p.out[2] + o, hy[r]);
        }
    }
    cluster_sync_all();  // no CTA may exit while a neighbour can still write into its shared memory
}

}  // namespace fdtd2d
