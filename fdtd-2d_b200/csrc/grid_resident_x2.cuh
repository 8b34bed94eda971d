// Cluster-resident kernel, packed form (fp32): the successor of grid_resident.cuh for the small independent grids of the
// batched mode (BASELINE configs[4]: 1024 x 256^2) and the reference demo grid (fdtd.py:14-19).  Same idea -- one
// thread-block CLUSTER holds one grid for ALL n leapfrog steps of a call, HBM is touched once to load and once to store --
// rebuilt around what limited the first kernel (profiles/r1_resident_cfg5_full.summary.txt and the FDTD2D_RES_DIAG
// timings): 550 instructions per warp and step of which 268 were arithmetic, 22 shared-memory transactions of 512 bytes
// per warp and step (the shared-memory pipe was as busy as the FP pipe), and three CTA-wide barriers per step in the
// first / last CTA of a cluster for the top / bottom ring.
//
//   * 6 rows x 8 columns per thread, 8 warps per CTA (one CTA per SM, up to 255 registers): half as many rows are
//     exchanged between warps per cell, and the arithmetic is sm_100a's two-wide fp32 (add/sub/fma.rn.f32x2 on the column
//     pairs a 16-byte access delivers; strip_wave.cuh explains why the product is an fma with a -0 addend);
//   * dt/(mu*dx) is a kernel argument when it is uniform (every material_init output): the map is not kept on chip;
//   * the top / bottom Mur ring needs NO CTA-wide barrier and no full-row frames: the six rows next to the edge belong to
//     ONE warp (the first warp of the first CTA; the last active warp of the last CTA, whose band is a multiple of six
//     rows), so main.py:43-51 is elementwise between that warp's own registers -- it rides along the Ez update, which
//     walks those rows away from the edge so that Ez_prev of the next row is still in its register -- and only the two
//     5 x 5 corners (main.py:53-61) go through the warp's ring frame, one lane per corner cell.
// A step has NO CTA-wide barrier: a warp waits only for the two warps whose rows it reads (one mbarrier per warp and step
// parity, 32 arrivals), at the last moment -- the H half-step of its other rows needs nobody -- so a warp that is late (the
// ring warps, the warps that wait for another CTA) does not stop the rest of the CTA at once.
//
// Work layout.  A cluster of n <= 8 CTAs splits the rows into bands: rows [0, CW) for the first CTA, CH rows for the
// others (multiples of 6; the last band is what is left and the host makes it a multiple of 6 as well; CW is arbitrary).
// Warp w owns rows [6w, 6w+6) of the band, lane l columns [4l, 4l+4) and [128+4l, 128+4l+4), each as two register PAIRS.
// The first band may end inside a warp: that warp parks the row it receives from the CTA below in the registers of its
// first unused row (a ghost row: zero coefficients keep it inert), so the row loops stay static.
// Neighbours: columns -> warp shuffles; rows across warps -> first / last Ez row of every warp in shared memory (double
// buffered by step parity) plus a redundant copy of the Hx row above (advanced locally); rows across CTAs -> DSMEM
// st.async with mbarrier complete_tx.  Index ranges of the reference's slices are imposed by zero coefficients (maps in
// shared memory) or, for the uniform dt/(mu*dx), by per-column masked pairs and a row count.
// Ring: every warp keeps the outermost columns of its rows in a small frame (S0 = Ez before the step, S1 = after the
// interior update); S2 (Mur left/right, main.py:33-41) one lane per ring cell in place; the finished frame is pulled
// back into registers.  Sources / probes outside the frames go through 4-cell slots exactly as in grid_resident.cuh.
// Results are bit-identical to every other kernel and to the reference (tests/test_gpu_resident.py).
#pragma once
#include <type_traits>

#include "grid_resident.cuh"
#include "strip_wave.cuh"

namespace fdtd2d {

constexpr int RX_MR = 6, RX_NW = 8, RX_TH = RX_MR * RX_NW;

__host__ __device__ constexpr size_t resident_x2_smem_floats() {
    return (size_t)2 * RX_TH * RES_TW      // coefficient maps
           + RES_TW                        // dt/(mu*dx) of the row above the band
           + 4 * RX_NW * RES_TW            // Ez rows exchanged between warps: 2 parities x (first, last)
           + 4 * RES_TW                    // Ez rows from the neighbour CTAs: 2 parities x (below, above)
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

// UCH: dt/(mu*dx) is `ch_uniform` in every cell.  p.k = steps of this launch; p.CW / p.CH = rows of the first / of every
// other CTA of a cluster; gridDim.x = batch * cluster size.
template <bool UCH>
__global__ void __launch_bounds__(RX_NW * 32, 1) grid_resident_x2_kernel(const PassParams<float> p, const float ch_uniform, const u64 negzero) {
    constexpr int MR = RX_MR, NW = RX_NW, TW = RES_TW, TH = RX_TH, NT = NW * 32, LW = RES_LW, ZW = RES_ZW;
    static_assert(resident_x2_smem_floats() * 4 + 256 <= 232448, "band does not fit shared memory");
    static_assert(MR * RING <= 32 && MR >= RING + 1, "S2: one lane per ring cell of the warp's rows; the ring rows fit one warp");
    constexpr unsigned FULL = 0xffffffffu;
    constexpr uint32_t ROW_BYTES = TW * sizeof(float);
    constexpr int oLR = 0, DELTA = TH * ZW, oSlot = 2 * DELTA;  // ring frames as float offsets from F: S0 | S1 | slots
    extern __shared__ __align__(16) unsigned char smem_res[];
    float* sCe = reinterpret_cast<float*>(smem_res);  // [TH][TW] dt/(eps*dx), zero where Ez is not updated
    float* sCh = sCe + TH * TW;                       // [TH][TW] dt/(mu*dx), zero where H is not updated (unused when UCH)
    float* sChA = sCh + TH * TW;                      // [TW] dt/(mu*dx) of the row above the band
    float* sEx = sChA + TW;                           // [2 parities][first | last][NW][TW]
    float* rEz = sEx + 4 * NW * TW;                   // [2 parities][from below | from above][TW]
    float* F = rEz + 4 * TW;                          // ring frames
    int* slotW = reinterpret_cast<int*>(F + oSlot + 256 * 4);
    int* plist = slotW + 256 * 4;                     // [RES_MAX_CELLS][3] probes: frame offset, trace column, owner warp
    int* rlist = plist + RES_MAX_CELLS * 3;           // [RES_MAX_CELLS][3] sources inside ring frames: offset, wave*amp_steps, owner warp
    unsigned char* slot_tbl = reinterpret_cast<unsigned char*>(rlist + RES_MAX_CELLS * 3);  // [TH][TW/4] slot or 0xFF
    __shared__ __align__(8) uint64_t barB[2], barA[2];
    __shared__ __align__(8) uint64_t wbar[2][NW];  // "warp w has published its first / last Ez row of this parity" (32 arrivals)
    __shared__ int s_counts[2];

    const int tid = threadIdx.x, w = tid >> 5, l = tid & 31;
    const int crank = (int)cluster_ctarank(), csize = (int)cluster_nctarank();
    const int b = blockIdx.x / csize;
    const int R = p.Rg, C = p.C, n_steps = p.k;
    const int row_lo = crank == 0 ? 0 : p.CW + (crank - 1) * p.CH;
    const int nrows = crank == 0 ? min(p.CW, R) : min(p.CH, R - row_lo);
    const bool isTop = crank == 0, isBot = crank == csize - 1;
    const bool has_above = !isTop, has_below = !isBot;
    const int wl = (nrows - 1) / MR;  // the warp that holds the band's last row
    const int li0 = w * MR;
    const int nact = max(0, min(MR, nrows - li0));                             // my real rows
    const int hrows = max(0, min(MR, min(nrows, R - 1 - row_lo) - li0));       // ... of which H is updated (main.py:70,74: rows 0..R-2)
    const int cR0 = ((C - 6) >> 2) << 2;
    const int cg[2] = {4 * l, 128 + 4 * l};
    const float coef = p.mur[b];
    const u64 coef2 = pack2(coef, coef);
    const long long gbase = (long long)b * p.grid_stride;

    // ---- load: fields -> register pairs, coefficient maps -> shared memory (masked) ------------------
    u64 e[MR][2][2], hx[MR][2][2], hy[MR][2][2];
#pragma unroll
    for (int r = 0; r < MR; ++r) {
        const int lr = li0 + r, gi = row_lo + lr;
        const bool hrow = gi <= R - 2, erow = gi >= 1 && gi <= R - 2;
#pragma unroll
        for (int g = 0; g < 2; ++g) {
            float e4[4], x4[4], y4[4], ce4[4], ch4[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) e4[q] = x4[q] = y4[q] = ce4[q] = ch4[q] = 0.0f;
            if (lr < nrows && cg[g] < p.pitch) {
                const long long o = gbase + (long long)gi * p.pitch + cg[g];
                ldg4(p.in[0] + o, e4);
                ldg4(p.in[1] + o, x4);
                ldg4(p.in[2] + o, y4);
                ldg4(p.ce + o, ce4);
                if (!UCH) ldg4(p.ch + o, ch4);
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int gj = cg[g] + q;
                if (!(hrow && gj <= C - 2)) ch4[q] = 0.0f;
                if (!(erow && gj >= 1 && gj <= C - 2)) ce4[q] = 0.0f;
            }
            e[r][g][0] = pack2(e4[0], e4[1]), e[r][g][1] = pack2(e4[2], e4[3]);
            hx[r][g][0] = pack2(x4[0], x4[1]), hx[r][g][1] = pack2(x4[2], x4[3]);
            hy[r][g][0] = pack2(y4[0], y4[1]), hy[r][g][1] = pack2(y4[2], y4[3]);
            store4(sCe + lr * TW + cg[g], ce4);
            if (!UCH) store4(sCh + lr * TW + cg[g], ch4);
        }
    }
    // the uniform dt/(mu*dx) of my columns, zero beyond column C-2 (main.py:70,74)
    u64 chm[2][2];
#pragma unroll
    for (int g = 0; g < 2; ++g)
#pragma unroll
        for (int h = 0; h < 2; ++h)
            chm[g][h] = pack2(cg[g] + 2 * h <= C - 2 ? ch_uniform : 0.0f, cg[g] + 2 * h + 1 <= C - 2 ? ch_uniform : 0.0f);

    // redundant copy of the Hx row just above my rows
    u64 hxa[2][2];
    const bool has_up_row = row_lo + li0 - 1 >= 0 && li0 < nrows;
    {
        const int ga = row_lo + li0 - 1;
#pragma unroll
        for (int g = 0; g < 2; ++g) {
            float x4[4], cha[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) x4[q] = cha[q] = 0.0f;
            if (has_up_row && cg[g] < p.pitch) {
                const long long o = gbase + (long long)ga * p.pitch + cg[g];
                ldg4(p.in[1] + o, x4);
                if (!UCH && w == 0) ldg4(p.ch + o, cha);
            }
            hxa[g][0] = pack2(x4[0], x4[1]), hxa[g][1] = pack2(x4[2], x4[3]);
            if (!UCH && w == 0) {
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (cg[g] + q > C - 2) cha[q] = 0.0f;
                store4(sChA + cg[g], cha);
            }
        }
    }
    const int qA = keep(w == 0 ? TH * TW + cg[0] : (li0 - 1) * TW + cg[0]);
    const int qC = keep(li0 * TW + cg[0]);
    const int qX = keep(w * TW + cg[0]);

    // ---- set-up: barriers; where the source / probe cells of this band live in shared memory --------
    for (int i = tid; i < TH * (TW / 4) / 4; i += NT) reinterpret_cast<unsigned*>(slot_tbl)[i] = 0xffffffffu;
    for (int i = tid; i < 256 * 4; i += NT) slotW[i] = -1;
    __syncthreads();
    // offset (from F) of the finished value of cell (gi, gj) of this band once a step's boundary stages are done
    auto ring_off = [&](int gi, int gj) -> int {
        const int lr = gi - row_lo;
        if (gj < LW) return DELTA + oLR + lr * ZW + gj;
        if (gj >= cR0) return DELTA + oLR + lr * ZW + LW + (gj - cR0);
        return -1;  // not in a ring frame: the cell gets a slot
    };
    if (tid == 0) {
        mbar_init(&barB[0], 1), mbar_init(&barB[1], 1), mbar_init(&barA[0], 1), mbar_init(&barA[1], 1);
        for (int i = 0; i < 2 * NW; ++i) mbar_init(&wbar[0][0] + i, 32);
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
                rlist[3 * n_rsrc] = off, rlist[3 * n_rsrc + 1] = c.wave * p.amp_steps, rlist[3 * n_rsrc + 2] = (c.row - row_lo) / MR, ++n_rsrc;
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
            plist[3 * n_prb] = off, plist[3 * n_prb + 1] = q, plist[3 * n_prb + 2] = (c.row - row_lo) / MR, ++n_prb;
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
    // my groups' columns inside the ring frame of a row (-1: not a ring group)
    const int zo[2] = {cg[0] < LW ? cg[0] : (cg[0] >= cR0 && cg[0] < cR0 + RES_RW ? LW + cg[0] - cR0 : -1),
                       cg[1] >= cR0 && cg[1] < cR0 + RES_RW ? LW + cg[1] - cR0 : -1};
    const bool z0 = zo[0] >= 0, z1 = zo[1] >= 0;
    float* const zp0 = F + keep(oLR + li0 * ZW + zo[0]);  // S0 slot of my group 0 in the frame of my first row (S1 is DELTA further)
    float* const zp1 = F + keep(oLR + li0 * ZW + zo[1]);
    // 1: my six rows are the top ring rows 0..5, 2: the bottom ring rows R-6..R-1 (the host makes the last band a multiple
    // of six rows, and a band that holds both has at least twelve), 0: neither
    const int tbmode = FDTD2D_RES_DIAG >= 1 ? 0 : (isTop && w == 0) ? 1 : (isBot && li0 + MR == nrows) ? 2 : 0;
    // S2 (Mur left/right) of my warp's rows: lane -> (row l / 5, ring cell l % 5)
    const int s2row = li0 + l / RING, s2k = l % RING, s2gi = row_lo + s2row;
    const bool s2act = l < MR * RING && s2row < nrows && s2gi >= 1 && s2gi <= R - 2;
    float* const s2L = F + keep(DELTA + oLR + s2row * ZW + s2k);                           // S1 of cell (row, k)
    float* const s2R = F + keep(DELTA + oLR + s2row * ZW + LW + (C - 1 - s2k - cR0));      // S1 of cell (row, C-1-k)
    // the corners (main.py:53-61), by the warp that holds the six rows: lane -> (depth d = l / 5 from the edge, j = l % 5
    // from the side); frame row of depth d is d (top) or 5 - d (bottom)
    const bool cact = tbmode != 0 && l < RING * RING;
    const int cd = l / RING, cj = l % RING;
    auto cfr = [&](int d) { return (tbmode == 1 ? d : MR - 1 - d) * ZW; };
    float* const cF = F + oLR + li0 * ZW;  // S0 frame of my first row

    auto park = [&](int delta) {
        if (FDTD2D_RES_DIAG >= 2) return;
#pragma unroll
        for (int r = 0; r < MR; ++r) {
            if (z0) store22(zp0 + delta + r * ZW, e[r][0]);
            if (z1) store22(zp1 + delta + r * ZW, e[r][1]);
        }
    };
    // the probes of my rows, from the finished frames / slots (by the rows' own warp: nobody else knows when they are final)
    auto sample_probes = [&](long long step) {
        __syncwarp();
        for (int i = l; i < n_prb; i += 32)
            if (plist[3 * i + 2] == w && step < p.trace_cap) p.trace[step * p.n_probe + plist[3 * i + 1]] = F[plist[3 * i]];
        __syncwarp();
    };
    // addresses in the neighbours' shared memory
    const uint32_t up_rank = has_above ? crank - 1 : crank, dn_rank = has_below ? crank + 1 : crank;
    const uint32_t up_rEz = map_to_rank(smem_u32(rEz), up_rank), up_barB = map_to_rank(smem_u32(&barB[0]), up_rank);
    const uint32_t dn_rEz = map_to_rank(smem_u32(rEz), dn_rank), dn_barA = map_to_rank(smem_u32(&barA[0]), dn_rank);
    const bool edge_dn = (w == wl) && has_below;  // my last real row needs the Ez row of the CTA below
    const bool edge_up = (w == 0) && has_above;   // my first row needs the Hx row of the CTA above
    const bool warp_on = li0 < nrows;
    const int rl = nact - 1;                      // my last real row
    const bool ghost_dn = edge_dn && rl < MR - 1;  // the row from below is parked in the registers of row rl + 1
    {  // rows nobody publishes are still read as a neighbour row: keep them finite
        const float z[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
        for (int i = 0; i < 4; ++i) store4(sEx + (i * NW + w) * TW + cg[0], z), store4(sEx + (i * NW + w) * TW + cg[1], z);
    }
    // one row of the H half-step (main.py:69-74); dn = the Ez row below it
    auto h_row = [&](const int r, const u64 (&dn)[2][2]) {
        u64 c[2][2];
        if (UCH) {
            c[0][0] = chm[0][0], c[0][1] = chm[0][1], c[1][0] = chm[1][0], c[1][1] = chm[1][1];
        } else {
            load22(sCh + qC + r * TW, c[0]);
            load22(sCh + qC + r * TW + 128, c[1]);
        }
        const float ra = __shfl_sync(FULL, lo2(e[r][0][0]), (l + 1) & 31);
        const float rb = __shfl_sync(FULL, lo2(e[r][1][0]), (l + 1) & 31);
        const float right3[2] = {l == 31 ? rb : ra, rb};
#pragma unroll
        for (int g = 0; g < 2; ++g) {
            const u64 e0 = e[r][g][0], e1 = e[r][g][1];
            const u64 dx0 = pack2(sub_rn(hi2(e0), lo2(e0)), sub_rn(lo2(e1), hi2(e0)));
            const u64 dx1 = pack2(sub_rn(hi2(e1), lo2(e1)), sub_rn(right3[g], hi2(e1)));
            hx[r][g][0] = sub2(hx[r][g][0], mul2(c[g][0], sub2(dn[g][0], e0), negzero));
            hx[r][g][1] = sub2(hx[r][g][1], mul2(c[g][1], sub2(dn[g][1], e1), negzero));
            hy[r][g][0] = add2(hy[r][g][0], mul2(c[g][0], dx0, negzero));
            hy[r][g][1] = add2(hy[r][g][1], mul2(c[g][1], dx1, negzero));
        }
    };
    // the interior Ez update of one row (main.py:21-27) into t; up = the Hx row above it
    auto e_row = [&](const int r, const u64 (&up)[2][2], u64 (&t)[2][2]) {
        u64 c[2][2];
        load22(sCe + qC + r * TW, c[0]);
        load22(sCe + qC + r * TW + 128, c[1]);
        const float la = __shfl_sync(FULL, hi2(hy[r][0][1]), (l + 31) & 31);
        const float lb = __shfl_sync(FULL, hi2(hy[r][1][1]), (l + 31) & 31);
        const float left0[2] = {la, l == 0 ? la : lb};
#pragma unroll
        for (int g = 0; g < 2; ++g) {
            const u64 y0 = hy[r][g][0], y1 = hy[r][g][1];
            const u64 dy0 = pack2(sub_rn(lo2(y0), left0[g]), sub_rn(hi2(y0), lo2(y0)));
            const u64 dy1 = pack2(sub_rn(lo2(y1), hi2(y0)), sub_rn(hi2(y1), lo2(y1)));
            const u64 curl0 = sub2(dy0, sub2(hx[r][g][0], up[g][0]));
            const u64 curl1 = sub2(dy1, sub2(hx[r][g][1], up[g][1]));
            t[g][0] = add2(e[r][g][0], mul2(curl0, c[g][0], negzero));
            t[g][1] = add2(e[r][g][1], mul2(curl1, c[g][1], negzero));
        }
    };
    // Ez update of a warp that holds the top (TOP) or bottom ring rows: the rows are walked AWAY from the edge's far side,
    // i.e. towards the edge (top: r = 5..0, bottom: r = 0..5), so that when row r has its interior update S1[r] the row nearer
    // the edge, rn, still holds Ez_prev; main.py:43-51 for row rn is S0[r] + coef * (S1[r] - S0[rn]) (on the ring columns S2
    // instead of S1: those cells are corner cells and are overwritten through the frame).  S1 of the ring groups is parked
    // on the way (S2 and the corners need it).
    auto e_rows_tb = [&](auto top_c) {
        constexpr bool TOP = decltype(top_c)::value;
        u64 pend[2][2] = {{0, 0}, {0, 0}};
#pragma unroll
        for (int i = 0; i < MR; ++i) {
            const int r = TOP ? MR - 1 - i : i;    // the row whose interior update is formed
            const int rn = TOP ? r - 1 : r + 1;    // the row nearer the edge
            u64 t[2][2], t3[2][2] = {{0, 0}, {0, 0}};
            if (r == 0)
                e_row(r, hxa, t);
            else
                e_row(r, hx[r > 0 ? r - 1 : 0], t);
            if (z0) store22(zp0 + DELTA + r * ZW, t[0]);
            if (z1) store22(zp1 + DELTA + r * ZW, t[1]);
            if (i < MR - 1) {
#pragma unroll
                for (int g = 0; g < 2; ++g)
#pragma unroll
                    for (int h = 0; h < 2; ++h) t3[g][h] = add2(e[r][g][h], mul2(coef2, sub2(t[g][h], e[rn < 0 ? 0 : (rn >= MR ? MR - 1 : rn)][g][h]), negzero));
            }
#pragma unroll
            for (int g = 0; g < 2; ++g)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    e[r][g][h] = i == 0 ? t[g][h] : pend[g][h];
                    pend[g][h] = t3[g][h];
                }
        }
    };
    cluster_sync_all();  // every CTA's barriers are initialised before anyone signals them

#ifdef FDTD2D_RES_TIMING
    long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tprev = clock64();
#define RX_STAMP(i) { const long long tn = clock64(); tacc[i] += tn - tprev; tprev = tn; }
#else
#define RX_STAMP(i)
#endif
#pragma unroll 1
    for (int s = 0; s < n_steps; ++s) {
        const int par = s & 1;
        const uint32_t ph = (uint32_t)(s >> 1) & 1u;
        const long long step = p.step0 + s;
        float* const xF = sEx + par * (2 * NW * TW) + qX;
        if (warp_on) {
            // ---- publish my first and last Ez rows; the band's first / last row also go to the neighbour CTAs ----
            store22(xF, e[0][0]);
            store22(xF + 128, e[0][1]);
            store22(xF + NW * TW, e[MR - 1][0]);
            store22(xF + NW * TW + 128, e[MR - 1][1]);
            if (edge_up) {
                st_async22(up_rEz + (uint32_t)(par * 2 * TW + cg[0]) * 4u, e[0][0], up_barB + 8u * par);
                st_async22(up_rEz + (uint32_t)(par * 2 * TW + cg[1]) * 4u, e[0][1], up_barB + 8u * par);
            }
            if (edge_dn) {
                const uint32_t a0 = dn_rEz + (uint32_t)((par * 2 + 1) * TW + cg[0]) * 4u, a1 = dn_rEz + (uint32_t)((par * 2 + 1) * TW + cg[1]) * 4u;
                const uint32_t bar = dn_barA + 8u * par;
                if (!ghost_dn) {
                    st_async22(a0, e[MR - 1][0], bar), st_async22(a1, e[MR - 1][1], bar);
                } else {
                    switch (rl) {
                        case 0: st_async22(a0, e[0][0], bar), st_async22(a1, e[0][1], bar); break;
                        case 1: st_async22(a0, e[1][0], bar), st_async22(a1, e[1][1], bar); break;
                        case 2: st_async22(a0, e[2][0], bar), st_async22(a1, e[2][1], bar); break;
                        case 3: st_async22(a0, e[3][0], bar), st_async22(a1, e[3][1], bar); break;
                        default: st_async22(a0, e[4][0], bar), st_async22(a1, e[4][1], bar); break;
                    }
                }
            }
            mbar_arrive(&wbar[par][w]);  // (release: the rows above are visible to whoever sees the phase complete)
            park(0);  // S0: Ez is not changed by the H half-step
        }
        RX_STAMP(0)
        RX_STAMP(1)
        if (warp_on) {
            // ---- H half-step; the row that needs a neighbour CTA's Ez row comes last ---------------------------
            if (ghost_dn) {  // (one warp of the first CTA) the row from below becomes the Ez of my first unused row
                if (l == 0) mbar_expect_tx(&barB[par], ROW_BYTES);
                mbar_wait(&barB[par], ph);
                u64 dn[2][2];
                load22(rEz + par * 2 * TW + cg[0], dn[0]);
                load22(rEz + par * 2 * TW + cg[0] + 128, dn[1]);
                switch (rl) {
                    case 0: e[1][0][0] = dn[0][0], e[1][0][1] = dn[0][1], e[1][1][0] = dn[1][0], e[1][1][1] = dn[1][1]; break;
                    case 1: e[2][0][0] = dn[0][0], e[2][0][1] = dn[0][1], e[2][1][0] = dn[1][0], e[2][1][1] = dn[1][1]; break;
                    case 2: e[3][0][0] = dn[0][0], e[3][0][1] = dn[0][1], e[3][1][0] = dn[1][0], e[3][1][1] = dn[1][1]; break;
                    case 3: e[4][0][0] = dn[0][0], e[4][0][1] = dn[0][1], e[4][1][0] = dn[1][0], e[4][1][1] = dn[1][1]; break;
                    default: e[5][0][0] = dn[0][0], e[5][0][1] = dn[0][1], e[5][1][0] = dn[1][0], e[5][1][1] = dn[1][1]; break;
                }
            }
            if (hrows == MR) {  // (all but a few warps: one straight line of independent rows)
#pragma unroll
                for (int r = 0; r + 1 < MR; ++r) h_row(r, e[r + 1 < MR ? r + 1 : r]);
            } else {
#pragma unroll
                for (int r = 0; r + 1 < MR; ++r)
                    if (r < hrows) h_row(r, e[r + 1 < MR ? r + 1 : r]);
            }
            RX_STAMP(2)
            if (MR - 1 < hrows) {
                u64 dn[2][2];
                const float* belowp = xF + (w + 1 < NW ? TW : 0);  // first row of the warp below
                if (edge_dn) {
                    if (l == 0) mbar_expect_tx(&barB[par], ROW_BYTES);
                    mbar_wait(&barB[par], ph);
                    belowp = rEz + par * 2 * TW + cg[0];
                } else {
                    mbar_wait(&wbar[par][w + 1 < NW ? w + 1 : w], ph);
                }
                load22(belowp, dn[0]);
                load22(belowp + 128, dn[1]);
                h_row(MR - 1, dn);
            }
            if (has_up_row) {  // my copy of the Hx row above my rows (main.py:69-70 for that row)
                u64 up[2][2], c[2][2];
                const float* abovep = xF + NW * TW - (w > 0 ? TW : 0);  // last row of the warp above
                if (edge_up) {
                    if (l == 0) mbar_expect_tx(&barA[par], ROW_BYTES);
                    mbar_wait(&barA[par], ph);
                    abovep = rEz + (par * 2 + 1) * TW + cg[0];
                } else {
                    mbar_wait(&wbar[par][w > 0 ? w - 1 : 0], ph);
                }
                load22(abovep, up[0]);
                load22(abovep + 128, up[1]);
                if (UCH) {
                    c[0][0] = chm[0][0], c[0][1] = chm[0][1], c[1][0] = chm[1][0], c[1][1] = chm[1][1];
                } else {
                    load22(sCh + qA, c[0]);
                    load22(sCh + qA + 128, c[1]);
                }
#pragma unroll
                for (int g = 0; g < 2; ++g)
#pragma unroll
                    for (int h = 0; h < 2; ++h) hxa[g][h] = sub2(hxa[g][h], mul2(c[g][h], sub2(e[0][g][h], up[g][h]), negzero));
            }
            RX_STAMP(3)
            // ---- interior Ez update (no barrier: every Hx row it reads is in my registers); S1 -> ring frame ------
            if (tbmode == 1) {
                e_rows_tb(std::true_type{});
            } else if (tbmode == 2) {
                e_rows_tb(std::false_type{});
            } else {
#pragma unroll
                for (int r = MR - 1; r > 0; --r) e_row(r, hx[r > 0 ? r - 1 : 0], e[r]);
                e_row(0, hxa, e[0]);
                park(DELTA);
            }
            RX_STAMP(4)
            // ---- source / probe cells outside the frames go through their slot ---------------------------------
            if (spmask) {
#pragma unroll
                for (int r = 0; r < MR; ++r)
#pragma unroll
                    for (int g = 0; g < 2; ++g)
                        if (spmask >> (r * 2 + g) & 1u) {
                            const int sl = slot_tbl[(li0 + r) * (TW / 4) + (cg[g] >> 2)];
                            float* f = F + oSlot + sl * 4;
                            store22(f, e[r][g]);
                            if (step < p.amp_steps)
                                for (int q = 0; q < 4; ++q) {  // source add (fdtd.py:34): float64 sum, then cast
                                    const int wv = slotW[sl * 4 + q];
                                    if (wv >= 0) f[q] = add_source(f[q], p.amp[(long long)wv * p.amp_steps + step]);
                                }
                            load22(f, e[r][g]);
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
            RX_STAMP(5)
            if (tbmode) {
                // ---- corners: Jacobi over values after S3 (every read is of a not-yet-processed cell) --------------
                __syncwarp();
                if (cact) {
                    // value after main.py:43-51 of the frame cell (depth d, frame column fc = global column gc)
                    auto S3v = [&](int d, int fc, int gc) -> float {
                        const float s2 = cF[DELTA + cfr(d) + fc];
                        if (d <= 4 && gc >= 1 && gc <= C - 2)
                            return add_rn(cF[cfr(d + 1) + fc], mul_rn(coef, sub_rn(cF[DELTA + cfr(d + 1) + fc], cF[cfr(d) + fc])));
                        return s2;
                    };
                    const int gr = C - 1 - cj, fr = LW + gr - cR0;
                    vl = mul_rn(add_rn(S3v(cd, cj + 1, cj + 1), S3v(cd + 1, cj, cj)), 0.5f);  // == sum / 2 exactly
                    vr = mul_rn(add_rn(S3v(cd, fr - 1, gr - 1), S3v(cd + 1, fr, gr)), 0.5f);
                }
                __syncwarp();
                // the rows S3 finished in registers -> frame (columns of the ring groups that are not corner cells)
#pragma unroll
                for (int r = 0; r < MR; ++r)
                    if (tbmode == 1 ? r < MR - 1 : r > 0) {
                        if (z0) store22(zp0 + DELTA + r * ZW, e[r][0]);
                        if (z1) store22(zp1 + DELTA + r * ZW, e[r][1]);
                    }
                __syncwarp();
                if (cact) cF[DELTA + cfr(cd) + cj] = vl, cF[DELTA + cfr(cd) + LW + (C - 1 - cj) - cR0] = vr;
            }
            RX_STAMP(6)
            if (n_rsrc) {  // sources inside a ring frame are added once the frame is finished (by the rows' own warp)
                __syncwarp();
                for (int i = l; i < n_rsrc; i += 32)
                    if (rlist[3 * i + 2] == w && step < p.amp_steps) {
                        float* f = F + rlist[3 * i];
                        *f = add_source(*f, p.amp[(long long)rlist[3 * i + 1] + step]);
                    }
            }
            __syncwarp();
            // ---- finished ring -> registers -------------------------------------------------------------------
            if (FDTD2D_RES_DIAG < 2)
#pragma unroll
            for (int r = 0; r < MR; ++r) {
                if (z0) load22(zp0 + DELTA + r * ZW, e[r][0]);
                if (z1) load22(zp1 + DELTA + r * ZW, e[r][1]);
            }
            if (n_prb) sample_probes(step);
            RX_STAMP(7)
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
#pragma unroll
        for (int g = 0; g < 2; ++g)
            if (lr < nrows && cg[g] < p.pitch) {
                const long long o = gbase + (long long)gi * p.pitch + cg[g];
                store22(p.out[0] + o, e[r][g]);
                store22(p.out[1] + o, hx[r][g]);
                store22(p.out[2] + o, hy[r][g]);
            }
    }
    cluster_sync_all();  // no CTA may exit while a neighbour can still write into its shared memory
}

}  // namespace fdtd2d
