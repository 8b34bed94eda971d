"""The pass planner on the CPU (fdtd2d_plan_host: host arithmetic only, no GPU): for whole grids and y-slabs, fp32 and
fp64, with sources and probes in awkward places, every owned cell must be produced by exactly one task (edge tile, TMA
tile or wavefront run), every wavefront run must stay inside the rows / columns it may read, band rows of a slab must
be produced by band tasks, and the band-task counts the kernels wait for must match the lists."""
import ctypes

import numpy as np
import pytest

import fdtd2d_b200 as fd
from fdtd2d_b200 import _lib

RING, SM = 5, 148


@pytest.fixture(scope="module")
def lib():
    fd.build()
    return _lib.lib()


def plan(lib, dtype, Rg, C, k, *, rows=None, halo=0, batch=1, src=(), probe=(), wave_min=-1, ring_min=-1, wavefront=1,
         ring_strips=1, variant=0, uniform=1):
    rb, re = rows if rows else (0, Rg)
    geom = np.array([dtype, batch, Rg, C, rb, re, halo, k, SM, variant, wave_min, ring_min, wavefront, ring_strips, uniform], np.int32)
    s = np.ascontiguousarray(np.array(src, np.int32).reshape(-1, 3))
    p = np.ascontiguousarray(np.array(probe, np.int32).reshape(-1, 3))
    out = np.zeros(16, np.int32)
    vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    rc = lib.fdtd2d_plan_host(vp(geom), len(s), vp(s), len(p), vp(p), vp(out), None, 0, None, 0)
    assert rc == 0, lib.fdtd2d_last_error()
    names = ("tiles_y", "tiles_x", "CH", "CW", "hx", "org", "Rl", "n_edge", "n_edge_band", "n_tma", "n_wave", "n_wave_band", "ring",
             "exp_top", "exp_bot", "pitch")
    info = dict(zip(names, (int(v) for v in out)))
    kind = np.zeros(max(1, batch * info["tiles_y"] * info["tiles_x"]), np.int32)
    tasks = np.zeros((max(1, info["n_wave"]), 8), np.int32)
    rc = lib.fdtd2d_plan_host(vp(geom), len(s), vp(s), len(p), vp(p), vp(out), vp(kind), len(kind), vp(tasks), len(tasks))
    assert rc == 0, lib.fdtd2d_last_error()
    info.update(kind=kind.reshape(batch, info["tiles_y"], info["tiles_x"]), tasks=tasks[:info["n_wave"]], k=k, Rg=Rg, C=C, rb=rb, re=re,
                halo=halo, batch=batch, dtype=dtype, src=s, probe=p)
    return info


def check(pl):
    k, Rg, C, rb, re, halo, B = pl["k"], pl["Rg"], pl["C"], pl["rb"], pl["re"], pl["halo"], pl["batch"]
    top, bot = rb > 0, re < Rg
    row0 = rb - (halo if top else 0)
    Rl, org, own_hi = pl["Rl"], pl["org"], pl["org"] + (re - rb)
    assert org == rb - row0 and Rl == (re + (halo if bot else 0)) - row0
    CH, CW, hx, pitch = pl["CH"], pl["CW"], pl["hx"], pl["pitch"]
    W = 128 if pl["dtype"] == 0 else 64  # strip width
    TH = 64 if pl["dtype"] == 0 else 32
    band = [(org, org + (halo if top else 0)), (own_hi - (halo if bot else 0), own_hi)]
    cover = np.zeros((B, Rl, pitch), np.int16)
    n_edge = n_edge_band = n_tma = 0
    exp = [0, 0]
    for b in range(B):
        for ty in range(pl["tiles_y"]):
            r0, r1 = org + ty * CH, min(org + (ty + 1) * CH, own_hi)
            for tx in range(pl["tiles_x"]):
                kd = pl["kind"][b, ty, tx]
                c0, c1 = tx * CW, min(tx * CW + CW, pitch)
                in_band = [r0 < band[s][1] and r1 > band[s][0] for s in (0, 1)]
                if kd in (0, 3):
                    cover[b, r0:r1, c0:c1] += 1
                if kd == 0:
                    n_edge += 1
                    n_edge_band += any(in_band)
                    exp[0] += in_band[0]
                    exp[1] += in_band[1]
                if kd == 3:  # TMA tile: the whole window exists, full core, plain, never a band tile
                    n_tma += 1
                    assert pl["dtype"] == 0 and not any(in_band) and r1 - r0 == CH
                    assert r0 - k >= 0 and r0 - k + TH <= Rl
                    assert row0 + r0 - k >= RING and row0 + r0 - k + TH <= Rg - RING
                    assert tx * CW - hx >= RING and tx * CW - hx + 128 <= C - RING
    assert (n_edge, n_edge_band, n_tma) == (pl["n_edge"], pl["n_edge_band"], pl["n_tma"])
    t = pl["tasks"]
    assert np.all(t[:pl["n_wave_band"], 7] != 0) and np.all(t[pl["n_wave_band"]:, 7] == 0), "band runs come first"
    for b_, x0, y0, y1, c0, c1, side, bnd in t:
        assert y0 < y1 and y0 - k >= 0 and y1 + k <= Rl, "a run reads k rows above and below what it stores"
        assert row0 + y0 - k >= RING and row0 + y1 + k <= Rg - RING, "no top / bottom ring within reach"
        assert org <= y0 and y1 <= own_hi, "only owned rows are stored"
        q = 4 if pl["dtype"] == 0 else 2
        assert x0 % q == 0 and c0 % q == 0 and c1 % q == 0 and 0 <= c0 < c1 <= W and x0 >= 0 and x0 + W <= pitch
        if side == 0:
            assert c0 >= k and W - c1 >= k, "column halo"
            assert x0 + c0 - k >= RING and x0 + c1 + k <= C - RING, "a left / right ring cell within k columns of what a plain strip stores"
        else:
            assert pl["ring"] == 1 and k == 8
            assert (x0 == 0 and c0 == 0 and W - c1 >= k) if side == 1 else (x0 == (C + q - 1) // q * q - W and c1 == W and c0 >= k)
        for g, r, c in pl["src"]:
            assert not (g == b_ and y0 - k <= r - row0 < y1 + k and x0 <= c < x0 + W), "a source inside a run's window"
        for g, r, c in pl["probe"]:
            assert not (g == b_ and y0 <= r - row0 < y1 and x0 + c0 <= c < x0 + c1), "a probe inside a run"
        if bnd:
            lo, hi = band[bnd - 1]
            assert lo <= y0 and y1 <= hi and hi > lo
            exp[bnd - 1] += 1
        else:
            assert all(not (y0 < hi and y1 > lo) for lo, hi in band), "band rows belong to band runs"
        cover[b_, y0:y1, x0 + c0:x0 + c1] += 1
    owned = cover[:, org:own_hi, :C]
    assert owned.min() == 1 and owned.max() == 1, f"cells produced {owned.min()}..{owned.max()} times"
    assert cover[:, :org].sum() == 0 and cover[:, own_hi:].sum() == 0, "ghost rows are never stored locally"
    if top or bot:
        assert exp == [pl["exp_top"], pl["exp_bot"]]
    else:
        assert pl["exp_top"] == pl["exp_bot"] == 0


GRIDS = [(300, 517), (1024, 1024), (203, 600), (2000, 260), (700, 1500), (97, 225), (11, 11), (4096, 4096), (129, 1000)]


@pytest.mark.parametrize("dtype", [0, 1])
@pytest.mark.parametrize("k", [1, 3, 4, 6, 8, 12])
def test_whole_grids_are_covered_once(lib, dtype, k):
    if dtype == 1 and k > 8:
        pytest.skip("fp64 tiles are 32 rows high: k <= 8")
    for R, C in GRIDS:
        src = [(0, R // 2, C // 2), (0, R // 3, C // 4), (0, 7, 9), (0, R - 1, 0)]
        probe = [(0, R // 2, C // 2 + 3), (0, 0, 0), (0, R - 1, C - 1), (0, R // 4, C // 3)]
        for wave_min in (-1, 0):
            for ring_min in (-1, 0):
                check(plan(lib, dtype, R, C, k, src=src, probe=probe, wave_min=wave_min, ring_min=ring_min))
    check(plan(lib, dtype, 400, 900, k, batch=3, src=[(b, 200 + 10 * b, 450 - 50 * b) for b in range(3)], probe=[(1, 133, 307)], wave_min=0, ring_min=0))
    check(plan(lib, dtype, 16384, 16384, k, src=[(0, 8192, 8192)], probe=[(0, 8192, 8208), (0, 8, 8192)]))


@pytest.mark.parametrize("dtype", [0, 1])
@pytest.mark.parametrize("world,k,halo", [(2, 8, 8), (3, 8, 8), (4, 5, 8), (2, 4, 8), (8, 8, 8), (2, 12, 12), (3, 6, 12), (2, 1, 1)])
def test_slabs_are_covered_once_and_bands_are_band_tasks(lib, dtype, world, k, halo):
    if dtype == 1 and k > 8:
        pytest.skip("fp64 tiles are 32 rows high: k <= 8")
    for R, C in [(700, 900), (4096, 3000), (65536 // 8 * world, 2048), (34 * world + 5, 640)]:
        for r in range(world):
            rb, re = fd.slab_rows(R, world, r)
            b1 = fd.slab_rows(R, world, 1)[0]
            src = [(0, R // 2, C // 2), (0, b1, 100), (0, b1 - 1, 300), (0, b1 + 3, 500), (0, 7, 7)]
            probe = [(0, b1, 101), (0, b1 - 2, 301), (0, R - 1, C - 1), (0, 0, 0), (0, R // 2, C // 2)]
            for wave_min in (-1, 0):
                pl = plan(lib, dtype, R, C, k, rows=(rb, re), halo=halo, src=src, probe=probe, wave_min=wave_min, ring_min=0 if wave_min == 0 else -1)
                check(pl)
                # every band row is produced by a task that knows it is a band task
                assert (pl["exp_top"] > 0) == (r > 0) and (pl["exp_bot"] > 0) == (r < world - 1)


def test_wavefront_takes_large_grids_and_both_dtypes(lib):
    """The default choice: 16384^2 fp32 runs on the wavefront with ring strips; a 8192-row slab of 65536 columns has its
    band rows on band runs; 8192^2 fp64 runs on 64-column strips, two per tile column."""
    pl = plan(lib, 0, 16384, 16384, 8)
    assert pl["n_wave"] > 1000 and pl["ring"] == 1 and pl["n_tma"] == 0
    pl = plan(lib, 0, 65536, 65536, 8, rows=(8192, 16384), halo=8)
    assert pl["n_wave_band"] >= 2 * (65536 // 112) and pl["n_edge"] == 0, "a middle slab without sources is all wavefront"
    pl = plan(lib, 1, 8192, 8192, 8)
    assert pl["n_wave"] > 1000 and pl["CW"] == 96 and pl["ring"] == 1 and pl["n_edge"] < 300, pl["n_edge"]
    t = pl["tasks"]
    assert set(np.unique((t[:, 5] - t[:, 4])[t[:, 6] == 0])) == {48}
    pl = plan(lib, 1, 8192, 8192, 8, wavefront=0)
    assert pl["n_wave"] == 0 and pl["n_tma"] == 0 and pl["n_edge"] == pl["tiles_y"] * pl["tiles_x"]
    pl = plan(lib, 0, 1024, 1024, 8)  # small grid: persistent TMA tiles
    assert pl["n_wave"] == 0 and pl["n_tma"] > 0


# ---- the fused double pass (two k = 8 passes per launch, the second fed from L2) ----------------------------------------
def plan_fused(lib, Rg, C, *, fuse=1, batch=1, src=(), probe=(), wave_min=-1, ring_min=-1):
    geom = np.array([0, batch, Rg, C, 0, Rg, 0, 8, SM, 0, wave_min, ring_min, 1, 1, 1], np.int32)
    s = np.ascontiguousarray(np.array(src, np.int32).reshape(-1, 3))
    p = np.ascontiguousarray(np.array(probe, np.int32).reshape(-1, 3))
    vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    counts = np.zeros(4, np.int32)
    assert lib.fdtd2d_plan_host_fused(vp(geom), fuse, len(s), vp(s), len(p), vp(p), vp(counts), None, 0, None, 0) == 0, lib.fdtd2d_last_error()
    fused = np.zeros((max(1, counts[0]), 12), np.int32)
    deferred = np.zeros((max(1, counts[1]), 12), np.int32)
    assert lib.fdtd2d_plan_host_fused(vp(geom), fuse, len(s), vp(s), len(p), vp(p), vp(counts), vp(fused), len(fused), vp(deferred),
                                      len(deferred)) == 0, lib.fdtd2d_last_error()
    return fused[:counts[0]], deferred[:counts[1]], int(counts[2]), int(counts[3])


@pytest.mark.parametrize("shape", [(4096, 4096), (3000, 4100), (16384, 2100), (1500, 2100)])
def test_fused_double_pass_plan(lib, shape):
    """Phase 0 of the fused launch is the single-pass plan with runs cut at 16-row blocks; phase 1 (fused + deferred pieces)
    stores exactly the same cells; every fused phase-1 run comes after all phase-0 runs it can read, reads only cells that
    phase-0 RUNS store (never an edge tile's), and names exactly the tile columns its window covers."""
    R, C = shape
    src = [(0, R // 2, C // 2), (0, R // 3, C // 4)]
    probe = [(0, R // 2, C // 2 + 16), (0, R // 4, C // 4), (0, 8, C // 2)]
    single = plan(lib, 0, R, C, 8, src=src, probe=probe, wave_min=0, ring_min=0)
    fused, deferred, nblk, n_single = plan_fused(lib, R, C, src=src, probe=probe, wave_min=0, ring_min=0)
    assert n_single == single["n_wave"] and nblk == -(-R // 16) and len(fused) > 0
    CW, K, pitch = single["CW"], 8, single["pitch"]
    p0, p1 = fused[fused[:, 8] == 0], fused[fused[:, 8] == 1]
    cover0, cover1 = np.zeros((R, pitch), np.int8), np.zeros((R, pitch), np.int8)
    for b_, x0, y0, y1, c0, c1, side, bnd, ph, tx, txlo, txhi in p0:
        assert y0 % 16 == 0 and y1 % 16 == 0 and y0 < y1 and bnd == 0
        assert tx * CW <= x0 + c0 and x0 + c1 <= max((tx + 1) * CW, pitch if tx == single["tiles_x"] - 1 else 0), "stores stay in its tile column"
        cover0[y0:y1, x0 + c0:x0 + c1] += 1
    for b_, x0, y0, y1, c0, c1, side, bnd, ph, tx, txlo, txhi in np.concatenate([p1, deferred]):
        assert ph == 1 and y0 < y1
        cover1[y0:y1, x0 + c0:x0 + c1] += 1
    want = np.zeros((R, pitch), np.int8)
    for b_, x0, y0, y1, c0, c1, side, bnd in single["tasks"]:
        want[y0:y1, x0 + c0:x0 + c1] += 1
    assert np.array_equal(cover0, want) and np.array_equal(cover1, want) and want.max() == 1
    # order + windows of the fused phase-1 runs: a run is PAIRED with the phase-0 run of its strip that starts 16 rows below
    # its own first row (they start together, so the second trails the first closely enough to read from L2), and every
    # phase-0 run it reads has a ticket at most a few hundred after its own (taken long before the waiting runs could
    # fill the 1184 warps of the GPU)
    p0_idx = [i for i, t in enumerate(fused) if t[8] == 0]
    worst = 0
    starts = {(t[1], t[2]) for t in p0}
    for i, (b_, x0, y0, y1, c0, c1, side, bnd, ph, tx, txlo, txhi) in enumerate(fused):
        if ph == 0:
            continue
        assert txlo == max(0, x0 // CW) and txhi == min(single["tiles_x"] - 1, (x0 + 127) // CW)
        win = want[y0 - K:y1 + K, x0:min(x0 + 128, C)]
        assert win.min() == 1, "the window holds a cell no phase-0 run stores (an edge tile's)"
        if (x0, y0 + 16) in starts:  # (the first piece of a fusable range starts at a tile row instead: its producer started earlier)
            prev = fused[i - 1]
            assert prev[8] == 0 and prev[1] == x0 and prev[2] == y0 + 16, "not paired with the phase-0 run of its strip"
        for a in p0_idx:
            u = fused[a]
            if txlo <= u[9] <= txhi and u[2] < y1 + K and u[3] > y0 - K:
                worst = max(worst, a - i)
    assert worst <= 2 * single["tiles_x"] + 8 and worst < SM * 8 // 2, worst
    assert len(deferred) < 0.2 * len(p1) + 4 * single["tiles_x"]


def test_fused_automatic_mode_takes_only_large_grids(lib):
    assert len(plan_fused(lib, 16384, 16384, fuse=-1)[0]) > 2000
    assert len(plan_fused(lib, 4096, 4096, fuse=-1)[0]) == 0
    assert len(plan_fused(lib, 16384, 16384, fuse=0)[0]) == 0


# ---- the band split of the cluster-resident kernels (fdtd2d_plan_resident: host arithmetic only) ----------------------
def resident_plan(lib, R, C, cfg=5, cluster=0):
    out = np.zeros(4, np.int32)
    rc = lib.fdtd2d_plan_resident(R, C, cfg, cluster, out.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0, lib.fdtd2d_last_error()
    return tuple(int(v) for v in out)


def test_resident_bands_of_the_packed_kernel(lib):
    """grid_resident_x2.cuh relies on: at most 8 bands of at most 48 rows; every band but the first a multiple of six rows
    (the six bottom ring rows are ONE row block's rows), the last one at least six; the first band 6..48 rows (the six top
    ring rows are its first block); a band that holds both rings has twelve rows or more; grids whose right ring straddles
    column 128 and grids of more than 384 rows go to the round-1 kernel (shape 0) or are not resident at all."""
    for C in (16, 100, 128, 129, 133, 134, 200, 256):
        straddle = C > 128 and ((C - 6) // 4) * 4 < 128
        for R in range(16, 420):
            shape, n, first, rpc = resident_plan(lib, R, C)
            if R > 384:
                assert shape == -1, (R, C)
                continue
            assert shape == (0 if straddle else 5), (R, C, shape)
            if shape != 5:
                continue
            assert 1 <= n <= 8 and n == max(-(-R // 48), 2 if (R % 6 or R < 12) and R <= 48 else 1), (R, C, n)
            if n == 1:
                assert first == R and R % 6 == 0 and R >= 12
                continue
            last = R - first - (n - 2) * rpc
            assert 6 <= first <= 48 and rpc % 6 == 0 and 6 <= rpc <= 48, (R, C, n, first, rpc)
            assert last % 6 == 0 and 6 <= last <= min(48, rpc), (R, C, n, first, rpc, last)  # (the kernel caps every band but the first at rpc rows)
            # warps w and w + 4 share a scheduler: the two slow row blocks of a CTA (first and last) are never blocks 0 and 4
            assert -(-last // 6) != 5 and (n == 2 or -(-rpc // 6) != 5) or R < 60, (R, C, n, first, rpc, last)
    # the cluster-size knob: more, thinner bands keep the same invariants (or the grid is not resident)
    for R in (96, 120, 200, 256):
        for n_req in range(2, 9):
            shape, n, first, rpc = resident_plan(lib, R, 200, cluster=n_req)
            if shape != 5:
                continue
            last = R - first - (n - 2) * rpc
            assert n >= n_req or n == -(-R // 48)
            assert 6 <= first <= 48 and rpc % 6 == 0 and last % 6 == 0 and 6 <= last <= min(48, rpc), (R, n_req, n, first, rpc, last)
    assert resident_plan(lib, 64, 300) == (-1, -1, -1, -1) and resident_plan(lib, 12, 64) == (-1, -1, -1, -1)


def test_edge_reserve_model(lib):
    """fdtd2d_plan_edge_reserve: SMs the wavefront kernel leaves to the edge tiles.  4096^2 fp32 (81 edge tiles next to
    ~158 k weighted rows): the smallest share that gets the tiles through in four rounds; 16384^2 and 8192^2 fp64: the six
    SMs the A/B on the GPU measured (profiles/r2_reserve_ab.txt); nothing below 2 % of the pass (65536^2) or above 25 % (small
    grids); never more than half the SMs."""
    f = lib.fdtd2d_plan_edge_reserve
    assert f(81, 158_000, SM, 8) == 21  # ceil(81 / 21) = 4 rounds of 38 row-times < (158000 / (127 * 8) + 16)
    assert f(301, 2_430_000, SM, 8) == 6 and f(296, 2_430_000, SM, 8) == 6 and f(180, 1_400_000, SM, 8) == 6
    assert f(1180, 38_400_000, SM, 8) == 0 and f(177, 72_000, SM, 8) == 0 and f(0, 100_000, SM, 8) == 0
    for n_edge, rows in ((60, 150_000), (109, 330_000), (149, 620_000), (40, 50_000)):
        r = f(n_edge, rows, SM, 8)
        assert 0 < r <= SM // 2
        # the edge rounds fit under the wavefront's own time on the remaining SMs (or it is the best compromise)
        t_edge, t_wave = -(-n_edge // r) * 38, -(-rows // ((SM - r) * 8)) + 16
        best = min(max(-(-n_edge // q) * 38, -(-rows // ((SM - q) * 8)) + 16) for q in range(1, SM // 2 + 1))
        assert max(t_edge, t_wave) == best


def test_reserved_sms_shorten_nothing_but_the_run_count(lib):
    """4096^2 fp32 with the bench's source and probes: the automatic reserve is on (edge tiles ~12 % of the pass), the runs fit
    the warps of the remaining SMs one each, ring runs are balanced against plain runs with their warm-up rows counted, and
    the coverage invariants hold as for every other plan."""
    R = 4096
    probes = [(0, R // 2, R // 2 + 5)] + [(0, R // 8 * i + 3, R // 8 * i + 7) for i in range(1, 8)]
    pl = plan(lib, 0, R, R, 8, src=[(0, R // 2, R // 2)], probe=probes)
    check(pl)
    t = pl["tasks"]
    rows, ring = t[:, 3] - t[:, 2], t[:, 6] != 0
    weighted = int(rows[~ring].sum() + (rows[ring].astype(np.int64) * 208 // 100).sum())
    r = lib.fdtd2d_plan_edge_reserve(pl["n_edge"], weighted, SM, 8)
    assert r > 0 and len(t) <= (SM - r) * 8
    assert ring.any() and abs((rows[ring].mean() + 16) * 2.2 - (rows[~ring].mean() + 16)) < 0.1 * (rows[~ring].mean() + 16)
