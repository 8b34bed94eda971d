"""GPU parity tests proper: the CUDA path (through the C ABI / ctypes) against
  (a) the golden vectors produced by the real reference (tests/golden/), and
  (b) the CPU oracle (oracle/) on the same seeded inputs.
The bar is BIT-EXACT for fp32 and fp64 (which implies the north-star tolerances: rel-L2 <= 1e-5 in
fp32, <= 1e-12 in fp64, at the final fields and at every probe); rel-L2 is reported on failure."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

DT, DX, FC = 5e-14, 1e-4, 30e9
TOL = {"float32": 1e-5, "float64": 1e-12}  # north-star tolerances (BASELINE.json); we demand 0.0


def rel_l2(a, b):
    d = np.linalg.norm(a.astype(np.float64) - b.astype(np.float64))
    n = np.linalg.norm(b.astype(np.float64))
    return d / n if n else d


def assert_bits(a, b, what):
    assert a.dtype == b.dtype and a.shape == b.shape, what
    if not np.array_equal(a, b):
        bad = np.argwhere(a != b)
        raise AssertionError(f"{what}: {len(bad)} cells differ, first at {bad[0]}, rel-L2 = {rel_l2(a, b):.3e} "
                             f"(north-star tolerance {TOL[a.dtype.name]})")


@pytest.fixture(scope="module")
def fd():
    import fdtd2d_b200

    return fdtd2d_b200


@pytest.fixture(autouse=True, params=["auto", "tiled"])
def engine(request, monkeypatch):
    """Every test runs twice: with the automatic kernel choice (small fp32 grids take the cluster-resident kernel)
    and with that kernel disabled, so the k-step tile kernels keep their coverage of small and ragged grids."""
    if request.param == "tiled":
        monkeypatch.setenv("FDTD2D_NO_RESIDENT", "1")
    else:
        monkeypatch.delenv("FDTD2D_NO_RESIDENT", raising=False)
    return request.param


@pytest.fixture(scope="module")
def oracle():
    from oracle import c_oracle, numpy_oracle

    c_oracle.build()
    return c_oracle, numpy_oracle


# --------------------------------------------------------------------------------------------
# per-function parity against the reference's own outputs
# --------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", ["float32", "float64"])
@pytest.mark.parametrize("shape", [(11, 11), (12, 13), (16, 11), (37, 53), (64, 48)])
def test_update_functions_vs_reference_golden(fd, golden_dir, dtype, shape):
    g = np.load(os.path.join(golden_dir, "single_call.npz"))
    k = f"{dtype}_{shape[0]}x{shape[1]}"
    eps, mu = g[k + "_eps"], g[k + "_mu"]
    Ez, Hx, Hy = g[k + "_Ez0"].copy(), g[k + "_Hx0"].copy(), g[k + "_Hy0"].copy()
    rHx, rHy = fd.update_Hx_Hy(Ez, Hx, Hy, mu, eps, DT, DX)
    assert rHx is Hx and rHy is Hy  # in place + returned, like main.py:76
    assert_bits(Hx, g[k + "_Hx1"], "Hx after update_Hx_Hy")
    assert_bits(Hy, g[k + "_Hy1"], "Hy after update_Hx_Hy")
    assert_bits(Ez, g[k + "_Ez0"], "Ez untouched by update_Hx_Hy")
    rEz = fd.update_Ez(Ez, Hx, Hy, mu, eps, DT, DX)
    assert rEz is Ez
    assert_bits(Ez, g[k + "_Ez1"], "Ez after update_Ez")


# --------------------------------------------------------------------------------------------
# the reference demo (fdtd.py defaults), every k
# --------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", ["float32", "float64"])
@pytest.mark.parametrize("k", [0, 1, 2, 3, 4, 8, 12])  # 0: the library's choice (fp32: cluster-resident; fp64: 12 steps per launch)
def test_demo200_vs_reference_golden(fd, golden_dir, dtype, k):
    g = np.load(os.path.join(golden_dir, f"demo200_vacuum_{dtype}.npz"))
    eps, mu = fd.material_init(None, 200, 200)
    with fd.Simulation(200, 200, dtype, dt=DT, dx=DX) as sim:
        sim.set_materials(eps, mu)
        sim.set_point_source(100, 100, 1000, FC)
        sim.set_probes([tuple(p) for p in g["probes"]], 1000)
        sim.step(1000, k)
        Ez, Hx, Hy = sim.state()
        trace = sim.read_probes()
    assert_bits(trace, g["trace"], "probe trace (every step)")
    assert_bits(Ez, g["Ez"], "Ez")
    assert_bits(Hx, g["Hx"], "Hx")
    assert_bits(Hy, g["Hy"], "Hy")


@pytest.mark.parametrize("dtype", ["float32", "float64"])
@pytest.mark.parametrize("case", [(37, 53, "ricker"), (96, 130, "sinusoidal"), (11, 11, "ricker")])
@pytest.mark.parametrize("k", [1, 4, 7])
def test_random_runs_vs_reference_golden(fd, golden_dir, dtype, case, k):
    """Random eps AND mu, non-zero initial state (incl. the never-updated Hx last row / Hy last
    column), ragged sizes, both waveforms."""
    g = np.load(os.path.join(golden_dir, "random_runs.npz"))
    R, C, kind = case
    key = f"{dtype}_{R}x{C}_{kind}"
    n = int(g[key + "_nsteps"])
    src = tuple(int(v) for v in g[key + "_src"])
    with fd.Simulation(R, C, dtype, dt=DT, dx=DX) as sim:
        sim.set_materials(g[key + "_eps"], g[key + "_mu"])
        sim.set_state(g[key + "_Ez0"], g[key + "_Hx0"], g[key + "_Hy0"])
        sim.set_point_source(src[0], src[1], n, FC, kind)
        sim.set_probes([tuple(p) for p in g[key + "_probes"]], n)
        sim.step(n, k)
        Ez, Hx, Hy = sim.state()
        trace = sim.read_probes()
    assert_bits(trace, g[key + "_trace"], "probe trace")
    assert_bits(Ez, g[key + "_Ez"], "Ez")
    assert_bits(Hx, g[key + "_Hx"], "Hx")
    assert_bits(Hy, g[key + "_Hy"], "Hy")


# --------------------------------------------------------------------------------------------
# seeded inputs against the CPU oracle at sizes it finishes in seconds
# --------------------------------------------------------------------------------------------
def _random_problem(rng, R, C, dtype, scale=1e-3):
    eps = (8.85418e-12 * (1 + 9 * rng.random((R, C)))).astype(dtype)
    mu = (4 * np.pi * 1e-7 * (1 + 0.5 * rng.random((R, C)))).astype(dtype)
    Ez = (scale * rng.standard_normal((R, C))).astype(dtype)
    Hx = (scale * 1e-3 * rng.standard_normal((R, C - 1))).astype(dtype)
    Hy = (scale * 1e-3 * rng.standard_normal((R - 1, C))).astype(dtype)
    return eps, mu, Ez, Hx, Hy


@pytest.mark.parametrize("dtype", ["float32", "float64"])
@pytest.mark.parametrize("shape,nsteps,k", [
    ((300, 517), 64, 4), ((300, 517), 33, 8), ((129, 1000), 50, 5), ((1000, 131), 50, 3),
    ((28, 120), 40, 4), ((29, 121), 40, 4), ((57, 241), 40, 4), ((64, 256), 24, 6),
    ((1024, 1024), 40, 4), ((1024, 1024), 17, 1), ((203, 215), 30, 2),
])
def test_seeded_runs_vs_oracle(fd, oracle, dtype, shape, nsteps, k):
    c_oracle, npo = oracle
    R, C = shape
    rng = np.random.default_rng(R * 100003 + C * 17 + nsteps + k)
    eps, mu, Ez, Hx, Hy = _random_problem(rng, R, C, dtype)
    ce, ch, coef = c_oracle.coefficients(eps, mu, DT, DX, np.dtype(dtype))
    cells = [(R // 2, C // 2), (7, 9), (R - 3, C - 2), (0, 0)]
    amp = npo.source_table("ricker", nsteps, DT, FC)
    probes = [(R // 2, C // 2), (0, 0), (R - 1, C - 1), (4, C - 5), (R - 5, 4), (R // 3, C // 5), (5, 5), (R - 6, C - 6)]
    oEz, oHx, oHy = Ez.copy(), Hx.copy(), Hy.copy()
    otrace = c_oracle.run(oEz, oHx, oHy, ce, ch, coef, nsteps, amp, cells, probes, omp=True)
    with fd.Simulation(R, C, dtype, dt=DT, dx=DX) as sim:
        sim.set_materials(eps, mu)
        gce, gch, gmur = sim.coefficients()
        assert_bits(gce, ce, "ce = dt/(eps*dx) formed on the device")
        assert_bits(gch, ch, "ch = dt/(mu*dx) formed on the device")
        assert gmur[0] == coef and gmur.dtype == np.dtype(dtype)
        sim.set_state(Ez, Hx, Hy)
        sim.set_sources([(0, r, c, 0) for r, c in cells], amp[None, :])
        sim.set_probes(probes, nsteps)
        sim.step(nsteps, k)
        gEz, gHx, gHy = sim.state()
        gtrace = sim.read_probes()
    assert_bits(gtrace, otrace, "probe trace")
    assert_bits(gEz, oEz, "Ez")
    assert_bits(gHx, oHx, "Hx")
    assert_bits(gHy, oHy, "Hy")


@pytest.mark.parametrize("dtype", ["float32", "float64"])
def test_batched_grids_vs_oracle(fd, oracle, dtype):
    """Batched mode: independent grids, per-grid media, Mur coefficient, sources (point and line)."""
    c_oracle, npo = oracle
    B, R, C, nsteps = 6, 72, 100, 48
    rng = np.random.default_rng(5)
    probs = [_random_problem(rng, R, C, dtype, scale=0.0) for _ in range(B)]
    eps = np.stack([p[0] for p in probs])
    mu = np.stack([p[1] for p in probs])
    fcs = [18e9 + 2e9 * b for b in range(B)]
    tables = np.stack([npo.source_table("ricker" if b % 2 == 0 else "sinusoidal", nsteps, DT, fcs[b]) for b in range(B)])
    cells, per_grid = [], []
    for b in range(B):
        if b % 3 == 0:
            mine = [(20 + b, 30 + j) for j in range(7)]  # horizontal line source
        elif b % 3 == 1:
            mine = [(10 + i, 50 + b) for i in range(5)]  # vertical line source
        else:
            mine = [(R // 2, C // 2)]
        per_grid.append(mine)
        cells += [(b, r, c, b) for r, c in mine]
    probes = [(b, R // 2, C // 2 + 3) for b in range(B)] + [(b, 1, 1) for b in range(B)]
    with fd.Simulation(R, C, dtype, dt=DT, dx=DX, batch=B) as sim:
        sim.set_materials(eps, mu)
        sim.set_sources(cells, tables)
        sim.set_probes(probes, nsteps)
        sim.step(nsteps, 4)
        gEz, gHx, gHy = sim.state()
        gtrace = sim.read_probes()
    for b in range(B):
        ce, ch, coef = c_oracle.coefficients(eps[b], mu[b], DT, DX, np.dtype(dtype))
        oEz, oHx, oHy = npo.grid_init(R, C, np.dtype(dtype))
        otr = c_oracle.run(oEz, oHx, oHy, ce, ch, coef, nsteps, tables[b], per_grid[b], [(R // 2, C // 2 + 3), (1, 1)])
        assert_bits(gEz[b], oEz, f"Ez grid {b}")
        assert_bits(gHx[b], oHx, f"Hx grid {b}")
        assert_bits(gHy[b], oHy, f"Hy grid {b}")
        assert_bits(gtrace[:, b], otr[:, 0], f"probe A grid {b}")
        assert_bits(gtrace[:, B + b], otr[:, 1], f"probe B grid {b}")


def test_step_in_pieces_equals_one_call(fd):
    """Stepping 1000 = 1+2+...: the step index (source phase) carries across calls and mixed k."""
    eps, mu = fd.material_init(None, 200, 200)
    outs = []
    for plan in ([(1000, 4)], [(1, 1), (2, 2), (333, 4), (64, 8), (600, 5)]):
        with fd.Simulation(200, 200, np.float32, dt=DT, dx=DX) as sim:
            sim.set_materials(eps, mu)
            sim.set_point_source(100, 100, 1000, FC)
            for n, k in plan:
                sim.step(n, k)
            assert sim.step_index == 1000
            outs.append(sim.state())
    for a, b in zip(*outs):
        assert_bits(a, b, "piecewise stepping")


# --------------------------------------------------------------------------------------------
# size-independent properties at sizes the oracle cannot reach
# --------------------------------------------------------------------------------------------
def test_temporal_blocking_invariance_large(fd):
    """k steps per HBM round trip must not change a single bit: 3000 x 4100 fp32, device-generated
    random medium, k in {1, 3, 4, 8}."""
    R, C, n = 3000, 4100, 24
    ref = None
    for k in (1, 3, 4, 8):
        with fd.Simulation(R, C, np.float32, dt=DT, dx=DX) as sim:
            sim.set_materials_random(seed=11, span=9.0)
            sim.set_point_source(R // 2, C // 2, 700, FC)
            sim.step_index = 640  # near the Ricker peak so the field is far from zero
            sim.step(n, k)
            Ez, Hx, Hy = sim.state()
        if ref is None:
            ref = (Ez, Hx, Hy)
            assert np.abs(Ez).max() > 0.1
        else:
            for a, b in zip((Ez, Hx, Hy), ref):
                assert_bits(a, b, f"k={k} vs k=1")


@pytest.mark.parametrize("k", [4, 8])
def test_window_locality_vs_oracle_large(fd, oracle, k):
    """Full-size check through locality: after n steps a window of a 4096 x 4096 fp32 run equals the
    oracle run on the window grown by n+8 cells (interior dependency radius is 1 cell per step).  k = 4 runs on the
    persistent TMA tiles, k = 8 on the wavefront strips."""
    c_oracle, npo = oracle
    R = C = 4096
    n, m = 20, 28
    rng = np.random.default_rng(3)
    with fd.Simulation(R, C, np.float32, dt=DT, dx=DX) as sim:
        sim.set_materials_random(seed=2026, span=9.0)
        Ez0 = (1e-3 * rng.standard_normal((R, C))).astype(np.float32)
        Hx0 = (1e-6 * rng.standard_normal((R, C - 1))).astype(np.float32)
        Hy0 = (1e-6 * rng.standard_normal((R - 1, C))).astype(np.float32)
        sim.set_state(Ez0, Hx0, Hy0)
        ce, ch, _ = sim.coefficients()
        assert (sim.plan_info(k)["wave_runs"] > 0) == (k == 8)
        sim.step(n, k)
        Ez, Hx, Hy = sim.state()
    for (r0, c0, h, w) in [(1000, 2000, 96, 160), (3000, 100, 64, 64), (2040, 2040, 40, 40)]:
        a, b_, c_, d = r0 - m, r0 + h + m, c0 - m, c0 + w + m
        wEz = Ez0[a:b_, c_:d].copy()
        wHx = Hx0[a:b_, c_:d - 1].copy()
        wHy = Hy0[a:b_ - 1, c_:d].copy()
        c_oracle.run(wEz, wHx, wHy, ce[a:b_, c_:d].copy(), ch[a:b_, c_:d].copy(), np.float32(0), n)
        assert_bits(Ez[r0:r0 + h, c0:c0 + w], wEz[m:m + h, m:m + w], "Ez window")
        assert_bits(Hx[r0:r0 + h, c0:c0 + w], wHx[m:m + h, m:m + w], "Hx window")
        assert_bits(Hy[r0:r0 + h, c0:c0 + w], wHy[m:m + h, m:m + w], "Hy window")


def test_random_medium_matches_host_definition(fd):
    """The device-generated medium is reproducible on the host (fdtd2d_hash_uniform)."""
    from fdtd2d_b200 import _lib

    R, C = 40, 70
    for dtype in (np.float32, np.float64):
        with fd.Simulation(R, C, dtype, dt=DT, dx=DX, batch=2) as sim:
            sim.set_materials_random(seed=77, span=9.0)
            ce, ch, mur = sim.coefficients()
        T = np.dtype(dtype).type
        for b in range(2):
            u = np.array([[_lib.lib().fdtd2d_hash_uniform(77, b, i, j) for j in range(C)] for i in range(R)]).astype(dtype)
            eps = T(8.85418e-12) * (T(1) + T(9.0) * u)
            mu = np.full((R, C), T(4 * np.pi * 1e-7))
            assert_bits(ce[b], DT / (eps * DX), "ce of the synthetic medium")
            assert_bits(ch[b], DT / (mu * DX), "ch of the synthetic medium")
            c = 1 / np.sqrt(mu[0, 0] * eps[0, 0])
            assert mur[b] == (c * DT - DX) / (c * DT + DX)


# --------------------------------------------------------------------------------------------
# error behaviour at the boundary
# --------------------------------------------------------------------------------------------
def test_errors(fd):
    with pytest.raises(fd.Fdtd2dError) as e:
        fd.Simulation(5, 200, np.float32, dt=DT, dx=DX)  # below 6 the reference itself indexes out of range
    assert e.value.code == -1
    with fd.Simulation(32, 32, np.float32, dt=DT, dx=DX) as sim:
        with pytest.raises(fd.Fdtd2dError) as e:
            sim.step(1)
        assert e.value.code == -4  # materials not set
        sim.set_materials(*fd.material_init(None, 32, 32))
        with pytest.raises(fd.Fdtd2dError):
            sim.step(1, 13)  # k > FDTD2D_MAX_K
        with pytest.raises(fd.Fdtd2dError):
            sim.set_sources([(0, 40, 3, 0)], np.zeros((1, 4)))  # outside the grid
        with pytest.raises(fd.Fdtd2dError):
            sim.set_sources([(0, 4, 3, 0), (0, 4, 3, 0)], np.zeros((1, 4)))  # duplicate cell
        with pytest.raises(ValueError):
            sim.set_state(np.zeros((32, 32), np.float32), np.zeros((32, 32), np.float32), np.zeros((31, 32), np.float32))
        sim.set_point_source(16, 16, 10, FC)
        sim.step(10)
        with pytest.raises(ValueError):  # the reference's source never ends; a table does: stepping past it is refused
            sim.step(1)
        sim.step(1, strict=False)
        with pytest.raises(fd.Fdtd2dError):
            sim.set_option("no_such_option", 1)
    with pytest.raises(ValueError):
        fd.update_Hx_Hy(np.zeros((20, 20)), np.zeros((20, 20)), np.zeros((19, 20)), np.ones((20, 20)), np.ones((20, 20)), DT, DX)


# --------------------------------------------------------------------------------------------
# ragged sizes: every tiling corner case (last-tile remainders, single tiles, thin grids)
# --------------------------------------------------------------------------------------------
def _ragged_cases():
    rng = np.random.default_rng(12345)
    cases = [(11, 11, 3, 5), (11, 300, 8, 9), (300, 11, 8, 9), (49, 113, 8, 8), (57, 121, 8, 9), (48, 112, 8, 8),
             (96, 224, 8, 16), (97, 225, 8, 10), (104, 232, 8, 9), (65, 129, 1, 3), (2 * 48 + 7, 2 * 112 + 7, 8, 8)]
    for _ in range(14):
        cases.append((int(rng.integers(11, 400)), int(rng.integers(11, 700)), int(rng.integers(1, 9)), int(rng.integers(1, 20))))
    return cases


@pytest.mark.parametrize("dtype", ["float32", "float64"])
def test_ragged_sizes_vs_oracle(fd, oracle, dtype):
    c_oracle, npo = oracle
    for (R, C, k, nsteps) in _ragged_cases():
        rng = np.random.default_rng(R * 7919 + C * 31 + k)
        eps, mu, Ez, Hx, Hy = _random_problem(rng, R, C, dtype)
        ce, ch, coef = c_oracle.coefficients(eps, mu, DT, DX, np.dtype(dtype))
        cells = [(R // 2, C // 2), (R - 1, 0), (0, C - 1)]
        amp = npo.source_table("sinusoidal", nsteps, DT, FC)
        probes = [(0, 0), (R - 1, C - 1), (R // 2, C // 2), (5, C - 6), (R - 6, 5)]
        oEz, oHx, oHy = Ez.copy(), Hx.copy(), Hy.copy()
        otrace = c_oracle.run(oEz, oHx, oHy, ce, ch, coef, nsteps, amp, cells, probes)
        with fd.Simulation(R, C, dtype, dt=DT, dx=DX) as sim:
            sim.set_coefficients(ce, ch, coef)  # host-precomputed maps this time
            sim.set_state(Ez, Hx, Hy)
            sim.set_sources([(0, r, c, 0) for r, c in cells], amp[None, :])
            sim.set_probes(probes, nsteps)
            sim.step(nsteps, k)
            gEz, gHx, gHy = sim.state()
            gtrace = sim.read_probes()
        what = f"{R}x{C} k={k} n={nsteps}"
        assert_bits(gtrace, otrace, "probe trace " + what)
        assert_bits(gEz, oEz, "Ez " + what)
        assert_bits(gHx, oHx, "Hx " + what)
        assert_bits(gHy, oHy, "Hy " + what)


@pytest.mark.parametrize("variant", [1, 3])
def test_kernel_variants_agree(fd, variant):
    """The shared-memory generic kernel (variant 1: everywhere, 3: edge tiles only) and the default
    register-resident kernels produce the same bits on a grid with many edge, source and probe tiles."""
    R, C, n = 700, 1500, 27
    outs = []
    for v in (0, variant):
        with fd.Simulation(R, C, np.float32, dt=DT, dx=DX) as sim:
            sim.set_kernel_variant(v)
            sim.set_materials_random(3, 9.0)
            amp = fd.source_table("ricker", 700, DT, FC)
            sim.set_sources([(0, 350, 750, 0), (0, 100, 100, 0), (0, 640, 1400, 0)], amp[None, :])
            sim.set_probes([(350, 760), (3, 3), (696, 1496), (64, 128)], 700)
            sim.step_index = 640
            sim.step(n, 8)
            outs.append(sim.state() + (sim.read_probes(640, n),))
    for a, b in zip(*outs):
        assert_bits(a, b, f"variant {variant} vs default")


# --------------------------------------------------------------------------------------------
# the row-streaming wavefront kernel (strip_wave.cuh): k = 8 / k = 12 passes of grids with many plain tiles
# --------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape,nsteps", [((300, 517), 40), ((1024, 1024), 24), ((203, 600), 17), ((2000, 260), 32),
                                          ((700, 1500), 8)])
@pytest.mark.parametrize("uniform_mu", [False, True])
@pytest.mark.parametrize("dtype,k", [("float32", 8), ("float32", 12), ("float64", 4), ("float64", 6), ("float64", 8)])  # (fp64 k = 6: tile kernel)
def test_wavefront_kernel_vs_oracle(fd, oracle, shape, nsteps, uniform_mu, dtype, k, monkeypatch):
    """Forced onto small grids (FDTD2D_WAVE_MIN_TILES=0) so the oracle can check it: runs of plain tiles broken by
    sources and probes, ragged sizes, the remainder pass (nsteps % k) on the tile kernel.  fp32: the packed (FADD2 /
    FFMA2) kernel with 8 levels and, for uniform permeability, 12 (otherwise those passes run on the tile kernels, which
    is checked all the same); fp64: the scalar kernel on 64-column strips with 4 and 8 levels (ring strips at 8)."""
    monkeypatch.setenv("FDTD2D_WAVE_MIN_TILES", "0")
    monkeypatch.setenv("FDTD2D_RING_MIN_TILES", "0")  # ring strips (left / right Mur ring on the wavefront) where C >= 512
    c_oracle, npo = oracle
    R, C = shape
    rng = np.random.default_rng(R * 31 + C)
    eps, mu, Ez, Hx, Hy = _random_problem(rng, R, C, dtype)
    if uniform_mu:  # every material_init output: the kernel then takes dt/(mu*dx) as a scalar and skips the map
        mu[...] = np.dtype(dtype).type(4 * np.pi * 1e-7)
    ce, ch, coef = c_oracle.coefficients(eps, mu, DT, DX, np.dtype(dtype))
    cells = [(R // 2, C // 2), (R // 3, C // 4), (7, 9)]
    amp = npo.source_table("ricker", nsteps, DT, FC) + 0.125
    probes = [(R // 2, C // 2 + 3), (0, 0), (R - 1, C - 1), (R // 4, C // 3), (3 * R // 4, 2 * C // 3)]
    oEz, oHx, oHy = Ez.copy(), Hx.copy(), Hy.copy()
    otrace = c_oracle.run(oEz, oHx, oHy, ce, ch, coef, nsteps, amp, cells, probes, omp=True)
    with fd.Simulation(R, C, dtype, dt=DT, dx=DX) as sim:
        sim.set_kernel_variant(2)  # tile kernels (the cluster-resident kernel would take the small ones)
        sim.set_coefficients(ce, ch, coef)
        sim.set_state(Ez, Hx, Hy)
        sim.set_sources([(0, r, c, 0) for r, c in cells], amp[None, :])
        sim.set_probes(probes, nsteps)
        info = sim.plan_info(k)
        if nsteps >= k and (k == 8 or (dtype == "float64" and k == 4) or (dtype == "float32" and uniform_mu)) and min(R, C) >= 260:
            assert info["wave_runs"] > 0, info  # the kernel under test really runs
            if k == 8:
                assert (info["ring_strips"] == 1) == (C >= 512), info  # the left / right ring rides along (fp32 and fp64)
        sim.step(nsteps, k)
        gEz, gHx, gHy = sim.state()
        gtrace = sim.read_probes()
    assert_bits(gtrace, otrace, "probe trace")
    assert_bits(gEz, oEz, "Ez")
    assert_bits(gHx, oHx, "Hx")
    assert_bits(gHy, oHy, "Hy")


def test_wavefront_batched_grids_vs_oracle(fd, oracle, monkeypatch):
    """Three independent 400 x 900 grids in one handle on the wavefront kernel (ring strips included): per-grid media and
    Mur coefficient, sources and probes in different places, each grid against the oracle."""
    monkeypatch.setenv("FDTD2D_WAVE_MIN_TILES", "0")
    monkeypatch.setenv("FDTD2D_RING_MIN_TILES", "0")
    c_oracle, npo = oracle
    B, R, C, nsteps = 3, 400, 900, 24
    rng = np.random.default_rng(21)
    probs = [_random_problem(rng, R, C, "float32") for _ in range(B)]
    for p_ in probs:
        p_[1][...] = np.float32(4 * np.pi * 1e-7)
    eps, mu, Ez, Hx, Hy = (np.stack([p_[i] for p_ in probs]) for i in range(5))
    tables = np.stack([npo.source_table("ricker", nsteps, DT, FC + 1e9 * b) + 0.25 for b in range(B)])
    per_grid = [[(R // 2 + 10 * b, C // 2 - 50 * b)] for b in range(B)]
    cells = [(b, r, c, b) for b in range(B) for r, c in per_grid[b]]
    pcell = [(R // 3, C // 3 + 7), (R // 2, 6)]
    probes = [(b, r, c) for r, c in pcell for b in range(B)]
    with fd.Simulation(R, C, np.float32, dt=DT, dx=DX, batch=B) as sim:
        sim.set_kernel_variant(2)
        sim.set_materials(eps, mu)
        sim.set_state(Ez, Hx, Hy)
        sim.set_sources(cells, tables)
        sim.set_probes(probes, nsteps)
        sim.step(nsteps, 8)
        gEz, gHx, gHy = sim.state()
        gtrace = sim.read_probes()
    for b in range(B):
        ce, ch, coef = c_oracle.coefficients(eps[b], mu[b], DT, DX, np.dtype(np.float32))
        oEz, oHx, oHy = Ez[b].copy(), Hx[b].copy(), Hy[b].copy()
        otr = c_oracle.run(oEz, oHx, oHy, ce, ch, coef, nsteps, tables[b], per_grid[b], pcell, omp=True)
        assert_bits(gEz[b], oEz, f"Ez grid {b}")
        assert_bits(gHx[b], oHx, f"Hx grid {b}")
        assert_bits(gHy[b], oHy, f"Hy grid {b}")
        assert_bits(gtrace[:, b], otr[:, 0], f"probe A grid {b}")
        assert_bits(gtrace[:, B + b], otr[:, 1], f"probe B grid {b}")


def test_wavefront_equals_tile_kernel_large(fd, monkeypatch):
    """6000 x 5000 fp32, 40 steps near the Ricker peak: the wavefront strips (8 and 12 levels, with and without the ring
    strips, and whatever k = 0 picks) and the persistent TMA tile kernel (wavefront option off) must agree bit for bit."""
    outs = []
    for wavefront, ring, k in ((0, 1, 8), (1, 1, 8), (1, 0, 8), (1, 1, 12), (1, 1, 0)):
        with fd.Simulation(6000, 5000, np.float32, dt=DT, dx=DX) as sim:
            sim.set_option("wavefront", wavefront)  # per-handle options instead of environment variables
            sim.set_option("ring_strips", ring)
            sim.set_materials_random(seed=5, span=9.0)
            sim.set_point_source(3000, 2500, 700, FC)
            sim.set_probes([(3000, 2510), (10, 10), (5990, 4990)], 700)
            sim.step_index = 640
            sim.step(40, k)
            outs.append(sim.state() + (sim.read_probes(640, 40),))
    assert np.abs(outs[0][0]).max() > 0.1
    for o in outs[1:]:
        for a, b in zip(outs[0], o):
            assert_bits(a, b, "wavefront vs tile kernel")


def test_two_handles_in_two_host_threads(fd, monkeypatch):
    """Two handles driven from two host threads at the same time (what bench.py's e2e does: one job's copies overlap
    the other's kernels): upload, step, download in a loop; both must give what a single handle gives alone.  The
    wavefront kernel is forced on so that its task list (uploaded when the plan is built) is in play."""
    import threading

    monkeypatch.setenv("FDTD2D_WAVE_MIN_TILES", "0")
    monkeypatch.setenv("FDTD2D_RING_MIN_TILES", "0")
    R, C, n = 1500, 2100, 400
    rng = np.random.default_rng(11)
    eps, mu, Ez, Hx, Hy = _random_problem(rng, R, C, "float32")
    mu[...] = np.float32(4 * np.pi * 1e-7)

    def job(sim):
        sim.step_index = 0
        sim.set_materials(eps, mu)
        sim.set_state(Ez, Hx, Hy)
        sim.step(n, 0)
        return sim.state() + (sim.read_probes(0, n),)

    def make():
        sim = fd.Simulation(R, C, np.float32, dt=DT, dx=DX)
        sim.set_kernel_variant(2)
        sim.set_point_source(R // 2, C // 2, n, FC)
        sim.set_probes([(R // 2, C // 2 + 16), (R // 4, C // 4), (R // 2, 8)], n)
        return sim

    with make() as sim:
        ref = job(sim)
    sims, res, errs = [make(), make()], [None, None], []

    def work(w):
        try:
            for _ in range(4):
                res[w] = job(sims[w])
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    ts = [threading.Thread(target=work, args=(w,)) for w in range(2)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    for s_ in sims:
        s_.close()
    assert not errs, errs
    for r in res:
        for a, b in zip(r, ref):
            assert_bits(a, b, "two threads vs one handle alone")


# --------------------------------------------------------------------------------------------
# the default path at the size the bench runs it
# --------------------------------------------------------------------------------------------
def _window_check(c_oracle, state0, state, ce, ch, coef, n, windows, m, what):
    """A window of a big run equals the oracle run on the window grown by m >= n + 6 cells.  Windows that touch the left
    or right edge of the grid keep that edge (Mur columns included: they read only inward); the other sides are cut m
    cells outside the window, far enough for nothing stale to arrive."""
    Ez0, Hx0, Hy0 = state0
    Ez, Hx, Hy = state
    R, C = Ez0.shape
    for (r0, c0, h, w) in windows:
        a, b_ = r0 - m, r0 + h + m
        c_, d = max(0, c0 - m), min(C, c0 + w + m)
        wEz, wHx, wHy = Ez0[a:b_, c_:d].copy(), Hx0[a:b_, c_:d - 1].copy(), Hy0[a:b_ - 1, c_:d].copy()
        c_oracle.run(wEz, wHx, wHy, ce[a:b_, c_:d].copy(), ch[a:b_, c_:d].copy(), coef, n)
        ro, co = r0 - a, c0 - c_
        assert_bits(Ez[r0:r0 + h, c0:c0 + w], wEz[ro:ro + h, co:co + w], f"Ez window {what} at {(r0, c0)}")
        wx = min(w, Hx.shape[1] - c0)
        assert_bits(Hx[r0:r0 + h, c0:c0 + wx], wHx[ro:ro + h, co:co + wx], f"Hx window {what} at {(r0, c0)}")
        assert_bits(Hy[r0:r0 + h, c0:c0 + w], wHy[ro:ro + h, co:co + w], f"Hy window {what} at {(r0, c0)}")


def test_default_path_at_bench_size_vs_oracle(fd, oracle, engine):
    """8192 x 16384 fp32, default options, k_temporal = 0, 24 steps from a random state: what bench.py's cfg3 runs (the
    packed wavefront with ring strips, automatic run lengths).  Three windows against the C oracle: one in the middle,
    one on the left ring strip (Mur columns included), one straddling the boundary between two runs of rows."""
    if engine == "tiled":
        pytest.skip("one kernel choice is enough for a 8192 x 16384 grid")
    c_oracle, npo = oracle
    R, C, n, m = 8192, 16384, 24, 32
    rng = np.random.default_rng(8)
    with fd.Simulation(R, C, np.float32, dt=DT, dx=DX) as sim:
        sim.set_materials_random(seed=2026, span=9.0)
        Ez0 = (1e-3 * rng.standard_normal((R, C), dtype=np.float32))
        Hx0 = (1e-6 * rng.standard_normal((R, C - 1), dtype=np.float32))
        Hy0 = (1e-6 * rng.standard_normal((R - 1, C), dtype=np.float32))
        sim.set_state(Ez0, Hx0, Hy0)
        ce, ch, mur = sim.coefficients()
        info = sim.plan_info(8)
        assert info["wave_runs"] > 1000 and info["ring_strips"] == 1 and info["tma_tiles"] == 0, info
        sim.step(n, 0)
        assert sim.pass_count == 3
        Ez, Hx, Hy = sim.state()
    # run boundaries: the planner cuts the ~8000 plain rows into equal runs; rows 2000..2100 and 4000..4100 hold several
    windows = [(5000, 9000, 64, 96), (3000, 0, 64, 40), (2000, 12000, 100, 64), (4000, 16384 - 40, 100, 40)]
    _window_check(c_oracle, (Ez0, Hx0, Hy0), (Ez, Hx, Hy), ce, ch, mur[0], n, windows, m, "default path")


class _DeviceView:
    """numpy-style window access to a field the library holds in HBM (fdtd2d_device_field), for grids whose state does
    not fit the host: rows x pitch elements, wrapped through __cuda_array_interface__."""

    def __init__(self, sim, field):
        import torch

        self.__cuda_array_interface__ = {"shape": (sim.local_rows, sim.pitch), "typestr": np.dtype(sim.dtype).str,
                                         "data": (sim.device_field(field), False), "version": 3}
        self.t = torch.as_tensor(self, device=f"cuda:{sim.device}")

    def get(self, r0, r1, c0, c1):
        return self.t[r0:r1, c0:c1].cpu().numpy()

    def put(self, r0, c0, a):
        import torch

        self.t[r0:r0 + a.shape[0], c0:c0 + a.shape[1]] = torch.from_numpy(a).to(self.t.device)


@pytest.mark.parametrize("R,C", [(16384, 16384), (65536, 65536)], ids=["cfg3_16384sq", "cfg4_65536sq"])
def test_baseline_sizes_on_one_gpu_vs_oracle_windows(fd, oracle, engine, R, C):
    """BASELINE configs[2] and configs[3] at their FULL sizes on one GPU (65536^2: 137 GB of HBM, element offsets past
    2^31), default options, 24 steps = 3 passes.  The state is zero except for random patches written straight into HBM
    around eight windows -- the four corners, the four edges away from the corners (top / bottom ring tiles, left / right
    ring strips), two interior spots far into the index range -- and each window is compared with the C oracle run on
    its patch (cut sides lie n + 8 cells outside the window; true edges of the grid stay edges)."""
    import torch

    if engine == "tiled":
        pytest.skip("one kernel choice is enough at this size")
    need = (8 * R * C * 4) * 1.05
    if torch.cuda.get_device_properties(0).total_memory < need:
        pytest.skip(f"needs {need / 2**30:.0f} GiB of HBM")
    c_oracle, npo = oracle
    n, m = 24, 32
    rng = np.random.default_rng(R)
    windows = [(0, 0, 48, 48), (0, C - 48, 48, 48), (R - 48, 0, 48, 48), (R - 48, C - 48, 48, 48),
               (0, C // 2 + 1000, 56, 96), (R - 56, C // 3, 56, 96), (R // 2 + 777, 0, 64, 40), (R // 3, C - 40, 64, 40),
               (R - 5000, C - 7000, 64, 96), (R // 2 - 20, C // 2 - 30, 64, 64)]
    with fd.Simulation(R, C, np.float32, dt=DT, dx=DX) as sim:
        sim.set_materials_random(seed=2026, span=9.0)
        sim.set_point_source(R // 2, C // 2, 40, FC)
        sim.zero_state()
        sim.synchronize()
        views = [_DeviceView(sim, f) for f in range(5)]
        patches = []
        for (r0, c0, h, w) in windows:
            a, b_, c_, d = max(0, r0 - m), min(R, r0 + h + m), max(0, c0 - m), min(C, c0 + w + m)
            Ez0 = (1e-3 * rng.standard_normal((b_ - a, d - c_))).astype(np.float32)
            Hx0 = (1e-6 * rng.standard_normal((b_ - a, d - c_ - 1))).astype(np.float32)
            Hy0 = (1e-6 * rng.standard_normal((b_ - a - 1, d - c_))).astype(np.float32)
            views[0].put(a, c_, Ez0), views[1].put(a, c_, Hx0), views[2].put(a, c_, Hy0)
            patches.append((a, b_, c_, d, Ez0, Hx0, Hy0, views[3].get(a, b_, c_, d), views[4].get(a, b_, c_, d)))
        torch.cuda.synchronize()
        info = sim.plan_info(8)
        assert info["wave_runs"] > 1000 and info["ring_strips"] == 1, info
        sim.step(n, 0)
        assert sim.pass_count == 3
        sim.synchronize()
        out = [_DeviceView(sim, f) for f in range(3)]  # (the current set has changed)
        u00 = np.float32(sim_hash(fd, 2026, 0, 0, 0))
        eps00 = np.float32(8.85418e-12) * (np.float32(1) + np.float32(9.0) * u00)
        cc = 1 / np.sqrt(np.float32(4 * np.pi * 1e-7) * eps00)
        coef = np.float32((cc * DT - DX) / (cc * DT + DX))
        amp = npo.source_table("ricker", n, DT, FC)
        for (r0, c0, h, w), (a, b_, c_, d, Ez0, Hx0, Hy0, ce, ch) in zip(windows, patches):
            src = [(R // 2 - a, C // 2 - c_)] if a <= R // 2 < b_ and c_ <= C // 2 < d else []
            c_oracle.run(Ez0, Hx0, Hy0, ce, ch, coef, n, amp if src else None, src, [])
            ro, co = r0 - a, c0 - c_
            assert_bits(out[0].get(r0, r0 + h, c0, c0 + w), Ez0[ro:ro + h, co:co + w], f"Ez window at {(r0, c0)}")
            wx, hy = min(w, C - 1 - c0), min(h, R - 1 - r0)
            assert_bits(out[1].get(r0, r0 + h, c0, c0 + wx), Hx0[ro:ro + h, co:co + wx], f"Hx window at {(r0, c0)}")
            assert_bits(out[2].get(r0, r0 + hy, c0, c0 + w), Hy0[ro:ro + hy, co:co + w], f"Hy window at {(r0, c0)}")
            assert np.count_nonzero(Ez0[ro:ro + h, co:co + w]) > h * w // 2


def sim_hash(fd, seed, grid, row, col):
    from fdtd2d_b200 import _lib

    return _lib.lib().fdtd2d_hash_uniform(seed, grid, row, col)


def test_fp64_wavefront_large_vs_oracle_windows(fd, oracle, engine):
    """4096 x 4096 fp64 with the default choice (8 levels on 64-column strips): windows against the C oracle, and the
    whole state against the tile kernel (wavefront option off)."""
    if engine == "tiled":
        pytest.skip("one kernel choice is enough")
    c_oracle, npo = oracle
    R = C = 4096
    n, m = 16, 24
    rng = np.random.default_rng(9)
    Ez0 = 1e-3 * rng.standard_normal((R, C))
    Hx0 = 1e-6 * rng.standard_normal((R, C - 1))
    Hy0 = 1e-6 * rng.standard_normal((R - 1, C))
    outs = []
    for wavefront in (1, 0):
        with fd.Simulation(R, C, np.float64, dt=DT, dx=DX) as sim:
            sim.set_option("wavefront", wavefront)
            sim.set_materials_random(seed=31, span=9.0)
            sim.set_state(Ez0, Hx0, Hy0)
            if wavefront:
                ce, ch, mur = sim.coefficients()
                assert sim.plan_info(8)["wave_runs"] > 1000
            sim.step(n, 0)
            assert sim.pass_count == (2 if wavefront else 4)
            outs.append(sim.state())
    for a, b in zip(*outs):
        assert_bits(a, b, "fp64 wavefront vs tile kernel")
    _window_check(c_oracle, (Ez0, Hx0, Hy0), outs[0], ce, ch, mur[0], n, [(2000, 2000, 64, 96), (1000, 0, 48, 40)], m, "fp64 wavefront")


# --------------------------------------------------------------------------------------------
# grids below 11 rows / columns: the reference's statement order (grid_small.cuh) against the reference's own outputs
# --------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", ["float32", "float64"])
def test_small_grids_vs_reference_golden(fd, golden_dir, dtype):
    g = np.load(os.path.join(golden_dir, "small_grids.npz"))
    for R, C in (tuple(int(v) for v in sz) for sz in g["sizes"]):
        k = f"{dtype}_{R}x{C}"
        eps, mu = g[k + "_eps"], g[k + "_mu"]
        Ez, Hx, Hy = g[k + "_Ez0"].copy(), g[k + "_Hx0"].copy(), g[k + "_Hy0"].copy()
        fd.update_Hx_Hy(Ez, Hx, Hy, mu, eps, DT, DX)
        assert_bits(Hx, g[k + "_Hx1"], f"Hx after update_Hx_Hy {R}x{C}")
        assert_bits(Hy, g[k + "_Hy1"], f"Hy after update_Hx_Hy {R}x{C}")
        fd.update_Ez(Ez, Hx, Hy, mu, eps, DT, DX)
        assert_bits(Ez, g[k + "_Ez1"], f"Ez after update_Ez {R}x{C}")
        src = tuple(int(v) for v in g[k + "_src"])
        for pieces in ((30,), (1, 7, 22)):
            with fd.Simulation(R, C, dtype, dt=DT, dx=DX) as sim:
                sim.set_materials(eps, mu)
                sim.set_state(g[k + "_Ez0"], g[k + "_Hx0"], g[k + "_Hy0"])
                sim.set_point_source(src[0], src[1], 680, FC)
                sim.set_probes([tuple(p) for p in g[k + "_probes"]], 680)
                sim.step_index = 650
                for n in pieces:
                    sim.step(n, 3)  # (k_temporal does not apply to these grids)
                Ez, Hx, Hy = sim.state()
                trace = sim.read_probes(650, 30)
            assert_bits(trace, g[k + "_trace"], f"probe trace {R}x{C}")
            assert_bits(Ez, g[k + "_Ez"], f"Ez {R}x{C}")
            assert_bits(Hx, g[k + "_Hx"], f"Hx {R}x{C}")
            assert_bits(Hy, g[k + "_Hy"], f"Hy {R}x{C}")
    fd.release_handles()


def test_small_grids_batched_vs_oracle(fd, oracle):
    c_oracle, npo = oracle
    B, R, C, n = 5, 9, 14, 25
    rng = np.random.default_rng(77)
    probs = [_random_problem(rng, R, C, "float32") for _ in range(B)]
    eps, mu, Ez, Hx, Hy = (np.stack([p_[i] for p_ in probs]) for i in range(5))
    tables = np.stack([npo.source_table("ricker", n, DT, FC + 1e9 * b) + 0.5 for b in range(B)])
    with fd.Simulation(R, C, np.float32, dt=DT, dx=DX, batch=B) as sim:
        sim.set_materials(eps, mu)
        sim.set_state(Ez, Hx, Hy)
        sim.set_sources([(b, 4, 5 + b, b) for b in range(B)], tables)
        sim.set_probes([(b, 4, 6) for b in range(B)], n)
        sim.step(n)
        gEz, gHx, gHy = sim.state()
        gtrace = sim.read_probes()
    for b in range(B):
        ce, ch, coef = c_oracle.coefficients(eps[b], mu[b], DT, DX, np.dtype(np.float32))
        oEz, oHx, oHy = Ez[b].copy(), Hx[b].copy(), Hy[b].copy()
        otr = c_oracle.run(oEz, oHx, oHy, ce, ch, coef, n, tables[b], [(4, 5 + b)], [(4, 6)])
        assert_bits(gEz[b], oEz, f"Ez grid {b}")
        assert_bits(gHx[b], oHx, f"Hx grid {b}")
        assert_bits(gHy[b], oHy, f"Hy grid {b}")
        assert_bits(gtrace[:, b], otr[:, 0], f"probe grid {b}")


# --------------------------------------------------------------------------------------------
# per-handle options (fdtd2d_set_option): they choose kernels, never result bits
# --------------------------------------------------------------------------------------------
def test_options_are_per_handle_and_change_only_the_kernel(fd, monkeypatch):
    monkeypatch.setenv("FDTD2D_WAVE_MIN_TILES", "7")  # defaults are read when a handle is created ...
    R, C, n = 1500, 2100, 24
    with fd.Simulation(R, C, np.float32, dt=DT, dx=DX) as a, fd.Simulation(R, C, np.float32, dt=DT, dx=DX) as b:
        monkeypatch.setenv("FDTD2D_WAVE_MIN_TILES", "100000000")  # ... not afterwards
        assert a.get_option("wave_min_tiles") == 7 and b.get_option("wave_min_tiles") == 7
        b.set_option("wavefront", 0)
        assert a.get_option("wavefront") == 1 and b.get_option("wavefront") == 0
        outs = []
        for sim in (a, b):
            sim.set_kernel_variant(2)
            sim.set_materials_random(4, 9.0)
            sim.set_point_source(R // 2, C // 2, 700, FC)
            sim.step_index = 640
            sim.step(n, 8)
            outs.append(sim.state())
        assert a.plan_info(8)["wave_runs"] > 0 and a.plan_info(8)["tma_tiles"] == 0
        assert b.plan_info(8)["wave_runs"] == 0 and b.plan_info(8)["tma_tiles"] > 0
    for x, y in zip(*outs):
        assert_bits(x, y, "wavefront option on vs off")


@pytest.mark.parametrize("dtype", ["float32", "float64"])
@pytest.mark.parametrize("reserve,ring_cost", [(0, 0), (20, 0), (60, 208), (-1, 240)])
def test_edge_tiles_on_reserved_sms_vs_oracle(fd, oracle, dtype, reserve, ring_cost):
    """Option edge_reserve: the wavefront kernel goes out first on fewer SMs, its runs cut for those, and the edge tiles follow
    on the side stream; ring_cost changes the length of the ring-strip runs.  Scheduling only -- the bits are the oracle's."""
    c_oracle, npo = oracle
    R, C, nsteps = 700, 1500, 24
    rng = np.random.default_rng(R * 11 + C + reserve)
    eps, mu, Ez, Hx, Hy = _random_problem(rng, R, C, dtype)
    mu[...] = np.dtype(dtype).type(4 * np.pi * 1e-7)
    ce, ch, coef = c_oracle.coefficients(eps, mu, DT, DX, np.dtype(dtype))
    cells = [(R // 2, C // 2), (7, 9)]
    amp = npo.source_table("ricker", nsteps, DT, FC) + 0.125
    probes = [(R // 2, C // 2 + 3), (0, 0), (R - 1, C - 1), (R // 4, C // 3)]
    oEz, oHx, oHy = Ez.copy(), Hx.copy(), Hy.copy()
    otrace = c_oracle.run(oEz, oHx, oHy, ce, ch, coef, nsteps, amp, cells, probes, omp=True)
    with fd.Simulation(R, C, np.dtype(dtype), dt=DT, dx=DX) as sim:
        sim.set_kernel_variant(2)
        for key, v in (("wave_min_tiles", 0), ("ring_min_tiles", 0), ("edge_reserve", reserve), ("ring_cost", ring_cost)):
            sim.set_option(key, v)
        sim.set_coefficients(ce, ch, coef)
        sim.set_state(Ez, Hx, Hy)
        sim.set_sources([(0, r, c, 0) for r, c in cells], amp[None, :])
        sim.set_probes(probes, nsteps)
        sim.step(nsteps, 8)
        info = sim.plan_info(8)
        assert info["wave_runs"] > 0 and info["edge_tiles"] > 0 and info["ring_strips"] == 1
        if reserve >= 0:
            assert info["reserve_sms"] == reserve
        gEz, gHx, gHy = sim.state()
        gtrace = sim.read_probes()
    assert_bits(gtrace, otrace, "probe trace")
    assert_bits(gEz, oEz, "Ez")
    assert_bits(gHx, oHx, "Hx")
    assert_bits(gHy, oHy, "Hy")


# --------------------------------------------------------------------------------------------
# the fused double pass: two k = 8 passes per launch, the second one reading the first one's rows from L2
# --------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape,nsteps", [((1024, 1024), 40), ((700, 1500), 33), ((2000, 640), 16)])
@pytest.mark.parametrize("uniform_mu", [False, True])
def test_fused_double_pass_vs_oracle(fd, oracle, shape, nsteps, uniform_mu):
    """Forced onto small grids so the oracle can check it: sources and probes break the runs (pieces next to their tiles
    go to the second launch), the remainder (nsteps % 16) runs as single passes."""
    c_oracle, npo = oracle
    R, C = shape
    rng = np.random.default_rng(R * 37 + C)
    eps, mu, Ez, Hx, Hy = _random_problem(rng, R, C, "float32")
    if uniform_mu:
        mu[...] = np.float32(4 * np.pi * 1e-7)
    ce, ch, coef = c_oracle.coefficients(eps, mu, DT, DX, np.dtype(np.float32))
    cells = [(R // 2, C // 2), (R // 3, C // 4), (7, 9)]
    amp = npo.source_table("ricker", nsteps, DT, FC) + 0.125
    probes = [(R // 2, C // 2 + 3), (0, 0), (R - 1, C - 1), (R // 4, C // 3), (3 * R // 4, 2 * C // 3)]
    oEz, oHx, oHy = Ez.copy(), Hx.copy(), Hy.copy()
    otrace = c_oracle.run(oEz, oHx, oHy, ce, ch, coef, nsteps, amp, cells, probes, omp=True)
    with fd.Simulation(R, C, np.float32, dt=DT, dx=DX) as sim:
        sim.set_kernel_variant(2)
        for key, v in (("wave_min_tiles", 0), ("ring_min_tiles", 0), ("fuse", 1)):
            sim.set_option(key, v)
        sim.set_coefficients(ce, ch, coef)
        sim.set_state(Ez, Hx, Hy)
        sim.set_sources([(0, r, c, 0) for r, c in cells], amp[None, :])
        sim.set_probes(probes, nsteps)
        before = sim.launch_count
        sim.step(nsteps, 0)
        sim.synchronize()
        assert sim.pass_count == -(-nsteps // 8)
        # a fused pair is 4 launches (2 x edge tiles, the fused runs, the deferred runs); a single pass 2
        assert sim.launch_count - before <= 4 * (nsteps // 16) + 2 * -(-(nsteps % 16) // 8) + 1
        gEz, gHx, gHy = sim.state()
        gtrace = sim.read_probes()
    assert_bits(gtrace, otrace, "probe trace")
    assert_bits(gEz, oEz, "Ez")
    assert_bits(gHx, oHx, "Hx")
    assert_bits(gHy, oHy, "Hy")


def test_fused_equals_single_passes_large(fd):
    """6000 x 5000 fp32, 48 steps near the Ricker peak, fused pairs against single passes: not a bit may differ."""
    outs = []
    for fuse in (0, 1):
        with fd.Simulation(6000, 5000, np.float32, dt=DT, dx=DX) as sim:
            sim.set_option("fuse", fuse)
            sim.set_materials_random(seed=5, span=9.0)
            sim.set_point_source(3000, 2500, 700, FC)
            sim.set_probes([(3000, 2510), (10, 10), (5990, 4990)], 700)
            sim.step_index = 640
            assert (sim.plan_info(8)["wave_runs"] > 0)
            sim.step(48, 0)
            sim.synchronize()
            assert sim.pass_count == 6
            outs.append(sim.state() + (sim.read_probes(640, 48),))
    assert np.abs(outs[0][0]).max() > 0.1
    for a, b in zip(*outs):
        assert_bits(a, b, "fused pairs vs single passes")


# --------------------------------------------------------------------------------------------
# the staged wavefront: 12 levels as three warps of four (strip_stage.cuh)
# --------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape,nsteps", [((300, 517), 40), ((1024, 1024), 24), ((203, 600), 17), ((2000, 260), 36), ((700, 1500), 12)])
@pytest.mark.parametrize("groups", [4, 5])
def test_staged_wavefront_vs_oracle(fd, oracle, shape, nsteps, groups):
    """Forced onto small grids: runs broken by sources and probes, ring strips where C >= 512, ragged sizes, the remainder
    pass (nsteps % 12) on the tile kernels; uniform permeability (what the 12-level kernels exist for)."""
    c_oracle, npo = oracle
    R, C = shape
    rng = np.random.default_rng(R * 41 + C)
    eps, mu, Ez, Hx, Hy = _random_problem(rng, R, C, "float32")
    mu[...] = np.float32(4 * np.pi * 1e-7)
    ce, ch, coef = c_oracle.coefficients(eps, mu, DT, DX, np.dtype(np.float32))
    cells = [(R // 2, C // 2), (R // 3, C // 4), (7, 9)]
    amp = npo.source_table("ricker", nsteps, DT, FC) + 0.125
    probes = [(R // 2, C // 2 + 3), (0, 0), (R - 1, C - 1), (R // 4, C // 3), (3 * R // 4, 2 * C // 3)]
    oEz, oHx, oHy = Ez.copy(), Hx.copy(), Hy.copy()
    otrace = c_oracle.run(oEz, oHx, oHy, ce, ch, coef, nsteps, amp, cells, probes, omp=True)
    with fd.Simulation(R, C, np.float32, dt=DT, dx=DX) as sim:
        sim.set_kernel_variant(2)
        for key, v in (("wave_min_tiles", 0), ("ring_min_tiles", 0), ("stage", groups)):
            sim.set_option(key, v)
        sim.set_coefficients(ce, ch, coef)
        sim.set_state(Ez, Hx, Hy)
        sim.set_sources([(0, r, c, 0) for r, c in cells], amp[None, :])
        sim.set_probes(probes, nsteps)
        info = sim.plan_info(12)
        assert info["wave_runs"] > 0 and (info["ring_strips"] == 1) == (C >= 512), info
        sim.step(nsteps, 12)
        gEz, gHx, gHy = sim.state()
        gtrace = sim.read_probes()
    assert_bits(gtrace, otrace, "probe trace")
    assert_bits(gEz, oEz, "Ez")
    assert_bits(gHx, oHx, "Hx")
    assert_bits(gHy, oHy, "Hy")


def test_staged_wavefront_equals_one_warp_kernels_large(fd):
    """6000 x 5000 fp32, 48 steps near the Ricker peak: 12 levels staged over three warps (4 and 5 groups per CTA) against the
    one-warp 8-level kernel."""
    outs = []
    for stage, k in ((0, 8), (4, 12), (5, 12)):
        with fd.Simulation(6000, 5000, np.float32, dt=DT, dx=DX) as sim:
            sim.set_option("stage", stage)
            sim.set_materials_random(seed=5, span=9.0)
            sim.set_point_source(3000, 2500, 700, FC)
            sim.set_probes([(3000, 2510), (10, 10), (5990, 4990)], 700)
            sim.step_index = 640
            sim.step(48, k)
            outs.append(sim.state() + (sim.read_probes(640, 48),))
    assert np.abs(outs[0][0]).max() > 0.1
    for o in outs[1:]:
        for a, b in zip(outs[0], o):
            assert_bits(a, b, "staged 12 levels vs 8 levels")


# --------------------------------------------------------------------------------------------
# non-blocking copies: two jobs in flight from one host thread (what bench.py's e2e does)
# --------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shared_stream", [False, True])
def test_async_copies_pipeline_two_handles_from_one_thread(fd, shared_stream):
    """Upload (materials + state), step, download with the *_async entry points on pinned arrays, two handles alternating so
    that one job's copies overlap the other's kernels; every job must give what the blocking calls give.  With a shared
    compute stream (how sharded jobs run) the handles must still only wait for their OWN work."""
    import torch

    R, C, n = 1500, 2100, 160
    rng = np.random.default_rng(13)
    eps, mu, Ez, Hx, Hy = _random_problem(rng, R, C, "float32")
    mu[...] = np.float32(4 * np.pi * 1e-7)

    def pin(a):
        t = torch.from_numpy(a.copy()).pin_memory()
        return t.numpy()

    peps, pmu, pEz, pHx, pHy = (pin(a) for a in (eps, mu, Ez, Hx, Hy))
    with fd.Simulation(R, C, np.float32, dt=DT, dx=DX) as ref:
        ref.set_point_source(R // 2, C // 2, n, FC)
        ref.set_probes([(R // 2, C // 2 + 16), (R // 4, C // 4)], n)
        ref.set_materials(eps, mu)
        ref.set_state(Ez, Hx, Hy)
        ref.step(n, 0)
        want_Ez, want_tr = ref.read_Ez(), ref.read_probes(0, n)
    sims = [fd.Simulation(R, C, np.float32, dt=DT, dx=DX) for _ in range(2)]
    stream = torch.cuda.Stream()
    outs = [pin(np.zeros((R, C), np.float32)) for _ in sims]
    try:
        for sim in sims:
            if shared_stream:
                sim.set_stream(stream.cuda_stream)
            sim.set_point_source(R // 2, C // 2, n, FC)
            sim.set_probes([(R // 2, C // 2 + 16), (R // 4, C // 4)], n)

        def issue(w):
            sims[w].step_index = 0
            sims[w].set_materials_async(peps, pmu)
            sims[w].set_state_async(pEz, pHx, pHy)
            sims[w].step(n, 0)
            sims[w].read_Ez_async(outs[w])

        queued, done = [], 0
        for j in range(6):
            if len(queued) == 2:
                w = queued.pop(0)
                sims[w].synchronize()
                assert_bits(outs[w], want_Ez, f"job {done} (handle {w})")
                assert_bits(sims[w].read_probes(0, n), want_tr, f"probes of job {done}")
                outs[w][...] = 0
                done += 1
            issue(j % 2)
            queued.append(j % 2)
        for w in queued:
            sims[w].synchronize()
            assert_bits(outs[w], want_Ez, f"job {done} (handle {w})")
            done += 1
        assert done == 6
        with pytest.raises(ValueError):
            sims[0].set_state_async(Ez.astype(np.float64), pHx, pHy)  # the caller's buffer is taken as it is: wrong dtype is refused
    finally:
        for sim in sims:
            sim.close()
