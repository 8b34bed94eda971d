"""GPU parity of the cluster-resident kernel (csrc/grid_resident.cuh): small fp32 grids that stay on chip
for a whole fdtd2d_step call, one thread-block cluster per grid, halo rows through distributed shared
memory.  Bit-exact against the CPU oracle and against the tiled path, through the C ABI."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

DT, DX, FC = 5e-14, 1e-4, 30e9


def assert_bits(a, b, what):
    assert a.dtype == b.dtype and a.shape == b.shape, what
    if not np.array_equal(a, b):
        bad = np.argwhere(a != b)
        d = np.linalg.norm(a.astype(np.float64) - b.astype(np.float64)) / max(np.linalg.norm(b.astype(np.float64)), 1e-300)
        raise AssertionError(f"{what}: {len(bad)} cells differ, first at {bad[0]}, rel-L2 = {d:.3e} (tolerance 1e-5)")


@pytest.fixture(scope="module")
def fd():
    import fdtd2d_b200

    return fdtd2d_b200


@pytest.fixture(scope="module")
def oracle():
    from oracle import c_oracle, numpy_oracle

    c_oracle.build()
    return c_oracle, numpy_oracle


def _problem(rng, R, C, scale=1e-3):
    eps = (8.85418e-12 * (1 + 9 * rng.random((R, C)))).astype(np.float32)
    mu = (4 * np.pi * 1e-7 * (1 + 0.5 * rng.random((R, C)))).astype(np.float32)
    Ez = (scale * rng.standard_normal((R, C))).astype(np.float32)
    Hx = (scale * 1e-3 * rng.standard_normal((R, C - 1))).astype(np.float32)
    Hy = (scale * 1e-3 * rng.standard_normal((R - 1, C))).astype(np.float32)
    return eps, mu, Ez, Hx, Hy


# rows choose the cluster size (48-row bands: 1 CTA up to 48 rows ... 8 CTAs up to 384) and the ragged last band;
# columns choose where the right ring frame falls (second half, first half, straddling column 128)
SHAPES = [(16, 16), (17, 33), (48, 256), (49, 255), (64, 256), (65, 255), (100, 128), (128, 129), (129, 131), (130, 136),
          (200, 200), (256, 256), (255, 140), (260, 64), (300, 250), (384, 256), (383, 17), (37, 53), (96, 130), (72, 100)]
# (kernel shape, grid shape): shape 0 = 3 rows per thread x 16 warps (48-row bands, the default), 1 = 4 x 12 (48 rows),
# 2 = 2 x 16 (32 rows), 3 = 4 x 8 (32 rows), 4 = 3 x 12 (36 rows); a cluster has at most 8 bands;
# 5 = the packed kernel of grid_resident_x2.cuh (6 x 8, 48 rows: first band of any size, the others multiples of six)
SMALL = [(256, 256), (33, 40), (250, 141), (100, 128), (17, 33), (200, 200)]
X2_MORE = [(18, 40), (24, 64), (30, 31), (47, 47), (54, 250), (97, 256), (101, 19), (144, 144), (250, 141), (33, 40),
           (60, 128), (61, 127), (90, 134), (77, 135), (120, 160), (50, 224), (52, 225)]
CASES = ([(0, s) for s in SHAPES] + [(1, s) for s in SHAPES[::2]] + [(c, s) for c in (2, 3, 4) for s in SMALL]
         + [(5, s) for s in SHAPES + X2_MORE])


# (the packed kernel takes dt/(mu*dx) as a kernel argument when it is uniform: both forms)
CASES = [(c, s, False) for c, s in CASES] + [(5, s, True) for s in SHAPES[::2] + X2_MORE[::2]]


@pytest.mark.parametrize("rcfg,shape,uniform_mu", CASES)
@pytest.mark.parametrize("nsteps", [1, 2, 37])
def test_resident_vs_oracle(fd, oracle, rcfg, shape, uniform_mu, nsteps, monkeypatch):
    c_oracle, npo = oracle
    R, C = shape
    monkeypatch.setenv("FDTD2D_RESIDENT_CFG", str(rcfg))
    rng = np.random.default_rng(R * 1009 + C * 13 + nsteps)
    eps, mu, Ez, Hx, Hy = _problem(rng, R, C)
    if uniform_mu:
        mu[:] = mu[0, 0]
    ce, ch, coef = c_oracle.coefficients(eps, mu, DT, DX, np.dtype(np.float32))
    # sources in the interior, inside every ring frame and on the corner cells
    cells = [(R // 2, C // 2), (R // 2, C // 2 + 1), (7, 9), (R - 3, C - 2), (0, 0), (R - 1, 0), (3, C - 1), (R // 3, 2),
             (R - 7, C // 3)]
    cells = sorted(set(cells))
    amp = npo.source_table("ricker", nsteps, DT, FC) + 0.25
    probes = [(R // 2, C // 2), (0, 0), (R - 1, C - 1), (4, C - 5), (R - 5, 4), (R // 3, C // 5), (5, 5), (R - 6, C - 6),
              (6, 6), (R - 7, C - 7), (R // 2, 0), (0, C // 2), (R // 2 + 1, C // 2)]
    oEz, oHx, oHy = Ez.copy(), Hx.copy(), Hy.copy()
    otrace = c_oracle.run(oEz, oHx, oHy, ce, ch, coef, nsteps, amp, cells, probes)
    with fd.Simulation(R, C, np.float32, dt=DT, dx=DX) as sim:
        sim.set_kernel_variant(4)  # the cluster-resident kernel or an error, never a silent detour
        sim.set_materials(eps, mu)
        sim.set_state(Ez, Hx, Hy)
        sim.set_sources([(0, r, c, 0) for r, c in cells], amp[None, :])
        sim.set_probes(probes, nsteps)
        before = sim.launch_count
        sim.step(nsteps)
        # the whole run is one launch (the packed kernel asks once whether dt/(mu*dx) is uniform: one small kernel more)
        # (grids whose right ring straddles column 128 stay with the first kernel)
        packed = rcfg == 5 and not (C > 128 and ((C - 6) // 4) * 4 < 128)
        assert sim.launch_count == before + (2 if packed else 1)
        gEz, gHx, gHy = sim.state()
        gtrace = sim.read_probes()
    what = f"{R}x{C} n={nsteps}"
    assert_bits(gtrace, otrace, "probe trace " + what)
    assert_bits(gEz, oEz, "Ez " + what)
    assert_bits(gHx, oHx, "Hx " + what)
    assert_bits(gHy, oHy, "Hy " + what)


@pytest.mark.parametrize("rcfg", [0, 5])
def test_resident_batched_dataset_like(fd, oracle, rcfg, monkeypatch):
    """Many independent 256 x 256 grids in one launch (BASELINE configs[4]): per-grid binary media, point and
    line sources with per-grid waveforms, probes; a sample of the grids is checked against the oracle and
    every grid against the tiled path."""
    c_oracle, npo = oracle
    monkeypatch.setenv("FDTD2D_RESIDENT_CFG", str(rcfg))
    B, R, C, nsteps = 40, 256, 256, 120
    rng = np.random.default_rng(8)
    eps = np.where(rng.random((B, R, C)) > 0.5, 5.0, 1.0).astype(np.float32) * np.float32(8.85418e-12)
    mu = np.full((B, R, C), 4 * np.pi * 1e-7, np.float32)
    tables = np.stack([npo.source_table("ricker", nsteps, DT, 18e9 + 3e8 * b) for b in range(B)])
    cells, per_grid = [], []
    for b in range(B):
        r0, c0 = int(rng.integers(26, 230)), int(rng.integers(26, 230))
        if b % 3 == 0:
            mine = [(r0, c0)]
        elif b % 3 == 1:
            mine = [(r0, min(c0 + j, 229)) for j in range(25)]
        else:
            mine = [(min(r0 + i, 229), c0) for i in range(25)]
        mine = sorted(set(mine))
        per_grid.append(mine)
        cells += [(b, r, c, b) for r, c in mine]
    probes = [(b, 128, 131) for b in range(B)] + [(b, 2, 250) for b in range(B)]
    outs = []
    for variant in (4, 2):
        with fd.Simulation(R, C, np.float32, dt=DT, dx=DX, batch=B) as sim:
            sim.set_kernel_variant(variant)
            sim.set_materials(eps, mu)
            sim.set_sources(cells, tables)
            sim.set_probes(probes, nsteps)
            sim.step(nsteps, 8)
            outs.append(sim.state() + (sim.read_probes(),))
    for a, b_, name in zip(outs[0], outs[1], ("Ez", "Hx", "Hy", "trace")):
        assert_bits(a, b_, f"resident vs tiled: {name}")
    gEz, gHx, gHy, gtrace = outs[0]
    for b in (0, 1, 2, 17, 39):
        ce, ch, coef = c_oracle.coefficients(eps[b], mu[b], DT, DX, np.dtype(np.float32))
        oEz, oHx, oHy = npo.grid_init(R, C, np.dtype(np.float32))
        otr = c_oracle.run(oEz, oHx, oHy, ce, ch, coef, nsteps, tables[b], per_grid[b], [(128, 131), (2, 250)])
        assert_bits(gEz[b], oEz, f"Ez grid {b}")
        assert_bits(gHx[b], oHx, f"Hx grid {b}")
        assert_bits(gHy[b], oHy, f"Hy grid {b}")
        assert_bits(gtrace[:, b], otr[:, 0], f"probe A grid {b}")
        assert_bits(gtrace[:, B + b], otr[:, 1], f"probe B grid {b}")


@pytest.mark.parametrize("rcfg", [0, 5])
def test_resident_demo_golden_and_pieces(fd, golden_dir, rcfg, monkeypatch):
    """The reference demo (fdtd.py defaults, fp32) in one launch and in pieces, against the reference's output."""
    monkeypatch.setenv("FDTD2D_RESIDENT_CFG", str(rcfg))
    g = np.load(os.path.join(golden_dir, "demo200_vacuum_float32.npz"))
    eps, mu = fd.material_init(None, 200, 200)
    for plan in ([1000], [1, 2, 333, 64, 600]):
        with fd.Simulation(200, 200, np.float32, dt=DT, dx=DX) as sim:
            sim.set_kernel_variant(4)
            sim.set_materials(eps, mu)
            sim.set_point_source(100, 100, 1000, FC)
            sim.set_probes([tuple(p) for p in g["probes"]], 1000)
            for n in plan:
                sim.step(n)
            Ez, Hx, Hy = sim.state()
            trace = sim.read_probes()
        assert_bits(trace, g["trace"], "probe trace (every step)")
        assert_bits(Ez, g["Ez"], "Ez")
        assert_bits(Hx, g["Hx"], "Hx")
        assert_bits(Hy, g["Hy"], "Hy")


@pytest.mark.parametrize("rcfg", [0, 5])
@pytest.mark.parametrize("cluster", [3, 4, 5, 7, 8])
def test_resident_cluster_size_knob(fd, oracle, cluster, rcfg, monkeypatch):
    """More, thinner bands per grid (FDTD2D_RESIDENT_CLUSTER) must not change a bit."""
    c_oracle, npo = oracle
    R, C, nsteps = 120, 200, 33
    rng = np.random.default_rng(cluster)
    eps, mu, Ez, Hx, Hy = _problem(rng, R, C)
    ce, ch, coef = c_oracle.coefficients(eps, mu, DT, DX, np.dtype(np.float32))
    amp = npo.source_table("sinusoidal", nsteps, DT, FC) + 0.5
    oEz, oHx, oHy = Ez.copy(), Hx.copy(), Hy.copy()
    otrace = c_oracle.run(oEz, oHx, oHy, ce, ch, coef, nsteps, amp, [(60, 100)], [(61, 100), (0, 0)])
    monkeypatch.setenv("FDTD2D_RESIDENT_CLUSTER", str(cluster))
    monkeypatch.setenv("FDTD2D_RESIDENT_CFG", str(rcfg))
    with fd.Simulation(R, C, np.float32, dt=DT, dx=DX) as sim:
        sim.set_kernel_variant(4)
        sim.set_coefficients(ce, ch, coef)
        sim.set_state(Ez, Hx, Hy)
        sim.set_sources([(0, 60, 100, 0)], amp[None, :])
        sim.set_probes([(61, 100), (0, 0)], nsteps)
        sim.step(nsteps)
        gEz, gHx, gHy = sim.state()
        gtrace = sim.read_probes()
    assert_bits(gtrace, otrace, "probe trace")
    assert_bits(gEz, oEz, "Ez")
    assert_bits(gHx, oHx, "Hx")
    assert_bits(gHy, oHy, "Hy")


def test_resident_not_eligible_is_an_error_when_forced(fd):
    for (R, C, dtype) in [(64, 300, np.float32), (600, 64, np.float32), (64, 64, np.float64), (12, 64, np.float32), (386, 64, np.float32)]:
        with fd.Simulation(R, C, dtype, dt=DT, dx=DX) as sim:
            sim.set_kernel_variant(4)
            sim.set_materials(*fd.material_init(None, R, C))
            with pytest.raises(fd.Fdtd2dError):
                sim.step(3)
            sim.set_kernel_variant(0)  # automatic: falls back to the tiled kernels
            sim.step(3)
