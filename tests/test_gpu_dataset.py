"""GPU tests of the rows around the hot path: device-side structure-image mapping, the batched dataset
generator, and the fdtd.py driver with device-rendered frames -- all through the C ABI, bit-exact vs the oracle."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def assert_bits(a, b, what):
    assert a.dtype == b.dtype and a.shape == b.shape, what
    if not np.array_equal(a, b):
        bad = np.argwhere(a != b)
        raise AssertionError(f"{what}: {len(bad)} cells differ, first at {bad[0]}")


@pytest.fixture(scope="module")
def fd():
    import fdtd2d_b200

    return fdtd2d_b200


@pytest.mark.parametrize("dtype", ["float32", "float64"])
@pytest.mark.parametrize("shape", [(200, 200), (97, 301)])
def test_structure_image_mapped_on_device(fd, golden_dir, dtype, shape):
    """set_materials_image == material_init (host, main.py:88-123) + cast + set_materials, bit for bit."""
    R, C = shape
    png = os.path.join(golden_dir, "structure.png")
    eps, mu = fd.material_init(png, R, C, black_point=7.5)
    with fd.Simulation(R, C, dtype, dt=5e-14, dx=1e-4) as a, fd.Simulation(R, C, dtype, dt=5e-14, dx=1e-4) as b:
        a.set_materials(eps, mu)
        b.set_materials_image(png, black_point=7.5)
        for x, y, name in zip(a.coefficients(), b.coefficients(), ("ce", "ch", "mur")):
            assert_bits(x, y, name)


@pytest.mark.parametrize("dtype", ["float32", "float64"])
def test_blob_media_vs_oracle(fd, dtype):
    """The device-generated two-phase media equal the oracle's restatement of generate_random_permittivity on the
    same hash-uniform field, cell for cell (float32 blur, row-major accumulation, no FMA)."""
    from oracle import dataset_oracle as do

    for (B, R, C, seed) in [(3, 64, 64, 2026), (2, 60, 100, 5), (2, 256, 256, 9), (1, 33, 47, 1)]:
        eps, mu, src, omega, Ez = fd.generate_data(B, (R, C), n_steps=1, seed=seed, dtype=dtype)
        assert eps.dtype == np.dtype(dtype) and eps.shape == (B, R, C)
        for b in range(B):
            assert_bits(eps[b], do.permittivity(seed, b, R, C, dtype), f"eps of sample {b} ({R}x{C}, seed {seed})")
        assert np.all(mu == np.dtype(dtype).type(do.MU_0))


def test_generate_data_vs_oracle_fdtd(fd):
    """A dataset of 24 samples of 256 x 256 (cluster-resident kernel): shapes like the reference's generate_data,
    sources as planned, and Ez of a few samples equal to the CPU oracle's leapfrog loop on the same inputs."""
    from oracle import c_oracle, dataset_oracle as do, numpy_oracle as npo

    c_oracle.build()
    N, R, C, n, seed = 24, 256, 256, 150, 77
    eps, mu, src, omega, Ez = fd.generate_data(N, (R, C), n_steps=n, seed=seed)
    assert all(a.shape == (N, R, C) and a.dtype == np.float32 for a in (eps, mu, src, Ez)) and omega.shape == (N,)
    plan = fd.dataset.sample_plan(N, (R, C), seed)
    dx = 1e-3
    dt = 0.5 * dx * float(np.sqrt(do.EPS_0 * do.MU_0))
    kinds = set()
    for b in (0, 1, 5, 11, 23):
        _, cells, w = plan[b]
        kinds.add(len(cells))
        want = np.zeros((R, C), np.float32)
        for r, c in cells:
            want[r, c] = 1.0
        assert_bits(src[b], want, f"src {b}")
        assert omega[b] == np.float32(w)
        ce, ch, coef = c_oracle.coefficients(eps[b], mu[b], dt, dx, np.dtype(np.float32))
        oEz, oHx, oHy = npo.grid_init(R, C, np.float32)
        c_oracle.run(oEz, oHx, oHy, ce, ch, coef, n, npo.source_table("ricker", n, dt, w), sorted(cells), None)
        assert_bits(Ez[b], oEz, f"Ez of sample {b}")
        assert np.abs(Ez[b]).max() > 0
    assert np.isfinite(Ez).all()


def test_driver_frames_and_final_state(fd, golden_dir):
    """fdtd.py's loop with snapshots every 5 steps: the device-rendered frames equal the oracle's restatement of
    capture_snapshot on the oracle's fields, and the final state equals the reference's own output."""
    from oracle import numpy_oracle as npo

    g = np.load(os.path.join(golden_dir, "demo200_vacuum_float64.npz"))
    frames = {}
    Ez, Hx, Hy = fd.driver.run(None, on_frame=lambda i, rgb: frames.__setitem__(i, rgb.copy()))
    assert sorted(frames) == list(range(200))
    assert_bits(Ez, g["Ez"], "Ez")
    assert_bits(Hx, g["Hx"], "Hx")
    assert_bits(Hy, g["Hy"], "Hy")
    eps, mu = npo.material_init(None, 200, 200)
    oEz, oHx, oHy = npo.grid_init(200, 200)
    done = 0
    for f in (0, 1, 60, 133, 199):
        n = 5 * f + 1
        npo.run(oEz, oHx, oHy, mu, eps, 5e-14, 1e-4, n - done, source=(100, 100, 30e9, "ricker"), step0=done, dense_source=False)
        done = n
        assert_bits(frames[f], npo.snapshot_rgb(oEz, eps, 1e-3, -1e-3), f"frame {f}")
