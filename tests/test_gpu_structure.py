"""Structure drawing on the device (fdtd2d_canvas_*, fdtd2d_b200.RegionDrawer) against its CPU restatement, bit for bit,
and -- for straight waveguides, couplers and discs -- against masks drawn by the reference's own RegionDrawer class; then
the canvas -> permittivity -> coefficient path against material_init on the saved picture."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
DT, DX = 5e-14, 1e-4


@pytest.mark.parametrize("dtype", ["float32", "float64"])
def test_device_canvas_equals_oracle_and_reference(golden_dir, dtype):
    import fdtd2d_b200 as fd
    from oracle import structure_oracle as so

    g = np.load(os.path.join(golden_dir, "structures.npz"))
    for name, (cols, rows, _, exact) in so.SCENES.items():
        want = so.draw_scene(so.RegionDrawer(cols, rows), name).image
        with fd.Simulation(rows, cols, dtype, dt=DT, dx=DX) as sim:
            got = so.draw_scene(fd.RegionDrawer(sim), name).image()
        assert got.dtype == np.uint8 and np.array_equal(got, want), f"{name}: {(got != want).sum()} cells differ from the CPU restatement"
        if exact:
            ref = np.unpackbits(g[name])[:rows * cols].reshape(rows, cols).astype(bool)
            assert np.array_equal(got == 0, ref), f"{name}: differs from the reference's RegionDrawer"


def test_canvas_becomes_the_medium_like_material_init(tmp_path):
    """draw -> apply() on the device == RegionDrawer.save + material_init(png) + set_materials on the host path."""
    import fdtd2d_b200 as fd
    from oracle import structure_oracle as so

    cols, rows = so.SCENES["device"][:2]
    png = str(tmp_path / "structure.png")
    for dtype in (np.float32, np.float64):
        with fd.Simulation(rows, cols, dtype, dt=DT, dx=DX) as sim:
            d = so.draw_scene(fd.RegionDrawer(sim), "device")
            d.save(png)
            d.apply(black_point=10.0)
            ce, ch, mur = sim.coefficients()
        eps, mu = fd.material_init(png, rows, cols, 10.0)  # same size: the LANCZOS resize is the identity
        with fd.Simulation(rows, cols, dtype, dt=DT, dx=DX) as ref:
            ref.set_materials(eps, mu)
            rce, rch, rmur = ref.coefficients()
        assert np.array_equal(ce, rce) and np.array_equal(ch, rch) and np.array_equal(mur, rmur)


def test_slabs_draw_their_rows_of_the_global_picture():
    import fdtd2d_b200 as fd
    from oracle import structure_oracle as so

    cols, rows = so.SCENES["device"][:2]
    want = so.draw_scene(so.RegionDrawer(cols, rows), "device").image
    for r in range(3):
        b, e = fd.slab_rows(rows, 3, r)
        with fd.Simulation(rows, cols, np.float32, dt=DT, dx=DX, slab=(rows, b, e, 8)) as sim:
            got = so.draw_scene(fd.RegionDrawer(sim), "device").image()
            assert np.array_equal(got, want[sim.row0:sim.row0 + sim.local_rows])
