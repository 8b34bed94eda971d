"""Device-side field readout (capture_snapshot, main.py:153-179) against the oracle's numpy restatement."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
DT, DX, FC = 5e-14, 1e-4, 30e9


@pytest.mark.parametrize("dtype", ["float32", "float64"])
@pytest.mark.parametrize("uniform", [False, True])
def test_render_matches_oracle(golden_dir, dtype, uniform, tmp_path):
    import fdtd2d_b200 as fd
    from oracle import numpy_oracle as npo

    R, C = 200, 200
    eps, mu = fd.material_init(None if uniform else os.path.join(golden_dir, "structure.png"), R, C)
    with fd.Simulation(R, C, dtype, dt=DT, dx=DX) as sim:
        sim.set_materials(eps, mu)
        sim.set_point_source(100, 100, 700, FC)
        sim.set_snapshot_background(eps)
        sim.step(650)
        for vmax, vmin in ((1e-3, -1e-3), (20, -20), (0.25, -0.1)):  # fdtd.py:38 passes 1e-3 / -1e-3
            frame = sim.render_snapshot(vmax, vmin)
            Ez = sim.read_Ez()
            want = npo.snapshot_rgb(Ez, eps, vmax, vmin)
            assert frame.dtype == np.uint8 and frame.shape == (R, C, 3)
            assert np.array_equal(frame, want), f"{(frame != want).sum()} bytes differ"
        assert len(np.unique(frame)) > 10  # not a blank image
        # the drop-in function writes the same pixels
        path = str(tmp_path / "frame.png")
        fd.capture_snapshot(Ez, eps, path, 1e-3, -1e-3)
        from PIL import Image

        assert np.array_equal(np.array(Image.open(path)), npo.snapshot_rgb(Ez, eps, 1e-3, -1e-3))
    fd.release_handles()
