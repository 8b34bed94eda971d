"""The structure rasteriser's CPU restatement (oracle/structure_oracle.py) against masks drawn by the REAL reference class
RegionDrawer (python-src/region_drawer.py, PIL) -- tests/golden/structures.npz, made by oracle/make_golden_structures.py:
straight waveguides, couplers and discs cell for cell; rings and slanted / curved waveguides within the stated bound (the
device rules are geometric where PIL rounds polygon corners to integers)."""
import os

import numpy as np
import pytest

import fdtd2d_b200 as fd
from oracle import structure_oracle as so


def _ref(golden_dir, name):
    g = np.load(os.path.join(golden_dir, "structures.npz"))
    cols, rows = so.SCENES[name][:2]
    return np.unpackbits(g[name])[:rows * cols].reshape(rows, cols).astype(bool)


@pytest.mark.parametrize("name", sorted(so.SCENES))
def test_oracle_vs_reference_region_drawer(golden_dir, name):
    cols, rows, _, exact = so.SCENES[name]
    ref = _ref(golden_dir, name)
    mine = so.draw_scene(so.RegionDrawer(cols, rows), name).image == 0
    diff = int((ref != mine).sum())
    if exact:
        assert diff == 0, f"{name}: {diff} cells differ from the reference's RegionDrawer"
    else:
        # rings: the inner edge of PIL's outline; slanted segments: PIL rounds the rectangle's corners to integers
        assert diff <= 0.05 * ref.sum(), f"{name}: {diff} of {int(ref.sum())} cells differ"
        assert diff > 0  # (if this ever becomes exact, mark the scene exact)


def test_filled_ellipses_equal_pil_for_every_small_box():
    """Pillow's quarter walk restated: every box up to 40 x 40 and a few large ones, against PIL itself."""
    from PIL import Image, ImageDraw

    def pil(box, shape):
        im = Image.new("L", shape[::-1], 255)
        ImageDraw.Draw(im).ellipse(box, fill=0)
        return np.array(im) == 0

    for w in range(0, 41):
        for h in range(0, 41, 3):
            box = (3, 2, 3 + w, 2 + h)
            assert np.array_equal(so.ellipse_mask((h + 6, w + 8), box), pil(box, (h + 6, w + 8))), box
    rng = np.random.default_rng(5)
    for _ in range(12):
        w, h = int(rng.integers(50, 700)), int(rng.integers(50, 700))
        box = (-7, 11, -7 + w, 11 + h)  # clipped on the left
        assert np.array_equal(so.ellipse_mask((h + 20, w + 5), box), pil(box, (h + 20, w + 5))), box


def test_wide_line_boxes_equal_pil():
    """The product's host arithmetic for horizontal / vertical waveguides (fdtd2d_b200.structure.wide_line_box)."""
    from PIL import Image, ImageDraw

    for (x0, y0, x1, y1) in [(10, 30, 80, 30), (80, 30, 10, 30), (40, 5, 40, 70), (40, 70, 40, 5)]:
        for width in range(1, 24):
            im = Image.new("L", (100, 100), 255)
            ImageDraw.Draw(im).line([(x0, y0), (x1, y1)], fill=0, width=width)
            ys, xs = np.nonzero(np.array(im) == 0)
            assert fd.structure.wide_line_box(x0, y0, x1, y1, width) == (xs.min(), ys.min(), xs.max(), ys.max()), (x0, y0, x1, y1, width)
            assert len(xs) == (xs.max() - xs.min() + 1) * (ys.max() - ys.min() + 1)  # a full rectangle
