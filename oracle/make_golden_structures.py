"""tests/golden/structures.npz: masks drawn by the REAL reference class RegionDrawer (python-src/region_drawer.py, PIL
ImageDraw) for the scenes of oracle/structure_oracle.py -- authoring-container script.

Run from the repo root:  python -m oracle.make_golden_structures
Every array saved is the output of the reference's own class (bit-packed: black = 1); nothing is produced by the oracle or by
the CUDA path.  Pillow version is recorded.
"""
from __future__ import annotations

import importlib.util
import os

import numpy as np
import PIL

from .structure_oracle import SCENES, draw_scene

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def main():
    spec = importlib.util.spec_from_file_location("ref_region_drawer", "/root/reference/python-src/region_drawer.py")
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    out = {"pillow_version": PIL.__version__}
    for name, (cols, rows, _, _) in SCENES.items():
        d = draw_scene(ref.RegionDrawer(cols, rows), name)
        img = np.array(d.image)
        assert img.shape == (rows, cols) and set(np.unique(img)) <= {0, 255}
        out[name] = np.packbits(img == 0)
        print(name, img.shape, int((img == 0).sum()), "black cells")
    np.savez_compressed(os.path.join(GOLDEN, "structures.npz"), **out)


if __name__ == "__main__":
    main()
