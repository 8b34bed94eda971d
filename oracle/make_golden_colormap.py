"""Pin for the "seismic" colour table behind capture_snapshot (python-src/main.py:171) -- authoring-container script.

Run from the repo root:  python -m oracle.make_golden_colormap
matplotlib is not installed here, so the table cannot be read from it.  The reference tree, however, ships images its
own colour pipeline produced: python-src/Ez.png, assets/ring_resonator.png and assets/Ez_tiled.png are 1000 x 1000 frames
written by utils.plot_Ez (python-src/utils.py:15-41), which is line for line the pipeline of capture_snapshot
(main.py:153-179): clip, matplotlib's seismic colormap, alpha 0.7 over the grayscale permittivity background, uint8.
This script stores the DISTINCT pixel colours of those images (with their counts) in tests/golden/seismic_pixels.npz;
tests/test_host_cpu.py then demands that every one of them is produced, bit for bit, by the rebuilt table and the blend
formula for some table index and some background gray level -- 240 of the 256 table entries are exercised on the white
background alone.  Nothing here is produced by the oracle or by the CUDA path.
"""
from __future__ import annotations

import os

import numpy as np
from PIL import Image

REF = "/root/reference"
IMAGES = ("python-src/Ez.png", "assets/ring_resonator.png", "assets/Ez_tiled.png")
GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def main():
    out = {"images": np.array(IMAGES)}
    for i, rel in enumerate(IMAGES):
        a = np.array(Image.open(os.path.join(REF, rel)))
        assert a.ndim == 3 and a.shape[2] == 4 and (a[..., 3] == 255).all(), "plt.imsave writes opaque RGBA"
        colours, counts = np.unique(a[..., :3].reshape(-1, 3), axis=0, return_counts=True)
        out[f"colours_{i}"], out[f"counts_{i}"] = colours, counts
        print(rel, a.shape, len(colours), "distinct colours")
    np.savez_compressed(os.path.join(GOLDEN, "seismic_pixels.npz"), **out)


if __name__ == "__main__":
    main()
