"""Field readout as images: the GPU counterpart of `capture_snapshot` / `make_video_from_frames`
(python-src/main.py:153-179 and :126-150), the step that follows the hot loop in the reference driver
(fdtd.py:36-40).

The reference maps Ez through matplotlib's "seismic" colormap; matplotlib is not a dependency here, so
the 256-entry lookup table is rebuilt from matplotlib's published construction
(`LinearSegmentedColormap.from_list` over the five seismic anchor colours -> `_create_lookup_table`,
N = 256, gamma = 1) and checked against frames the reference's own pipeline wrote (every distinct pixel colour of three
1000 x 1000 images in the reference tree is reproduced bit for bit, tests/test_host_cpu.py).  The colour lookup, the alpha blend over the permittivity background and the uint8
conversion run on the device (fdtd2d_render_snapshot), in the reference's operation order and precision.
"""
from __future__ import annotations

import ctypes
import subprocess

import numpy as np

from ._lib import check, lib

# matplotlib/_cm.py `_seismic_data`: evenly spaced anchors at 0, 1/4, 1/2, 3/4, 1
SEISMIC_ANCHORS = ((0.0, 0.0, 0.3), (0.0, 0.0, 1.0), (1.0, 1.0, 1.0), (1.0, 0.0, 0.0), (0.5, 0.0, 0.0))
EPS_MIN = 8.85418e-12  # main.py:158


def seismic_lut(n: int = 256) -> np.ndarray:
    """(n, 3) float64 table: matplotlib's `_create_lookup_table(n, data, gamma=1)` for each channel."""
    anchors = np.asarray(SEISMIC_ANCHORS, dtype=float)
    pos = np.linspace(0, 1, len(anchors))
    lut = np.empty((n, 3))
    x = pos * (n - 1)
    xind = (n - 1) * np.linspace(0, 1, n)
    ind = np.searchsorted(x, xind)[1:-1]
    for ch in range(3):
        y = anchors[:, ch]
        distance = (xind[1:-1] - x[ind - 1]) / (x[ind] - x[ind - 1])
        lut[:, ch] = np.concatenate([[y[0]], distance * (y[ind] - y[ind - 1]) + y[ind - 1], [y[-1]]])
    return np.clip(lut, 0.0, 1.0)


def eps_background(eps) -> np.ndarray:
    """uint8 grayscale permittivity background (main.py:157-165): white if uniform, else 128..255."""
    eps = np.asarray(eps)
    eps_max = np.max(eps)
    if eps_max == EPS_MIN:
        return np.full_like(eps, 255, dtype=np.uint8)
    eps_normed = (eps - EPS_MIN) / (eps_max - EPS_MIN)
    return ((1 - eps_normed) * 127 + 128).astype(np.uint8)


def set_background(sim, eps) -> None:
    """Upload the background of `eps` (host map as returned by material_init) and the colormap."""
    gray = np.ascontiguousarray(eps_background(eps))
    lut = np.ascontiguousarray(seismic_lut())
    check(lib().fdtd2d_set_snapshot_background(sim._h, gray.ctypes.data_as(ctypes.c_void_p),
                                               lut.ctypes.data_as(ctypes.c_void_p)))


def render(sim, vmax=20, vmin=-20, grid: int = 0, out=None) -> np.ndarray:
    """(rows, cols, 3) uint8 frame of the current Ez, rendered on the device."""
    if out is None:
        out = np.empty((sim.local_rows, sim.cols, 3), np.uint8)
    check(lib().fdtd2d_render_snapshot(sim._h, grid, float(vmin), float(vmax), out.ctypes.data_as(ctypes.c_void_p)))
    return out


def make_video_from_frames():
    """frames/frame_%04d.png -> animation.mp4 at 15 fps with ffmpeg/libx264 (main.py:126-150)."""
    cmd = ["ffmpeg", "-y", "-framerate", "15", "-i", "frames/frame_%04d.png", "-c:v", "libx264", "-pix_fmt", "yuv420p",
           "animation.mp4"]
    try:
        subprocess.run(cmd, check=True, capture_output=True)
    except subprocess.CalledProcessError as e:  # the reference prints and carries on
        print(f"Error creating video: {e.stderr.decode()}")
