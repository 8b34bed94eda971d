"""Structure drawing on the device: the reference's `RegionDrawer` (python-src/region_drawer.py:5-87) over a
`Simulation`.

The reference draws waveguides, rings, discs and couplers with PIL on a host canvas, saves a PNG and lets
`material_init` (main.py:109-123) resize it and map gray levels to permittivity.  Here the canvas lives in HBM, one cell per
grid cell, the primitives are rasterised by CUDA kernels (fdtd2d_canvas_*) and `apply()` forms eps, the coefficient maps
and the Mur coefficient on the device -- nothing of the size of the grid ever exists on the host, which is what a
65536 x 65536 structure needs.  Same method names and arguments as the reference class; coordinates are (x, y) = (column,
global row) as in PIL.

Rasterisation rules (tests/test_structure_cpu.py, tests/test_gpu_structure.py):
  * horizontal / vertical waveguides and couplers: the rectangle PIL's wide line paints, cell for cell;
  * discs: PIL's filled ellipse, cell for cell;
  * rings: the disc of the box minus the disc of the box shrunk by the ring width (PIL's outline differs in a few cells along
    the inner edge);
  * slanted and curved waveguides: cells whose centre lies within width / 2 of the segment (PIL rounds the corners of the
    same rectangle to integers first).
"""
from __future__ import annotations

import math

import numpy as np

from ._lib import check, lib


def _round_up(f: float) -> int:  # Pillow's ROUND_UP / ROUND_DOWN (libImaging/Draw.c)
    return int(math.floor(f + 0.5)) if f >= 0 else -int(math.floor(abs(f) + 0.5))


def _round_down(f: float) -> int:
    return int(math.ceil(f - 0.5)) if f >= 0 else -int(math.ceil(abs(f) - 0.5))


def wide_line_box(x0: int, y0: int, x1: int, y1: int, width: int):
    """Inclusive (xa, ya, xb, yb) of the rectangle ImageDraw.line([(x0, y0), (x1, y1)], width=width) paints for a
    horizontal or vertical segment: the four corners Pillow's ImagingDrawWideLine computes."""
    if width <= 1 or (x0 == x1 and y0 == y1):
        return min(x0, x1), min(y0, y1), max(x0, x1), max(y0, y1)
    dx, dy = x1 - x0, y1 - y0
    hyp = math.hypot(dx, dy)
    small = (width - 1) / 2.0
    rmax, rmin = _round_up(small) / hyp, _round_down(small) / hyp
    dxmin, dxmax = _round_down(rmin * dy), _round_down(rmax * dy)
    dymin, dymax = _round_up(rmin * dx), _round_up(rmax * dx)
    xs = (x0 - dxmin, x1 - dxmin, x1 + dxmax, x0 + dxmax)
    ys = (y0 + dymax, y1 + dymax, y1 - dymin, y0 - dymin)
    return min(xs), min(ys), max(xs), max(ys)


class RegionDrawer:
    """`RegionDrawer(width, height)` of the reference, drawing into the canvas of `sim` (cols x global rows)."""

    def __init__(self, sim, grid: int = 0):
        self.sim, self.grid = sim, grid
        self.width, self.height = sim.cols, sim.global_rows
        check(lib().fdtd2d_canvas_clear(sim._h, 255))  # white background (region_drawer.py:10)

    # -- primitives ---------------------------------------------------------------------------
    def _segment(self, p0, p1, width):
        x0, y0, x1, y1 = int(p0[0]), int(p0[1]), int(p1[0]), int(p1[1])  # PIL truncates line coordinates
        if x0 == x1 or y0 == y1:
            check(lib().fdtd2d_canvas_rect(self.sim._h, self.grid, *wide_line_box(x0, y0, x1, y1, int(width)), 0))
        else:
            check(lib().fdtd2d_canvas_segment(self.sim._h, self.grid, float(x0), float(y0), float(x1), float(y1), float(width), 0))

    def draw_waveguide(self, start, end, width: int):
        """Straight waveguide between two points (region_drawer.py:13-15)."""
        self._segment(start, end, width)

    def _box(self, center, radius, w):
        return (center[0] - radius - w // 2, center[1] - radius - w // 2, center[0] + radius + w // 2, center[1] + radius + w // 2)

    def draw_ring_resonator(self, center, radius: int, ring_width: int):
        """Ring centred at a point (region_drawer.py:17-28)."""
        check(lib().fdtd2d_canvas_ellipse(self.sim._h, self.grid, *self._box(center, radius, ring_width), int(ring_width), 0))

    def draw_sphere(self, center, radius: int, sphere_width: int):
        """Filled disc (region_drawer.py:30-39)."""
        check(lib().fdtd2d_canvas_ellipse(self.sim._h, self.grid, *self._box(center, radius, sphere_width), 0, 0))

    def draw_curved_waveguide(self, start, end, control_point, width: int):
        """Quadratic Bezier waveguide: 100 points, 99 thick segments (region_drawer.py:41-64)."""
        pts = []
        for t in np.linspace(0, 1, 100):
            x = (1 - t) ** 2 * start[0] + 2 * (1 - t) * t * control_point[0] + t**2 * end[0]
            y = (1 - t) ** 2 * start[1] + 2 * (1 - t) * t * control_point[1] + t**2 * end[1]
            pts.append((x, y))
        for p0, p1 in zip(pts, pts[1:]):
            self._segment(p0, p1, width)

    def draw_directional_coupler(self, start, length: int, gap: int, waveguide_width: int):
        """Two parallel waveguides (region_drawer.py:66-82)."""
        y_offset = gap // 2 + waveguide_width // 2
        self.draw_waveguide((start[0], start[1] - y_offset), (start[0] + length, start[1] - y_offset), waveguide_width)
        self.draw_waveguide((start[0], start[1] + y_offset), (start[0] + length, start[1] + y_offset), waveguide_width)

    # -- results ------------------------------------------------------------------------------
    def image(self) -> np.ndarray:
        """The canvas as a uint8 array (local rows x cols; batch-major for batched handles)."""
        sim = self.sim
        out = np.empty(sim._shape(sim.local_rows, sim.cols), np.uint8)
        check(lib().fdtd2d_canvas_download(sim._h, out.ctypes.data))
        return out

    def save(self, filename: str):
        """Save the drawn structure as an image, like the reference (region_drawer.py:84-86)."""
        from PIL import Image

        img = self.image()
        Image.fromarray(img if img.ndim == 2 else img[self.grid]).save(filename)

    def apply(self, black_point: float = 10.0):
        """The canvas becomes the simulation's medium: what `material_init(png, rows, cols, black_point)` + set_materials
        would give for the saved picture, without the picture ever leaving the GPU."""
        sim = self.sim
        check(lib().fdtd2d_canvas_apply(sim._h, float(black_point), sim.dt, sim.dx))
