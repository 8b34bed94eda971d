"""CPU restatement of the structure rasteriser (fdtd-2d_b200/csrc/structure.cuh) in numpy -- TEST INFRASTRUCTURE ONLY.

What it follows: the reference's RegionDrawer (python-src/region_drawer.py:5-87) draws with PIL's ImageDraw; the
rasterisation itself therefore lives in Pillow (third-party, 12.2.0 in the authoring container), not in the reference
tree.  Restated here from Pillow's published algorithms (libImaging/Draw.c: ImagingDrawWideLine's corner arithmetic,
quarter_next's integer ellipse walk) and pinned by masks the reference's own class produced in the authoring container
(oracle/make_golden_structures.py -> tests/golden/structures.npz):
  * horizontal / vertical waveguides, couplers and discs: equal to the reference's output cell for cell;
  * rings and slanted / curved waveguides: the rules documented in structure.cuh, which differ from PIL in a few boundary
    cells -- the tests bound the difference.
"""
from __future__ import annotations

import math

import numpy as np


def _round_up(f):
    return int(math.floor(f + 0.5)) if f >= 0 else -int(math.floor(abs(f) + 0.5))


def _round_down(f):
    return int(math.ceil(f - 0.5)) if f >= 0 else -int(math.ceil(abs(f) - 0.5))


def quarter_half_widths(a: int, b: int) -> np.ndarray:
    """Pillow's quarter walk (doubled coordinates) -> x of the first point visited in each row y = b%2, b%2+2, .., b."""
    half = np.zeros(b // 2 + 1, np.int64)
    a2, b2 = a * a, b * b
    a2b2 = a2 * b2
    delta = lambda x, y: abs(a2 * y * y + b2 * x * x - a2b2)
    cx, cy, ex, ey = a, b % 2, a % 2, b
    seen = -1
    while True:
        row = (cy - b % 2) // 2
        if row != seen:
            half[row], seen = cx, row
        if cx == ex and cy == ey:
            return half
        nx, ny = cx, cy + 2
        nd = delta(nx, ny)
        if nx > 1:
            d = delta(cx - 2, cy + 2)
            if nd > d:
                nx, ny, nd = cx - 2, cy + 2, d
            d = delta(cx - 2, cy)
            if nd > d:
                nx, ny = cx - 2, cy
        cx, cy = nx, ny


def ellipse_mask(shape, box) -> np.ndarray:
    """ImageDraw.ellipse(box, fill=...) as a boolean mask of `shape` = (rows, cols)."""
    x0, y0, x1, y1 = box
    m = np.zeros(shape, bool)
    if x1 < x0 or y1 < y0 or (x1 - x0) + (y1 - y0) < 1:  # (Pillow fills with "width" a + b: a one-cell box draws nothing)
        return m
    a, b = x1 - x0, y1 - y0
    half = quarter_half_widths(a, b)
    for py in range(max(0, y0), min(shape[0], y1 + 1)):
        Y = 2 * (py - y0) - b
        hw = int(half[(abs(Y) - b % 2) // 2])
        # cells with |2 (px - x0) - a| <= hw
        lo, hi = x0 + (a - hw + 1) // 2, x0 + (a + hw) // 2
        m[py, max(0, lo):max(0, min(shape[1], hi + 1))] = True
    return m


class Canvas:
    """uint8 canvas, (rows, cols), white; the same primitives as the device."""

    def __init__(self, cols, rows):
        self.img = np.full((rows, cols), 255, np.uint8)

    def rect(self, x0, y0, x1, y1):
        R, C = self.img.shape
        x0, y0, x1, y1 = max(x0, 0), max(y0, 0), min(x1, C - 1), min(y1, R - 1)
        if x0 <= x1 and y0 <= y1:
            self.img[y0:y1 + 1, x0:x1 + 1] = 0

    def ellipse(self, box, width=0):
        m = ellipse_mask(self.img.shape, box)
        x0, y0, x1, y1 = box
        if width > 0 and (x1 - x0) - 2 * width >= 0 and (y1 - y0) - 2 * width >= 0:
            m &= ~ellipse_mask(self.img.shape, (x0 + width, y0 + width, x1 - width, y1 - width))
        self.img[m] = 0

    def segment(self, x0, y0, x1, y1, width):
        R, C = self.img.shape
        dx, dy = x1 - x0, y1 - y0
        ln = math.sqrt(dx * dx + dy * dy)
        tx, ty = (dx / ln, dy / ln) if ln > 0 else (1.0, 0.0)
        hw, pad = 0.5 * width, 0.5 * width + 1.0
        bx0, bx1 = max(0, math.floor(min(x0, x1) - pad)), min(C - 1, math.ceil(max(x0, x1) + pad))
        by0, by1 = max(0, math.floor(min(y0, y1) - pad)), min(R - 1, math.ceil(max(y0, y1) + pad))
        if bx0 > bx1 or by0 > by1:
            return
        yy, xx = np.mgrid[by0:by1 + 1, bx0:bx1 + 1].astype(np.float64)
        rx, ry = xx - x0, yy - y0
        along, across = rx * tx + ry * ty, ry * tx - rx * ty
        self.img[by0:by1 + 1, bx0:bx1 + 1][(along >= 0.0) & (along <= ln) & (np.abs(across) <= hw)] = 0


class RegionDrawer:
    """The reference class's interface (region_drawer.py:5-87) over `Canvas`."""

    def __init__(self, width, height):
        self.c = Canvas(width, height)

    def _segment(self, p0, p1, width):
        x0, y0, x1, y1 = int(p0[0]), int(p0[1]), int(p1[0]), int(p1[1])
        if x0 == x1 or y0 == y1:
            if width <= 1 or (x0 == x1 and y0 == y1):
                self.c.rect(min(x0, x1), min(y0, y1), max(x0, x1), max(y0, y1))
                return
            dx, dy = x1 - x0, y1 - y0
            hyp = math.hypot(dx, dy)
            small = (width - 1) / 2.0
            rmax, rmin = _round_up(small) / hyp, _round_down(small) / hyp
            dxmin, dxmax = _round_down(rmin * dy), _round_down(rmax * dy)
            dymin, dymax = _round_up(rmin * dx), _round_up(rmax * dx)
            xs = (x0 - dxmin, x1 - dxmin, x1 + dxmax, x0 + dxmax)
            ys = (y0 + dymax, y1 + dymax, y1 - dymin, y0 - dymin)
            self.c.rect(min(xs), min(ys), max(xs), max(ys))
        else:
            self.c.segment(float(x0), float(y0), float(x1), float(y1), float(width))

    def draw_waveguide(self, start, end, width):
        self._segment(start, end, width)

    def _box(self, center, radius, w):
        return (center[0] - radius - w // 2, center[1] - radius - w // 2, center[0] + radius + w // 2, center[1] + radius + w // 2)

    def draw_ring_resonator(self, center, radius, ring_width):
        self.c.ellipse(self._box(center, radius, ring_width), ring_width)

    def draw_sphere(self, center, radius, sphere_width):
        self.c.ellipse(self._box(center, radius, sphere_width), 0)

    def draw_curved_waveguide(self, start, end, control_point, width):
        pts = []
        for t in np.linspace(0, 1, 100):
            x = (1 - t) ** 2 * start[0] + 2 * (1 - t) * t * control_point[0] + t**2 * end[0]
            y = (1 - t) ** 2 * start[1] + 2 * (1 - t) * t * control_point[1] + t**2 * end[1]
            pts.append((x, y))
        for p0, p1 in zip(pts, pts[1:]):
            self._segment(p0, p1, width)

    def draw_directional_coupler(self, start, length, gap, waveguide_width):
        y_offset = gap // 2 + waveguide_width // 2
        self.draw_waveguide((start[0], start[1] - y_offset), (start[0] + length, start[1] - y_offset), waveguide_width)
        self.draw_waveguide((start[0], start[1] + y_offset), (start[0] + length, start[1] + y_offset), waveguide_width)

    @property
    def image(self):
        return self.c.img


# the scenes of tests/golden/structures.npz: name -> (cols, rows, [(method, args), ...], exact?)
SCENES = {
    "box": (1000, 1000, [("draw_waveguide", ((80, 100), (920, 100), 40)), ("draw_waveguide", ((80, 900), (920, 900), 40)),
                         ("draw_waveguide", ((100, 80), (100, 920), 40)), ("draw_waveguide", ((900, 80), (900, 920), 40))], True),
    "waveguides": (300, 200, [("draw_waveguide", ((10, 20), (250, 20), 7)), ("draw_waveguide", ((280, 60), (30, 60), 12)),
                              ("draw_waveguide", ((40, 90), (40, 190), 9)), ("draw_waveguide", ((90, 185), (90, 95), 16)),
                              ("draw_waveguide", ((150, 100), (290, 100), 1)), ("draw_waveguide", ((200, 120), (200, 199), 2)),
                              ("draw_waveguide", ((-20, 150), (70, 150), 5))], True),
    "coupler": (400, 160, [("draw_directional_coupler", ((30, 80), 330, 10, 14)), ("draw_directional_coupler", ((50, 30), 100, 5, 3))], True),
    "discs": (320, 260, [("draw_sphere", ((60, 60), 40, 6)), ("draw_sphere", ((200, 70), 55, 9)), ("draw_sphere", ((90, 190), 3, 1)),
                         ("draw_sphere", ((250, 200), 70, 0)), ("draw_sphere", ((160, 130), 1, 0))], True),
    "rings": (360, 300, [("draw_ring_resonator", ((100, 100), 70, 12)), ("draw_ring_resonator", ((250, 180), 90, 25)),
                         ("draw_ring_resonator", ((60, 240), 30, 3))], False),
    "curved": (400, 300, [("draw_curved_waveguide", ((20, 250), (380, 260), (200, -60), 14)), ("draw_waveguide", ((30, 30), (300, 120), 10)),
                          ("draw_waveguide", ((350, 40), (250, 280), 21))], False),
    "device": (500, 400, [("draw_waveguide", ((0, 60), (499, 60), 16)), ("draw_ring_resonator", ((250, 170), 90, 16)),
                          ("draw_waveguide", ((0, 280), (499, 280), 16)), ("draw_sphere", ((420, 350), 25, 4))], False),
}


def draw_scene(drawer, name):
    for method, args in SCENES[name][2]:
        getattr(drawer, method)(*args)
    return drawer
