// Structure drawing on the device: the primitives of the reference's RegionDrawer (python-src/region_drawer.py:5-87 --
// straight waveguide, disc, ring, curved waveguide, directional coupler, all PIL ImageDraw calls on an "L" canvas that
// material_init then turns into a permittivity map, main.py:109-123) rasterised straight into a uint8 canvas in HBM, so a
// 65536 x 65536 structure never exists on the host.  One canvas cell per grid cell: x = column, y = GLOBAL row.
//   * rectangle: what PIL's wide line gives for a horizontal or vertical segment (region_drawer.py:13-15).
//   * filled ellipse: PIL's own integer rasterisation (Pillow's quarter-ellipse walk: from (a, b mod 2) to (a mod 2, b) in
//     doubled coordinates, each step to the neighbour -- up, up-left or left -- that lies closest to the true curve),
//     restated in ellipse_walk_kernel: one thread walks the quarter once and records the half-width of every row, the fill
//     kernel mirrors the spans.  Bit-identical to ImageDraw.ellipse(fill=...) (tests/golden/structures.npz).
//   * ring: the filled ellipse of the box minus the filled ellipse of the box shrunk by the ring width on every side.
//     (PIL's outline rule differs from this in a few pixels along the inner edge; the tests measure it.)
//   * slanted segment: the cells whose centre lies inside the rectangle of the given width around the segment
//     (ImageDraw.line(width=...) is the same rectangle with its corners rounded to integers first).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace fdtd2d {

// canvas layout: [grid][local row][col], dense; row0 = global row of local row 0
struct Canvas {
    unsigned char* px;
    int Rl, C, row0;
    long long grid_stride;
};

__global__ void canvas_fill_kernel(unsigned char* px, long long n, unsigned char v) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) px[i] = v;
}

// inclusive rectangle [x0, x1] x [y0, y1] (global rows), clipped to the canvas
__global__ void canvas_rect_kernel(Canvas cv, int grid, int x0, int y0, int x1, int y1, unsigned char v) {
    const int w = x1 - x0 + 1;
    const long long n = (long long)w * (y1 - y0 + 1);
    unsigned char* base = cv.px + (long long)grid * cv.grid_stride;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int y = y0 + (int)(i / w) - cv.row0, x = x0 + (int)(i % w);
        if (y >= 0 && y < cv.Rl && x >= 0 && x < cv.C) base[(long long)y * cv.C + x] = v;
    }
}

// Pillow's quarter walk for an ellipse with box extents a = x1 - x0, b = y1 - y0 (doubled coordinates: the centre is
// (0, 0), cells sit at x = -a, -a + 2, .., a and y = -b, .., b).  half[(y - b % 2) / 2] receives the x of the FIRST point the
// walk visits in row y, which is the largest: the filled ellipse covers -x .. x there.
__global__ void ellipse_walk_kernel(int a, int b, int* half) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    const long long a2 = (long long)a * a, b2 = (long long)b * b, a2b2 = a2 * b2;
    auto delta = [&](long long x, long long y) { return llabs(a2 * y * y + b2 * x * x - a2b2); };
    int cx = a, cy = b % 2;
    const int ex = a % 2, ey = b;
    int last_row = -1;
    for (;;) {
        const int row = (cy - b % 2) / 2;
        if (row != last_row) half[row] = cx, last_row = row;
        if (cx == ex && cy == ey) break;
        int nx = cx, ny = cy + 2;
        long long nd = delta(nx, ny);
        if (nx > 1) {
            long long d = delta(cx - 2, cy + 2);
            if (nd > d) nx = cx - 2, ny = cy + 2, nd = d;
            d = delta(cx - 2, cy);
            if (nd > d) nx = cx - 2, ny = cy;
        }
        cx = nx, cy = ny;
    }
}

// Fill the ellipse of box (x0, y0)-(x1, y1) from its row half-widths; with an inner box (ix0 <= ix1) the cells of the
// inner ellipse stay as they are (ring).
__global__ void canvas_ellipse_kernel(Canvas cv, int grid, int x0, int y0, int x1, int y1, const int* half, int ix0, int iy0, int ix1, int iy1,
                                      const int* ihalf, unsigned char v) {
    const int a = x1 - x0, b = y1 - y0, w = a + 1;
    const int ia = ix1 - ix0, ib = iy1 - iy0;
    const bool ring = ia >= 0 && ib >= 0;
    const long long n = (long long)w * (b + 1);
    unsigned char* base = cv.px + (long long)grid * cv.grid_stride;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int py = y0 + (int)(i / w), px = x0 + (int)(i % w);
        const int Y = 2 * (py - y0) - b, X = 2 * (px - x0) - a;  // doubled coordinates about the centre
        const int hw = half[((Y < 0 ? -Y : Y) - b % 2) / 2];
        if ((X < 0 ? -X : X) > hw) continue;
        if (ring && px >= ix0 && px <= ix1 && py >= iy0 && py <= iy1) {
            const int IY = 2 * (py - iy0) - ib, IX = 2 * (px - ix0) - ia;
            if ((IX < 0 ? -IX : IX) <= ihalf[((IY < 0 ? -IY : IY) - ib % 2) / 2]) continue;
        }
        const int ly = py - cv.row0;
        if (ly >= 0 && ly < cv.Rl && px >= 0 && px < cv.C) base[(long long)ly * cv.C + px] = v;
    }
}

// Thick slanted segment: cells (x, y) with |(p - p0) . n| <= width / 2 and 0 <= (p - p0) . t <= length, t the unit
// direction, n its normal; evaluated in float64 over the bounding box.
__global__ void canvas_segment_kernel(Canvas cv, int grid, double x0, double y0, double x1, double y1, double width, int bx0, int by0, int bx1,
                                      int by1, unsigned char v) {
    const int w = bx1 - bx0 + 1;
    const long long n = (long long)w * (by1 - by0 + 1);
    const double dx = x1 - x0, dy = y1 - y0, len = sqrt(dx * dx + dy * dy);
    const double tx = len > 0 ? dx / len : 1.0, ty = len > 0 ? dy / len : 0.0, hw = 0.5 * width;
    unsigned char* base = cv.px + (long long)grid * cv.grid_stride;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int py = by0 + (int)(i / w), px = bx0 + (int)(i % w);
        const double rx = px - x0, ry = py - y0;
        const double along = rx * tx + ry * ty, across = ry * tx - rx * ty;
        if (along < 0.0 || along > len || fabs(across) > hw) continue;
        const int ly = py - cv.row0;
        if (ly >= 0 && ly < cv.Rl && px >= 0 && px < cv.C) base[(long long)ly * cv.C + px] = v;
    }
}

}  // namespace fdtd2d
