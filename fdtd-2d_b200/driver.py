"""The reference driver python-src/fdtd.py:14-40 on the device-resident engine.

Same parameters, same order of operations and the same outputs (frames/frame_%04d.png every
nsteps // nframes steps, then animation.mp4), but the loop body runs on the GPU: the fields never leave
the device between frames, and a frame leaves it as 3 bytes per cell (device-side colour mapping).

    python -m fdtd2d_b200.driver [structure.png]
"""
from __future__ import annotations

import os
import sys

import numpy as np

from .api import material_init
from .simulation import Simulation, courant_number
from .snapshot import make_video_from_frames


def run(structure=None, rows=200, cols=200, dt=5e-14, dx=1e-4, nsteps=1000, nframes=200, fc=30e9, dtype=np.float64, vmax=1e-3,
        vmin=-1e-3, frames_dir="frames", video=True, device=0, on_frame=None):
    """fdtd.py's `__main__`.  Returns the final (Ez, Hx, Hy).  `on_frame(frame_num, rgb)` replaces writing PNGs
    when given (tests use it); `structure=None` is the uniform medium."""
    eps, mu = material_init(structure, rows, cols)
    courant = courant_number(eps, mu, dt, dx)  # fdtd.py:25-28
    print(courant)
    assert courant <= 1.0, f"Courant stability condition not met: {courant} > 1.0"
    every = nsteps // nframes
    if on_frame is None:
        from PIL import Image

        os.makedirs(frames_dir, exist_ok=True)
    with Simulation(rows, cols, dtype, dt=dt, dx=dx, device=device) as sim:
        sim.set_materials(eps, mu)
        sim.set_point_source(rows // 2, cols // 2, nsteps, fc)  # fdtd.py:34
        sim.set_snapshot_background(eps)
        done = 0
        for i in range(0, nsteps, every):  # the steps after which fdtd.py:36 takes a snapshot: i % every == 0
            sim.step(i + 1 - done)
            done = i + 1
            rgb = sim.render_snapshot(vmax, vmin)
            if on_frame is not None:
                on_frame(i // every, rgb)
            else:
                Image.fromarray(rgb).save(os.path.join(frames_dir, f"frame_{i // every:04d}.png"))
        sim.step(nsteps - done)
        out = sim.state()
    if video and on_frame is None:
        make_video_from_frames()
    return out


if __name__ == "__main__":
    run(sys.argv[1] if len(sys.argv) > 1 else None)
