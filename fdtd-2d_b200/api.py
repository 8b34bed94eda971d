"""The reference's Python call surface for the FDTD path, backed by the CUDA library.

python-src/fdtd.py:1-9 imports these names from the reference's `main` module:
    grid_init, material_init, update_Hx_Hy, update_Ez, ricker (+ sinusoidal, main.py:190)
Same names, argument order, in-place mutation and return values here, so the driver loop
fdtd.py:30-34 runs unchanged with `from fdtd2d_b200 import ...`.  The array-in/array-out kernels are
parity shims (upload, one GPU pass, download -- they exist so each reference function can be checked in
isolation and so existing scripts keep working); the fast path is `Simulation`.

No CPU fallback: `update_Hx_Hy` / `update_Ez` raise if libfdtd2d.so is missing or no GPU is present.
Unlike the reference module, importing this one has no filesystem side effect (main.py:7-9 wipes
./frames on import; deliberately not replicated).
"""
from __future__ import annotations

import numpy as np

from . import _lib
from .simulation import Simulation, ricker_amplitude, sinusoidal_amplitude

EPSILON0 = 8.85418e-12  # main.py:100
MU0 = 4 * np.pi * 1e-7  # main.py:101

_handles: dict = {}


def grid_init(rows: int, cols: int):
    """Zero fields (Ez, Hx, Hy) of shapes (R,C), (R,C-1), (R-1,C), float64 (main.py:79-85)."""
    return np.zeros((rows, cols)), np.zeros((rows, cols - 1)), np.zeros((rows - 1, cols))


def material_init(path, rows: int, cols: int, black_point: float = 10.0):
    """(eps, mu) float64 maps (main.py:88-123).  path None -> vacuum; else a grayscale image where
    black maps to black_point*eps0 and white to eps0.  Host-side setup, as in the reference."""
    mu = np.ones((rows, cols)) * MU0
    if path is None:
        return np.ones((rows, cols)) * EPSILON0, mu
    from PIL import Image

    gray = np.array(Image.open(path).convert("L").resize((cols, rows), Image.LANCZOS), dtype=float) / 255.0
    eps = (1 + (black_point - 1) * (1.0 - gray)) * EPSILON0
    return eps, mu


def ricker(rows, cols, x_pos, y_pos, t, fc):
    """Dense float64 (rows, cols) source array with one non-zero cell (main.py:182-187)."""
    src = np.zeros((rows, cols), dtype=float)
    src[x_pos, y_pos] = ricker_amplitude(t, fc)
    return src


def sinusoidal(rows, cols, x_pos, y_pos, t, fc):
    """Dense float64 (rows, cols) ramped-sine source array (main.py:190-195)."""
    src = np.zeros((rows, cols), dtype=float)
    src[x_pos, y_pos] = sinusoidal_amplitude(t, fc)
    return src


def _sim_for(Ez, dt, dx) -> Simulation:
    key = (Ez.shape, Ez.dtype.str, float(dt), float(dx))
    sim = _handles.get(key)
    if sim is None:
        if len(_handles) >= 4:  # keep the shim's device footprint bounded
            _handles.pop(next(iter(_handles))).close()
        sim = Simulation(Ez.shape[0], Ez.shape[1], Ez.dtype, dt=dt, dx=dx)
        _handles[key] = sim
    return sim


def _check_arrays(Ez, Hx, Hy, mu, eps):
    R, C = Ez.shape
    if Hx.shape != (R, C - 1) or Hy.shape != (R - 1, C) or mu.shape != (R, C) or eps.shape != (R, C):
        raise ValueError("operands could not be broadcast together: expected Ez (R,C), Hx (R,C-1), Hy (R-1,C), "
                         f"mu/eps (R,C); got {Ez.shape}, {Hx.shape}, {Hy.shape}, {mu.shape}, {eps.shape}")
    if Ez.dtype not in (np.float32, np.float64) or Hx.dtype != Ez.dtype or Hy.dtype != Ez.dtype:
        raise TypeError("Ez, Hx, Hy must share dtype float32 or float64")


def _one_pass(Ez, Hx, Hy, mu, eps, dt, dx, phases):
    _check_arrays(Ez, Hx, Hy, mu, eps)
    sim = _sim_for(Ez, dt, dx)
    sim.set_materials(eps, mu)
    sim.set_state(Ez, Hx, Hy)
    sim.step_phases(phases)
    return sim


def update_Hx_Hy(Ez, Hx, Hy, mu, eps, dt, dx):
    """H half-step on the GPU (main.py:66-76): mutates Hx, Hy in place and returns them."""
    sim = _one_pass(Ez, Hx, Hy, mu, eps, dt, dx, _lib.PHASE_H)
    _, hx, hy = sim.state()
    Hx[...] = hx
    Hy[...] = hy
    return Hx, Hy


def update_Ez(Ez, Hx, Hy, mu, eps, dt, dx):
    """Ez step on the GPU -- interior update, 5-px Mur ABC, corner means (main.py:12-63): mutates Ez in
    place and returns it."""
    sim = _one_pass(Ez, Hx, Hy, mu, eps, dt, dx, _lib.PHASE_E)
    Ez[...] = sim.read_Ez()
    return Ez


def capture_snapshot(Ez, eps, path, vmax=20, vmin=-20):
    """Render Ez over the permittivity background and save it as an image (main.py:153-179).  The colour
    mapping and blending run on the GPU; only the encoded image is written by PIL, as in the reference."""
    from PIL import Image

    from . import snapshot

    Ez = np.asarray(Ez)
    if Ez.dtype not in (np.float32, np.float64):
        Ez = Ez.astype(np.float64)
    sim = _sim_for(Ez, 0.0, 0.0)
    R, C = Ez.shape
    check_state = (np.zeros((R, C - 1), Ez.dtype), np.zeros((R - 1, C), Ez.dtype))
    sim.set_state(Ez, *check_state)
    snapshot.set_background(sim, eps)
    Image.fromarray(snapshot.render(sim, vmax, vmin)).save(path)


def release_handles():
    """Free the device buffers cached by the array-in/array-out shims."""
    while _handles:
        _handles.popitem()[1].close()
