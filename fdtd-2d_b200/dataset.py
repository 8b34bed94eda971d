"""Batched dataset generation on the GPU: the consumer named by BASELINE configs[4].

Mirrors the sample generator of the reference's diffusion pipeline
(python-src/diffusion_training.py:54-193): per sample a random two-phase permittivity map (uniform noise ->
15 x 15 Gaussian blur with sigma in [2, 6) -> threshold 0.5 -> eps0 / 5*eps0, uniform mu;
`generate_random_permittivity`, :54-93), a random point or short line source inside the middle 80 % of
the grid (`generate_random_source`, :96-146) and a random frequency in [18, 30) GHz (:177); the tuple
returned is shaped like `generate_data`'s: (eps, mu, src, omega, Ez), each (N, R, C) except omega (N,).

The reference obtains Ez from its frequency-domain solver; this package is the time-domain path, so Ez is
the field after `n_steps` leapfrog steps of all N grids at once (one batched `Simulation`; grids of up to
256 columns x 384 rows run on the cluster-resident kernel, i.e. one launch for the whole run), driven by a
Ricker wavelet of centre frequency omega on the sample's source cells (fdtd.py:34).

The media are generated ON THE DEVICE (fdtd2d_generate_materials_blobs) from a counter-based hash, so a
dataset is reproducible from (seed, N, R, C) alone and a CPU checker can rebuild every sample.  The random
draws that the reference takes from torch's global RNG come from `numpy.random.default_rng(seed)` and the
hash here; the distributions are the reference's.
"""
from __future__ import annotations

import ctypes

import numpy as np

from ._lib import check, lib
from .simulation import Simulation, source_table

EPS_0 = 8.85418782e-12  # diffusion_training.py:69 (note: not main.py's 8.85418e-12)
MU_0 = 1.25663706e-6    # diffusion_training.py:71
KERNEL_SIZE = 15        # diffusion_training.py:75
SIGMA_SALT = 0x5BD1E995


def blur_weights(sigma: float) -> np.ndarray:
    """The normalised 15 x 15 Gaussian of diffusion_training.py:77-81, in float32 like the reference's tensors."""
    coords = np.arange(KERNEL_SIZE, dtype=np.float32) - (KERNEL_SIZE // 2)
    xg, yg = np.meshgrid(coords, coords, indexing="ij")
    kernel = np.exp(-(xg**2 + yg**2) / (2 * sigma**2))
    kernel /= kernel.sum()
    return kernel.astype(np.float32)


def sample_sigma(seed: int, grid: int) -> float:
    """sigma = u * 4 + 2 (diffusion_training.py:76) with u from the library's counter-based hash."""
    return lib().fdtd2d_hash_uniform(seed ^ SIGMA_SALT, grid, 0xFFFFF, 0xFFFFF) * 4.0 + 2.0


def phase_values(dtype=np.float32):
    """(eps_lo, eps_hi): `mask.float() * (eps_max - eps_0) + eps_0` evaluated in float32 for mask = 0 / 1
    (diffusion_training.py:90), then cast to the run dtype."""
    f = np.float32
    lo = f(0.0) * f(5 * EPS_0 - EPS_0) + f(EPS_0)
    hi = f(1.0) * f(5 * EPS_0 - EPS_0) + f(EPS_0)
    return np.dtype(dtype).type(lo), np.dtype(dtype).type(hi)


def random_source_cells(rng: np.random.Generator, dimension) -> list[tuple[int, int]]:
    """Cells of one random source (diffusion_training.py:96-146): a point, or a horizontal / vertical line of
    at most 10 % of the valid extent, outside the outer 5 pixels and inside the middle 80 % of the grid."""
    R, C = dimension
    start_x, end_x = max(5, int(R * 0.1)), min(R - 5, R - int(R * 0.1))
    start_y, end_y = max(5, int(C * 0.1)), min(C - 5, C - int(C * 0.1))
    max_len = min(end_x - start_x, end_y - start_y) // 10
    if rng.random() < 0.5:
        if rng.random() < 0.5:  # horizontal line
            row = int(rng.integers(start_x, end_x))
            start = int(rng.integers(start_y, end_y - max_len))
            return [(row, start + j) for j in range(max_len)]
        col = int(rng.integers(start_y, end_y))
        start = int(rng.integers(start_x, end_x - max_len))
        return [(start + i, col) for i in range(max_len)]
    return [(int(rng.integers(start_x, end_x)), int(rng.integers(start_y, end_y)))]


def sample_plan(num_samples: int, dimension, seed: int):
    """Host-side random draws of a dataset: per sample (sigma, source cells, omega)."""
    rng = np.random.default_rng(seed)
    plan = []
    for b in range(num_samples):
        cells = random_source_cells(rng, dimension)
        omega = float(rng.random(dtype=np.float32)) * (30e9 - 18e9) + 18e9  # diffusion_training.py:177
        plan.append((sample_sigma(seed, b), cells, omega))
    return plan


def generate_data(num_samples: int, dimension, n_steps: int = 400, *, seed: int = 0, dx: float = 1e-3, dt: float | None = None,
                  device: int = 0, dtype=np.float32, k: int = 0, return_sim: bool = False):
    """(eps, mu, src, omega, Ez) like the reference's `generate_data` (diffusion_training.py:149-193), all
    samples stepped together on the GPU.  dx defaults to the reference's 1e-3 (:179); dt to half the vacuum
    Courant limit.  Arrays are numpy, float32 unless `dtype` says float64 (validation runs)."""
    R, C = dimension
    dtype = np.dtype(dtype)
    if dt is None:
        dt = 0.5 * dx * float(np.sqrt(EPS_0 * MU_0))
    plan = sample_plan(num_samples, dimension, seed)
    weights = np.ascontiguousarray(np.stack([blur_weights(sig) for sig, _, _ in plan]))
    eps_lo, eps_hi = phase_values(dtype)
    sim = Simulation(R, C, dtype, dt=dt, dx=dx, device=device, batch=num_samples)
    try:
        eps = np.empty((num_samples, R, C), dtype)
        check(lib().fdtd2d_generate_materials_blobs(sim._h, seed, weights.ctypes.data_as(ctypes.c_void_p), float(eps_lo), float(eps_hi),
                                                    float(dtype.type(MU_0)), sim.dt, sim.dx, eps.ctypes.data_as(ctypes.c_void_p)))
        mu = np.full((num_samples, R, C), MU_0, dtype)
        src = np.zeros((num_samples, R, C), dtype)
        cells, tables = [], np.empty((num_samples, n_steps))
        for b, (_, mine, omega) in enumerate(plan):
            for r, c in mine:
                src[b, r, c] = 1.0
                cells.append((b, r, c, b))
            tables[b] = source_table("ricker", n_steps, dt, omega)
        sim.set_sources(cells, tables)
        sim.step(n_steps, k)
        Ez = sim.read_Ez().reshape(num_samples, R, C)
        omega = np.array([w for _, _, w in plan], dtype=np.float32)
        if return_sim:
            out, sim = (eps, mu, src, omega, Ez, sim), None
            return out
        return eps, mu, src, omega, Ez
    finally:
        if sim is not None:
            sim.close()
