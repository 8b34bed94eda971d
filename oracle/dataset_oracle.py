"""CPU restatement of the dataset generator's media and sources -- TEST INFRASTRUCTURE ONLY.

Follows (does not copy) the reference:
  generate_random_permittivity   python-src/diffusion_training.py:54-93
  generate_random_source         python-src/diffusion_training.py:96-146
and the product's own definition of where the random numbers come from (the counter-based hash of
csrc/common.cuh, restated here in numpy) and of the blur's accumulation order (row-major over the 15 x 15
taps, float32, one rounding per multiply and per add).

Parity pin: tests/golden/dataset.npz holds outputs of the REAL reference functions run in the authoring
container on injected draws (oracle/make_golden_dataset.py); tests/test_dataset_oracle.py checks this module
against them.
"""
from __future__ import annotations

import numpy as np

EPS_0 = 8.85418782e-12  # diffusion_training.py:69
MU_0 = 1.25663706e-6    # diffusion_training.py:71
K = 15                  # diffusion_training.py:75
SIGMA_SALT = 0x5BD1E995
M64 = (1 << 64) - 1


def hash_uniform(seed: int, grid, row, col) -> np.ndarray:
    """splitmix64 finaliser over (seed, grid, row, col) -> 24 random bits / 2^24 (csrc/common.cuh), vectorised."""
    with np.errstate(over="ignore"):
        key = (np.asarray(grid, np.uint64) << np.uint64(40)) ^ (np.asarray(row, np.uint64) << np.uint64(20)) ^ np.asarray(col, np.uint64)
        z = key + np.uint64((seed * 0x9E3779B97F4A7C15 + 0x632BE59BD9B4E019) & M64)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(40)).astype(np.float64) * (1.0 / 16777216.0)


def uniform_field(seed: int, grid: int, R: int, C: int) -> np.ndarray:
    i, j = np.meshgrid(np.arange(R, dtype=np.uint64), np.arange(C, dtype=np.uint64), indexing="ij")
    return hash_uniform(seed, grid, i, j).astype(np.float32)


def sigma_of(seed: int, grid: int) -> float:
    return float(hash_uniform(seed ^ SIGMA_SALT, grid, 0xFFFFF, 0xFFFFF)) * 4.0 + 2.0  # diffusion_training.py:76


def blur_weights(sigma: float) -> np.ndarray:
    coords = np.arange(K, dtype=np.float32) - (K // 2)  # diffusion_training.py:77-81
    xg, yg = np.meshgrid(coords, coords, indexing="ij")
    kernel = np.exp(-(xg**2 + yg**2) / (2 * sigma**2))
    kernel /= kernel.sum()
    return kernel.astype(np.float32)


def blur(u: np.ndarray, w: np.ndarray) -> np.ndarray:
    """15 x 15 correlation with zero padding (F.conv2d(..., padding=7), diffusion_training.py:86-89), float32,
    taps accumulated row-major."""
    R, C = u.shape
    pad = np.zeros((R + K - 1, C + K - 1), np.float32)
    pad[K // 2:K // 2 + R, K // 2:K // 2 + C] = u
    acc = np.zeros((R, C), np.float32)
    for ky in range(K):
        for kx in range(K):
            acc = acc + w[ky, kx] * pad[ky:ky + R, kx:kx + C]
    return acc


def phase_values(dtype=np.float32):
    f = np.float32  # `mask.float() * (eps_max - eps_0) + eps_0` in float32 (diffusion_training.py:90)
    lo = f(0.0) * f(5 * EPS_0 - EPS_0) + f(EPS_0)
    hi = f(1.0) * f(5 * EPS_0 - EPS_0) + f(EPS_0)
    return np.dtype(dtype).type(lo), np.dtype(dtype).type(hi)


def permittivity(seed: int, grid: int, R: int, C: int, dtype=np.float32) -> np.ndarray:
    lo, hi = phase_values(dtype)
    b = blur(uniform_field(seed, grid, R, C), blur_weights(sigma_of(seed, grid)))
    return np.where(b > np.float32(0.5), hi, lo).astype(dtype)


def random_source_cells(draw, dimension):
    """diffusion_training.py:96-146 with the draws taken from `draw.random()` / `draw.integers(lo, hi)`."""
    R, C = dimension
    margin = 5
    start_x, end_x, start_y, end_y = margin, R - margin, margin, C - margin
    mid_x, mid_y = int(R * 0.1), int(C * 0.1)
    start_x, end_x = max(start_x, mid_x), min(end_x, R - mid_x)
    start_y, end_y = max(start_y, mid_y), min(end_y, C - mid_y)
    max_len = min(end_x - start_x, end_y - start_y) // 10
    if draw.random() < 0.5:
        if draw.random() < 0.5:
            row = int(draw.integers(start_x, end_x))
            start = int(draw.integers(start_y, end_y - max_len))
            return [(row, c) for c in range(start, start + max_len)]
        col = int(draw.integers(start_y, end_y))
        start = int(draw.integers(start_x, end_x - max_len))
        return [(r, col) for r in range(start, start + max_len)]
    row = int(draw.integers(start_x, end_x))
    col = int(draw.integers(start_y, end_y))
    return [(row, col)]
