"""CPU tests of the dataset generator's host logic and of its oracle (no GPU needed).

The golden file tests/golden/dataset.npz was produced by the REAL reference functions
(diffusion_training.py:54-146) on injected draws (oracle/make_golden_dataset.py)."""
import os

import numpy as np
import pytest

from oracle import dataset_oracle as do


class Script:
    """Scripted draws with numpy.random.Generator's two method names."""

    def __init__(self, scalars, ints):
        self.scalars, self.ints = list(scalars), list(ints)

    def random(self, dtype=None):
        return self.scalars.pop(0)

    def integers(self, lo, hi):
        v = self.ints.pop(0)
        assert lo <= v < hi
        return v


@pytest.fixture(scope="module")
def golden(golden_dir):
    return np.load(os.path.join(golden_dir, "dataset.npz"))


def test_permittivity_oracle_vs_reference_output(golden):
    """Same uniform field and sigma through the reference's F.conv2d + threshold and through the oracle's
    row-major float32 blur: the two-phase maps agree except where the blur sits within rounding of 0.5."""
    seed = int(golden["seed"])
    lo, hi = do.phase_values(np.float32)
    for g in range(3):
        R, C = (int(v) for v in golden[f"shape_{g}"])
        ref = golden[f"eps_{g}"]
        assert ref.dtype == np.float32 and set(np.unique(ref)) <= {lo, hi}  # bit-identical phase values
        mine = do.permittivity(seed, g, R, C)
        blurred = do.blur(do.uniform_field(seed, g, R, C), do.blur_weights(do.sigma_of(seed, g)))
        differ = mine != ref
        assert differ.mean() <= 1e-3
        assert np.all(np.abs(blurred[differ] - 0.5) < 1e-6)  # only ties at the threshold may flip
        assert np.all(golden[f"mu_{g}"] == np.float32(do.MU_0))


def test_source_oracle_vs_reference_output(golden):
    i = 0
    while f"src_{i}" in golden:
        dim = tuple(int(v) for v in golden[f"src_{i}_dim"])
        cells = do.random_source_cells(Script(golden[f"src_{i}_scalars"], golden[f"src_{i}_ints"]), dim)
        mine = np.zeros(dim, np.float32)
        for r, c in cells:
            mine[r, c] = 1.0
        assert np.array_equal(mine, golden[f"src_{i}"]), f"source script {i}"
        i += 1
    assert i >= 6


def test_hash_matches_library_definition():
    from fdtd2d_b200 import _lib

    rng = np.random.default_rng(0)
    for _ in range(200):
        seed, g, r, c = (int(v) for v in rng.integers(0, 1 << 20, 4))
        assert float(do.hash_uniform(seed, g, r, c)) == _lib.lib().fdtd2d_hash_uniform(seed, g, r, c)
    f = do.uniform_field(7, 3, 5, 9)
    assert f.dtype == np.float32 and f[2, 4] == np.float32(_lib.lib().fdtd2d_hash_uniform(7, 3, 2, 4))


def test_product_host_logic_matches_oracle():
    """fdtd2d_b200.dataset's host-side pieces (weights, phase values, sigma, source placement) restate the same
    reference lines as the oracle: they must agree exactly."""
    from fdtd2d_b200 import dataset as ds

    for sigma in (2.0, 3.3, 5.999):
        assert np.array_equal(ds.blur_weights(sigma), do.blur_weights(sigma))
    assert ds.phase_values(np.float32) == do.phase_values(np.float32)
    assert ds.phase_values(np.float64) == do.phase_values(np.float64)
    assert ds.sample_sigma(11, 5) == do.sigma_of(11, 5)
    rng_a, rng_b = np.random.default_rng(5), np.random.default_rng(5)
    for dim in [(256, 256), (60, 100), (64, 64)] * 20:
        assert ds.random_source_cells(rng_a, dim) == do.random_source_cells(rng_b, dim)
    plan = ds.sample_plan(50, (256, 256), 3)
    assert all(18e9 <= w < 30e9 and 2.0 <= s < 6.0 for s, _, w in plan)
    assert {len(c) for _, c, _ in plan} <= {1, 20}  # a point or a line of 10 % of the valid extent
    for _, cells, _ in plan:
        assert all(25 <= r < 231 and 25 <= c < 231 for r, c in cells)  # middle 80 %, outside the Mur ring
