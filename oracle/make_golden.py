"""Generate tests/golden/*.npz by running the REAL reference -- authoring-container script.

Run from the repo root:  python -m oracle.make_golden
Needs /root/reference (see oracle/ref_loader.py).  Every array saved here is an
output of the reference's own functions (python-src/main.py) called in the
reference driver's order (python-src/fdtd.py:30-34); nothing is produced by the
oracle or by the CUDA path.  numpy version is recorded in each file.
"""
from __future__ import annotations

import hashlib
import os

import numpy as np

from .ref_loader import load_reference_main

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
DT, DX, FC = 5e-14, 1e-4, 30e9  # fdtd.py:16-17,34


def sha(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def random_case(ref, rng, R, C, dtype):
    eps = (8.85418e-12 * (1 + 9 * rng.random((R, C)))).astype(dtype)
    mu = (np.ones((R, C)) * (4 * np.pi * 1e-7) * (1 + rng.random((R, C)))).astype(dtype)
    Ez = rng.standard_normal((R, C)).astype(dtype)
    Hx = (1e-3 * rng.standard_normal((R, C - 1))).astype(dtype)
    Hy = (1e-3 * rng.standard_normal((R - 1, C))).astype(dtype)
    return eps, mu, Ez, Hx, Hy


def make_single_call(ref):
    """One update_Hx_Hy call and one update_Ez call on random state (isolates boundary logic)."""
    rng = np.random.default_rng(20261018)
    out = {"numpy_version": np.__version__, "dt": DT, "dx": DX}
    for dtype in (np.float32, np.float64):
        for (R, C) in [(11, 11), (12, 13), (16, 11), (37, 53), (64, 48)]:
            key = f"{np.dtype(dtype).name}_{R}x{C}"
            eps, mu, Ez, Hx, Hy = random_case(ref, rng, R, C, dtype)
            for n, a in zip(("eps", "mu", "Ez0", "Hx0", "Hy0"), (eps, mu, Ez, Hx, Hy)):
                out[f"{key}_{n}"] = a.copy()
            ref.update_Hx_Hy(Ez, Hx, Hy, mu, eps, DT, DX)
            out[f"{key}_Hx1"], out[f"{key}_Hy1"] = Hx.copy(), Hy.copy()
            Ez = ref.update_Ez(Ez, Hx, Hy, mu, eps, DT, DX)
            out[f"{key}_Ez1"] = Ez.copy()
    np.savez_compressed(os.path.join(GOLDEN, "single_call.npz"), **out)


def drive(ref, Ez, Hx, Hy, mu, eps, nsteps, src, probes, kind="ricker", step0=0):
    """fdtd.py:30-34 replayed with the reference's functions (no snapshots)."""
    R, C = Ez.shape
    fn = ref.ricker if kind == "ricker" else ref.sinusoidal
    trace = np.zeros((nsteps, len(probes)), dtype=Ez.dtype)
    for n in range(nsteps):
        i = step0 + n
        Hx, Hy = ref.update_Hx_Hy(Ez, Hx, Hy, mu, eps, DT, DX)
        Ez = ref.update_Ez(Ez, Hx, Hy, mu, eps, DT, DX)
        Ez += fn(R, C, src[0], src[1], i * DT, FC)
        for p, (r, c) in enumerate(probes):
            trace[n, p] = Ez[r, c]
    return Ez, Hx, Hy, trace


def make_demo(ref):
    """fdtd.py defaults (200x200, 1000 steps, Ricker at centre), vacuum, fp64 and fp32.
    Matches SURVEY.md Appendix B."""
    R = C = 200
    probes = [(100, 100), (100, 150), (3, 3), (0, 0), (199, 199), (2, 100), (100, 197), (57, 31)]
    for dtype in (np.float64, np.float32):
        Ez, Hx, Hy = (a.astype(dtype) for a in ref.grid_init(R, C))
        eps, mu = (a.astype(dtype) for a in ref.material_init(None, R, C))
        Ez, Hx, Hy, trace = drive(ref, Ez, Hx, Hy, mu, eps, 1000, (R // 2, C // 2), probes)
        name = np.dtype(dtype).name
        np.savez_compressed(
            os.path.join(GOLDEN, f"demo200_vacuum_{name}.npz"),
            numpy_version=np.__version__, probes=np.array(probes), trace=trace, Ez=Ez, Hx=Hx, Hy=Hy,
            sha_Ez=sha(Ez), sha_Hx=sha(Hx), sha_Hy=sha(Hy),
        )
        print(name, "Ez[100,100] =", repr(Ez[100, 100]), "sha", sha(Ez)[:16], sha(Hx)[:16], sha(Hy)[:16])


def make_random_runs(ref):
    """Multi-step runs on random-eps grids with non-trivial mu, non-zero initial state and
    ragged sizes; Ricker and sinusoidal sources; nonzero step offset for the second leg."""
    rng = np.random.default_rng(7)
    out = {"numpy_version": np.__version__}
    for dtype in (np.float32, np.float64):
        for (R, C, nsteps, kind) in [(37, 53, 300, "ricker"), (96, 130, 200, "sinusoidal"), (11, 11, 40, "ricker")]:
            key = f"{np.dtype(dtype).name}_{R}x{C}_{kind}"
            eps, mu, Ez, Hx, Hy = random_case(ref, rng, R, C, dtype)
            Ez *= dtype(1e-3)
            probes = [(R // 2, C // 2), (0, 0), (R - 1, C - 1), (2, C - 3), (R - 4, 1), (5, 5), (R // 3, C // 4)]
            src = (R // 2, C // 2)
            for n, a in zip(("eps", "mu", "Ez0", "Hx0", "Hy0"), (eps, mu, Ez, Hx, Hy)):
                out[f"{key}_{n}"] = a.copy()
            Ez, Hx, Hy, trace = drive(ref, Ez, Hx, Hy, mu, eps, nsteps, src, probes, kind)
            out[f"{key}_probes"] = np.array(probes)
            out[f"{key}_src"] = np.array(src)
            out[f"{key}_nsteps"] = nsteps
            out[f"{key}_trace"] = trace
            out[f"{key}_Ez"], out[f"{key}_Hx"], out[f"{key}_Hy"] = Ez.copy(), Hx.copy(), Hy.copy()
    np.savez_compressed(os.path.join(GOLDEN, "random_runs.npz"), **out)


def make_material_and_sources(ref):
    """material_init on a PNG (main.py:109-123) and source amplitudes (main.py:182-195)."""
    from PIL import Image, ImageDraw

    png = os.path.join(GOLDEN, "structure.png")
    img = Image.new("L", (150, 110), 255)
    d = ImageDraw.Draw(img)
    d.rectangle([0, 45, 149, 62], fill=0)  # a straight waveguide
    d.ellipse([40, 5, 90, 40], fill=64)  # a grey disc
    d.polygon([(100, 70), (140, 100), (95, 105)], fill=128)
    img.save(png)
    out = {"numpy_version": np.__version__}
    for (R, C, bp) in [(64, 80, 10.0), (200, 200, 10.0), (37, 53, 4.0)]:
        eps, mu = ref.material_init(png, R, C, bp)
        out[f"eps_{R}x{C}_bp{bp:g}"] = eps
        out[f"mu_{R}x{C}_bp{bp:g}"] = mu
    eps, mu = ref.material_init(None, 23, 17)
    out["eps_none_23x17"], out["mu_none_23x17"] = eps, mu
    steps = np.array([0, 1, 2, 10, 100, 500, 666, 667, 668, 999, 5000, 123456])
    out["steps"] = steps
    out["ricker_amp"] = np.array([ref.ricker(4, 5, 1, 2, int(i) * DT, FC)[1, 2] for i in steps])
    out["sinus_amp"] = np.array([ref.sinusoidal(4, 5, 3, 4, int(i) * DT, FC)[3, 4] for i in steps])
    dense = ref.ricker(6, 7, 2, 3, 667 * DT, FC)
    out["ricker_dense_6x7"] = dense
    np.savez_compressed(os.path.join(GOLDEN, "material_sources.npz"), **out)


SMALL_SIZES = [(6, 6), (6, 23), (7, 9), (10, 10), (8, 40), (40, 7), (10, 11), (11, 10), (9, 300)]


def make_small_grids(ref):
    """Grids with fewer than 11 rows or columns: the Mur strips of opposite sides overlap, so the reference's statements
    read what earlier statements of the same step wrote.  6 is the smallest size the reference's own
    indexing allows.  One update_Hx_Hy / update_Ez call and a 30-step run per size and dtype."""
    rng = np.random.default_rng(611)
    out = {"numpy_version": np.__version__, "sizes": np.array(SMALL_SIZES)}
    for dtype in (np.float32, np.float64):
        for (R, C) in SMALL_SIZES:
            key = f"{np.dtype(dtype).name}_{R}x{C}"
            eps, mu, Ez, Hx, Hy = random_case(ref, rng, R, C, dtype)
            Ez *= dtype(1e-2)
            for n, a in zip(("eps", "mu", "Ez0", "Hx0", "Hy0"), (eps, mu, Ez, Hx, Hy)):
                out[f"{key}_{n}"] = a.copy()
            e1, h1, h2 = Ez.copy(), Hx.copy(), Hy.copy()
            ref.update_Hx_Hy(e1, h1, h2, mu, eps, DT, DX)
            out[f"{key}_Hx1"], out[f"{key}_Hy1"] = h1.copy(), h2.copy()
            e1 = ref.update_Ez(e1, h1, h2, mu, eps, DT, DX)
            out[f"{key}_Ez1"] = e1.copy()
            probes = [(R // 2, C // 2), (0, 0), (R - 1, C - 1), (1, C - 2), (R - 2, 1)]
            src = (R // 2, C // 3)
            Ez, Hx, Hy, trace = drive(ref, Ez, Hx, Hy, mu, eps, 30, src, probes, "ricker", step0=650)
            out[f"{key}_probes"], out[f"{key}_src"], out[f"{key}_trace"] = np.array(probes), np.array(src), trace
            out[f"{key}_Ez"], out[f"{key}_Hx"], out[f"{key}_Hy"] = Ez.copy(), Hx.copy(), Hy.copy()
    np.savez_compressed(os.path.join(GOLDEN, "small_grids.npz"), **out)


def main():
    import sys

    ref = load_reference_main()
    os.makedirs(GOLDEN, exist_ok=True)
    if "--small-only" not in sys.argv:
        make_single_call(ref)
        make_demo(ref)
        make_random_runs(ref)
        make_material_and_sources(ref)
    make_small_grids(ref)
    for f in sorted(os.listdir(GOLDEN)):
        print(f, os.path.getsize(os.path.join(GOLDEN, f)))


if __name__ == "__main__":
    main()
