"""numpy restatement of the reference's FDTD leapfrog path -- TEST INFRASTRUCTURE ONLY.

Follows (does not copy) the reference:
  update_Hx_Hy   python-src/main.py:66-76
  update_Ez      python-src/main.py:12-63   (interior 18-27, Mur 29-51, corners 53-61)
  grid_init      python-src/main.py:79-85
  material_init  python-src/main.py:88-123
  ricker         python-src/main.py:182-187
  sinusoidal     python-src/main.py:190-195
  leapfrog loop  python-src/fdtd.py:30-34   (H -> Ez(+Mur+corners) -> Ez += source(i*dt))

The reference writes the Mur boundary and the corner averaging as sequential
in-place loops.  For rows, cols >= 11 every read inside those loops sees a value
that has not yet been overwritten by the same loop, so each loop is one
whole-slice assignment whose right-hand side is evaluated before the store
(stages S0..S4).  That is the form used here; for
rows or cols < 11 the statement-by-statement form is used instead so the
function stays faithful for every size.  Arithmetic order, parenthesisation and
dtype promotion (NEP 50: python-float ``dt``/``dx`` are weak scalars) are the
reference's.  Requires numpy >= 2.

Parity pin: tests/golden/*.npz were produced by running the real reference
(oracle/make_golden.py); tests/test_oracle_golden.py checks this module against
them bit for bit.
"""
from __future__ import annotations

import numpy as np

if int(np.__version__.split(".")[0]) < 2:  # pragma: no cover
    raise ImportError("numpy >= 2 (NEP 50 scalar promotion) is required by the oracle")

EPSILON0 = 8.85418e-12  # main.py:100
MU0 = 4 * np.pi * 1e-7  # main.py:101
RING = 5  # depth of the Mur boundary, main.py:33,38,43,48


# ---------------------------------------------------------------- setup ----
def grid_init(rows: int, cols: int, dtype=np.float64):
    """Zero state in the reference's three shapes (main.py:79-85)."""
    return (
        np.zeros((rows, cols), dtype=dtype),
        np.zeros((rows, cols - 1), dtype=dtype),
        np.zeros((rows - 1, cols), dtype=dtype),
    )


def material_init(path, rows: int, cols: int, black_point: float = 10.0):
    """eps/mu maps (main.py:88-123): uniform vacuum, or grayscale image -> eps."""
    mu = np.ones((rows, cols)) * MU0
    if path is None:
        return np.ones((rows, cols)) * EPSILON0, mu
    from PIL import Image

    img = Image.open(path).convert("L").resize((cols, rows), Image.LANCZOS)
    arr = np.array(img, dtype=float) / 255.0
    inv = 1.0 - arr
    factor = 1 + (black_point - 1) * inv
    return factor * EPSILON0, mu


# ---------------------------------------------------------- coefficients ----
def h_coeff(mu, dt, dx):
    """dt/(mu*dx) exactly as formed at main.py:70,74 (product first, then divide)."""
    return dt / (mu * dx)


def e_coeff(eps, dt, dx):
    """dt/(eps*dx) exactly as formed at main.py:27."""
    return dt / (eps * dx)


def mur_coef(mu, eps, dt, dx):
    """(c*dt-dx)/(c*dt+dx) with c from the corner cell only (main.py:30-31)."""
    c = 1 / np.sqrt(mu[0, 0] * eps[0, 0])
    return (c * dt - dx) / (c * dt + dx)


# --------------------------------------------------------------- kernels ----
def update_Hx_Hy(Ez, Hx, Hy, mu, eps, dt, dx):
    """H half-step (main.py:66-76). In place; returns (Hx, Hy). ``eps`` unused."""
    ch = h_coeff(mu[:-1, :-1], dt, dx)
    Hx[:-1, :] -= ch * (Ez[1:, :-1] - Ez[:-1, :-1])
    Hy[:, :-1] += ch * (Ez[:-1, 1:] - Ez[:-1, :-1])
    return Hx, Hy


def _update_Ez_sequential(Ez, S0, coef):
    """Mur + corners, statement by statement (main.py:33-61); used for tiny grids."""
    for k in range(RING):
        Ez[1:-1, k] = S0[1:-1, k + 1] + coef * (Ez[1:-1, k + 1] - S0[1:-1, k])
    for k in range(RING):
        Ez[1:-1, -(k + 1)] = S0[1:-1, -(k + 2)] + coef * (Ez[1:-1, -(k + 2)] - S0[1:-1, -(k + 1)])
    for k in range(RING):
        Ez[k, 1:-1] = S0[k + 1, 1:-1] + coef * (Ez[k + 1, 1:-1] - S0[k, 1:-1])
    for k in range(RING):
        Ez[-(k + 1), 1:-1] = S0[-(k + 2), 1:-1] + coef * (Ez[-(k + 2), 1:-1] - S0[-(k + 1), 1:-1])
    for a in range(RING):
        for b in range(RING):
            Ez[a, b] = (Ez[a, b + 1] + Ez[a + 1, b]) / 2
            Ez[a, -b - 1] = (Ez[a, -b - 2] + Ez[a + 1, -b - 1]) / 2
            Ez[-a - 1, b] = (Ez[-a - 2, b] + Ez[-a - 1, b + 1]) / 2
            Ez[-a - 1, -b - 1] = (Ez[-a - 2, -b - 1] + Ez[-a - 1, -b - 2]) / 2


def update_Ez(Ez, Hx, Hy, mu, eps, dt, dx):
    """Ez step: interior curl-H update, 5-px Mur ABC (L,R,T,B), 5x5 corner means
    (main.py:12-63). In place; returns Ez."""
    R, C = Ez.shape
    S0 = Ez.copy()  # main.py:18
    curl = (Hy[1:, 1:-1] - Hy[1:, :-2]) - (Hx[1:-1, 1:] - Hx[:-2, 1:])
    Ez[1:-1, 1:-1] += curl * e_coeff(eps[1:-1, 1:-1], dt, dx)  # S1
    coef = mur_coef(mu, eps, dt, dx)
    if R < 2 * RING + 1 or C < 2 * RING + 1:
        _update_Ez_sequential(Ez, S0, coef)
        return Ez
    n = RING
    # S2: left then right columns (rows 1..R-2). RHS is evaluated before the store.
    Ez[1:-1, 0:n] = S0[1:-1, 1 : n + 1] + coef * (Ez[1:-1, 1 : n + 1] - S0[1:-1, 0:n])
    Ez[1:-1, C - n : C] = S0[1:-1, C - n - 1 : C - 1] + coef * (
        Ez[1:-1, C - n - 1 : C - 1] - S0[1:-1, C - n : C]
    )
    # S3: top then bottom rows (cols 1..C-2)
    Ez[0:n, 1:-1] = S0[1 : n + 1, 1:-1] + coef * (Ez[1 : n + 1, 1:-1] - S0[0:n, 1:-1])
    Ez[R - n : R, 1:-1] = S0[R - n - 1 : R - 1, 1:-1] + coef * (
        Ez[R - n - 1 : R - 1, 1:-1] - S0[R - n : R, 1:-1]
    )
    # S4: four 5x5 corner blocks, each the mean of the two inward neighbours
    Ez[0:n, 0:n] = (Ez[0:n, 1 : n + 1] + Ez[1 : n + 1, 0:n]) / 2
    Ez[0:n, C - n : C] = (Ez[0:n, C - n - 1 : C - 1] + Ez[1 : n + 1, C - n : C]) / 2
    Ez[R - n : R, 0:n] = (Ez[R - n - 1 : R - 1, 0:n] + Ez[R - n : R, 1 : n + 1]) / 2
    Ez[R - n : R, C - n : C] = (Ez[R - n - 1 : R - 1, C - n : C] + Ez[R - n : R, C - n - 1 : C - 1]) / 2
    return Ez


# --------------------------------------------------------------- sources ----
def ricker_amplitude(t, fc):
    """Ricker wavelet value, float64 (main.py:183-184)."""
    tau = np.pi * fc * (t - 1 / fc)
    return (1 - 2 * tau**2) * np.exp(-(tau**2))


def sinusoidal_amplitude(t, fc):
    """Ramped sine value, float64 (main.py:193-194)."""
    envelope = 1 - np.exp(-((t - 3000 / fc) ** 2) / (2 * (2 / fc) ** 2))
    return envelope * np.sin(2 * np.pi * fc * t)


def ricker(rows, cols, x_pos, y_pos, t, fc):
    """Dense float64 (rows, cols) array with one non-zero cell (main.py:182-187)."""
    src = np.zeros((rows, cols), dtype=float)
    src[x_pos, y_pos] = ricker_amplitude(t, fc)
    return src


def sinusoidal(rows, cols, x_pos, y_pos, t, fc):
    """Dense float64 (rows, cols) array with one non-zero cell (main.py:190-195)."""
    src = np.zeros((rows, cols), dtype=float)
    src[x_pos, y_pos] = sinusoidal_amplitude(t, fc)
    return src


def source_table(kind, nsteps, dt, fc, step0=0):
    """float64 amplitudes amp[i] = source((step0+i)*dt), evaluated one step at a time
    exactly as the driver does (fdtd.py:34: ``i * dt`` with python ints/floats)."""
    fn = {"ricker": ricker_amplitude, "sinusoidal": sinusoidal_amplitude}[kind]
    return np.array([fn((step0 + i) * dt, fc) for i in range(nsteps)], dtype=np.float64)


# ---------------------------------------------------------------- driver ----
def run(Ez, Hx, Hy, mu, eps, dt, dx, nsteps, source=None, step0=0, probes=None, dense_source=True):
    """Replay fdtd.py:30-34 for ``nsteps`` steps starting at step index ``step0``.

    source: None or (row, col, fc, kind) with kind in {"ricker", "sinusoidal"}.
    probes: optional list of (row, col); returns an (nsteps, nprobes) array of Ez
            sampled after the source add of each step, in Ez's dtype.
    dense_source: True adds a dense float64 array as the reference does
            (``Ez += ricker(...)``); False adds to the one cell only -- identical
            bits (float64 add then cast), far cheaper, used for long oracle runs.
    """
    R, C = Ez.shape
    trace = None
    if probes is not None:
        trace = np.zeros((nsteps, len(probes)), dtype=Ez.dtype)
    for n in range(nsteps):
        i = step0 + n
        update_Hx_Hy(Ez, Hx, Hy, mu, eps, dt, dx)
        update_Ez(Ez, Hx, Hy, mu, eps, dt, dx)
        if source is not None:
            sr, sc, fc, kind = source
            if dense_source:
                fn = ricker if kind == "ricker" else sinusoidal
                Ez += fn(R, C, sr, sc, i * dt, fc)
            else:
                fn = ricker_amplitude if kind == "ricker" else sinusoidal_amplitude
                Ez[sr, sc] = Ez.dtype.type(np.float64(Ez[sr, sc]) + fn(i * dt, fc))
        if trace is not None:
            for p, (pr, pc) in enumerate(probes):
                trace[n, p] = Ez[pr, pc]
    return Ez, Hx, Hy, trace


def courant(eps, mu, dt, dx):
    """Courant number as the driver computes it (fdtd.py:25-26)."""
    c = 1 / np.sqrt(eps.min() * mu.min())
    return (c * dt) / dx


# -------------------------------------------------------------- readout ----
def seismic_lut(n: int = 256):
    """matplotlib's "seismic" lookup table rebuilt from its published construction: five anchor colours
    (matplotlib/_cm.py `_seismic_data`) evenly spaced on [0, 1], `LinearSegmentedColormap.from_list`
    -> `_create_lookup_table(N=256, gamma=1)`.  matplotlib itself is absent from the authoring container; the table
    is PINNED by images the reference's own colour pipeline wrote (python-src/Ez.png, assets/ring_resonator.png,
    assets/Ez_tiled.png): every distinct pixel colour in them is reproduced bit for bit by this table and the blend of
    main.py:171-177, 240 of the 256 entries being exercised (oracle/make_golden_colormap.py ->
    tests/golden/seismic_pixels.npz -> tests/test_host_cpu.py); tests also compare with matplotlib when it is importable."""
    anchors = np.array([(0.0, 0.0, 0.3), (0.0, 0.0, 1.0), (1.0, 1.0, 1.0), (1.0, 0.0, 0.0), (0.5, 0.0, 0.0)])
    x = np.linspace(0, 1, len(anchors)) * (n - 1)
    xind = (n - 1) * np.linspace(0, 1, n)
    ind = np.searchsorted(x, xind)[1:-1]
    lut = np.empty((n, 3))
    for ch in range(3):
        y = anchors[:, ch]
        distance = (xind[1:-1] - x[ind - 1]) / (x[ind] - x[ind - 1])
        lut[:, ch] = np.concatenate([[y[0]], distance * (y[ind] - y[ind - 1]) + y[ind - 1], [y[-1]]])
    return np.clip(lut, 0, 1)


def snapshot_rgb(Ez, eps, vmax=20, vmin=-20, lut=None):
    """The uint8 (R, C, 3) array `capture_snapshot` hands to PIL (main.py:153-177), with the colormap
    call `cmap(X)` spelled out as matplotlib evaluates it for float input: index = int(X * 256), the
    value 256 mapped to 255."""
    lut = seismic_lut() if lut is None else lut
    normed = np.clip(Ez, vmin, vmax)
    eps_min = 8.85418e-12
    eps_max = np.max(eps)
    if eps_max == eps_min:
        eps_gray = np.full_like(eps, 255, dtype=np.uint8)
    else:
        eps_normed = (eps - eps_min) / (eps_max - eps_min)
        eps_gray = ((1 - eps_normed) * 127 + 128).astype(np.uint8)
    background = np.stack([eps_gray] * 3, axis=-1)
    X = (normed - vmin) / (vmax - vmin)
    xa = np.array(X, copy=True)
    xa *= 256
    xa[xa == 256] = 255
    idx = np.clip(xa.astype(int), 0, 255)
    rgba = np.concatenate([lut[idx], np.ones(idx.shape + (1,))], axis=-1)
    rgba[..., 3] = 0.7
    rgb_float = rgba[..., :3] * rgba[..., 3:] + (background / 255) * (1 - rgba[..., 3:])
    return (rgb_float * 255).astype(np.uint8)
