"""Device-resident FDTD simulation: the performance surface of the package.

`Simulation` owns a libfdtd2d handle (device buffers + stream) and mirrors the reference driver
python-src/fdtd.py:14-38: build grid and materials, check Courant, then loop
H-update -> Ez-update(+Mur+corners) -> source add -> readout.  The loop itself runs on the GPU,
`k` leapfrog steps per HBM round trip; only setup and readout touch the host.
"""
from __future__ import annotations

import ctypes
from typing import Iterable, Sequence

import numpy as np

from . import _lib
from ._lib import check, lib

_DT = {np.dtype(np.float32): _lib.F32, np.dtype(np.float64): _lib.F64}


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def ricker_amplitude(t, fc):
    """Ricker wavelet value at time t, float64 -- the scalar expression of main.py:183-184."""
    tau = np.pi * fc * (t - 1 / fc)
    return (1 - 2 * tau**2) * np.exp(-(tau**2))


def sinusoidal_amplitude(t, fc):
    """Ramped sine value at time t, float64 -- the scalar expression of main.py:193-194."""
    envelope = 1 - np.exp(-((t - 3000 / fc) ** 2) / (2 * (2 / fc) ** 2))
    return envelope * np.sin(2 * np.pi * fc * t)


_WAVEFORMS = {"ricker": ricker_amplitude, "sinusoidal": sinusoidal_amplitude}


def source_table(kind: str, nsteps: int, dt: float, fc: float) -> np.ndarray:
    """amp[i] = waveform(i*dt, fc) for i in 0..nsteps-1, evaluated per step with python scalars exactly
    as the driver does (fdtd.py:34 passes ``i * dt``), so the table is bit-identical to the reference's
    per-step values."""
    fn = _WAVEFORMS[kind]
    return np.array([fn(i * dt, fc) for i in range(nsteps)], dtype=np.float64)


def courant_number(eps, mu, dt, dx):
    """c_max*dt/dx with c_max = 1/sqrt(eps.min()*mu.min()) (fdtd.py:25-26)."""
    c = 1 / np.sqrt(eps.min() * mu.min())
    return (c * dt) / dx


class Simulation:
    """`batch` independent rows x cols TM-mode Yee grids resident on one GPU.

    dtype follows the reference convention: float64 is what `grid_init` returns (main.py:79-85);
    float32 is the reference run with all five arrays cast to float32.
    """

    def __init__(self, rows: int, cols: int, dtype=np.float32, *, dt: float, dx: float, device: int = 0,
                 batch: int = 1, slab: tuple[int, int, int, int] | None = None):
        self.dtype = np.dtype(dtype)
        if self.dtype not in _DT:
            raise TypeError("dtype must be float32 or float64")
        self.dt, self.dx = float(dt), float(dx)
        self.batch = int(batch)
        self._h = ctypes.c_void_p()
        if slab is None:
            check(lib().fdtd2d_create(ctypes.byref(self._h), rows, cols, _DT[self.dtype], device, batch))
        else:
            global_rows, row_begin, row_end, halo = slab
            check(lib().fdtd2d_create_slab(ctypes.byref(self._h), global_rows, cols, row_begin, row_end, halo,
                                           _DT[self.dtype], device))
        lr, c, r0, gr, b, dtc = (ctypes.c_int() for _ in range(6))
        pitch = ctypes.c_size_t()
        check(lib().fdtd2d_geometry(self._h, lr, c, r0, gr, b, dtc, pitch))
        self.local_rows, self.cols, self.row0, self.global_rows = lr.value, c.value, r0.value, gr.value
        self.pitch = pitch.value
        self.rows = self.global_rows
        self.device = device
        self._hy_rows = self.local_rows - 1 if self.row0 + self.local_rows == self.global_rows else self.local_rows
        self._n_probes = 0
        self._keep = []  # host arrays that must outlive async copies

    # ---- lifetime ---------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            lib().fdtd2d_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # ---- helpers ----------------------------------------------------------------------------
    def _shape(self, rows, cols):
        return (rows, cols) if self.batch == 1 else (self.batch, rows, cols)

    def _as(self, a, rows, cols, name):
        a = np.ascontiguousarray(a, dtype=self.dtype)
        if a.shape != self._shape(rows, cols) and a.shape != (self.batch, rows, cols):
            raise ValueError(f"{name} has shape {a.shape}, expected {self._shape(rows, cols)}")
        return a

    def set_stream(self, cuda_stream: int | None):
        """Order all work of this simulation on an external CUDA stream handle (e.g.
        ``torch.cuda.current_stream().cuda_stream``; 0 is the legacy default stream).
        None goes back to the simulation's own stream."""
        if cuda_stream is None:
            check(lib().fdtd2d_reset_stream(self._h))
        else:
            check(lib().fdtd2d_set_stream(self._h, ctypes.c_void_p(cuda_stream)))

    def synchronize(self):
        check(lib().fdtd2d_sync(self._h))

    @property
    def cuda_stream(self) -> int:
        """The CUDA stream handle the simulation's work is ordered on."""
        p = ctypes.c_void_p()
        check(lib().fdtd2d_get_stream(self._h, ctypes.byref(p)))
        return p.value or 0

    # ---- tuning options (per handle; defaults come from FDTD2D_<KEY> when the handle is created) ----
    def set_option(self, key: str, value: int):
        check(lib().fdtd2d_set_option(self._h, key.encode(), int(value)))

    def get_option(self, key: str) -> int:
        v = ctypes.c_int()
        check(lib().fdtd2d_get_option(self._h, key.encode(), ctypes.byref(v)))
        return v.value

    def plan_info(self, k: int = 8) -> dict:
        """What a k-step pass of this handle consists of (see fdtd2d_plan_info)."""
        a = (ctypes.c_int32 * _lib.PLAN_INFO_WORDS)()
        check(lib().fdtd2d_plan_info(self._h, k, a, _lib.PLAN_INFO_WORDS))
        names = ("tiles_y", "tiles_x", "core_rows", "core_cols", "edge_tiles", "edge_band_tiles", "tma_tiles", "wave_runs",
                 "wave_band_runs", "ring_strips", "band_tasks_top", "band_tasks_bottom", "reserve_sms")
        return dict(zip(names, (int(v) for v in a)))

    # ---- state ------------------------------------------------------------------------------
    def set_state(self, Ez, Hx, Hy):
        """Upload Ez (R,C), Hx (R,C-1), Hy (R-1,C) -- the shapes grid_init returns (main.py:79-85)."""
        Ez = self._as(Ez, self.local_rows, self.cols, "Ez")
        Hx = self._as(Hx, self.local_rows, self.cols - 1, "Hx")
        Hy = self._as(Hy, self._hy_rows, self.cols, "Hy")
        check(lib().fdtd2d_upload_state(self._h, _p(Ez), _p(Hx), _p(Hy)))
        self.synchronize()

    def state(self, out=None):
        """Download (Ez, Hx, Hy) in the reference's shapes."""
        if out is None:
            out = (np.empty(self._shape(self.local_rows, self.cols), self.dtype),
                   np.empty(self._shape(self.local_rows, self.cols - 1), self.dtype),
                   np.empty(self._shape(self._hy_rows, self.cols), self.dtype))
        Ez, Hx, Hy = out
        check(lib().fdtd2d_download_state(self._h, _p(Ez), _p(Hx), _p(Hy)))
        return Ez, Hx, Hy

    def read_Ez(self, out=None):
        if out is None:
            out = np.empty(self._shape(self.local_rows, self.cols), self.dtype)
        check(lib().fdtd2d_download_state(self._h, _p(out), None, None))
        return out

    def zero_state(self):
        check(lib().fdtd2d_zero_state(self._h))

    # ---- non-blocking copies (PINNED host arrays; see fdtd2d_upload_state_async) -------------------
    def set_state_async(self, Ez, Hx, Hy):
        """Like set_state, but returns at once: the arrays must be pinned, of the run dtype, and stay alive until
        copy_wait() / synchronize()."""
        for a, (r, c), name in ((Ez, (self.local_rows, self.cols), "Ez"), (Hx, (self.local_rows, self.cols - 1), "Hx"),
                                (Hy, (self._hy_rows, self.cols), "Hy")):
            self._check_raw(a, r, c, name)
        check(lib().fdtd2d_upload_state_async(self._h, _p(Ez), _p(Hx), _p(Hy)))

    def read_Ez_async(self, out):
        self._check_raw(out, self.local_rows, self.cols, "out")
        check(lib().fdtd2d_download_state_async(self._h, _p(out), None, None))
        return out

    def set_materials_async(self, eps, mu):
        self._check_raw(eps, self.local_rows, self.cols, "eps")
        self._check_raw(mu, self.local_rows, self.cols, "mu")
        check(lib().fdtd2d_set_materials_async(self._h, _p(eps), _p(mu), self.dt, self.dx))

    def copy_wait(self):
        check(lib().fdtd2d_copy_wait(self._h))

    def _check_raw(self, a, rows, cols, name):
        if not (isinstance(a, np.ndarray) and a.dtype == self.dtype and a.flags.c_contiguous and
                a.shape in (self._shape(rows, cols), (self.batch, rows, cols))):
            raise ValueError(f"{name} must be a C-contiguous {self.dtype} array of shape {self._shape(rows, cols)} "
                             "(the non-blocking copies take the caller's buffer as it is)")

    # ---- materials --------------------------------------------------------------------------
    def set_materials(self, eps, mu):
        """eps, mu maps as `material_init` returns them (main.py:88-123), cast to the run dtype.  The
        device forms dt/(eps*dx), dt/(mu*dx) and the Mur coefficient with the reference's op order."""
        eps = self._as(eps, self.local_rows, self.cols, "eps")
        mu = self._as(mu, self.local_rows, self.cols, "mu")
        check(lib().fdtd2d_set_materials(self._h, _p(eps), _p(mu), self.dt, self.dx))

    def set_coefficients(self, ce, ch, mur_coef):
        """Host-precomputed maps ce = dt/(eps*dx), ch = dt/(mu*dx) and Mur coefficient(s)."""
        ce = self._as(ce, self.local_rows, self.cols, "ce")
        ch = self._as(ch, self.local_rows, self.cols, "ch")
        mc = np.ascontiguousarray(np.broadcast_to(np.asarray(mur_coef, dtype=self.dtype), (self.batch,)))
        check(lib().fdtd2d_set_coeffs(self._h, _p(ce), _p(ch), _p(mc)))

    def set_mur_coef(self, mur_coef):
        mc = np.ascontiguousarray(np.broadcast_to(np.asarray(mur_coef, dtype=self.dtype), (self.batch,)))
        check(lib().fdtd2d_set_mur_coef(self._h, _p(mc)))

    def set_materials_gray(self, gray, black_point: float = 10.0):
        """Structure image -> materials on the device (main.py:109-123).  gray: uint8 (rows, cols) "L" image
        already resized (what `Image.open(p).convert("L").resize((cols, rows), LANCZOS)` gives); the device forms
        eps = (1 + (black_point-1)(1-gray/255)) eps0 in float64, casts to the run dtype and sets mu = mu0."""
        gray = np.ascontiguousarray(gray, dtype=np.uint8)
        if gray.shape != self._shape(self.local_rows, self.cols):
            raise ValueError(f"gray has shape {gray.shape}, expected {self._shape(self.local_rows, self.cols)}")
        check(lib().fdtd2d_set_materials_gray(self._h, _p(gray), float(black_point), self.dt, self.dx))

    def set_materials_image(self, path, black_point: float = 10.0):
        """`material_init(path, rows, cols, black_point)` + set_materials with the mapping done on the device."""
        from PIL import Image

        img = Image.open(path).convert("L").resize((self.cols, self.local_rows), Image.LANCZOS)
        self.set_materials_gray(np.array(img, dtype=np.uint8), black_point)

    def set_materials_random(self, seed: int, span: float = 9.0):
        """Synthetic medium generated on the device: eps = eps0*(1 + span*u), mu = mu0."""
        check(lib().fdtd2d_set_materials_random(self._h, seed, span, self.dt, self.dx))

    def coefficients(self):
        ce = np.empty(self._shape(self.local_rows, self.cols), self.dtype)
        ch = np.empty_like(ce)
        mur = np.empty((self.batch,), self.dtype)
        check(lib().fdtd2d_download_coeffs(self._h, _p(ce), _p(ch), _p(mur)))
        return ce, ch, mur

    # ---- sources and probes -----------------------------------------------------------------
    def set_sources(self, cells: Sequence[tuple], tables: np.ndarray):
        """cells: (grid, row, col, wave) tuples (global rows); tables: float64 [n_waves][n_steps]."""
        tables = np.ascontiguousarray(np.atleast_2d(tables), dtype=np.float64)
        if len(cells) == 0:
            check(lib().fdtd2d_set_sources(self._h, 0, None, None, None, None, 0, 0, None))
            return
        arr = np.ascontiguousarray(np.asarray(cells, dtype=np.int32).reshape(-1, 4))
        g, r, c, w = (np.ascontiguousarray(arr[:, i]) for i in range(4))
        check(lib().fdtd2d_set_sources(self._h, len(arr), _p(g), _p(r), _p(c), _p(w), tables.shape[0],
                                       tables.shape[1], _p(tables)))

    def set_point_source(self, row: int, col: int, nsteps: int, fc: float = 30e9, kind: str = "ricker", grid: int = 0):
        """The driver's source: ``Ez += ricker(rows, cols, row, col, i*dt, fc)`` (fdtd.py:34)."""
        self.set_sources([(grid, row, col, 0)], source_table(kind, nsteps, self.dt, fc)[None, :])

    def set_probes(self, cells: Iterable[tuple], capacity_steps: int):
        """cells: (row, col) or (grid, row, col); Ez there is recorded after every step."""
        cells = [(0, *c) if len(c) == 2 else tuple(c) for c in cells]
        self._n_probes = len(cells)
        if not cells:
            check(lib().fdtd2d_set_probes(self._h, 0, None, None, None, 0))
            return
        arr = np.ascontiguousarray(np.asarray(cells, dtype=np.int32).reshape(-1, 3))
        g, r, c = (np.ascontiguousarray(arr[:, i]) for i in range(3))
        check(lib().fdtd2d_set_probes(self._h, len(arr), _p(g), _p(r), _p(c), capacity_steps))

    def read_probes(self, first_step: int = 0, n_steps: int | None = None):
        if n_steps is None:
            n_steps = self.step_index - first_step
        out = np.empty((n_steps, self._n_probes), self.dtype)
        check(lib().fdtd2d_read_probes(self._h, _p(out), first_step, n_steps))
        return out

    # ---- time stepping ----------------------------------------------------------------------
    def step(self, n_steps: int = 1, k: int = 0, strict: bool = True):
        """Advance n_steps leapfrog steps (asynchronous), k steps per HBM round trip (0 = default).

        The reference injects its source for as long as its loop runs (fdtd.py:34) and a readout has no end; here the
        waveform tables and probe traces have the length they were given, and the library adds nothing / records nothing
        beyond it.  strict=True (default) refuses to step past either instead of silently running on."""
        if strict and n_steps > 0:
            nc, ns, cap = ctypes.c_int(), ctypes.c_int(), ctypes.c_int64()
            check(lib().fdtd2d_source_steps(self._h, ctypes.byref(nc), ctypes.byref(ns), ctypes.byref(cap)))
            end = self.step_index + n_steps
            if nc.value and end > ns.value:
                raise ValueError(f"stepping to step {end} but the source tables hold {ns.value} steps: pass a longer table to "
                                 "set_sources / set_point_source (or strict=False to run on without a source)")
            if cap.value and end > cap.value:
                raise ValueError(f"stepping to step {end} but the probe traces hold {cap.value} steps: raise capacity_steps "
                                 "in set_probes (or strict=False to stop recording)")
        check(lib().fdtd2d_step(self._h, n_steps, k))

    run = step

    def step_phases(self, phases: int):
        check(lib().fdtd2d_step_phases(self._h, phases))

    @property
    def step_index(self) -> int:
        v = ctypes.c_int64()
        check(lib().fdtd2d_get_step_index(self._h, ctypes.byref(v)))
        return v.value

    @step_index.setter
    def step_index(self, v: int):
        check(lib().fdtd2d_set_step_index(self._h, v))

    def set_kernel_variant(self, variant: int):
        check(lib().fdtd2d_set_kernel_variant(self._h, variant))

    @property
    def pass_count(self) -> int:
        """Stepping passes so far (HBM round trips of the fields)."""
        v = ctypes.c_int64()
        check(lib().fdtd2d_pass_count(self._h, ctypes.byref(v)))
        return v.value

    @property
    def launch_count(self) -> int:
        v = ctypes.c_int64()
        check(lib().fdtd2d_launch_count(self._h, ctypes.byref(v)))
        return v.value

    # ---- field readout as an image (capture_snapshot, main.py:153-179) ----------------------
    def set_snapshot_background(self, eps):
        """eps: the host permittivity map (as material_init returns it) behind the frames."""
        from . import snapshot

        snapshot.set_background(self, eps)

    def render_snapshot(self, vmax=20, vmin=-20, grid: int = 0, out=None):
        """uint8 (rows, cols, 3) RGB frame of the current Ez, colour-mapped and blended on the device."""
        from . import snapshot

        return snapshot.render(self, vmax, vmin, grid, out)

    # ---- multi-GPU plumbing -----------------------------------------------------------------
    def halo_block(self, field: int, side: int, next_state: bool = False):
        """(send_ptr, recv_ptr, nbytes) of the halo block of the current state, or of the state an open
        pass is writing (see fdtd2d_halo_block / fdtd2d_halo_block_next)."""
        sp, rp, nb = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_size_t()
        fn = lib().fdtd2d_halo_block_next if next_state else lib().fdtd2d_halo_block
        check(fn(self._h, field, side, ctypes.byref(sp), ctypes.byref(rp), ctypes.byref(nb)))
        return sp.value, rp.value, nb.value

    def pass_begin(self, k: int):
        """Launch the halo-producing tiles of a k-step pass (see fdtd2d_pass_begin)."""
        check(lib().fdtd2d_pass_begin(self._h, k))

    def pass_end(self):
        """Launch the rest of the open pass and make its result the current state."""
        check(lib().fdtd2d_pass_end(self._h))

    def peer_export(self) -> bytes:
        """This slab's blob for its neighbours' peer_attach (geometry, device pointers, CUDA IPC handles)."""
        buf = ctypes.create_string_buffer(_lib.PEER_BLOB_BYTES)
        check(lib().fdtd2d_peer_export(self._h, buf))
        return buf.raw

    def peer_attach(self, side: int, blob: bytes):
        """Link the neighbour slab on `side` (0 top, 1 bottom): the halo exchange then happens inside the kernels."""
        check(lib().fdtd2d_peer_attach(self._h, side, ctypes.create_string_buffer(blob, _lib.PEER_BLOB_BYTES)))

    def peer_detach(self):
        check(lib().fdtd2d_peer_detach(self._h))

    def peer_status(self) -> dict:
        a = (ctypes.c_uint32 * 8)()
        check(lib().fdtd2d_peer_status(self._h, a))
        return {"attached_top": bool(a[0] & 1), "attached_bottom": bool(a[0] & 2), "passes": int(a[1]), "in_top": int(a[2]),
                "in_bottom": int(a[3]), "error": int(a[4])}

    def device_field(self, field: int) -> int:
        p = ctypes.c_void_p()
        check(lib().fdtd2d_device_field(self._h, field, ctypes.byref(p)))
        return p.value
