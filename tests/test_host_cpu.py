"""CPU-only tests: the C-ABI library loads and exports every symbol include/fdtd2d.h declares, host
logic of the Python surface, and hygiene rules (no oracle / CPU fallback in the product)."""
import ctypes
import os
import re
import sys

import numpy as np
import pytest

import fdtd2d_b200 as fd
from fdtd2d_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DT, DX, FC = 5e-14, 1e-4, 30e9


@pytest.fixture(scope="module")
def lib():
    fd.build()
    return _lib.lib()


def test_library_exports_every_declared_symbol(lib):
    hdr = open(os.path.join(ROOT, "include", "fdtd2d.h")).read()
    declared = set(re.findall(r"^(?:int|double|const char\*)\s+(fdtd2d_\w+)\s*\(", hdr, flags=re.M))
    assert len(declared) >= 28
    assert declared == set(_lib.EXPORTED_SYMBOLS), declared ^ set(_lib.EXPORTED_SYMBOLS)
    raw = ctypes.CDLL(fd.LIB_PATH)
    for name in declared:
        assert hasattr(raw, name), f"{name} declared in include/fdtd2d.h but not exported"
    assert lib.fdtd2d_abi_version() == 1


def test_no_cpu_fallback_without_gpu(lib):
    n = ctypes.c_int(-1)
    rc = lib.fdtd2d_device_count(ctypes.byref(n))
    if rc == 0 and n.value > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(fd.Fdtd2dError) as e:
        fd.Simulation(64, 64, np.float32, dt=DT, dx=DX)
    assert e.value.code == -2  # FDTD2D_ECUDA: fails loudly, nothing is computed on the CPU
    Ez, Hx, Hy = fd.grid_init(20, 20)
    eps, mu = fd.material_init(None, 20, 20)
    with pytest.raises(fd.Fdtd2dError):
        fd.update_Hx_Hy(Ez, Hx, Hy, mu, eps, DT, DX)
    with pytest.raises(fd.Fdtd2dError):
        fd.update_Ez(Ez, Hx, Hy, mu, eps, DT, DX)


def test_argument_validation_needs_no_gpu(lib):
    h = ctypes.c_void_p()
    assert lib.fdtd2d_create(ctypes.byref(h), 5, 64, 0, 0, 1) == -1  # rows < 6: the reference itself indexes out of range
    assert b">= 6" in lib.fdtd2d_last_error()
    assert lib.fdtd2d_create_slab(ctypes.byref(h), 10, 64, 0, 5, 4, 0, 0) == -1  # slabs need the staged form (>= 11)
    assert lib.fdtd2d_set_option(None, b"wavefront", 0) == -1 and lib.fdtd2d_peer_detach(None) == -1
    assert lib.fdtd2d_create(ctypes.byref(h), 64, 64, 7, 0, 1) == -1  # bad dtype
    assert lib.fdtd2d_create_slab(ctypes.byref(h), 64, 64, 10, 5, 4, 0, 0) == -1  # empty slab
    assert lib.fdtd2d_step(None, 1, 1) == -1 and lib.fdtd2d_sync(None) == -1


def test_product_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "fdtd-2d_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.lower(), f"{f} mentions the oracle"
    assert "oracle" not in open(os.path.join(ROOT, "fdtd2d_b200.py")).read()


def test_call_surface_matches_reference_names():
    # fdtd.py:1-9 imports these from `main`
    for name in ("grid_init", "material_init", "update_Hx_Hy", "update_Ez", "ricker", "sinusoidal"):
        assert callable(getattr(fd, name))
    Ez, Hx, Hy = fd.grid_init(7, 9)
    assert Ez.shape == (7, 9) and Hx.shape == (7, 8) and Hy.shape == (6, 9) and Ez.dtype == np.float64


def test_material_init_and_sources_vs_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "material_sources.npz"))
    png = os.path.join(golden_dir, "structure.png")
    for (R, C, bp) in [(64, 80, 10.0), (200, 200, 10.0), (37, 53, 4.0)]:
        eps, mu = fd.material_init(png, R, C, bp)
        assert np.array_equal(eps, g[f"eps_{R}x{C}_bp{bp:g}"]) and np.array_equal(mu, g[f"mu_{R}x{C}_bp{bp:g}"])
    eps, mu = fd.material_init(None, 23, 17)
    assert np.array_equal(eps, g["eps_none_23x17"]) and np.array_equal(mu, g["mu_none_23x17"])
    steps = [int(i) for i in g["steps"]]
    assert np.array_equal(np.array([fd.ricker_amplitude(i * DT, FC) for i in steps]), g["ricker_amp"])
    assert np.array_equal(np.array([fd.sinusoidal_amplitude(i * DT, FC) for i in steps]), g["sinus_amp"])
    assert np.array_equal(fd.ricker(6, 7, 2, 3, 667 * DT, FC), g["ricker_dense_6x7"])
    tab = fd.source_table("ricker", 669, DT, FC)
    assert np.array_equal(tab[[0, 1, 2, 10, 100, 500, 666, 667, 668]], g["ricker_amp"][:9])
    assert fd.courant_number(eps, mu, DT, DX) == 0.14989629517391773  # SURVEY A.6 / fdtd.py:25-27


def test_hash_uniform_host_definition(lib):
    def ref(seed, g, r, c):
        M = (1 << 64) - 1
        z = ((((g << 40) ^ (r << 20) ^ c) + seed * 0x9E3779B97F4A7C15 + 0x632BE59BD9B4E019)) & M
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M
        z ^= z >> 31
        return (z >> 40) / 16777216.0

    vals = []
    for seed, g, r, c in [(0, 0, 0, 0), (2026, 0, 5, 7), (77, 3, 65535, 65535), (2**63 + 5, 1023, 12345, 54321)]:
        v = lib.fdtd2d_hash_uniform(seed, g, r, c)
        assert v == ref(seed, g, r, c) and 0.0 <= v < 1.0
        vals.append(v)
    assert len(set(vals)) == len(vals)
    u = np.array([lib.fdtd2d_hash_uniform(9, 0, i, j) for i in range(64) for j in range(64)])
    assert abs(u.mean() - 0.5) < 0.02 and u.min() >= 0 and u.max() < 1


def test_slab_partition():
    for rows, world in [(65536, 8), (16384, 3), (1000, 7), (200, 2)]:
        spans = [fd.slab_rows(rows, world, r) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == rows
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1


def test_seismic_lut_construction():
    from oracle import numpy_oracle as npo

    lut = fd.seismic_lut()
    assert lut.shape == (256, 3) and np.array_equal(lut, npo.seismic_lut())
    assert np.allclose(lut[0], (0, 0, 0.3)) and np.allclose(lut[255], (0.5, 0, 0))
    assert np.allclose(lut[127], (0.99215686, 0.99215686, 1.0), atol=1e-6)  # just below the white anchor
    try:
        import matplotlib

        cm = matplotlib.colormaps["seismic"]
        assert np.array_equal(cm(np.arange(256))[:, :3], lut)
    except ImportError:
        pass  # matplotlib is not installed here: the table is pinned by the reference's own images (next test)
    g = fd.eps_background(np.array([[8.85418e-12, 2 * 8.85418e-12], [10 * 8.85418e-12, 8.85418e-12]]))
    assert g.dtype == np.uint8 and g[0, 0] == 255 and g[1, 0] == 128
    assert (fd.eps_background(np.full((3, 3), 8.85418e-12)) == 255).all()


def _blend_colours(lut, grays):
    """Every colour the reference's pipeline can write: trunc((lut * 0.7 + gray / 255 * (1 - 0.7)) * 255), main.py:171-177."""
    out = set()
    for g in grays:
        c = ((lut * 0.7 + (g / 255) * (1 - 0.7)) * 255).astype(np.uint8)
        out.update(map(tuple, c.tolist()))
    return out


def test_seismic_lut_pinned_by_the_references_own_images(golden_dir):
    """matplotlib is not installed, but the reference ships frames written by its colour pipeline (utils.plot_Ez, the
    twin of capture_snapshot): every distinct pixel colour of three such 1000 x 1000 images (tests/golden/
    seismic_pixels.npz, made by oracle/make_golden_colormap.py) must come out of the rebuilt table + blend formula, bit for
    bit.  The white background alone exercises 240 of the 256 entries; plausible wrong constructions of the table fail."""
    g = np.load(os.path.join(golden_dir, "seismic_pixels.npz"))
    lut = fd.seismic_lut()
    everything = _blend_colours(lut, range(128, 256))  # backgrounds are 128..255 (main.py:165), 255 where eps is uniform
    white = {c: i for i, c in enumerate(map(tuple, ((lut * 0.7 + 0.3) * 255).astype(np.uint8).tolist()))}
    used = set()
    for i in range(len(g["images"])):
        cols = list(map(tuple, g[f"colours_{i}"].tolist()))
        missing = [c for c in cols if c not in everything]
        assert not missing, f"{g['images'][i]}: {len(missing)} of {len(cols)} colours cannot come from this table, e.g. {missing[:5]}"
        used |= {white[c] for c in cols if c in white}
        # the commonest colour is a zero field on white: table index 128 (x = 0.5 -> 0.5 * 256 = 128)
        top = tuple(g[f"colours_{i}"][np.argmax(g[f"counts_{i}"])].tolist())
        assert top == tuple(((lut[128] * 0.7 + 0.3) * 255).astype(np.uint8).tolist())
    assert len(used) >= 240, len(used)
    # negative controls: tables sampled half a step off, or over 255 intervals with a wrong end point, do not explain the images
    anchors = np.asarray(fd.snapshot.SEISMIC_ANCHORS)
    for xs in ((np.arange(256) + 0.5) / 256, np.arange(256) / 256):
        wrong = np.stack([np.interp(xs, np.linspace(0, 1, 5), anchors[:, ch]) for ch in range(3)], axis=1)
        bad = _blend_colours(wrong, range(128, 256))
        cols = list(map(tuple, g["colours_2"].tolist()))
        assert sum(c not in bad for c in cols) > 50


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the reference's CPU path, here the bit-identical numpy port) needs no GPU and prints
    one JSON line with the keys the driver reads."""
    import json
    import subprocess

    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "small",
                          "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-500:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "Gcell-updates/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["steps"] == 2 and line["warmup"] == 1 and line["gpu_launches"] == 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] == 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0


def test_bench_fails_loudly_without_a_gpu():
    import subprocess

    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--workload", "small", "--steps", "1",
                          "--warmup", "1"], capture_output=True, text=True, timeout=300)
    assert out.returncode != 0 and "no CPU fallback" in (out.stderr + out.stdout)


def _plan(lib, rows, ring, warps=1184, cap=640, k=8):
    rows = np.ascontiguousarray(rows, np.int32)
    ring = np.ascontiguousarray(ring, np.uint8)
    parts = np.zeros(len(rows), np.int32)
    length = ctypes.c_int32()
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    _lib.check(lib.fdtd2d_plan_wave_runs(len(rows), p(rows), p(ring), warps, cap, k, p(parts), ctypes.cast(ctypes.byref(length), ctypes.c_void_p)))
    return parts, length.value


@pytest.mark.parametrize("n", [4096, 6000, 8192, 16384])
def test_wave_run_planning_balances_the_warps(lib, n):
    """The host logic that cuts the wavefront kernel's work into runs (fdtd2d_plan_wave_runs, no GPU needed): an n x n
    grid's strips fall into at most m x 1184 runs of (nearly) equal cost -- m as small as the ~640-row cap allows -- with
    the two ring strips cut to half the length; nothing shorter than 4k rows, no stretch lost."""
    k, CH, CW, W = 8, 48, 112, 148 * 8
    strips = -(-n // CW) - 2                       # plain tile columns
    rows_plain = (n // CH - 2) * CH                # rows between the top and the bottom ring tiles
    rows = [rows_plain] * strips + [rows_plain] * 2
    ring = [0] * strips + [1] * 2
    parts, L = _plan(lib, rows, ring, W, 640, k)
    weighted = rows_plain * (strips + 4)           # ring rows count double
    m = max(1, -(-weighted // (W * 640)))
    assert parts.sum() <= m * W
    assert parts.sum() > 0.9 * m * W or m == 1     # and the budget is used: the last round is not mostly idle warps
    assert L >= 4 * k
    plain_len, ring_len = rows_plain / parts[0], rows_plain / parts[-1]
    assert plain_len <= L and plain_len > 0.8 * L
    assert 0.4 * L <= ring_len <= max(4 * k, L / 2)
    # a shorter run length would not fit the budget
    if L > 4 * k:
        shorter = sum(-(-r // ((L - 1) // 2 if g else (L - 1))) for r, g in zip(rows, ring))
        assert shorter > m * W


def test_wave_run_planning_edge_cases(lib):
    parts, L = _plan(lib, [], [])
    assert len(parts) == 0 and L >= 1
    parts, L = _plan(lib, [5], [0])                # a stretch shorter than 4k rows is one run
    assert list(parts) == [1]
    parts, L = _plan(lib, [48, 4800, 96], [0, 0, 1], warps=16, cap=640, k=8)
    assert all(parts >= 1) and parts[1] >= parts[0]
    with pytest.raises(fd.Fdtd2dError):
        _plan(lib, [0], [0])
