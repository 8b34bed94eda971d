"""y-slab decomposition on the GPU: 2, 3 and 4 slabs must reproduce the single-domain result bit for bit
(no reduction is involved).  Runs the slabs of one grid from one process (InProcessSlabs) so that it
also works on a one-GPU box; with several GPUs visible the slabs are spread over them.  Every test runs with both
ways of moving the halo rows: "p2p" -- peer links, the band tasks of the stepping kernels store into the neighbour's
ghost rows and raise its flag, exactly the production path of the one-process-per-GPU layout -- and "copy" -- device
copies of the halo blocks between passes."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
DT, DX, FC = 5e-14, 1e-4, 30e9


def _devices():
    import torch

    return tuple(range(torch.cuda.device_count()))


@pytest.fixture(autouse=True, params=["default", "wavefront"])
def plain_tile_kernel(request, monkeypatch):
    """Every slab test runs with the default kernel choice and with the wavefront strips forced onto these small
    grids (in a slab pass the band tiles stay on the tile kernel, the rest of the plain tiles become strip runs)."""
    if request.param == "wavefront":
        monkeypatch.setenv("FDTD2D_WAVE_MIN_TILES", "0")
        monkeypatch.setenv("FDTD2D_RING_MIN_TILES", "0")
    return request.param


@pytest.fixture(params=["p2p", "copy"])
def exchange(request):
    return request.param


@pytest.mark.parametrize("dtype", ["float32", "float64"])
@pytest.mark.parametrize("world,k", [(2, 4), (3, 8), (4, 5)])
def test_slabs_match_single_domain_and_oracle(world, k, dtype, exchange):
    import fdtd2d_b200 as fd
    from oracle import c_oracle, numpy_oracle as npo

    R, C, n = 700, 900, 43
    rng = np.random.default_rng(world * 10 + k)
    eps = (8.85418e-12 * (1 + 9 * rng.random((R, C)))).astype(dtype)
    mu = (4 * np.pi * 1e-7 * (1 + 0.3 * rng.random((R, C)))).astype(dtype)
    Ez = (1e-3 * rng.standard_normal((R, C))).astype(dtype)
    Hx = (1e-6 * rng.standard_normal((R, C - 1))).astype(dtype)
    Hy = (1e-6 * rng.standard_normal((R - 1, C))).astype(dtype)
    ce, ch, coef = c_oracle.coefficients(eps, mu, DT, DX, np.dtype(dtype))
    amp = npo.source_table("ricker", n, DT, FC)
    # sources and probes next to / on slab boundaries on purpose
    b1 = fd.slab_rows(R, world, 1)[0]
    cells = [(R // 2, C // 2), (b1, 100), (b1 - 1, 300), (b1 + 3, 500), (7, 7)]
    probes = [(b1, 101), (b1 - 2, 301), (R - 1, C - 1), (0, 0), (R // 2, C // 2)]
    oEz, oHx, oHy = Ez.copy(), Hx.copy(), Hy.copy()
    otrace = c_oracle.run(oEz, oHx, oHy, ce, ch, coef, n, amp, cells, probes, omp=True)

    grp = fd.InProcessSlabs(R, C, dtype, dt=DT, dx=DX, world=world, devices=_devices(), halo=8, exchange=exchange)
    try:
        for s in grp.slabs:
            lo, hi = s.row0, s.row0 + s.local_rows
            s.set_materials(eps[lo:hi], mu[lo:hi], coef)
            s.set_state(Ez[lo:hi], Hx[lo:hi], Hy[lo:min(hi, R - 1)])
            s.set_sources([(0, r, c, 0) for r, c in cells], amp[None, :])
            s.set_probes(probes, n)
        grp.step(n, k)
        gEz, gHx, gHy = grp.gather()
        traces = [s.read_probes(0, n) for s in grp.slabs]
        if exchange == "p2p":
            for s in grp.slabs:
                st = s.sim.peer_status()
                assert st["error"] == 0 and st["passes"] == -(-n // k), st
    finally:
        grp.close()
    assert np.array_equal(gEz, oEz) and np.array_equal(gHx, oHx) and np.array_equal(gHy, oHy)
    # every probe is recorded by exactly its owner; the others leave zeros
    total = np.zeros_like(otrace)
    for p, (r, c) in enumerate(probes):
        owners = [i for i, s in enumerate(grp.slabs) if s.row_begin <= r < s.row_end]
        assert len(owners) == 1
        total[:, p] = traces[owners[0]][:, p]
    assert np.array_equal(total, otrace)


def test_slabs_zero_state_restarts_a_job(exchange):
    """A second job on the same slab handles (zero_state = grid_init on the device, the source restarted) must equal a
    fresh single-domain run: with peer links only the current field set is cleared, the ghost rows of the other one
    belong to the neighbours (bench.py's e2e restarts its slab jobs this way)."""
    import fdtd2d_b200 as fd

    R, C, n = 1500, 1100, 41
    with fd.Simulation(R, C, np.float32, dt=DT, dx=DX) as sim:
        sim.set_materials_random(5, 9.0)
        sim.set_sources([(0, R // 3, C // 2, 0), (0, 2 * R // 3 + 1, 40, 0)], (1e-2 * np.ones((1, n))))
        sim.step(n, 8)
        ref = sim.state()
    grp = fd.InProcessSlabs(R, C, np.float32, dt=DT, dx=DX, world=3, devices=_devices(), halo=8, exchange=exchange)
    try:
        for s in grp.slabs:
            s.set_materials_random(5, 9.0)
            s.set_sources([(0, R // 3, C // 2, 0), (0, 2 * R // 3 + 1, 40, 0)], (1e-2 * np.ones((1, n))))
        for job in range(3):  # 41 steps at k = 8 = 6 passes: the field sets swap roles from job to job
            for s in grp.slabs:
                s.zero_state()
            grp.step(n, 8)
            got = grp.gather()
            for a, b in zip(got, ref):
                assert np.array_equal(a, b), f"job {job}"
    finally:
        grp.close()


def test_large_slabs_vs_single_domain_fast_path(exchange, plain_tile_kernel):
    """4 slabs of a 4096 x 3000 fp32 grid (wavefront runs incl. the band runs, or TMA tiles) == single domain, k = 8,
    with a remainder pass (44 = 5 x 8 + 4)."""
    import fdtd2d_b200 as fd

    R, C, n = 4096, 3000, 44
    with fd.Simulation(R, C, np.float32, dt=DT, dx=DX) as sim:
        sim.set_materials_random(9, 9.0)
        sim.set_point_source(R // 2, C // 2, 700, FC)
        sim.step_index = 640
        sim.step(n, 8)
        ref = sim.state()
    grp = fd.InProcessSlabs(R, C, np.float32, dt=DT, dx=DX, world=4, devices=_devices(), halo=8, exchange=exchange)
    try:
        for s in grp.slabs:
            s.set_materials_random(9, 9.0)
            s.set_point_source(R // 2, C // 2, 700, FC)
            s.step_index = 640
        grp.step(n, 8)
        got = grp.gather()
        mid = grp.slabs[1].sim.plan_info(8)
        if plain_tile_kernel == "wavefront":  # a middle slab: every band row on a band run (the source tile aside)
            assert mid["wave_band_runs"] > 0 and mid["band_tasks_top"] > 0 and mid["band_tasks_bottom"] > 0, mid
    finally:
        grp.close()
    for a, b in zip(got, ref):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("mode", ["p2p", "nccl"])
def test_multiprocess_slabs(mode):
    """One rank per GPU (the production layout): peer links over CUDA IPC, or NCCL send/recv.  Needs >= 2 GPUs; skipped
    on a 1-GPU box (bench.py prints the same check as `slab_parity` on whatever box it runs on)."""
    import os
    import subprocess
    import sys

    import torch

    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    n = 2 if n < 4 else 4
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr",
           "127.0.0.1", "--master-port", "29611" if mode == "p2p" else "29612", os.path.join(root, "tests", "mp_slab_check.py"), mode]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "OK bit-exact" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
