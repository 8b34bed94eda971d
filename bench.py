#!/usr/bin/env python
"""bench.py -- Gcell-updates/s of the 2D FDTD leapfrog path (BASELINE.json metric) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg3] [--impl ours|reference]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...      (one rank per GPU)

A bench "step" is ONE call of the hot path over the workload grid: `inner` leapfrog steps
(H -> Ez+Mur+corners -> source -> probes; fdtd.py:31-34) advanced `k` steps per HBM round trip.
One cell-update = Hx, Hy and Ez of one cell advanced one leapfrog step.

  value   device-resident throughput: inputs already in HBM, CUDA-event timed, max over ranks.
  e2e     the same call through the public API with HOST (pinned) inputs: every step uploads
          eps, mu, Ez, Hx, Hy, forms the coefficient maps on the device, runs `inner` leapfrog steps and
          reads Ez and the probe traces back.
  roofline  HBM roofline of the tile kernel at the algorithmic 32 B per fp32 cell-update
          (SURVEY 8d); temporal blocking may legitimately exceed 1.0 -- `traffic` is the DRAM bytes ncu
          saw per launch (profiles/), which is what actually bounds the kernel.
  cpu_baseline  the oracle's numpy restatement of the reference (1 core: numpy elementwise) on a bounded
          sample of the same workload, timed on this box's host.

`--impl reference` times the reference's own CPU implementation of the path.  The reference is plain
numpy and `/root/reference` does not travel to the GPU box, so this is the oracle's numpy port
(oracle/numpy_oracle.py, bit-identical to the reference -- tests/test_oracle_golden.py).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DT, DX, FC = 5e-14, 1e-4, 30e9  # fdtd.py:16-17,34
BYTES_PER_UPDATE_F32 = 32  # read Ez,Hx,Hy,ce,ch + write Ez,Hx,Hy

WORKLOADS = {
    # name: rows-per-GPU, cols, inner leapfrog steps per bench step, description
    "cfg2": dict(rows=4096, cols=4096, inner=10000, desc="4096x4096 fp32, random permittivity, 10k steps (BASELINE configs[1])"),
    "cfg3": dict(rows=16384, cols=16384, inner=1000,
                 desc="16384x16384 fp32, random permittivity, temporal-blocked kernel (BASELINE configs[2]; the "
                      "grid the >=70%-of-roofline target is quoted on)"),
    "cfg4": dict(rows=65536, cols=65536, inner=16, desc="65536x65536 fp32 y-slab sharded (BASELINE configs[3]), strong scaling"),
    "cfg5": dict(rows=256, cols=256, inner=400, batch=1024,
                 desc="batched 1024 x (256x256) independent fp32 grids per GPU (BASELINE configs[4], dataset generation)"),
    "small": dict(rows=1024, cols=1024, inner=64, desc="1024x1024 fp32 (debug)"),
}


class BatchedRunner:
    """`batch` independent grids on one GPU behind the interface the bench uses for slabs (cfg5).  With
    several ranks every rank runs its own `batch` grids: no exchange, weak scaling by construction."""

    def __init__(self, fd, rows, cols, batch, device):
        self.sim = fd.Simulation(rows, cols, np.float32, dt=DT, dx=DX, device=device, batch=batch)
        self.fd, self.batch, self.rows, self.cols = fd, batch, rows, cols
        self.row0, self.local_rows, self.hy_rows = 0, rows, rows - 1
        self.tile_launch_count = 0
        self._k = fd.DEFAULT_K

    def set_stream(self, s):
        self.sim.set_stream(s)

    def set_kernel_variant(self, v):
        self.sim.set_kernel_variant(v)

    def set_materials_random(self, seed, span=9.0):
        self.sim.set_materials_random(seed, span)

    def set_materials(self, eps, mu, mur=None):
        self.sim.set_materials(eps, mu)

    def set_state(self, Ez, Hx, Hy):
        self.sim.set_state(Ez, Hx, Hy)

    def set_point_source(self, row, col, nsteps, fc=FC):
        # one point source per grid, frequencies spread over 18..30 GHz like the dataset generator's omega range
        fcs = np.linspace(18e9, 30e9, self.batch)
        tables = np.stack([self.fd.source_table("ricker", nsteps, DT, f) for f in fcs[:16]])
        cells = [(b, self.rows // 2 + (b % 7) - 3, self.cols // 2 + (b % 5) - 2, b % 16) for b in range(self.batch)]
        self.sim.set_sources(cells, tables)

    def set_probes(self, cells, cap):
        self.sim.set_probes([(b, self.rows // 2, self.cols // 2 + 20) for b in range(0, self.batch, max(1, self.batch // 8))], cap)

    def step(self, n, k=0):
        before = self.sim.pass_count
        self.sim.step(n, k)
        # passes = HBM round trips of the fields: one per step call for the cluster-resident kernel
        self.tile_launch_count += self.sim.pass_count - before

    def read_Ez(self, out=None):
        return self.sim.read_Ez(out)

    def read_probes(self, a=0, n=None):
        return self.sim.read_probes(a, n)

    @property
    def launch_count(self):
        return self.sim.launch_count

    @property
    def step_index(self):
        return self.sim.step_index

    @step_index.setter
    def step_index(self, v):
        self.sim.step_index = v

    def close(self):
        self.sim.close()


def make_sim(fd, wl, grows, cols, rank, world, local_rank, k, exchange="p2p"):
    if wl.get("batch"):
        return BatchedRunner(fd, wl["rows"], cols, wl["batch"], local_rank)
    return fd.SlabSimulation(grows, cols, np.float32, dt=DT, dx=DX, rank=rank, world=world, device=local_rank, halo=k or 8,
                             exchange=exchange)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc, self.thread = index, [], None, None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            return
        self.thread = threading.Thread(target=self._pump, daemon=True)
        self.thread.start()

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def bind_to_gpu_cpus(index: int):
    """Pin this process to the host cores next to GPU `index` (`nvidia-smi topo -m`, column "CPU Affinity") BEFORE any pinned
    memory is allocated: with one rank per GPU and first-touch page placement the e2e copies then stay on the GPU's own
    socket instead of crossing the inter-socket link.  Host tuning only; returns the affinity string or None."""
    try:
        out = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=20).stdout
        lines = [l for l in out.splitlines() if l.strip()]
        head = next(l for l in lines if "CPU Affinity" in l)
        cols = [c.strip() for c in head.replace("\x1b[4m", "").replace("\x1b[0m", "").split("\t")]
        row = next(l for l in lines if l.replace("\x1b[4m", "").startswith(f"GPU{index}\t") or l.startswith(f"GPU{index} "))
        cells = [c.strip() for c in row.split("\t")]
        aff = cells[cols.index("CPU Affinity")]
        cpus = set()
        for part in aff.split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return aff
    except Exception:
        pass
    return None


def synthetic_eps(rows, cols, seed, row0=0):
    """eps = eps0*(1+9*U[0,1)) (SURVEY 8d cfg2/cfg3 recipe), float32, generated band-wise."""
    rng = np.random.default_rng(seed + row0)
    out = np.empty((rows, cols), np.float32)
    band = 2048
    for a in range(0, rows, band):
        b = min(rows, a + band)
        out[a:b] = 8.85418e-12 * (1 + 9 * rng.random((b - a, cols), dtype=np.float32))
    return out


def cpu_reference_rate(rows, cols, nsteps, seed=7):
    """Time the oracle's numpy restatement of the reference loop (fdtd.py:30-34) on rows x cols fp32."""
    from oracle import numpy_oracle as npo

    eps = synthetic_eps(rows, cols, seed)
    mu = np.full((rows, cols), np.float32(4 * np.pi * 1e-7))
    Ez, Hx, Hy = npo.grid_init(rows, cols, np.float32)
    npo.run(Ez, Hx, Hy, mu, eps, DT, DX, 1, source=(rows // 2, cols // 2, FC, "ricker"))  # touch pages
    t0 = time.perf_counter()
    npo.run(Ez, Hx, Hy, mu, eps, DT, DX, nsteps, source=(rows // 2, cols // 2, FC, "ricker"), step0=1)
    dt = time.perf_counter() - t0
    return rows * cols * nsteps / dt / 1e9, dt


def cpu_openmp_rate(rows, cols, seconds=4.0):
    """The oracle's C restatement with OpenMP on every host core: not a reference artefact (the reference is
    single-threaded numpy), reported for context only."""
    try:
        from oracle import c_oracle, numpy_oracle as npo

        c_oracle.build()
        eps = synthetic_eps(rows, cols, 7)
        mu = np.full((rows, cols), np.float32(4 * np.pi * 1e-7))
        ce, ch, coef = c_oracle.coefficients(eps, mu, DT, DX, np.dtype(np.float32))
        Ez, Hx, Hy = npo.grid_init(rows, cols, np.float32)
        amp = npo.source_table("ricker", 4096, DT, FC)
        c_oracle.run(Ez, Hx, Hy, ce, ch, coef, 2, amp, [(rows // 2, cols // 2)], None, omp=True)
        n, t0 = 0, time.perf_counter()
        while time.perf_counter() - t0 < seconds:
            c_oracle.run(Ez, Hx, Hy, ce, ch, coef, 4, amp, [(rows // 2, cols // 2)], None, omp=True)
            n += 4
        el = time.perf_counter() - t0
        return {"value": rows * cols * n / el / 1e9, "unit": "Gcell-updates/s", "cores": os.cpu_count(),
                "what": "C + OpenMP restatement (oracle/fdtd_oracle.c), not a reference artefact"}
    except Exception as e:  # the baseline is informative; never fail the bench over it
        return {"unavailable": str(e)[:200]}


def run_reference_arm(args, wl):
    """--impl reference: the reference's CPU path (numpy port), rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    rows, cols = min(wl["rows"], 2048), wl["cols"]
    sample = f"1 leapfrog step per bench step on a {rows}x{cols} fp32 band of the workload grid (numpy, 1 core)"
    from oracle import numpy_oracle as npo

    eps = synthetic_eps(rows, cols, 7)
    mu = np.full((rows, cols), np.float32(4 * np.pi * 1e-7))
    Ez, Hx, Hy = npo.grid_init(rows, cols, np.float32)
    src = (rows // 2, cols // 2, FC, "ricker")
    for w in range(args.warmup):
        npo.run(Ez, Hx, Hy, mu, eps, DT, DX, 1, source=src, step0=w)
    t0 = time.perf_counter()
    for s in range(args.steps):
        npo.run(Ez, Hx, Hy, mu, eps, DT, DX, 1, source=src, step0=args.warmup + s)
    el = time.perf_counter() - t0
    val = rows * cols * args.steps / el / 1e9
    line = {
        "impl": "reference", "metric": "Gcell-updates/s (fp32 E+H step)", "value": val, "unit": "Gcell-updates/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": el / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {wl['desc']}", "sample": sample},
        "cpu_baseline": {"value": val, "unit": "Gcell-updates/s", "cores": 1, "kind": "port", "sample": sample,
                         "host_cpus": os.cpu_count()},
        "e2e": {"value": val, "unit": "Gcell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def timed_steps(torch, dist, world, sim, inner, k, steps, warmup):
    """`warmup` untimed bench steps, then `steps` timed ones bracketed by barrier + synchronize; device-timed with CUDA
    events on the stream the kernels run on, max over ranks.  Returns (ms, stepping passes, kernel launches)."""
    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        sim.step(inner, k)
    barrier()
    l0, tl0 = sim.launch_count, sim.tile_launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(steps):
        sim.step(inner, k)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms, sim.tile_launch_count - tl0, sim.launch_count - l0


def traffic_table():
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        with open(tp) as f:
            return json.load(f)
    return {}


def roofline(cells_per_rank, leapfrog_steps, passes, ms, bytes_per_update, traffic_per_pass, note=None):
    """HBM roofline of the stepping kernel.  achieved = algorithmic bytes per pass / mean pass time (32 B per
    fp32 cell-update, 64 B per fp64 one); frac > 1 is what keeping the fields on chip for several steps buys.  frac_dram =
    the DRAM bytes ncu counted for one pass of this workload (profiles/traffic.json) / mean pass time / peak: how close the
    kernel is to the copy bandwidth with the bytes it really moves."""
    peak, peak_src = peaks()
    n_pass = max(1, int(passes))
    pass_s = ms * 1e-3 / n_pass
    achieved = bytes_per_update * cells_per_rank * leapfrog_steps / n_pass / pass_s / 1e9
    out = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic_per_pass,
           "frac_dram": (traffic_per_pass / pass_s / 1e9 / peak) if traffic_per_pass else None, "peak_source": peak_src,
           "passes": n_pass, "ms_per_pass": pass_s * 1e3}
    if note:
        out["note"] = note
    return out


def slab_parity(fd, torch, dist, rank, world, local_rank):
    """Correctness evidence for the y-slab path on the box the numbers come from, outside every timed region: a
    (2048 x slabs) x 3000 fp32 grid, 44 steps around the Ricker peak, halo rows moved by the kernels over peer links,
    against the same run on one undivided grid.  One process per GPU when launched with several ranks (every rank checks
    its own rows); on one GPU two slabs are driven from this process."""
    slabs = world if world > 1 else 2
    R, C, n, k = 2048 * slabs, 3000, 44, 8
    probes = [(R // 2, C // 2 + 3), (R // 2 - 1, 10), (5, 5), (R - 3, C - 3)]

    def setup(sm):
        sm.set_materials_random(9, 9.0)
        sm.set_point_source(R // 2, C // 2, 700, FC)
        sm.set_probes(probes, 700)
        sm.step_index = 640

    with fd.Simulation(R, C, np.float32, dt=DT, dx=DX, device=local_rank) as ref:
        setup(ref)
        ref.step(n, k)
        full = ref.state()
        rtr = ref.read_probes(640, n)
    ok = True
    if world > 1:
        sim = fd.SlabSimulation(R, C, np.float32, dt=DT, dx=DX, rank=rank, world=world, device=local_rank, halo=8)
        setup(sim)
        sim.step(n, k)
        sim.synchronize()
        got = [sim.owned(a) for a in sim.state()]
        b, e = sim.row_begin, sim.row_end
        ok &= np.array_equal(got[0], full[0][b:e]) and np.array_equal(got[1], full[1][b:e])
        ok &= np.array_equal(got[2][:min(e, R - 1) - b], full[2][b:min(e, R - 1)])
        tr = sim.read_probes(640, n)
        for p, (r, _) in enumerate(probes):
            if b <= r < e:
                ok &= np.array_equal(tr[:, p], rtr[:, p])
        sim.close()
        t = torch.tensor([1 if ok else 0], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        ok = bool(t.item())
        how = f"{world} ranks, one per GPU, peer links over CUDA IPC"
    else:
        grp = fd.InProcessSlabs(R, C, np.float32, dt=DT, dx=DX, world=2, devices=(local_rank,), halo=8)
        try:
            for s in grp.slabs:
                setup(s)
            grp.step(n, k)
            got = grp.gather()
            ok = all(np.array_equal(a, b_) for a, b_ in zip(got, full))
            for p, (r, _) in enumerate(probes):
                own = [s for s in grp.slabs if s.row_begin <= r < s.row_end][0]
                ok &= np.array_equal(own.read_probes(640, n)[:, p], rtr[:, p])
        finally:
            grp.close()
        how = "2 slabs on this GPU driven from one process, peer links"
    return {"result": "bit-exact" if ok else "MISMATCH", "grid": [R, C], "steps": n, "k": k, "how": how,
            "against": "the same run on one undivided grid (which the -m gpu tests check against the CPU oracle)"}


def strong_scaling(args, fd, torch, dist, rank, world, local_rank):
    """BASELINE configs[3]: 65536 x 65536 fp32 cut into y-slabs over the ranks of this run (the whole grid on one GPU at
    N = 1: 8 x 17.2 GB), 16 leapfrog steps per bench step.  The driver's 1/2/4/8 runs give the strong-scaling curve."""
    wl = WORKLOADS["cfg4"]
    R, C, inner = wl["rows"], wl["cols"], wl["inner"]
    steps, warmup = max(3, min(args.steps, 5)), 3
    sim = fd.SlabSimulation(R, C, np.float32, dt=DT, dx=DX, rank=rank, world=world, device=local_rank, halo=8)
    try:
        sim.set_stream(torch.cuda.current_stream().cuda_stream)
        sim.set_materials_random(seed=2026, span=9.0)
        sim.set_point_source(R // 2, C // 2, (warmup + steps + 1) * inner, FC)
        sim.set_probes([(R // 2, C // 2 + 16), (R // 4, C // 4), (3 * R // 4, C // 3), (R - 9, C // 2)], (warmup + steps + 1) * inner)
        ms, passes, launches = timed_steps(torch, dist, world, sim, inner, args.k, steps, warmup)
    finally:
        sim.close()
    cells = R * C
    return {"workload": f"cfg4: {wl['desc']}", "value": cells * inner * steps / (ms * 1e-3) / 1e9, "unit": "Gcell-updates/s", "n_gpus": world,
            "ms_per_step": ms / steps, "steps": steps, "warmup": warmup, "scaling": "strong", "global_rows": R, "cols": C,
            "inner_leapfrog_steps_per_step": inner, "gpu_launches": int(launches),
            "roofline": roofline(cells / world, inner * steps, passes, ms, BYTES_PER_UPDATE_F32, None)}


def other_configs(args, fd, torch):
    """The other BASELINE configurations on one GPU (rank 0), short runs after the headline: device-resident, CUDA-event
    timed like `value`, each with its own roofline.  cfg2 = one 10 000-step call; cfg5 = 1024 grids x 400 steps per call;
    fp64 = 8192^2 float64 (the reference's native dtype at a size that needs HBM); cfg1 = the reference demo itself
    (200 x 200 float64, 1000 steps, fdtd.py:14-21)."""
    traffic = traffic_table()
    out = {}

    class One:  # the interface timed_steps uses, over a plain Simulation
        def __init__(self, sim):
            self.sim, self.tile_launch_count = sim, 0

        def step(self, n, k=0):
            before = self.sim.pass_count
            self.sim.step(n, k)
            self.tile_launch_count += self.sim.pass_count - before

        launch_count = property(lambda self: self.sim.launch_count)

    def run(name, rows, cols, dtype, inner, steps, warmup, batch=1, desc=""):
        total = (steps + warmup + 1) * inner
        with fd.Simulation(rows, cols, dtype, dt=DT, dx=DX, batch=batch) as sim:
            sim.set_stream(torch.cuda.current_stream().cuda_stream)
            sim.set_materials_random(seed=2026, span=9.0)
            if batch > 1:
                fcs = np.linspace(18e9, 30e9, 16)
                sim.set_sources([(b, rows // 2 + (b % 7) - 3, cols // 2 + (b % 5) - 2, b % 16) for b in range(batch)],
                                np.stack([fd.source_table("ricker", total, DT, f) for f in fcs]))
                sim.set_probes([(b, rows // 2, cols // 2 + 20) for b in range(0, batch, max(1, batch // 8))], total)
            else:
                sim.set_point_source(rows // 2, cols // 2, total, FC)
                sim.set_probes([(rows // 2, cols // 2 + 16), (rows // 4, cols // 4), (8, cols // 2), (rows // 2, 8)], total)
            ms, passes, launches = timed_steps(torch, None, 1, One(sim), inner, 0, steps, warmup)
        cells = rows * cols * batch
        bpu = BYTES_PER_UPDATE_F32 * (2 if np.dtype(dtype) == np.float64 else 1)
        out[name] = {"workload": desc, "value": cells * inner * steps / (ms * 1e-3) / 1e9, "unit": "Gcell-updates/s", "dtype": np.dtype(dtype).name,
                     "ms_per_step": ms / steps, "steps": steps, "warmup": warmup, "inner_leapfrog_steps_per_step": inner,
                     "leapfrog_steps_per_s": inner * steps / (ms * 1e-3), "gpu_launches": int(launches),
                     "roofline": roofline(cells, inner * steps, passes, ms, bpu, traffic.get(name))}

    run("cfg2", 4096, 4096, np.float32, 10000, 3, 3, desc=WORKLOADS["cfg2"]["desc"])
    run("cfg5", 256, 256, np.float32, 400, 5, 3, batch=1024, desc=WORKLOADS["cfg5"]["desc"])
    run("fp64_8192", 8192, 8192, np.float64, 96, 3, 3, desc="8192x8192 fp64 (the reference's native dtype), random permittivity, 96 steps per bench step")
    run("cfg1", 200, 200, np.float64, 1000, 5, 3, desc="python-src/fdtd.py defaults: 200x200 float64, 1000 steps, Ricker at the centre (BASELINE configs[0])")
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=sorted(WORKLOADS))
    ap.add_argument("--inner", type=int, default=0, help="leapfrog steps per bench step (0 = workload default)")
    ap.add_argument("--k", type=int, default=0, help="leapfrog steps per HBM round trip (0 = library default)")
    ap.add_argument("--variant", type=int, default=0, help="kernel variant (0 auto, 1 generic, 2 fast+generic)")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"], help="how y-slabs move their halo rows")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip strong_scaling, slab_parity and other_configs")
    args = ap.parse_args()
    wl = dict(WORKLOADS[args.workload])
    if args.inner:
        wl["inner"] = args.inner
    if args.impl == "reference":
        run_reference_arm(args, wl)
        return

    import torch
    import torch.distributed as dist

    import fdtd2d_b200 as fd
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    cpu_affinity = bind_to_gpu_cpus(local_rank)
    torch.cuda.set_device(local_rank)
    if world > 1:
        if args.exchange == "nccl":
            # the halo exchange must not queue behind the persistent stepping kernel of the same pass: NCCL's stream gets
            # high priority, so its few CTAs are placed first when the band tasks are done (SlabSimulation's side stream too)
            os.environ.setdefault("TORCH_NCCL_HIGH_PRIORITY", "1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world} (launch with torchrun for N>1)"

    strong = args.workload == "cfg4"
    batch = wl.get("batch", 0)
    cols = wl["cols"]
    grows = wl["rows"] if (strong or batch) else wl["rows"] * world  # weak scaling: fixed rows per GPU
    inner = wl["inner"]
    k = args.k  # 0: the library picks (fp32: 8)
    stream = torch.cuda.current_stream().cuda_stream

    sim = make_sim(fd, wl, grows, cols, rank, world, local_rank, k, args.exchange)
    sim.set_stream(stream)
    if args.variant:
        sim.set_kernel_variant(args.variant)
    sim.set_materials_random(seed=2026, span=9.0)
    total_steps = (args.warmup + args.steps + 2) * inner
    sim.set_point_source(grows // 2, cols // 2, total_steps, FC)
    probes = [(grows // 2, cols // 2 + 16), (grows // 4, cols // 4), (3 * grows // 4, cols // 3), (grows // 2, 8),
              (8, cols // 2), (grows - 9, cols // 2), (grows // 3, cols - 9), (grows // 2 + 100, cols // 2 + 100)]
    sim.set_probes(probes, total_steps)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    for _ in range(args.warmup):
        sim.step(inner, k)
    barrier()
    if rank == 0:
        sampler.start()
    ms, tile_launches, launches = timed_steps(torch, dist, world, sim, inner, k, args.steps, 0)
    clocks = sampler.stop() if rank == 0 else None
    cells = grows * cols * (batch * world if batch else 1)  # whole job
    value = cells * inner * args.steps / (ms * 1e-3) / 1e9
    n_pass = max(1, int(tile_launches))  # stepping passes (HBM round trips of the fields) per rank in the timed region
    resident = bool(batch) and tile_launches == args.steps  # cfg5: one cluster-resident launch per bench step
    k_eff = None if resident else round(inner * args.steps / n_pass)
    roof = roofline(cells / world, inner * args.steps, n_pass, ms, BYTES_PER_UPDATE_F32, traffic_table().get(args.workload),
                    note="achieved = 32 B x cell-updates per stepping-kernel pass / mean pass time; keeping the fields on chip for k "
                         "steps (wavefront / tiles) or for the whole call (cluster-resident) moves fewer DRAM bytes than the algorithmic "
                         "count, so frac > 1 is the point; frac_dram = ncu's DRAM bytes per pass / pass time / peak")
    sim.close()

    # ---- e2e: host (pinned) inputs, H2D + coefficient formation + inner steps + D2H, every step ----
    e2e = e2e_up = None
    if not args.no_e2e:
        e2e = run_e2e(args, wl, fd, torch, dist, rank, world, local_rank, grows, cols, inner, k, barrier)
        if not args.no_extras:  # the heavier job (initial fields uploaded too), for comparison with the earlier lines
            barrier()
            e2e_up = run_e2e(args, wl, fd, torch, dist, rank, world, local_rank, grows, cols, inner, k, barrier, upload_state=True)

    parity = strong_line = others = None
    if not args.no_extras and args.workload == "cfg3":
        barrier()
        parity = slab_parity(fd, torch, dist, rank, world, local_rank)
        barrier()
        strong_line = strong_scaling(args, fd, torch, dist, rank, world, local_rank)
        barrier()
        if world == 1:  # (the other configurations are single-GPU workloads: measured by the N = 1 run only)
            others = other_configs(args, fd, torch)

    cpu = None
    if world == 1 and not args.no_cpu:  # the CPU baseline is timed by the N = 1 run only
        crow = min(4096, wl["rows"])
        n = max(2, int(6e8 / (crow * cols)))  # ~10 s of numpy work at ~60-90 Mcell/s
        v, el = cpu_reference_rate(crow, cols, n)
        cpu = {"value": v, "unit": "Gcell-updates/s", "cores": 1, "kind": "port", "host_cpus": os.cpu_count(),
               "sample": f"{n} leapfrog steps on a {crow}x{cols} fp32 band/grid of the workload, numpy port of the "
                         f"reference loop ({el:.1f} s); numpy elementwise ops use 1 core",
               "c_openmp_port": cpu_openmp_rate(crow, cols)}
    if rank == 0:
        line = {
            "metric": "Gcell-updates/s (fp32 E+H step)", "value": value, "unit": "Gcell-updates/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.workload}: {wl['desc']}", "global_rows": grows, "cols": cols,
                       "batch_per_gpu": batch or 1, "inner_leapfrog_steps_per_step": inner,
                       "k_temporal": k_eff,
                       "kernel": ("cluster-resident (grid on chip for the whole step call, 1 launch per bench step)" if resident
                                  else f"{k_eff} leapfrog steps per HBM round trip: row-streaming wavefront strips (packed fp32x2 arithmetic; the "
                                       "left / right Mur ring rides along) from ~2000^2, else persistent TMA-fed tiles; "
                                       "edge-capable tiles on the top / bottom ring, corners, sources, probes"),
                       "parallelism": ("independent grids per rank" if batch else
                                       f"y-slabs x{world}, halo rows stored into the neighbour GPU by the stepping kernels (NVLink peer stores + flags)"
                                       if args.exchange == "p2p" else f"y-slabs x{world}, NCCL send/recv per pass") if world > 1 else "single GPU",
                       "l2": "state is larger than L2 (inputs larger than L2; no flush needed)",
                       "seed": 2026, "source": "ricker fc=30e9 at centre", "probes": len(probes),
                       "host_cpu_affinity": cpu_affinity},
            "roofline": roof, "cpu_baseline": cpu, "e2e": e2e, "e2e_state_upload": e2e_up, "gpu_launches": int(launches), "tile_kernel_launches": int(tile_launches),
            "clocks": clocks, "slab_parity": parity, "strong_scaling": strong_line, "other_configs": others,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_e2e(args, wl, fd, torch, dist, rank, world, local_rank, grows, cols, inner, k, barrier, upload_state=False):
    """Public-API path with host buffers.  One job is what fdtd.py:21-38 does with one structure: the material maps eps, mu
    come from the host (pinned memory), the fields start from grid_init's zeros (made on the device: zero_state), the
    coefficients are formed on the device, `inner` leapfrog steps run, Ez and the probe traces go back to the host.
    upload_state=True is the heavier job of the earlier bench lines: the initial Ez, Hx, Hy are host arrays as well and are
    uploaded with the maps (20 instead of 8 bytes per cell host -> device).

    A job that is uploaded, stepped and downloaded strictly in turn leaves the GPU idle while PCIe moves 5 arrays in and
    one out, so the bench keeps TWO jobs in flight from ONE host thread with the library's non-blocking copies
    (fdtd2d_set_materials_async / _upload_state_async / _download_state_async): two handles (two slab groups when the
    grid is sharded), each with its own copy stream, so the copies of one job overlap the stepping kernels of the
    other.  The jobs share one compute stream (set_stream): job after job on the device, and on slabs the same kernel order
    on every rank (a kernel that waits for a neighbour rank's flag can then never wait for a kernel queued behind another
    waiting kernel).
    Every job still uploads all its inputs and downloads its results inside the timed region; the figure is jobs
    finished per wall-clock second, pipeline fill and drain included."""
    batch = wl.get("batch", 0)
    inflight = 2
    if os.environ.get("BENCH_E2E_INFLIGHT"):
        inflight = max(1, int(os.environ["BENCH_E2E_INFLIGHT"]))
    steps = max(4, min(args.steps, 16))
    sims = [make_sim(fd, wl, grows, cols, rank, world, local_rank, k, args.exchange) for _ in range(inflight)]
    if not batch:
        # One compute stream for the jobs in flight: their stepping kernels run job after job instead of pass by pass in
        # turn, so a job is finished -- and its download and the next upload under way -- while the next one computes;
        # on slabs it also gives every rank the same kernel order.
        for sm in sims:
            sm.set_stream(torch.cuda.current_stream().cuda_stream)
    sim = sims[0]
    raw = [sm.sim for sm in sims]  # the Simulation under the slab / batch wrapper
    lr, hyr = sim.local_rows, sim.hy_rows
    pre = (batch,) if batch else ()

    def pinned(shape):
        return torch.zeros(pre + shape, dtype=torch.float32, pin_memory=True).numpy()

    eps, mu = pinned((lr, cols)), pinned((lr, cols))
    eps[...] = synthetic_eps(lr * max(1, batch), cols, 2026, sim.row0).reshape(eps.shape)
    mu[...] = np.float32(4 * np.pi * 1e-7)
    Ez, Hx, Hy = (pinned((lr, cols)), pinned((lr, cols - 1)), pinned((hyr, cols))) if upload_state else (None, None, None)
    outs = [pinned((lr, cols)) for _ in sims]
    mur = None if batch else fd_mur_coef(eps if sim.row0 == 0 else None, mu, dist, world, torch)
    for sm, r in zip(sims, raw):
        sm.set_point_source(grows // 2, cols // 2, inner, FC)
        sm.set_probes([(grows // 2, cols // 2 + 16), (grows // 4, cols // 4)], inner)
        if mur is not None:
            r.set_mur_coef(mur)  # (a slab that does not hold cell (0,0) cannot form it from its own rows)
    h2d = (eps.nbytes + mu.nbytes + (Ez.nbytes + Hx.nbytes + Hy.nbytes if upload_state else 0)) * world
    d2h = (outs[0].nbytes + inner * (8 if batch else 2) * 4) * world

    def issue(w):  # everything asynchronous: returns as soon as the work is queued
        sm, r = sims[w], raw[w]
        r.step_index = 0
        r.set_materials_async(eps, mu)
        if upload_state:
            r.set_state_async(Ez, Hx, Hy)
        else:
            r.zero_state()  # grid_init (main.py:79-85) on the device
        sm.step(inner, k)
        r.read_Ez_async(outs[w])

    def finish(w):
        raw[w].synchronize()  # kernels, copies and (slabs) the neighbours' last halo rows
        return raw[w].read_probes(0, inner)

    def run_jobs(n):
        queued = []
        for j in range(n):
            if len(queued) == inflight:
                finish(queued.pop(0))
            w = j % inflight
            issue(w)
            queued.append(w)
        while queued:
            finish(queued.pop(0))

    run_jobs(inflight)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    run_jobs(steps)
    e1.record()
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    # jobs overlap on several streams, so the wall clock (>= any one stream's device time) is the honest figure
    ms = max(e0.elapsed_time(e1), wall_ms)
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    for sm in sims:
        sm.close()
    cells = grows * cols * (batch * world if batch else 1)
    return {"value": cells * inner * steps / (ms * 1e-3) / 1e9, "unit": "Gcell-updates/s",
            "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "steps": steps,
            "ms_per_step": ms / steps, "jobs_in_flight": inflight,
            "what": "Simulation API with pinned host arrays: set_materials(eps, mu) + "
                    + ("set_state(Ez, Hx, Hy)" if upload_state else "zero_state (= grid_init, on the device)")
                    + " + step(inner) + read_Ez + read_probes, every step; two jobs in flight from one host thread "
                    "(non-blocking copies on per-handle copy streams) so one job's PCIe copies overlap the other's kernels; "
                    "wall clock over all jobs"}


def fd_mur_coef(eps_rank0, mu, dist, world, torch):
    """Mur coefficient from global cell (0,0) (main.py:30-31), shared with every slab."""
    if eps_rank0 is not None:
        c = 1 / np.sqrt(mu[0, 0] * eps_rank0[0, 0])
        coef = np.float32((c * DT - DX) / (c * DT + DX))
    else:
        coef = np.float32(0)
    if world > 1:
        t = torch.tensor([float(coef)], device="cuda", dtype=torch.float32)
        dist.broadcast(t, src=0)
        coef = np.float32(t.item())
    return coef


if __name__ == "__main__":
    main()
