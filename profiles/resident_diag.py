"""Timing of the cfg5 workload (1024 x 256^2 fp32, 400 steps per call) on a given build of the library (env LIB) for a few
cluster / trim settings.  Used with -DFDTD2D_RES_DIAG builds (wrong results) to see what the ring work costs."""
import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fdtd2d_b200 as fd, torch
import fdtd2d_b200._lib as L
if os.environ.get("LIB"):
    L.LIB_PATH = os.path.abspath(os.environ["LIB"])
DT, DX = 5e-14, 1e-3
B, R, C, n = int(os.environ.get("B", 1024)), int(os.environ.get("R", 256)), int(os.environ.get("C", 256)), 400
for cfg in [int(x) for x in os.environ.get("CFGS", "0").split(",")]:
  for trim in [int(x) for x in os.environ.get("TRIMS", "-1").split(",")]:
    with fd.Simulation(R, C, np.float32, dt=DT, dx=DX, batch=B) as sim:
        sim.set_stream(torch.cuda.current_stream().cuda_stream)
        sim.set_option("resident_cfg", cfg)
        sim.set_option("resident_trim", trim)
        if os.environ.get("CLUSTER"): sim.set_option("resident_cluster", int(os.environ["CLUSTER"]))
        if os.environ.get("ROWS"): sim.set_option("resident_rows", int(os.environ["ROWS"]))
        sim.set_materials_random(1, 4.0)
        amp = fd.source_table("ricker", 4000, DT, 20e9)
        sim.set_sources([(b, R // 2, C // 2, 0) for b in range(B)], amp[None, :])
        sim.set_probes([(b, R // 2, C // 2 + 4) for b in range(B)], 4000)
        sim.step(n, 0); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); sim.step(n, 0); sim.step(n, 0); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 2
        print(f"{os.path.basename(os.environ.get('LIB','default'))} cfg {cfg} trim {trim} B={B}: {ms:.2f} ms, {B*R*C*n/ms/1e6:.1f} Gcell/s", flush=True)
