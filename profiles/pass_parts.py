"""Where a pass spends its time, without a profiler: the default pass, the pass without its edge tiles, without its runs,
and with the edge tiles serialised (option measure_skip; the fields are wrong in those runs).  env: CASES."""
import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fdtd2d_b200 as fd
cases = [("cfg2 4096^2 f32", 4096, np.float32, 2000), ("cfg3 16384^2 f32", 16384, np.float32, 400), ("8192^2 f64", 8192, np.float64, 400)]
for name, R, dt, n in cases:
    with fd.Simulation(R, R, dt, dt=5e-14, dx=1e-4) as sim:
        sim.set_stream(torch.cuda.current_stream().cuda_stream)
        sim.set_materials_random(1, 9.0)
        sim.set_point_source(R // 2, R // 2, 20000, 30e9)
        sim.set_probes([(R // 2, R // 2 + 16), (R // 4, R // 4)], 20000)
        for label, skip in (("whole pass", 0), ("no edge tiles", 1), ("edge tiles only", 2), ("edge tiles serialised", 4)):
            sim.set_option("measure_skip", skip)
            sim.zero_state()
            sim.step(64); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            p0 = sim.pass_count
            e0.record(); sim.step(n); e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1); passes = sim.pass_count - p0
            print(f"{name}: {label:22s} {ms / passes * 1e3:8.1f} us per pass  ({R * R * n / ms / 1e6:7.1f} Gcell/s)  {sim.plan_info(8)}", flush=True)
