"""fp64 throughput at R x C (default 8192^2) for k in K (comma list), wavefront on / off; prints Gcell/s per setting."""
import os, sys, time, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fdtd2d_b200 as fd
R = int(os.environ.get("R", 8192)); C = int(os.environ.get("C", R)); n = int(os.environ.get("STEPS", 96))
for wave in (1, 0):
    for k in [int(v) for v in os.environ.get("K", "4,6,8").split(",")]:
        with fd.Simulation(R, C, np.float64, dt=5e-14, dx=1e-4) as sim:
            sim.set_option("wavefront", wave)
            sim.set_stream(torch.cuda.current_stream().cuda_stream)
            sim.set_materials_random(1, 9.0)
            sim.set_point_source(R // 2, C // 2, 10 * n, 30e9)
            sim.step(n, k); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                sim.step(n, k)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 3
            print(f"fp64 {R}x{C} wavefront={wave} k={k}: {R*C*n/ms/1e6:8.1f} Gcell/s  ({ms/n*1e3:.1f} us/step) plan {sim.plan_info(k)}", flush=True)
