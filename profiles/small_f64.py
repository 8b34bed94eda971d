"""cfg1 (200 x 200 fp64, 1000 steps, the reference demo) for k in 4, 6, 8: steps per second."""
import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fdtd2d_b200 as fd
for R in (200, 1000):
    for k in (8, 9, 10, 11, 12):
        with fd.Simulation(R, R, np.float64, dt=5e-14, dx=1e-4) as sim:
            sim.set_stream(torch.cuda.current_stream().cuda_stream)
            eps, mu = fd.material_init(None, R, R)
            sim.set_materials(eps, mu)
            sim.set_point_source(R // 2, R // 2, 5000, 30e9)
            sim.step(1000, k); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                sim.step(1000, k)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 3
            print(f"fp64 {R}^2 k={k}: {1000/ms*1e3:9.0f} steps/s ({ms:.2f} ms per 1000 steps) {sim.plan_info(k)['edge_tiles']} tiles", flush=True)
