"""fp64 rate of the tile kernels (every tile runs on tile_edge<double>).  usage: python profiles/fp64_bench.py"""
import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fdtd2d_b200 as fd, torch
for R in (8192, 4096):
    with fd.Simulation(R, R, np.float64, dt=5e-14, dx=1e-4) as sim:
        sim.set_stream(torch.cuda.current_stream().cuda_stream)
        sim.set_materials_random(1, 9.0); sim.set_point_source(R // 2, R // 2, 2000, 30e9)
        sim.step(8, 4); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); sim.step(64, 4); e1.record(); torch.cuda.synchronize()
        print(f"fp64 {R}^2 k=4: {R * R * 64 / e0.elapsed_time(e1) / 1e6:.1f} Gcell/s", flush=True)
