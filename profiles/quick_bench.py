"""k sweep on one GPU (k = 0: library default).  usage: [KS=8] [SIZES=4096,16384] [FDTD2D_WAVEFRONT=0] python profiles/quick_bench.py"""
import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fdtd2d_b200 as fd, torch
DT, DX = 5e-14, 1e-4
for R in [int(x) for x in os.environ.get("SIZES", "4096,16384").split(",")]:
    C = R
    for k in [int(x) for x in os.environ.get("KS", "1,2,4,6,8").split(",")]:
        with fd.Simulation(R, C, np.float32, dt=DT, dx=DX) as sim:
            sim.set_stream(torch.cuda.current_stream().cuda_stream)
            sim.set_materials_random(1, 9.0)
            sim.set_point_source(R // 2, C // 2, 2000, 30e9)
            sim.step(2 * (k or 12), k); torch.cuda.synchronize()
            n = 16 * (k or 12)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); sim.step(n, k); e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            print(f"{R}x{C} k={k}: {ms/n:.3f} ms/step, {R*C*n/ms/1e6:.1f} Gcell/s", flush=True)
