"""A few passes (K levels each, default 8) at 16384^2 for ncu (FDTD2D_FAST_CFG selects the plain-tile kernel)."""
import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fdtd2d_b200 as fd
R = C = int(os.environ.get("R", 16384))
with fd.Simulation(R, C, np.float32, dt=5e-14, dx=1e-4) as sim:
    sim.set_materials_random(1, 9.0)
    sim.set_point_source(R // 2, C // 2, 2000, 30e9)
    K = int(os.environ.get("K", 8))
    sim.step(3 * K, K)
    sim.synchronize()
print("done")
