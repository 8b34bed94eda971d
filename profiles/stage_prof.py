"""A few 12-level passes of the staged wavefront at 16384^2 for ncu (STAGE = groups per CTA)."""
import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fdtd2d_b200 as fd
R = C = int(os.environ.get("R", 16384))
with fd.Simulation(R, C, np.float32, dt=5e-14, dx=1e-4) as sim:
    sim.set_option("stage", int(os.environ.get("STAGE", 4)))
    sim.set_materials_random(1, 9.0)
    sim.set_point_source(R // 2, C // 2, 2000, 30e9)
    sim.set_probes([(R // 2, C // 2 + 16), (R // 4, C // 4), (8, C // 2)], 2000)
    sim.step(36, 12)
    sim.synchronize()
    print("done", sim.plan_info(12), sim.launch_count)
