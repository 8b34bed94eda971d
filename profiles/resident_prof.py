"""One launch of the cluster-resident kernel for ncu: B grids of R x C, n steps."""
import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fdtd2d_b200 as fd
DT, DX = 5e-14, 1e-3
B, R, C, n = int(os.environ.get("B", 33)), int(os.environ.get("R", 256)), int(os.environ.get("C", 256)), int(os.environ.get("N", 200))
with fd.Simulation(R, C, np.float32, dt=DT, dx=DX, batch=B) as sim:
    sim.set_kernel_variant(4)
    sim.set_materials_random(1, 4.0)
    amp = fd.source_table("ricker", 4000, DT, 20e9)
    sim.set_sources([(b, R // 2, C // 2, 0) for b in range(B)], amp[None, :])
    sim.set_probes([(b, R // 2, C // 2 + 4) for b in range(B)], 4000)
    sim.step(n)
    sim.synchronize()
print("done")
