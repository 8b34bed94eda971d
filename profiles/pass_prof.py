"""A few passes of the default path at R x C in DTYPE (f32 / f64) for ncu.  env: R, C, DTYPE, STEPS, K (0 = library default), OPTS ("key=v,key=v")."""
import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fdtd2d_b200 as fd
R = int(os.environ.get("R", 16384)); C = int(os.environ.get("C", R))
dtype = np.float64 if os.environ.get("DTYPE", "f32") == "f64" else np.float32
with fd.Simulation(R, C, dtype, dt=5e-14, dx=1e-4) as sim:
    for kv in filter(None, os.environ.get("OPTS", "").split(",")):
        key, v = kv.split("=")
        sim.set_option(key, int(v))
    sim.set_materials_random(1, 9.0)
    sim.set_point_source(R // 2, C // 2, 4000, 30e9)
    sim.set_probes([(R // 2, C // 2 + 16), (R // 4, C // 4), (8, C // 2), (R // 2, 8)], 4000)
    sim.step(int(os.environ.get("STEPS", 24)), int(os.environ.get("K", 0)))
    sim.synchronize()
    print("done", sim.pass_count, sim.launch_count)
