"""Per-warp phase timing of a cluster-resident kernel (needs a library built with -DFDTD2D_RES_TIMING, env LIB):
mean cycles per step spent in each phase, per CTA of the cluster and warp.  FDTD2D_RESIDENT_CFG picks the kernel."""
import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fdtd2d_b200 as fd
import fdtd2d_b200._lib as L
if os.environ.get("LIB"):
    L.LIB_PATH = os.path.abspath(os.environ["LIB"])
B, R, C, n = int(os.environ.get("B", 1)), 256, 256, 1000
x2 = os.environ.get("FDTD2D_RESIDENT_CFG", "0") == "5"
NW = 16
with fd.Simulation(R, C, np.float32, dt=5e-14, dx=1e-3, batch=B) as sim:
    sim.set_kernel_variant(4)
    sim.set_materials_random(1, 4.0)
    sim.set_probes([(0, 128, 132)], 4000)
    sim.step(n)
    t = sim.read_probes(0, 2000).reshape(-1)[:8 * NW * 8].reshape(8, NW, 8)
names = (["publish+park", "wait barrier", "H rows 0-4", "H last+hxa(remote)", "E rows", "slots+S2", "corners", "ring src+loadback"] if x2 else
         ["pre-A(loadback..park)", "wait A", "H rows", "H last+hxa(remote)", "E rows", "park+S2", "wait C'", "pass+D'(or warp)"])
np.set_printoptions(linewidth=200, suppress=True)
for c in range(6):
    print(f"CTA {c}: per-warp cycles per step, columns = {names}")
    print(np.round(t[c]).astype(int))
    print("  sum per warp:", np.round(t[c].sum(1)).astype(int))
