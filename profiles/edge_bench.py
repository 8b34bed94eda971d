"""Time of the edge-tile kernel per tile class: a tall thin grid (every tile touches the left or right ring), a flat wide
one (top / bottom ring), both with the plain tiles kept off the wavefront kernel.  usage: python profiles/edge_bench.py"""
import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("FDTD2D_WAVE_MIN_TILES", "100000000")
import fdtd2d_b200 as fd, torch
DT, DX = 5e-14, 1e-4
for R, C, what in ((32768, 200, "left+right ring in every tile"), (32768, 330, "left | plain | right"), (200, 32768, "top+bottom ring in every tile"),
                   (330, 32768, "top | plain | bottom")):
    with fd.Simulation(R, C, np.float32, dt=DT, dx=DX) as sim:
        sim.set_stream(torch.cuda.current_stream().cuda_stream)
        sim.set_kernel_variant(2)
        sim.set_materials_random(1, 9.0)
        sim.set_point_source(R // 2, C // 2, 2000, 30e9)
        sim.step(16, 8); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 128
        e0.record(); sim.step(n, 8); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        tiles = -(-R // 48) * -(-C // 112)
        print(f"{R}x{C} ({what}): {ms / (n / 8) * 1e3:.1f} us per pass, ~{tiles} tiles, {ms / (n / 8) * 1e3 * 148 / tiles:.1f} us per tile and SM, {R * C * n / ms / 1e6:.1f} Gcell/s", flush=True)
