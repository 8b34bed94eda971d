"""Edge tiles alone (option measure_skip = 2; wrong fields) at 4096^2 fp32 for k = 1..8: the slope is the cost of one step of
the slowest tile, the intercept the launch + load + store."""
import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fdtd2d_b200 as fd
R = int(os.environ.get("R", 4096))
for dt in (np.float32, np.float64):
  for src, prb in ((0, 0), (1, 0), (1, 2)):
    with fd.Simulation(R, R, dt, dt=5e-14, dx=1e-4) as sim:
        sim.set_stream(torch.cuda.current_stream().cuda_stream)
        sim.set_materials_random(1, 9.0)
        if src: sim.set_point_source(R // 2, R // 2, 200000, 30e9)
        if prb: sim.set_probes([(R // 2, R // 2 + 16), (R // 4, R // 4)], 200000)
        sim.set_option("measure_skip", 2)
        sim.set_option("wave_min_tiles", 0)
        for k in (1, 2, 4, 6, 8):
            n = 200 * k
            sim.zero_state()
            try:
                sim.step(8 * k, k); torch.cuda.synchronize()
            except Exception as e:
                print(k, str(e)[:100]); continue
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            p0 = sim.pass_count
            e0.record(); sim.step(n, k); e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1); passes = sim.pass_count - p0
            info = sim.plan_info(k)
            print(f"{np.dtype(dt).name} src={src} probes={prb} k={k}: {ms / passes * 1e3:7.1f} us per pass, {info['edge_tiles']} edge tiles, {info['wave_runs']} runs", flush=True)
