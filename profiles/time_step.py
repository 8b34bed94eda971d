"""Time Simulation.step on the GPU.  usage: timeit.py ROWS COLS [k=8] [variant=0] [dtype=f32] [batch=1] [nsteps=10*k]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fdtd2d_b200 as fd
a = sys.argv[1:]
R, C = int(a[0]), int(a[1])
k = int(a[2]) if len(a) > 2 else 8
variant = int(a[3]) if len(a) > 3 else 0
dtype = {"f32": np.float32, "f64": np.float64}[a[4] if len(a) > 4 else "f32"]
batch = int(a[5]) if len(a) > 5 else 1
n = int(a[6]) if len(a) > 6 else 10 * k
with fd.Simulation(R, C, dtype, dt=5e-14, dx=1e-4, batch=batch) as sim:
    sim.set_stream(torch.cuda.current_stream().cuda_stream)
    sim.set_kernel_variant(variant)
    sim.set_materials_random(1, 9.0)
    sim.set_sources([(b, R // 2, C // 2, 0) for b in range(batch)], fd.source_table("ricker", 4000, 5e-14, 30e9)[None, :])
    sim.step(2 * k, k); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); sim.step(n, k); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"{R}x{C} x{batch} {np.dtype(dtype).name} k={k} variant={variant} cfg={os.environ.get('FDTD2D_FAST_CFG','default')}: "
          f"{ms/n*1e3:.1f} us/step  {batch*R*C*n/ms/1e6:.1f} Gcell/s", flush=True)
