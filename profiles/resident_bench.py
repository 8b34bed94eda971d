"""Quick timing of the batched small-grid workload (cfg5: 1024 x 256^2 fp32): cluster-resident vs tiled path."""
import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fdtd2d_b200 as fd, torch
DT, DX = 5e-14, 1e-3
R, C = int(os.environ.get("R", 256)), int(os.environ.get("C", 256))
runs = [(4, b, 1000) for b in [int(x) for x in os.environ.get("BS", "1,33,1024").split(",")]]
for variant, B, n in runs:
    with fd.Simulation(R, C, np.float32, dt=DT, dx=DX, batch=B) as sim:
        sim.set_stream(torch.cuda.current_stream().cuda_stream)
        sim.set_kernel_variant(variant)
        sim.set_materials_random(1, 4.0)
        amp = fd.source_table("ricker", 4000, DT, 20e9)
        if not os.environ.get("NOSRC"):
            sim.set_sources([(b, R // 2, C // 2, 0) for b in range(B)], amp[None, :])
            sim.set_probes([(b, R // 2, C // 2 + 4) for b in range(B)], 4000)
        sim.step(16, 8); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); sim.step(n, 8); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print(f"variant {variant} B={B} {R}x{C} n={n}: {ms:.2f} ms, {ms/n*1e3:.2f} us/step, {B*R*C*n/ms/1e6:.1f} Gcell/s", flush=True)
