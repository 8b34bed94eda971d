"""cfg5 (1024 x 256^2 fp32, 400 steps) on the cluster-resident kernel: sweep the kernel shape (resident_cfg),
the CTAs per grid (resident_cluster) and the rows taken off the first / last band (resident_trim)."""
import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fdtd2d_b200 as fd, torch
DT, DX = 5e-14, 1e-3
R = C = 256
B, n = 1024, 400
shapes = {0: (3, 16), 1: (4, 12), 2: (2, 16), 3: (4, 8), 4: (3, 12)}
res = []
with fd.Simulation(R, C, np.float32, dt=DT, dx=DX, batch=B) as sim:
    sim.set_stream(torch.cuda.current_stream().cuda_stream)
    sim.set_materials_random(1, 4.0)
    amp = fd.source_table("ricker", 4000, DT, 20e9)
    sim.set_sources([(b, R // 2, C // 2, 0) for b in range(B)], amp[None, :])
    sim.set_probes([(b, R // 2, C // 2 + 4) for b in range(B)], 4000)
    for cfg in (0, 1, 2, 4):
        for cl in (0, 6, 7, 8):
            for trim in (0, 6, 12, 18, 24, 30, 36):
                try:
                    sim.set_option("resident_cfg", cfg); sim.set_option("resident_cluster", cl); sim.set_option("resident_trim", trim)
                    sim.set_kernel_variant(4)
                    sim.step_index = 0
                    sim.step(16); torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(); sim.step(n); e1.record(); torch.cuda.synchronize()
                except Exception as e:
                    print(f"cfg {cfg} {shapes[cfg]} cluster {cl} trim {trim}: {str(e)[:80]}", flush=True)
                    continue
                ms = e0.elapsed_time(e1)
                g = B * R * C * n / ms / 1e6
                res.append((g, cfg, cl, trim))
                print(f"cfg {cfg} {shapes[cfg]} cluster {cl} trim {trim}: {ms:.2f} ms {g:.1f} Gcell/s", flush=True)
print("best:", sorted(res, reverse=True)[:8])
