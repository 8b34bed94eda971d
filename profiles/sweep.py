"""Sweep fast-kernel tile shapes (FDTD2D_FAST_CFG) and k on a 16384^2 fp32 grid; check each shape
bit-for-bit against the generic kernel on a smaller grid first.  Run under gpurun."""
import os, sys, subprocess, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
DT, DX = 5e-14, 1e-4

def child(cfg):
    import torch
    import fdtd2d_b200 as fd
    # parity vs generic
    R, C, n = 1500, 2100, 24
    outs = []
    for variant, k in ((1, 4), (2, 3), (2, 4), (2, 8)):
        with fd.Simulation(R, C, np.float32, dt=DT, dx=DX) as sim:
            sim.set_kernel_variant(variant)
            sim.set_materials_random(5, 9.0)
            sim.set_point_source(R // 2, C // 2, 700, 30e9)
            sim.step_index = 640
            sim.step(n, k)
            outs.append(sim.state())
    ok = all(np.array_equal(a, b) for o in outs[1:] for a, b in zip(o, outs[0]))
    res = {"cfg": cfg, "parity": ok}
    R = C = 16384
    with fd.Simulation(R, C, np.float32, dt=DT, dx=DX) as sim:
        sim.set_stream(torch.cuda.current_stream().cuda_stream)
        sim.set_materials_random(1, 9.0)
        sim.set_point_source(R // 2, C // 2, 4000, 30e9)
        for k in (2, 3, 4, 5, 6, 7, 8):
            sim.step(2 * k, k); torch.cuda.synchronize()
            n = 10 * k
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); sim.step(n, k); e1.record(); torch.cuda.synchronize()
            res[f"k{k}"] = round(R * C * n / e0.elapsed_time(e1) / 1e6, 1)
    print(json.dumps(res), flush=True)

if __name__ == "__main__":
    if len(sys.argv) > 1:
        child(int(sys.argv[1]))
    else:
        for cfg in [int(c) for c in os.environ.get("SWEEP_CFGS", "0,1,2,3,4,5,6").split(",")]:
            env = dict(os.environ, FDTD2D_FAST_CFG=str(cfg))
            subprocess.run([sys.executable, __file__, str(cfg)], env=env)
