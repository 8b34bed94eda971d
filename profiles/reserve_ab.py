"""A/B of the options edge_reserve / ring_cost: pass time of the default path with the edge tiles first on
every SM (edge_reserve = 0, the behaviour until now) against the wavefront first on fewer SMs, and a CRC of Ez after the
same number of steps for every setting -- scheduling must not change a bit.  Host clock around synchronised runs of
>= 50 ms (no torch).  usage: python profiles/reserve_ab.py [cfg2,f32_8192,cfg3,f64]"""
import os
import sys
import time
import zlib

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fdtd2d_b200 as fd  # noqa: E402

T0 = time.time()
CASES = {
    # name: (rows, dtype, timed steps, [(edge_reserve, ring_cost, check Ez?)])
    "cfg2": (4096, np.float32, 4000, [(0, 0, 1), (-1, 0, 1), (21, 220, 1), (24, 0, 1), (27, 0, 0), (0, 208, 1), (0, 0, 0)]),
    "f32_8192": (8192, np.float32, 1600, [(0, 0, 1), (-1, 0, 1)]),
    "cfg3": (16384, np.float32, 400, [(0, 0, 1), (6, 0, 1), (0, 208, 0), (0, 0, 0)]),
    "f64": (8192, np.float64, 400, [(0, 0, 1), (6, 0, 1)]),
}
for name in (sys.argv[1].split(",") if len(sys.argv) > 1 else list(CASES)):
    R, dtype, n, combos = CASES[name]
    with fd.Simulation(R, R, dtype, dt=5e-14, dx=1e-4) as sim:
        sim.set_materials_random(2026, 9.0)
        sim.set_point_source(R // 2, R // 2, n + 64, 30e9)
        sim.set_probes([(R // 2, R // 2 + 5)] + [(R // 8 * i + 3, R // 8 * i + 7) for i in range(1, 8)], n + 64)
        crc0 = None
        for reserve, ring_cost, chk in combos:
            sim.set_option("edge_reserve", reserve)
            sim.set_option("ring_cost", ring_cost)
            best = None
            for rep in range(2):
                sim.zero_state()
                sim.step(64)
                sim.synchronize()
                p0, t0 = sim.pass_count, time.perf_counter()
                sim.step(n)
                sim.synchronize()
                dt_s, passes = time.perf_counter() - t0, sim.pass_count - p0
                best = dt_s if best is None else min(best, dt_s)
            info = sim.plan_info(8)
            line = (f"[{time.time() - T0:5.1f}s] {name}: edge_reserve {reserve:3d} ring_cost {ring_cost:3d} -> reserved SMs {info['reserve_sms']:3d}, "
                    f"runs {info['wave_runs']:5d}, edge tiles {info['edge_tiles']:4d}: {best / passes * 1e6:8.1f} us per pass, "
                    f"{R * R * n / best / 1e9:7.1f} Gcell/s")
            if chk:
                crc = zlib.crc32(sim.read_Ez().view(np.uint8).reshape(-1))
                crc0 = crc if crc0 is None else crc0
                line += f"  Ez crc {crc:08x} {'== baseline' if crc == crc0 else '!= BASELINE: MISMATCH'}"
            print(line, flush=True)
fd.release_handles()
print("reserve_ab done", flush=True)
