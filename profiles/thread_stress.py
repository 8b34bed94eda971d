"""Two host threads, one handle each, the e2e sequence (upload, step, download) in a loop: looks for faults that only
show when two handles work at the same time.  usage: [R=2048] [N=2000] [JOBS=6] python profiles/thread_stress.py"""
import os, sys, threading, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fdtd2d_b200 as fd, torch
R = C = int(os.environ.get("R", 2048)); N = int(os.environ.get("N", 2000)); JOBS = int(os.environ.get("JOBS", 6))
T = int(os.environ.get("THREADS", 2))
rng = np.random.default_rng(0)
eps = (fd.EPSILON0 * (1 + 9 * rng.random((R, C)))).astype(np.float32); mu = np.full((R, C), np.float32(fd.MU0))
Ez, Hx, Hy = np.zeros((R, C), np.float32), np.zeros((R, C - 1), np.float32), np.zeros((R - 1, C), np.float32)
sims = [fd.Simulation(R, C, np.float32, dt=5e-14, dx=1e-4) for _ in range(T)]
for s in sims:
    s.set_point_source(R // 2, C // 2, N, 30e9); s.set_probes([(R // 2, C // 2 + 16), (R // 4, C // 4)], N)
errs, res = [], [None] * T
def work(w):
    try:
        torch.cuda.set_device(0)
        for _ in range(JOBS):
            s = sims[w]; s.step_index = 0; s.set_materials(eps, mu); s.set_state(Ez, Hx, Hy); s.step(N, 0)
            res[w] = (s.read_Ez(), s.read_probes(0, N))
    except Exception as e:
        errs.append(e)
ts = [threading.Thread(target=work, args=(w,)) for w in range(T)]
[t.start() for t in ts]; [t.join() for t in ts]
if errs: print("FAIL", str(errs[0])[:200]); sys.exit(1)
ok = all(np.array_equal(res[0][0], r[0]) and np.array_equal(res[0][1], r[1]) for r in res[1:])
print("ok, handles agree:", ok, float(np.abs(res[0][0]).max()))
