"""Timeline of bench.py's e2e pipeline on N ranks (torchrun): host time stamps around every call of a job and device
events on the shared compute stream, relative to a common start.  env: INFLIGHT (2), JOBS (6), UPLOAD_STATE (0)."""
import os, sys, time, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import bench
import fdtd2d_b200 as fd
rank, world, lr_ = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr_)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr_))
R = C = 16384; grows = R * world; inner = 1000; k = 8
inflight, jobs, up = int(os.environ.get("INFLIGHT", 2)), int(os.environ.get("JOBS", 6)), int(os.environ.get("UPLOAD_STATE", 0))
wl = {"rows": R, "cols": C}
sims = [bench.make_sim(fd, wl, grows, C, rank, world, lr_, k, "p2p") for _ in range(inflight)]
cs = torch.cuda.current_stream()
for sm in sims: sm.set_stream(cs.cuda_stream)
raw = [sm.sim for sm in sims]
lr, hyr = sims[0].local_rows, sims[0].hy_rows
pin = lambda shape: torch.zeros(shape, dtype=torch.float32, pin_memory=True).numpy()
eps, mu = pin((lr, C)), pin((lr, C))
eps[...] = bench.synthetic_eps(lr, C, 2026, sims[0].row0); mu[...] = np.float32(4 * np.pi * 1e-7)
st = (pin((lr, C)), pin((lr, C - 1)), pin((hyr, C))) if up else None
outs = [pin((lr, C)) for _ in sims]
mur = bench.fd_mur_coef(eps if sims[0].row0 == 0 else None, mu, dist, world, torch)
for sm, r in zip(sims, raw):
    sm.set_point_source(grows // 2, C // 2, inner, bench.FC); r.set_mur_coef(mur)
log = []
def stamp(what, j): log.append((time.perf_counter(), what, j))
evs = []
def issue(w, j):
    sm, r = sims[w], raw[w]
    stamp("issue", j); r.step_index = 0
    r.set_materials_async(eps, mu); stamp(" materials_async queued", j)
    if up: r.set_state_async(*st)
    else: r.zero_state()
    stamp(" state queued", j)
    strm = cs
    e0 = torch.cuda.Event(enable_timing=True); e0.record(strm)
    sm.step(inner, k); stamp(" step queued", j)
    e1 = torch.cuda.Event(enable_timing=True); e1.record(strm)
    r.read_Ez_async(outs[w]); stamp(" read queued", j)
    evs.append((j, e0, e1))
def finish(w, j):
    stamp("finish", j); raw[w].synchronize(); stamp(" finished", j)
def run(n, base):
    q = []
    for j in range(n):
        if len(q) == inflight: finish(*q.pop(0))
        issue(j % inflight, base + j); q.append((j % inflight, base + j))
    while q: finish(*q.pop(0))
run(inflight, -inflight)
torch.cuda.synchronize()
if world > 1: dist.barrier()
log.clear(); evs.clear()
ref = torch.cuda.Event(enable_timing=True); ref.record(cs)
t0 = time.perf_counter()
run(jobs, 0)
torch.cuda.synchronize()
t1 = time.perf_counter()
if rank in (0, world - 1):
    print(f"rank {rank}: {jobs} jobs in {(t1 - t0) * 1e3:.1f} ms = {(t1 - t0) * 1e3 / jobs:.1f} ms per job", flush=True)
    for t, what, j in log: print(f"rank {rank} {(t - t0) * 1e3:8.1f} ms  job {j} {what}")
    for j, e0, e1 in evs: print(f"rank {rank} job {j}: stepping kernels on the device {ref.elapsed_time(e0):8.1f} .. {ref.elapsed_time(e1):8.1f} ms")
for sm in sims: sm.close()
if world > 1: dist.destroy_process_group()
