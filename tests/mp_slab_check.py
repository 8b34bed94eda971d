"""Multi-process y-slab parity check (run under torchrun, one rank per GPU):
    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tests/mp_slab_check.py [p2p|nccl]
Every rank steps its slab -- halo rows moved by the kernels themselves over peer links (p2p, default) or by NCCL
send/recv -- rank 0 also runs the single-domain simulation and compares the gathered owned rows bit for bit."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fdtd2d_b200 as fd  # noqa: E402

DT, DX, FC = 5e-14, 1e-4, 30e9


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    mode = sys.argv[1] if len(sys.argv) > 1 else "p2p"
    if mode == "nccl":
        os.environ.setdefault("TORCH_NCCL_HIGH_PRIORITY", "1")
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    R, C, n, k = 2048 * world, 3000, 44, 8
    sim = fd.SlabSimulation(R, C, np.float32, dt=DT, dx=DX, rank=rank, world=world, device=local, halo=8, exchange=mode)
    sim.set_stream(torch.cuda.current_stream().cuda_stream)
    sim.set_materials_random(9, 9.0)
    sim.set_point_source(R // 2, C // 2, 700, FC)
    sim.set_probes([(R // 2, C // 2 + 3), (R // 2 - 1, 10), (5, 5), (R - 3, C - 3)], 700)
    sim.step_index = 640
    sim.step(n // 2, k, overlap=False)  # first half: exchange after each pass
    sim.step(n - n // 2, k, overlap=True)  # second half: exchange overlapped with the rest of the pass
    sim.synchronize()
    torch.cuda.synchronize()
    Ez, Hx, Hy = sim.state()
    lo, cnt = sim.row_begin - sim.row0, sim.row_end - sim.row_begin
    mine = [a[lo:lo + cnt] for a in (Ez, Hx, Hy)]
    tr = torch.from_numpy(sim.read_probes(640, n).astype(np.float32)).cuda()
    dist.all_reduce(tr)  # each probe is recorded by its owner only; the others hold zeros
    ok = True
    if rank == 0:
        with fd.Simulation(R, C, np.float32, dt=DT, dx=DX, device=local) as ref:
            ref.set_materials_random(9, 9.0)
            ref.set_point_source(R // 2, C // 2, 700, FC)
            ref.set_probes([(R // 2, C // 2 + 3), (R // 2 - 1, 10), (5, 5), (R - 3, C - 3)], 700)
            ref.step_index = 640
            ref.step(n, k)
            full = ref.state()
            rtr = ref.read_probes(640, n)
        ok = np.array_equal(tr.cpu().numpy(), rtr)
    else:
        full = None
    # compare every rank's owned rows against the single-domain arrays held by rank 0
    gathered = [None] * world
    dist.all_gather_object(gathered, [sim.row_begin, sim.row_end] + mine)
    if rank == 0:
        for b, e, gEz, gHx, gHy in gathered:
            ok &= np.array_equal(gEz, full[0][b:e]) and np.array_equal(gHx, full[1][b:e])
            ok &= np.array_equal(gHy, full[2][b:min(e, R - 1)])
        print(f"mp_slab_check world={world} exchange={mode}: {'OK bit-exact' if ok else 'MISMATCH'}", flush=True)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    sim.close()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
