"""world_size 2 and 3 over gloo on CPU: the y-slab halo-exchange plumbing (fdtd2d_b200.HaloExchange +
slab_rows) reproduces the single-domain result bit for bit.  The stepping of each slab is done by the
CPU oracle here (it is only the stand-in stepper of this host-logic test): each rank advances its local
window (owned rows + `halo` ghost rows) k steps, then exchanges halos, exactly the loop
SlabSimulation.step runs around the CUDA kernel."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

DT, DX, FC = 5e-14, 1e-4, 30e9
R, C, HALO, K, PASSES = 120, 70, 8, 3, 6


def _problem():
    rng = np.random.default_rng(42)
    eps = (8.85418e-12 * (1 + 9 * rng.random((R, C)))).astype(np.float32)
    mu = (4 * np.pi * 1e-7 * (1 + rng.random((R, C)))).astype(np.float32)
    Ez = (1e-3 * rng.standard_normal((R, C))).astype(np.float32)
    Hx = (1e-6 * rng.standard_normal((R, C - 1))).astype(np.float32)
    Hy = (1e-6 * rng.standard_normal((R - 1, C))).astype(np.float32)
    return eps, mu, Ez, Hx, Hy


def _worker(rank, world, port, out_dir):
    import sys

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import fdtd2d_b200 as fd
    from oracle import c_oracle, numpy_oracle as npo

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    eps, mu, Ez, Hx, Hy = _problem()
    ce, ch, coef = c_oracle.coefficients(eps, mu, DT, DX, np.dtype(np.float32))
    begin, end = fd.slab_rows(R, world, rank)
    lo = begin - (HALO if rank > 0 else 0)
    hi = end + (HALO if rank < world - 1 else 0)
    last = hi == R
    loc = [Ez[lo:hi].copy(), Hx[lo:hi].copy(), Hy[lo:(hi - 1 if last else hi)].copy()]
    lce, lch = ce[lo:hi].copy(), ch[lo:hi].copy()
    own_first, own_last = begin - lo, end - lo  # local rows of the owned band

    def blocks(field, side):
        a = loc[field]
        if side == 0:
            send, recv = a[own_first:own_first + HALO], a[0:own_first]
        else:
            send, recv = a[own_last - HALO:own_last], a[own_last:own_last + HALO]
        assert send.shape[0] == HALO and recv.shape[0] == HALO
        return torch.from_numpy(send), torch.from_numpy(recv)

    xchg = fd.HaloExchange(rank, world, blocks)
    amp = npo.source_table("ricker", K * PASSES, DT, FC)
    src_global = (R // 2, C // 2)
    for p in range(PASSES):
        # the local window is stepped as a stand-alone grid; rows within K+4 of an internal edge go
        # stale (K + 4 < HALO) and are refreshed by the exchange
        src = [(src_global[0] - lo, src_global[1])] if lo <= src_global[0] < hi else []
        if not last:
            # a window that does not hold the global last row has a full-height Hy; the oracle wants (rows-1, C)
            hy_full = loc[2]
            hy = hy_full[:-1].copy()
            c_oracle.run(loc[0], loc[1], hy, lce, lch, coef, K, amp[p * K:(p + 1) * K] if src else None, src or None)
            hy_full[:-1] = hy
        else:
            c_oracle.run(loc[0], loc[1], loc[2], lce, lch, coef, K, amp[p * K:(p + 1) * K] if src else None, src or None)
        xchg.exchange()
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), Ez=loc[0][own_first:own_last], Hx=loc[1][own_first:own_last],
             Hy=loc[2][own_first:min(own_last, loc[2].shape[0])], begin=begin, end=end)
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world", [2, 3])
def test_slab_exchange_matches_single_domain(tmp_path, world):
    from oracle import c_oracle, numpy_oracle as npo

    c_oracle.build()
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    eps, mu, Ez, Hx, Hy = _problem()
    ce, ch, coef = c_oracle.coefficients(eps, mu, DT, DX, np.dtype(np.float32))
    amp = npo.source_table("ricker", K * PASSES, DT, FC)
    c_oracle.run(Ez, Hx, Hy, ce, ch, coef, K * PASSES, amp, [(R // 2, C // 2)])
    for r in range(world):
        g = np.load(os.path.join(str(tmp_path), f"rank{r}.npz"))
        b, e = int(g["begin"]), int(g["end"])
        assert np.array_equal(g["Ez"], Ez[b:e]), f"Ez rows of rank {r}"
        assert np.array_equal(g["Hx"], Hx[b:e]), f"Hx rows of rank {r}"
        assert np.array_equal(g["Hy"], Hy[b:min(e, R - 1)]), f"Hy rows of rank {r}"


def _blob_worker(rank, world, port, out_dir):
    import sys

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import fdtd2d_b200 as fd

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = bytes([rank]) * 640  # stands in for fdtd2d_peer_export's blob
    got = fd.exchange_peer_blobs(mine, rank, world)
    np.save(os.path.join(out_dir, f"blob{rank}.npy"), np.array([got.get(0, b"\xff")[0], got.get(1, b"\xff")[0], len(got)]))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_peer_blobs_reach_the_right_neighbours(tmp_path, world):
    """The one collective of the peer-link set-up (SlabSimulation, exchange='p2p'): every rank's blob is all-gathered and
    each rank keeps the blobs of the slab above (side 0) and below (side 1) -- checked over gloo, no GPU."""
    mp.spawn(_blob_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        top, bottom, n = np.load(os.path.join(str(tmp_path), f"blob{r}.npy"))
        assert top == (r - 1 if r > 0 else 255) and bottom == (r + 1 if r < world - 1 else 255)
        assert n == (r > 0) + (r < world - 1)
