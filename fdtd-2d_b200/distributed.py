"""y-slab domain decomposition over the GPUs of one box.

The reference is single-process; this is the new multi-GPU capability named by BASELINE.json
(configs[3]).  One process per GPU (torchrun), rank r owns a contiguous band of rows of all five arrays
(rows are the reference's axis 0: `Ez[1:, :] - Ez[:-1, :]`, main.py:69) plus `halo` ghost rows on each
side that has a neighbour.  A pass advances k <= halo leapfrog steps on chip; ghost rows go stale one row
per step (the same argument as for tile halos), so after each pass the `halo` owned rows next to each
internal boundary must reach the neighbour's ghost rows, for Ez, Hx and Hy.  There is no reduction
anywhere, so the result is bit-identical to the single-GPU run.

Two ways to move the halo rows:

* ``exchange="p2p"`` (default on GPUs): PEER LINKS.  At set-up every rank exports a blob (CUDA IPC handles of its field
  buffers and flag block), the blobs are all-gathered once over torch.distributed, and each rank maps its neighbours'
  buffers (`fdtd2d_peer_attach`).  From then on the stepping kernels do the exchange themselves: the first tasks of a pass
  compute the band rows, store them straight into the neighbour's ghost rows over NVLink and raise its flag; the tasks
  of the next pass that read ghost rows wait for their own flag.  No host work, no collective and no extra kernel per
  pass: `step(n)` is one library call.
* ``exchange="nccl"``: one batched group of NCCL send/recv per pass over tensors that alias the library's buffers
  (torch.distributed P2P ops), issued from Python between `fdtd2d_pass_begin` (band tasks) and `fdtd2d_pass_end` (the
  rest) so that it overlaps the rest of the pass.  The stepping kernel of a pass is persistent, so NCCL's stream has to
  be a high-priority one: `TORCH_NCCL_HIGH_PRIORITY=1` is set by SlabSimulation (world > 1, this mode) unless the
  caller already chose -- it only takes effect if the process group is created afterwards.

The exchange plumbing (`HaloExchange`) is independent of CUDA so that it is covered by world_size-2 gloo tests on CPU.
"""
from __future__ import annotations

import os
from typing import Callable

import numpy as np

from .simulation import Simulation

TOP, BOTTOM = 0, 1
FIELDS = (0, 1, 2)  # Ez, Hx, Hy


def slab_rows(global_rows: int, world: int, rank: int) -> tuple[int, int]:
    """Balanced contiguous partition of rows: [begin, end) owned by `rank`."""
    base, extra = divmod(global_rows, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


class HaloExchange:
    """Send/recv of halo blocks with the two neighbours of `rank` in a 1-D chain of `world` ranks.

    blocks(field, side) must return (send_tensor, recv_tensor): the `halo` owned rows next to `side` and
    the ghost rows on that side, as torch tensors on the communication device."""

    def __init__(self, rank: int, world: int, blocks: Callable, group=None):
        self.rank, self.world, self.blocks, self.group = rank, world, blocks, group

    def neighbours(self):
        out = []
        if self.rank > 0:
            out.append((TOP, self.rank - 1))
        if self.rank < self.world - 1:
            out.append((BOTTOM, self.rank + 1))
        return out

    def exchange(self):
        import torch.distributed as dist

        ops = []
        for side, peer in self.neighbours():
            for f in FIELDS:
                send, recv = self.blocks(f, side)
                ops.append(dist.P2POp(dist.isend, send, peer, group=self.group))
                ops.append(dist.P2POp(dist.irecv, recv, peer, group=self.group))
        if not ops:
            return
        for req in dist.batch_isend_irecv(ops):
            req.wait()


def exchange_peer_blobs(my_blob: bytes, rank: int, world: int, group=None, gather=None) -> dict:
    """{side: blob of the neighbour on that side} for `rank` in a chain of `world` ranks.  `gather(blob) -> list of
    every rank's blob` defaults to torch.distributed.all_gather_object; the CPU tests pass their own."""
    if gather is None:
        import torch.distributed as dist

        def gather(b):
            out = [None] * world
            dist.all_gather_object(out, b, group=group)
            return out

    blobs = gather(my_blob)
    if len(blobs) != world:
        raise RuntimeError(f"gathered {len(blobs)} blobs for a world of {world}")
    nb = {}
    if rank > 0:
        nb[TOP] = blobs[rank - 1]
    if rank < world - 1:
        nb[BOTTOM] = blobs[rank + 1]
    return nb


class _DeviceBlock:
    """A device memory range exposed through __cuda_array_interface__ so torch can alias it."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3,
                                         "strides": None}


class SlabSimulation:
    """A global_rows x cols simulation sharded into y-slabs, one per rank.  With world == 1 it is a thin
    wrapper over `Simulation`.  Mirrors `Simulation`'s interface with GLOBAL row indices.

    exchange: "p2p" (peer links, see the module docstring), "nccl", or "none" (the caller moves the halo rows, e.g.
    InProcessSlabs).  group: the torch.distributed group of the slabs (default: the world)."""

    def __init__(self, global_rows: int, cols: int, dtype=np.float32, *, dt: float, dx: float, rank: int = 0,
                 world: int = 1, device: int = 0, halo: int = 8, group=None, exchange: str = "p2p"):
        if exchange not in ("p2p", "nccl", "none"):
            raise ValueError("exchange must be 'p2p', 'nccl' or 'none'")
        self.rank, self.world, self.halo = rank, world, halo
        self.global_rows, self.cols = global_rows, cols
        self.row_begin, self.row_end = slab_rows(global_rows, world, rank)
        self.exchange, self.group = (exchange if world > 1 else "none"), group
        if self.exchange == "nccl":
            # (read by torch when the process group is created: set it before init_process_group to have an effect)
            os.environ.setdefault("TORCH_NCCL_HIGH_PRIORITY", "1")
        slab = None if world == 1 else (global_rows, self.row_begin, self.row_end, halo)
        self.sim = Simulation(global_rows, cols, dtype, dt=dt, dx=dx, device=device, slab=slab)
        self.dtype = self.sim.dtype
        self.row0, self.local_rows, self.hy_rows = self.sim.row0, self.sim.local_rows, self.sim._hy_rows
        self.tile_launch_count = 0
        self._tensors = {}
        self._xchg = self._xchg_next = self._comm = self._main = None
        if self.exchange == "nccl":
            import torch

            self._xchg = HaloExchange(rank, world, self._blocks, group)
            self._xchg_next = HaloExchange(rank, world, lambda f, side: self._blocks(f, side, True), group)
            # kernels and the exchange are ordered through torch streams: adopt the current one
            self.set_stream(torch.cuda.current_stream(device).cuda_stream)
            self._comm = torch.cuda.Stream(device, priority=-1)  # the exchange goes ahead of the rest of the pass
        elif self.exchange == "p2p":
            for side, blob in exchange_peer_blobs(self.sim.peer_export(), rank, world, group).items():
                self.sim.peer_attach(side, blob)
            self.barrier()  # nobody steps (and writes into a neighbour) before every link is in place

    # -- halo plumbing ------------------------------------------------------------------------
    def _blocks(self, field, side, next_state=False):
        import torch

        sp, rp, nb = self.sim.halo_block(field, side, next_state)
        out = []
        for ptr in (sp, rp):
            t = self._tensors.get(ptr)
            if t is None:
                t = torch.as_tensor(_DeviceBlock(ptr, nb), device=torch.device("cuda", self.sim.device))
                self._tensors[ptr] = t
            out.append(t)
        return out

    def exchange_halos(self):
        if self._xchg is not None:
            self._xchg.exchange()

    def barrier(self):
        """All slabs have finished everything issued so far (device and host).  With peer links this must separate the
        last step of one run from the upload of the next state: a neighbour's kernels write into this slab's ghost rows."""
        self.sim.synchronize()
        if self.world > 1 and self.exchange != "none":
            import torch.distributed as dist

            dist.barrier(group=self.group)

    # -- forwarding ---------------------------------------------------------------------------
    def set_stream(self, s):
        self.sim.set_stream(s)
        if self.exchange == "nccl":
            import torch

            dev = self.sim.device
            self._main = torch.cuda.default_stream(dev) if not s else torch.cuda.ExternalStream(s, dev)

    def set_kernel_variant(self, v):
        self.sim.set_kernel_variant(v)

    def set_option(self, key, value):
        self.sim.set_option(key, value)

    def set_materials_random(self, seed, span=9.0):
        self.sim.set_materials_random(seed, span)

    def set_materials(self, eps_local, mu_local, mur_coef=None):
        """eps/mu for the LOCAL rows (ghost rows included).  Slabs that do not hold global cell (0,0)
        need `mur_coef` (main.py:30-31 uses that cell's materials only)."""
        self.sim.set_materials(eps_local, mu_local)
        if mur_coef is not None:
            self.sim.set_mur_coef(mur_coef)

    def set_state(self, Ez, Hx, Hy):
        """Upload the LOCAL rows (ghost rows included).  With peer links every slab must have finished its previous run
        first (the neighbours' kernels store into this slab's ghost rows): call barrier() between runs."""
        self.sim.set_state(Ez, Hx, Hy)

    def zero_state(self):
        """grid_init (main.py:79-85) for this slab, on the device; the step index returns to 0."""
        self.sim.zero_state()

    def set_point_source(self, row, col, nsteps, fc=30e9, kind="ricker"):
        self.sim.set_point_source(row, col, nsteps, fc, kind)

    def set_sources(self, cells, tables):
        self.sim.set_sources(cells, tables)

    def set_probes(self, cells, capacity_steps):
        self.sim.set_probes(cells, capacity_steps)

    def read_probes(self, first_step=0, n_steps=None):
        """Traces of the probes this rank owns; columns of probes owned elsewhere are zero."""
        return self.sim.read_probes(first_step, n_steps)

    def read_Ez(self, out=None):
        return self.sim.read_Ez(out)

    def state(self):
        return self.sim.state()

    def owned(self, a):
        """Slice the owned rows out of a local (ghost-including) array."""
        lo = self.row_begin - self.row0
        return a[..., lo:lo + (self.row_end - self.row_begin), :]

    @property
    def step_index(self):
        return self.sim.step_index

    @step_index.setter
    def step_index(self, v):
        self.sim.step_index = v

    @property
    def launch_count(self):
        return self.sim.launch_count

    def synchronize(self):
        self.sim.synchronize()

    def close(self):
        if self.exchange == "p2p" and getattr(self.sim, "_h", None) is not None and self.sim._h.value:
            self.barrier()  # no neighbour may still be writing into buffers that are about to be freed
            self.sim.peer_detach()
            self.barrier()
        self._tensors.clear()
        self.sim.close()

    # -- time stepping ------------------------------------------------------------------------
    def step(self, n_steps: int, k: int = 0, overlap: bool = True):
        """n_steps leapfrog steps.  With peer links this is one library call (the kernels exchange the halo rows).  With
        NCCL, halos are exchanged after every pass of k steps; overlap=True starts each exchange as soon as the tasks
        that produce the boundary rows are done (fdtd2d_pass_begin) and runs it on a side stream while the rest of the
        pass computes."""
        from . import DEFAULT_K

        if self.world == 1 or self.exchange == "p2p":
            before = self.sim.pass_count
            self.sim.step(n_steps, min(k, self.halo) if self.world > 1 else k)  # k = 0: the library picks (fp32: 8)
            self.tile_launch_count += self.sim.pass_count - before
            return
        if self.exchange == "none":
            raise RuntimeError("exchange='none': step the slab with sim.step(k) and move the halo rows yourself")
        k = k or DEFAULT_K
        import torch

        k = min(k, self.halo)
        left = n_steps
        while left > 0:
            kk = min(k, left)
            if overlap:
                self.sim.pass_begin(kk)
                ev = torch.cuda.Event()
                ev.record(self._main)
                self._comm.wait_event(ev)
                with torch.cuda.stream(self._comm):
                    self._xchg_next.exchange()
                self.sim.pass_end()
                self._main.wait_stream(self._comm)
            else:
                self.sim.step(kk, kk)
                self.exchange_halos()
            self.tile_launch_count += 1
            left -= kk


class InProcessSlabs:
    """All slabs of a global_rows x cols grid driven by ONE process: slab i lives on devices[i % len].

    exchange="p2p" (default): the slabs are linked as peers inside the process (plain device pointers; peer access is
    enabled between different GPUs) and the kernels exchange the halo rows themselves, exactly as in the one-process-per-
    GPU layout.  Slabs that share a GPU share one stream and are stepped pass by pass in turn, so a slab's kernel never
    waits for a kernel queued behind it.  exchange="copy": device-to-device copies of the halo blocks after every pass,
    with full synchronisation (the round-1 flavour; also exercises fdtd2d_halo_block).
    This is the single-process flavour of the decomposition; it lets the whole slab logic (row offsets, ghost rows,
    band tasks, peer stores, flags) be checked on a one-GPU box."""

    def __init__(self, global_rows, cols, dtype=np.float32, *, dt, dx, world, devices=(0,), halo=8, exchange="p2p"):
        if exchange not in ("p2p", "copy"):
            raise ValueError("exchange must be 'p2p' or 'copy'")
        self.world, self.halo, self.exchange = world, halo, exchange if world > 1 else "copy"
        self.slabs = [SlabSimulation(global_rows, cols, dtype, dt=dt, dx=dx, rank=r, world=world,
                                     device=devices[r % len(devices)], halo=halo, exchange="none") for r in range(world)]
        if self.exchange == "p2p":
            first_on = {}
            for s in self.slabs:  # one stream per GPU
                lead = first_on.setdefault(s.sim.device, s)
                if lead is not s:
                    s.sim.set_stream(lead.sim.cuda_stream)
            blobs = [s.sim.peer_export() for s in self.slabs]
            for r, s in enumerate(self.slabs):
                for side, blob in exchange_peer_blobs(blobs[r], r, world, gather=lambda _b: blobs).items():
                    s.sim.peer_attach(side, blob)

    def each(self, fn):
        return [fn(s) for s in self.slabs]

    def exchange_halos(self):
        self.each(lambda s: s.synchronize())
        for r in range(self.world - 1):
            up, dn = self.slabs[r], self.slabs[r + 1]
            for f in FIELDS:
                up_send, up_recv = up._blocks(f, BOTTOM)
                dn_send, dn_recv = dn._blocks(f, TOP)
                dn_recv.copy_(up_send)
                up_recv.copy_(dn_send)
        import torch

        for d in {s.sim.device for s in self.slabs}:
            torch.cuda.synchronize(d)

    def step(self, n_steps, k=0):
        k = min(k or 8, self.halo) if self.world > 1 else (k or 0)
        left = n_steps
        while left > 0:
            kk = min(k, left) if k else left
            self.each(lambda s: s.sim.step(kk, kk if self.world > 1 else k))  # pass by pass, slab by slab
            if self.world > 1 and self.exchange == "copy":
                self.exchange_halos()
            left -= kk

    def synchronize(self):
        self.each(lambda s: s.synchronize())

    def gather(self):
        """(Ez, Hx, Hy) of the whole grid in the reference's shapes, assembled from the owned rows
        (the last slab's Hy is one row short, main.py:84; slicing clamps)."""
        self.synchronize()
        out = [[], [], []]
        for s in self.slabs:
            lo, n = s.row_begin - s.row0, s.row_end - s.row_begin
            for f, a in enumerate(s.state()):
                out[f].append(a[..., lo:lo + n, :])
        return tuple(np.concatenate(parts, axis=-2) for parts in out)

    def close(self):
        live = [s for s in self.slabs if getattr(s.sim, "_h", None) is not None and s.sim._h.value]
        if self.exchange == "p2p" and len(live) == len(self.slabs):
            for s in live:
                s.sim.synchronize()
            for s in live:
                s.sim.peer_detach()
                s.sim.set_stream(None)  # back to its own stream before the shared one is destroyed with its owner
        self.each(lambda s: s.close())
