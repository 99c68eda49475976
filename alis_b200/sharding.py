"""z-slab domain decomposition of the time-stepping path.

One process per GPU (torch.distributed), each owning a contiguous range of z-planes of the
global grid as its own engine instance.  The only exchange the path needs is one plane of
the two tangential components per interface per half step (reference update.c:110-111):

    after updateE:  E.x,E.y of my top plane    -> upper neighbour's ghost plane c=0
    after updateH:  H.x,H.y of my bottom plane -> lower neighbour's ghost plane c=nz+1

The planes are contiguous in the engine's x-fastest/z-slowest layout, so the exchange is a
plain pairwise send/recv of device buffers handed out by alisb_halo_info -- NCCL over
NVLink on the GPUs, gloo in the CPU tests (with the oracle engine).  The engine runs on
torch's current stream, so the NCCL work is ordered after the pass that produced the plane
and before the pass that consumes it without any host synchronisation.
"""
import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from . import capi


class _DevBuf:
    def __init__(self, ptr, nfloat):
        self.__cuda_array_interface__ = {"shape": (nfloat,), "typestr": "<f4", "data": (ptr, False), "version": 2}


def _as_tensor(ptr, nbytes, cuda):
    n = nbytes // 4
    if cuda:
        return torch.as_tensor(_DevBuf(ptr, n), device="cuda")
    arr = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_float)), shape=(n,))
    return torch.from_numpy(arr)


class SlabRunner:
    """Steps one z-slab engine and exchanges halos with its neighbours."""

    def __init__(self, engine: capi.Engine, rank, world, cuda, complex_field=0):
        self.e, self.rank, self.world, self.cuda = engine, rank, world, cuda
        self.ncf = complex_field + 1
        self.send, self.recv = {}, {}
        for s in range(2):
            for m in range(self.ncf):
                for comp in range(2):
                    sp, rp, nb = engine.halo_info(s, m, comp)
                    self.send[(s, m, comp)] = _as_tensor(sp, nb, cuda)
                    self.recv[(s, m, comp)] = _as_tensor(rp, nb, cuda)
        if cuda:
            engine.set_stream(torch.cuda.current_stream().cuda_stream)
        self.exchanged_bytes = 0

    def exchange(self, set_):
        if self.world == 1:
            return
        # E planes travel up (to rank+1), H planes travel down (to rank-1)
        dst = self.rank + 1 if set_ == capi.SET_E else self.rank - 1
        src = self.rank - 1 if set_ == capi.SET_E else self.rank + 1
        if 0 <= dst < self.world:
            self.e.halo_pack(set_)
        ops = []
        for m in range(self.ncf):
            for comp in range(2):
                if 0 <= dst < self.world:
                    ops.append(dist.P2POp(dist.isend, self.send[(set_, m, comp)], dst))
                    self.exchanged_bytes += self.send[(set_, m, comp)].numel() * 4
                if 0 <= src < self.world:
                    ops.append(dist.P2POp(dist.irecv, self.recv[(set_, m, comp)], src))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        if 0 <= src < self.world:
            self.e.halo_unpack(set_)

    def step(self):
        self.e.step_E()
        self.exchange(capi.SET_E)
        self.e.step_H()
        self.exchange(capi.SET_H)

    def run(self, steps):
        for _ in range(steps):
            self.step()
