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

    def __init__(self, engine: capi.Engine, rank, world, cuda, complex_field=0, overlap=True, fuse=False):
        self.e, self.rank, self.world, self.cuda = engine, rank, world, cuda
        self.overlap, self.pending_H = overlap, None
        self.fuse = fuse
        self.ncf = complex_field + 1
        self._tensors = {}
        if cuda:
            engine.set_stream(torch.cuda.current_stream().cuda_stream)
        if fuse:
            engine.set_fusion(True)
        self.exchanged_bytes = 0

    def _planes(self, set_, m, comp):
        """(send, recv) tensors over the halo planes the engine currently exposes for this field set;
        with fused stepping the fields are double buffered, so the pointers alternate between passes"""
        sp, rp, nb = self.e.halo_info(set_, m, comp)
        out = []
        for ptr in (sp, rp):
            t = self._tensors.get(ptr)
            if t is None:
                t = self._tensors[ptr] = _as_tensor(ptr, nb, self.cuda)
            out.append(t)
        return out

    def _post(self, set_):
        """post the send of my plane and the receive of the neighbour's plane for field set `set_`;
        returns the NCCL/gloo work handles (not waited for)"""
        # E planes travel up (to rank+1), H planes travel down (to rank-1)
        dst = self.rank + 1 if set_ == capi.SET_E else self.rank - 1
        src = self.rank - 1 if set_ == capi.SET_E else self.rank + 1
        if 0 <= dst < self.world:
            self.e.halo_pack(set_)
        ops = []
        for m in range(self.ncf):
            for comp in range(2):
                send, recv = self._planes(set_, m, comp)
                if 0 <= dst < self.world:
                    ops.append(dist.P2POp(dist.isend, send, dst))
                    self.exchanged_bytes += send.numel() * 4
                if 0 <= src < self.world:
                    ops.append(dist.P2POp(dist.irecv, recv, src))
        works = dist.batch_isend_irecv(ops) if ops else []
        return works, (0 <= src < self.world)

    def _finish(self, set_, pending):
        works, has_src = pending
        for w in works:
            w.wait()                      # orders the engine's stream after the transfer (no host sync on CUDA)
        if has_src:
            self.e.halo_unpack(set_)

    def exchange(self, set_):
        if self.world > 1:
            self._finish(set_, self._post(set_))

    def step(self):
        """one full step; with overlap=True the plane a neighbour needs is computed first and travels
        while the rest of the pass runs (the next pass waits for it just before its own send plane)"""
        if self.world == 1:
            self.e.step_E()
            self.e.step_H()
            return
        if not self.overlap:
            self.e.step_E()
            if self.fuse:
                self.e.flush(True)                       # the send plane must exist before it travels
            self.exchange(capi.SET_E)
            self.e.step_H()
            self.exchange(capi.SET_H)
            return
        up, down = self.rank + 1 < self.world, self.rank > 0
        # E pass: its top plane needs the H halo posted in the previous H pass
        if self.fuse:
            self.e.begin_pass(capi.SET_E)                # buffers swap now: halo_info is about the planes this pass writes
        if self.pending_H is not None:
            self._finish(capi.SET_H, self.pending_H)
            self.pending_H = None
        if up:
            self.e.step_part(capi.SET_E, capi.PART_SEND_PLANE)
            pending_E = self._post(capi.SET_E)
            self.e.step_part(capi.SET_E, capi.PART_REST)
        else:
            pending_E = self._post(capi.SET_E)           # only receives from below
            self.e.step_E()
        if self.fuse:
            self.e.flush(True)                           # the deferred bulk of the E pass, fused with H where possible
            self.e.begin_pass(capi.SET_H)
        # H pass: its bottom plane needs the E halo that is in flight
        self._finish(capi.SET_E, pending_E)
        if down:
            self.e.step_part(capi.SET_H, capi.PART_SEND_PLANE)
            self.pending_H = self._post(capi.SET_H)
            self.e.step_part(capi.SET_H, capi.PART_REST)
        else:
            self.pending_H = self._post(capi.SET_H)      # only receives from above
            self.e.step_H()

    def flush(self):
        """complete the halo transfer still in flight (call before reading fields)"""
        if self.pending_H is not None:
            self._finish(capi.SET_H, self.pending_H)
            self.pending_H = None

    def run(self, steps):
        for _ in range(steps):
            self.step()
        self.flush()
