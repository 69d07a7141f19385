"""z-slab domain decomposition over the GPUs of one box.

The reference is single-GPU; this is new work whose oracle is "N slabs == one domain, bit for bit".
One process per GPU (torchrun); rank r owns global planes [z_begin, z_end) plus one halo plane per
interior face.  Per time step each rank sends, per face, the 5 populations that cross it
(c_z = +1: q in {5,11,13,15,16} upward, c_z = -1: q in {6,12,14,17,18} downward) of its outermost
owned plane -- one contiguous staging buffer per face, packed by the library right after the face
planes were computed so the transfer overlaps the interior update -- with
torch.distributed P2P ops (NCCL over NVLink on GPUs; gloo on CPU for the host-logic tests).

Nothing here computes: the kernels live in liblbm_b200.so.  `exchange_halos` and `slab_ranges` /
`compact_offsets` are backend-agnostic so the same code runs under gloo in the CPU tests.
"""
from __future__ import annotations


import numpy as np

from . import api


def slab_ranges(nz: int, world: int):
    """contiguous, near-equal z ranges; every slab gets at least one plane"""
    if world > nz:
        raise ValueError("more slabs than planes")
    base, rem = divmod(nz, world)
    out, z = [], 0
    for r in range(world):
        n = base + (1 if r < rem else 0)
        out.append((z, z + n))
        z += n
    return out


def compact_offsets(local_counts):
    """exclusive running sum of the per-slab stored-node counts: the single-domain z,y,x numbering
    of index_transform (bifurcation.cu:241-252) continues across slabs"""
    offs, run = [], 0
    for c in local_counts:
        offs.append(run)
        run += int(c)
    return offs, run


def exchange_halos(send_lo, recv_lo, send_hi, recv_hi, rank: int, world: int, group=None):
    """Post the (up to four) point-to-point transfers of one step and return the work handles.

    send_lo -> rank-1's recv_hi,  send_hi -> rank+1's recv_lo.  Any of the tensors may be None at
    the outer faces of the domain.  Works for CUDA tensors (NCCL) and CPU tensors (gloo)."""
    import torch.distributed as dist

    ops = []
    if rank > 0 and send_lo is not None:
        ops.append(dist.P2POp(dist.isend, send_lo, rank - 1, group))
        ops.append(dist.P2POp(dist.irecv, recv_lo, rank - 1, group))
    if rank < world - 1 and send_hi is not None:
        ops.append(dist.P2POp(dist.isend, send_hi, rank + 1, group))
        ops.append(dist.P2POp(dist.irecv, recv_hi, rank + 1, group))
    if not ops:
        return []
    return dist.batch_isend_irecv(ops)


class _DevBuf:
    """zero-copy view of a library-owned device buffer for torch (CUDA array interface v2)"""

    def __init__(self, ptr: int, nbytes: int, dtype: np.dtype):
        self.__cuda_array_interface__ = {
            "shape": (nbytes // dtype.itemsize,), "typestr": dtype.str, "data": (ptr, False), "version": 2,
            "strides": None,
        }


class SlabCase(api.Case):
    """A `Case` that owns one z-slab of a larger domain and steps in lock-step with its neighbours."""

    def __init__(self, desc: api.CaseDesc, group=None):
        import torch.distributed as dist

        super().__init__(desc)
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self._bufs = None
        self._p2p = False
        self._tick = None
        self._mapped = []        # IPC handles this rank holds a mapping of
        self.p2p_error = None    # why enable_p2p fell back, if it did

    def setup(self, flag=None, bc_planes=None, flag_slab=None):
        """geo_pre -> (all-gather of stored counts) -> index_transform -> read_vel -> initialize"""
        import torch
        import torch.distributed as dist

        if flag is not None:
            self.set_flag(flag)
        if flag_slab is not None:  # only the planes this rank needs: (uint8 array, z_first)
            self.set_flag_slab(*flag_slab)
        self.geo_pre()
        mine = torch.tensor([self.local_stored_count()], dtype=torch.int64, device="cuda")
        allc = [torch.zeros_like(mine) for _ in range(self.world)]
        dist.all_gather(allc, mine, group=self.group)
        offs, total = compact_offsets([int(t.item()) for t in allc])
        self.set_compact_offset(offs[self.rank], total)
        self.index_transform()
        if bc_planes is not None:
            self.set_bc_planes(*bc_planes)
        self.initialize()
        self._wrap_buffers()

    def enable_p2p(self):
        """Fused halo exchange: every rank maps its neighbours' population buffers (CUDA IPC over
        NVLink) and the step kernel stores the crossing populations there itself; what is left of the
        transport is one tiny all-reduce per step that keeps the slabs in lock-step."""
        import torch
        import torch.distributed as dist

        mine = self.p2p_export()
        mine.pop("ptrs")  # raw pointers mean nothing in another process
        everyone = [None] * self.world
        dist.all_gather_object(everyone, mine, group=self.group)
        ok, attached = 1, []
        try:
            for side, nb in ((0, self.rank - 1), (1, self.rank + 1)):
                if 0 <= nb < self.world:
                    ptrs = []
                    for h, o in zip(everyone[nb]["handles"], everyone[nb]["boff"]):
                        ptrs.append(api.p2p_open(h) + o)
                        self._mapped.append(h)
                    pa, pb = ptrs
                    # my low face feeds the neighbour's HIGH halo plane and vice versa
                    self.p2p_attach(side, pa, pb, everyone[nb]["qs"], everyone[nb]["halo_c0"][1 - side],
                                    everyone[nb]["face_c0"][1 - side])
                    attached.append(side)
        except api.LbmError as e:
            ok, self.p2p_error = 0, str(e)
        # all ranks or none: a rank that cannot map its neighbour (no peer access / IPC) sends everyone
        # back to the pack + NCCL path
        flag = torch.tensor([ok], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        if int(flag.item()) == 0:
            for side in attached:
                self.p2p_attach(side, None, None)
            self._release_mappings()
            return False
        self._tick = torch.zeros(1, device="cuda")
        self._p2p = True
        return True

    def _release_mappings(self):
        for h in self._mapped:
            api.p2p_release(h)
        self._mapped = []

    def close(self):
        """collective when peer mappings exist: every rank unmaps its neighbours' buffers before anyone
        frees them"""
        if self._mapped:
            import torch.distributed as dist

            self.sync()
            self._release_mappings()
            if dist.is_initialized():
                dist.barrier(group=self.group)
        self._p2p = False
        super().close()

    def __del__(self):  # never collective: a garbage-collected case only drops its own mappings
        try:
            self._release_mappings()
            api.Case.close(self)
        except Exception:
            pass

    def _wrap_buffers(self):
        import torch

        self._ext = torch.cuda.ExternalStream(self.stream)
        bufs = []
        for side in (0, 1):
            s, r, ns, nr = self.halo_buffers(side)
            if not s:
                bufs.append((None, None))
            else:  # an empty plane (sparse storage) still takes part in the exchange with 0-size-safe tensors
                bufs.append((torch.as_tensor(_DevBuf(s, max(ns, self.dtype.itemsize), self.dtype), device="cuda")[: ns // self.dtype.itemsize],
                             torch.as_tensor(_DevBuf(r, max(nr, self.dtype.itemsize), self.dtype), device="cuda")[: nr // self.dtype.itemsize]))
        self._bufs = bufs

    def _one_step(self, flags=0):
        import torch

        if self._p2p:
            import torch.distributed as dist

            self.step_begin(flags)  # face planes first: their peer stores start crossing NVLink ...
            self.step_interior()    # ... while the interior planes are updated
            self.step_end()
            with torch.cuda.stream(self._ext):  # in-stream barrier: nobody starts t+1 before all finished t
                dist.all_reduce(self._tick, group=self.group)
            return
        self.step_begin(flags)  # face planes + pack, queued on the library's stream
        (s_lo, r_lo), (s_hi, r_hi) = self._bufs
        # torch.distributed orders NCCL's stream after what `self._ext` holds so far (faces + pack)
        with torch.cuda.stream(self._ext):
            works = exchange_halos(s_lo, r_lo, s_hi, r_hi, self.rank, self.world, self.group)
        self.step_interior()  # interior planes run while the faces travel over NVLink
        with torch.cuda.stream(self._ext):
            for w in works:
                w.wait()  # stream-ordered wait, no host sync
        self.step_end()  # unpack + buffer swap

    def step(self, n: int = 1):
        for i in range(int(n)):
            self._one_step(api.STEP_MOMENTS if i == n - 1 else 0)
        self.sync()

    def step_timed(self, n: int) -> float:
        import torch

        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(self._ext):
            e0.record()
        for i in range(int(n)):
            self._one_step(api.STEP_MOMENTS if i == n - 1 else 0)
        with torch.cuda.stream(self._ext):
            e1.record()
        e1.synchronize()
        return e0.elapsed_time(e1)
