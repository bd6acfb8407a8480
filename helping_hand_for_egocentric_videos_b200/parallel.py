"""Data-parallel plumbing: clips are sharded across ranks (one process per GPU); the only exchange of the path is an
all-gather of the per-rank video / text embeddings so every rank can form its rows of the cross-rank similarity
matrix (reference run/train.py:31-47,126-136 and utils/train_utils.py:51-59).

On CUDA tensors the gather is ONE ncclAllGather of a packed byte buffer issued through the C ABI (hh_allgather).
CPU tensors (the gloo unit tests of the host logic) are moved with torch.distributed.

  all_gather_packed(tensors)          differentiable (backward = this rank's slice of the gradient, exactly what the
                                      reference's AllGather_multi.backward returns), on the current stream
  all_gather_packed_async(tensors)    the same collective on a side stream; .wait() hands the result to the current
                                      stream -- a step can issue its gather and consume it one step later, so the
                                      collective never sits on the compute stream's critical path
"""
from __future__ import annotations

import ctypes as C
from typing import List, Sequence

import torch
import torch.distributed as dist

from . import _lib as L
from . import ops

_comm = None
_bufs = {}          # (device, nbytes, world, slot) -> (send, recv): allocated once, reused every step
_side_stream = {}   # device -> torch.cuda.Stream


def shard_range(total: int, rank: int, world: int):
    """Contiguous shard [lo, hi) of `total` units for `rank` (first total % world ranks get one extra)."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _nccl_comm():
    """Lazily create this process's NCCL communicator for the library (unique id travels over torch.distributed)."""
    global _comm
    if _comm is not None:
        return _comm
    lib = L.load()
    world, rank = dist.get_world_size(), dist.get_rank()
    buf = (C.c_char * 128)()
    if rank == 0:
        L.check(lib.hh_comm_unique_id(buf), "hh_comm_unique_id")
    obj = [bytes(buf.raw)]
    dist.broadcast_object_list(obj, src=0)
    idbuf = (C.c_char * 128).from_buffer_copy(obj[0])
    comm = C.c_void_p()
    L.check(lib.hh_comm_create(C.byref(comm), world, rank, idbuf), "hh_comm_create")
    _comm = comm
    return comm


def pack_layout(tensors: Sequence[torch.Tensor], align: int = 16):
    """Byte offsets of each tensor inside the packed send buffer (16-byte aligned) and the total size."""
    offs, cur = [], 0
    for t in tensors:
        cur = (cur + align - 1) // align * align
        offs.append(cur)
        cur += t.numel() * t.element_size()
    return offs, (cur + align - 1) // align * align


def _buffers(dev, nbytes, world, slot=0):
    key = (str(dev), nbytes, world, slot)
    b = _bufs.get(key)
    if b is None:
        b = (torch.zeros(nbytes, dtype=torch.uint8, device=dev), torch.empty(world * nbytes, dtype=torch.uint8, device=dev))
        _bufs[key] = b
    return b


def _gather_raw(tensors, slot=0):
    """Pack -> one collective -> unpack, on the current stream.  Returns detached tensors."""
    world = dist.get_world_size()
    tensors = [t.detach().contiguous() for t in tensors]
    offs, nbytes = pack_layout(tensors)
    dev = tensors[0].device
    send, recv = _buffers(dev, nbytes, world, slot)
    for t, o in zip(tensors, offs):
        send[o:o + t.numel() * t.element_size()].copy_(t.reshape(-1).view(torch.uint8))
    if dev.type == "cuda":
        L.check(L.load().hh_allgather(_nccl_comm(), L.ptr(send), L.ptr(recv), nbytes, L.stream_ptr()), "hh_allgather")
    else:
        dist.all_gather_into_tensor(recv, send)
    rv = recv.view(world, nbytes)
    outs = []
    for t, o in zip(tensors, offs):
        nb = t.numel() * t.element_size()
        # .contiguous() copies out of the (reused) receive buffer on the same stream the collective ran on
        outs.append(rv[:, o:o + nb].contiguous().view(t.dtype).view(world * t.shape[0], *t.shape[1:]))
    return outs


class _AllGatherPacked(torch.autograd.Function):
    """Forward: rank-order concatenation of every input over all ranks (one packed collective).  Backward: the slice of
    each output gradient that belongs to this rank's rows -- reference AllGather_multi.backward (run/train.py:42-47);
    like the reference, gradients are NOT summed across ranks here (DDP averages parameter gradients afterwards)."""

    @staticmethod
    def forward(ctx, *tensors):
        ctx.rank = dist.get_rank()
        ctx.rows = [t.shape[0] for t in tensors]
        outs = _gather_raw(tensors)
        nondiff = [o for o, t in zip(outs, tensors) if not t.is_floating_point()]
        if nondiff:
            ctx.mark_non_differentiable(*nondiff)
        return tuple(outs)

    @staticmethod
    def backward(ctx, *grads):
        res = []
        for g, n in zip(grads, ctx.rows):
            res.append(None if g is None else g[ctx.rank * n:(ctx.rank + 1) * n])
        return tuple(res)


def _is_dist():
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def all_gather_packed(tensors: Sequence[torch.Tensor]) -> List[torch.Tensor]:
    """All-gather several same-shaped-across-ranks tensors with ONE collective.
    Returns, per input, the concatenation over ranks along dim 0 (what AllGather_multi + torch.cat produce).
    Differentiable for floating-point inputs (see _AllGatherPacked)."""
    if not _is_dist():
        return [t.clone() for t in tensors]
    return list(_AllGatherPacked.apply(*tensors))


class GatherHandle:
    """Result of all_gather_packed_async: wait() makes the current stream wait for the side-stream collective."""

    def __init__(self, outs, event):
        self._outs, self._event = outs, event

    def wait(self) -> List[torch.Tensor]:
        if self._event is not None:
            torch.cuda.current_stream().wait_event(self._event)
            self._event = None
        return self._outs


def all_gather_packed_async(tensors: Sequence[torch.Tensor], slot: int = 0) -> GatherHandle:
    """The packed all-gather on a side stream (no autograd: inference / metric exchange).  The side stream first waits
    for the work already enqueued on the current stream (the producers of `tensors`); the caller consumes the result
    with handle.wait(), typically one step later.  `slot` selects one of several buffer sets so that a gather may still
    be in flight while the next one is issued (use slot = step % 2)."""
    if not _is_dist():
        return GatherHandle([t.clone() for t in tensors], None)
    dev = tensors[0].device
    if dev.type != "cuda":
        return GatherHandle(_gather_raw(tensors, slot), None)
    side = _side_stream.get(dev)
    if side is None:
        side = _side_stream[dev] = torch.cuda.Stream(device=dev)
    ready = torch.cuda.Event()
    ready.record()
    with torch.cuda.stream(side):
        side.wait_event(ready)
        for t in tensors:
            t.record_stream(side)
        outs = _gather_raw(tensors, slot)
        done = torch.cuda.Event()
        done.record(side)
    for o in outs:
        o.record_stream(torch.cuda.current_stream())
    return GatherHandle(outs, done)


def sharded_sim_operands(text_local: torch.Tensor, video_local: torch.Tensor):
    """Host logic of the sharded similarity: this rank's text rows stay local, the video embeddings of all ranks are
    gathered once (rank order).  Returns (text_local, video_all)."""
    (video_all,) = all_gather_packed([video_local])
    return text_local, video_all


def sharded_sim_matrix(text_local: torch.Tensor, video_local: torch.Tensor) -> torch.Tensor:
    """Rows of sim_matrix(text_all, video_all) owned by this rank: gather the video embeddings once, score the local
    text rows against all of them on the device (EPIC-MIR over 8 GPUs, BASELINE config 4).  CUDA tensors only -- the
    scoring kernel has no CPU fallback (ops.sim_matrix raises)."""
    a, b = sharded_sim_operands(text_local, video_local)
    return ops.sim_matrix(a, b)
