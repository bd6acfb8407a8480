"""Data-parallel plumbing: clips are sharded across ranks (one process per GPU); the only exchange of the path is an
all-gather of the per-rank video / text embeddings so every rank can form its rows of the cross-rank similarity
matrix (reference run/train.py:31-47,126-136 and utils/train_utils.py:51-59).

On CUDA tensors the gather is ONE ncclAllGather of a packed byte buffer issued through the C ABI (hh_allgather) on
the current stream.  CPU tensors (the gloo unit tests of the host logic) are moved with torch.distributed.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Sequence

import torch
import torch.distributed as dist

from . import _lib as L
from . import ops

_comm = None


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


def all_gather_packed(tensors: Sequence[torch.Tensor]) -> List[torch.Tensor]:
    """All-gather several same-shaped-across-ranks tensors with ONE collective.
    Returns, per input, the concatenation over ranks along dim 0 (what AllGather_multi + torch.cat produce)."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return [t.clone() for t in tensors]
    world = dist.get_world_size()
    tensors = [t.contiguous() for t in tensors]
    offs, nbytes = pack_layout(tensors)
    dev = tensors[0].device
    send = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
    for t, o in zip(tensors, offs):
        send[o:o + t.numel() * t.element_size()] = t.reshape(-1).view(torch.uint8)
    recv = torch.empty(world * nbytes, dtype=torch.uint8, device=dev)
    if dev.type == "cuda":
        L.check(L.load().hh_allgather(_nccl_comm(), L.ptr(send), L.ptr(recv), nbytes, L.stream_ptr()), "hh_allgather")
    else:
        dist.all_gather_into_tensor(recv, send)
    recv = recv.view(world, nbytes)
    outs = []
    for t, o in zip(tensors, offs):
        nb = t.numel() * t.element_size()
        g = recv[:, o:o + nb].contiguous().view(t.dtype).view(world * t.shape[0], *t.shape[1:])
        outs.append(g)
    return outs


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
