"""CPU, world_size 2 over gloo: the host logic of the data-parallel exchange (shard ranges, packed all-gather layout,
rank-order concatenation, sharded similarity rows) against single-process results."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helping_hand_for_egocentric_videos_b200 import parallel


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(5)
        vid_all = torch.randn(world * 6, 256, generator=g)
        txt_all = torch.randn(world * 6, 256, generator=g)
        tok_all = torch.randint(0, 1000, (world * 6, 77), generator=g)
        lo, hi = parallel.shard_range(world * 6, rank, world)
        outs = parallel.all_gather_packed([vid_all[lo:hi], tok_all[lo:hi], txt_all[lo:hi, :3]])
        ok = torch.equal(outs[0], vid_all) and torch.equal(outs[1], tok_all) and torch.equal(outs[2], txt_all[:, :3])
        # sharded similarity: the host logic hands the scoring kernel (CUDA only) this rank's text rows and ALL videos
        t_loc, v_all = parallel.sharded_sim_operands(txt_all[lo:hi], vid_all[lo:hi])
        ok = ok and torch.equal(t_loc, txt_all[lo:hi]) and torch.equal(v_all, vid_all)
        try:
            parallel.sharded_sim_matrix(txt_all[lo:hi], vid_all[lo:hi])      # CPU tensors: must refuse, not fall back
            ok = False
        except RuntimeError:
            pass
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_packed_allgather_and_sharded_similarity_world2():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}


class _RefAllGather(torch.autograd.Function):
    """The reference's AllGather_multi (run/train.py:31-47), restated for the comparison."""

    @staticmethod
    def forward(ctx, tensor, rank, world):
        output = [torch.empty_like(tensor) for _ in range(world)]
        dist.all_gather(output, tensor)
        ctx.rank, ctx.batch_size = rank, tensor.shape[0]
        return torch.cat(output, 0)

    @staticmethod
    def backward(ctx, grad_output):
        return grad_output[ctx.batch_size * ctx.rank: ctx.batch_size * (ctx.rank + 1)], None, None


def _grad_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(11 + rank)
        vid = torch.randn(5, 16, generator=g)
        txt = torch.randn(5, 16, generator=g)
        tok = torch.randint(0, 99, (5, 7), generator=g)
        wgt = torch.randn(world * 5, world * 5, generator=torch.Generator().manual_seed(3))   # same on every rank

        def loss_of(gather):
            v = vid.clone().requires_grad_(True)
            t = txt.clone().requires_grad_(True)
            va, ta, ka = gather(v, t, tok)
            loss = ((ta @ va.t()) * wgt).sum() + (va * va).sum() * (1 + ka.float().mean())
            loss.backward()
            return loss.detach(), v.grad, t.grad, ka

        ours = loss_of(lambda v, t, k: parallel.all_gather_packed([v, t, k]))
        ref = loss_of(lambda v, t, k: (_RefAllGather.apply(v, rank, world), _RefAllGather.apply(t, rank, world),
                                       _RefAllGather.apply(k, rank, world)))
        ok = torch.allclose(ours[0], ref[0]) and ours[1] is not None and ours[2] is not None
        ok = ok and torch.allclose(ours[1], ref[1]) and torch.allclose(ours[2], ref[2]) and torch.equal(ours[3], ref[3])
        ok = ok and not ours[3].requires_grad
        # the side-stream form is the same exchange (CPU tensors: executed inline)
        h = parallel.all_gather_packed_async([vid, tok], slot=1)
        va, ka = h.wait()
        ok = ok and torch.equal(va, ours_cat(vid, world)) and torch.equal(ka, ours_cat(tok, world))
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def ours_cat(t, world):
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    return torch.cat(out, 0)


def test_allgather_backward_matches_reference_world2():
    """the gather must carry gradients exactly like AllGather_multi (local slice of the output gradient)."""
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_grad_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}


def test_shard_range_covers_everything_once():
    for total in (0, 1, 7, 64, 9668):
        for world in (1, 2, 3, 8):
            spans = [parallel.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_pack_layout_alignment():
    ts = [torch.zeros(3, 5), torch.zeros(7, dtype=torch.int64), torch.zeros(2, 2, dtype=torch.bfloat16)]
    offs, total = parallel.pack_layout(ts)
    assert offs == [0, 64, 128] and total == 144
    assert all(o % 16 == 0 for o in offs) and total % 16 == 0


def test_single_process_gather_is_identity():
    x = torch.arange(12.).view(3, 4)
    (y,) = parallel.all_gather_packed([x])
    assert torch.equal(x, y)
