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
