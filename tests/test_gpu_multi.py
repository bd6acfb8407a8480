"""The path's one collective on real GPUs (needs >= 2 CUDA devices; one process per GPU over NCCL, like torchrun):
the packed all-gather through hh_allgather, its side-stream form, its backward, and the sharded EPIC-MIR similarity
(reference run/train.py:31-47,126-136, run/test_epic.py:262) against single-process results at MIR size
(N_v = N_t = 9728).

  gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -q
"""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

MIR_ROWS = 9728


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _world():
    n = torch.cuda.device_count() if torch.cuda.is_available() else 0
    return 8 if n >= 8 else 4 if n >= 4 else 2 if n >= 2 else 0


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from helping_hand_for_egocentric_videos_b200 import ops, parallel
        res = {}
        # ---- 1. packed gather of mixed dtypes / widths == the full tensors every rank can regenerate from the seed
        g = torch.Generator().manual_seed(5)
        per = 7
        vid_all = torch.randn(world * per, 256, generator=g)
        txt_all = torch.randn(world * per, 256, generator=g)
        tok_all = torch.randint(0, 49408, (world * per, 77), generator=g)
        lo, hi = parallel.shard_range(world * per, rank, world)
        mine = [vid_all[lo:hi].to(dev), tok_all[lo:hi].to(dev), txt_all[lo:hi, :3].contiguous().to(dev)]
        outs = parallel.all_gather_packed(mine)
        res["gather"] = bool(torch.equal(outs[0].cpu(), vid_all) and torch.equal(outs[1].cpu(), tok_all)
                             and torch.equal(outs[2].cpu(), txt_all[:, :3]))
        # ---- 2. the side-stream form, two gathers in flight on alternating buffer sets
        h0 = parallel.all_gather_packed_async([mine[0]], slot=0)
        h1 = parallel.all_gather_packed_async([mine[0] * 2], slot=1)
        (a0,), (a1,) = h0.wait(), h1.wait()
        torch.cuda.synchronize()
        res["async"] = bool(torch.equal(a0.cpu(), vid_all) and torch.equal(a1.cpu(), vid_all * 2))
        # ---- 3. backward = this rank's slice of the output gradient (AllGather_multi.backward)
        v = mine[0].clone().requires_grad_(True)
        (allv,) = parallel.all_gather_packed([v])
        wgt = torch.randn(world * per, 256, generator=torch.Generator().manual_seed(3)).to(dev)
        (allv * wgt).sum().backward()
        res["grad"] = bool(torch.equal(v.grad, wgt[lo:hi]))
        # ---- 4. EPIC-MIR: each rank holds 9728 / world text rows and video rows; rows of the full similarity matrix
        gm = torch.Generator().manual_seed(7)
        t_all = torch.randn(MIR_ROWS, 256, generator=gm)
        v_all = torch.randn(MIR_ROWS, 256, generator=gm)
        lo, hi = parallel.shard_range(MIR_ROWS, rank, world)
        sim_rows = parallel.sharded_sim_matrix(t_all[lo:hi].to(dev), v_all[lo:hi].to(dev))
        full = ops.sim_matrix(t_all.to(dev), v_all.to(dev))            # single-process statement on this GPU
        res["mir_shape"] = tuple(sim_rows.shape) == (hi - lo, MIR_ROWS)
        res["mir_rows_bit_equal"] = bool(torch.equal(sim_rows, full[lo:hi]))
        want = torch.nn.functional.normalize(t_all[lo:hi].double(), dim=-1) @ torch.nn.functional.normalize(v_all.double(), dim=-1).t()
        res["mir_vs_fp64"] = float((sim_rows.double().cpu() - want).abs().max())
        top2 = want.topk(2, -1).values
        decided = (top2[:, 0] - top2[:, 1]) > 1e-5                       # rows whose fp64 winner is not a near tie
        res["mir_argmax_equal"] = bool(torch.equal(ops.row_argmax(sim_rows).cpu()[decided], want.argmax(-1)[decided]))
        res["mir_undecided_rows"] = int((~decided).sum())
        ret[rank] = res
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(_world() < 2, reason="needs >= 2 CUDA devices")
def test_collective_and_sharded_similarity_on_gpus():
    world = _world()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    out = dict(ret)
    assert sorted(out) == list(range(world))
    for r, res in out.items():
        assert res["gather"] and res["async"] and res["grad"], (r, res)
        assert res["mir_shape"] and res["mir_rows_bit_equal"], (r, res)
        assert res["mir_vs_fp64"] <= 1e-5, (r, res)
        assert res["mir_argmax_equal"], (r, res)
