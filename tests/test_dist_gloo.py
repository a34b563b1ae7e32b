"""CPU, world_size 2 (gloo): the multi-GPU host logic -- contiguous batch shards, one all-gather of the
packed output blocks, merge in rank order -- reproduces the unsharded result."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from das_b200 import dist as ddist
from das_b200 import head as H
from das_b200 import synth
from oracle import das_oracle as O

TC = dict(nms_pre=6, nms_post=6, nms_thr=0.9, score_thr=0.0)
CFG = synth.PANOPTIC
B, HH, WW = 4, 12, 16


def _case():
    levels = synth.make_levels(CFG, B, HH, WW, seed=77, peaks=6)
    layers = synth.make_layers(CFG, seed=78)
    metas = synth.make_metas(B, HH, WW, seed=79)
    return levels, layers, metas


def _pack(results, nb, P, J):
    """Write oracle results into the plan's packed output-block layout (what a rank's GPU would hold)."""
    lay, total = H.block_layout(nb, P, J)
    block = torch.zeros(total, dtype=torch.uint8)
    v = H.block_views(block, nb, P, J)
    for b, r in enumerate(results):
        n = len(r["scores"])
        v["out_count"][b] = n
        v["out_score"][b, :n] = r["scores"]
        v["out_pose"][b, :n] = r["poses"]
        v["out_center"][b, :n] = r["centers"]
        v["out_cam"][b, :n] = torch.from_numpy(r["poses_cam"])
        v["out_world"][b, :n] = torch.from_numpy(r["poses_world"])
    return block


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    levels, layers, metas = _case()
    lo, hi = ddist.shard_bounds(B, world, rank)
    shard = [dict(lv, cls=lv["cls"][lo:hi], ctr=lv["ctr"][lo:hi], pose_raw=lv["pose_raw"][lo:hi],
                  feats=[f[lo:hi] for f in lv["feats"]]) for lv in levels]
    res, _ = O.decode_full(shard, layers, ddist.shard_list(metas, world, rank), CFG.as_dict(), TC)
    block = _pack(res, hi - lo, TC["nms_post"], CFG.num_joints)
    gathered = ddist.gather_blocks(block, world)
    if rank == 0:
        out = []
        for r in range(world):
            v = H.block_views(gathered[r], hi - lo, TC["nms_post"], CFG.num_joints)
            out.append({k: t.clone().numpy() for k, t in v.items()})
        q.put(out)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shard_gather_equals_unsharded():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    levels, layers, metas = _case()
    full, _ = O.decode_full(levels, layers, metas, CFG.as_dict(), TC)
    b = 0
    for rank_views in out:
        for i in range(rank_views["out_count"].shape[0]):
            n = int(rank_views["out_count"][i])
            assert n == len(full[b]["scores"])
            # the oracle's own convs sum in a batch/thread-count dependent order: last-bit differences only
            np.testing.assert_allclose(rank_views["out_pose"][i, :n], full[b]["poses"].numpy(), rtol=1e-4, atol=1e-3)
            np.testing.assert_allclose(rank_views["out_cam"][i, :n], full[b]["poses_cam"], rtol=1e-4, atol=1e-2)
            np.testing.assert_allclose(rank_views["out_score"][i, :n], np.asarray(full[b]["scores_list"], np.float32), rtol=1e-6)
            b += 1
    assert b == B


class _FakePlan:
    """Host stand-in for DecodePlan in the collection logic: same block layout / views / result construction."""

    def __init__(self, batch, P, J):
        self.batch, self.out_slots, self.J = batch, P, J

    def views_of_block(self, block):
        return H.block_views(block, self.batch, self.out_slots, self.J)

    def results(self, metas, src):
        return H.DecodePlan.results(self, metas, src=src)


def _worker_uneven(rank, world, port, q, total):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    levels = synth.make_levels(CFG, total, HH, WW, seed=91, peaks=6)
    layers = synth.make_layers(CFG, seed=92)
    metas = synth.make_metas(total, HH, WW, seed=93)
    lo, hi = ddist.shard_bounds(total, world, rank)
    per = ddist.padded_shard_size(total, world)
    idx = ddist.pad_shard(list(range(lo, hi)), per)              # the short rank repeats its last image
    shard = [dict(lv, cls=lv["cls"][idx], ctr=lv["ctr"][idx], pose_raw=lv["pose_raw"][idx],
                  feats=[f[idx] for f in lv["feats"]]) for lv in levels]
    res, _ = O.decode_full(shard, layers, [metas[i] for i in idx], CFG.as_dict(), TC)
    block = _pack(res, per, TC["nms_post"], CFG.num_joints)
    plan = _FakePlan(per, TC["nms_post"], CFG.num_joints)
    all_metas = [[metas[i] for i in range(*ddist.shard_bounds(total, world, r))] for r in range(world)]
    merged = ddist.collect_results(plan, block, [metas[i] for i in idx], world, rank, total, all_metas=all_metas,
                                   n_real=hi - lo, order="contiguous")
    # blocks of different size must raise on every rank instead of hanging the collective
    raised = False
    try:
        ddist.gather_blocks(block[: block.numel() - 256 * rank], world)
    except ValueError:
        raised = True
    if rank == 0:
        q.put((raised, [(r["image_paths"], r["scores"], r["poses"].numpy()) for r in merged]))
    else:
        assert raised and merged is None
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_with_a_batch_that_does_not_divide():
    """B % world != 0 (the last partial batch of a dataset): the short rank pads its plan to ceil(B / world) images, one
    all-gather moves equally sized blocks, the padding rows are dropped (unequal blocks hung the collective)."""
    total = 5
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_uneven, args=(r, 2, port, q, total)) for r in range(2)]
    for p in procs:
        p.start()
    raised, merged = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert raised and len(merged) == total
    levels = synth.make_levels(CFG, total, HH, WW, seed=91, peaks=6)
    full, _ = O.decode_full(levels, synth.make_layers(CFG, seed=92), synth.make_metas(total, HH, WW, seed=93), CFG.as_dict(), TC)
    for b, ((paths, scores, poses), o) in enumerate(zip(merged, full)):
        assert paths == o["image_paths"] == [f"synthetic_{b:05d}.jpg"]
        np.testing.assert_allclose(np.asarray(scores, np.float32), np.asarray(o["scores_list"], np.float32), rtol=1e-6)
        np.testing.assert_allclose(poses, o["poses"].numpy(), rtol=1e-4, atol=1e-3)
