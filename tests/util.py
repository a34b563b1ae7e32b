"""Shared helpers of the parity tests: build a synthetic case, run the CUDA path through the
C-ABI, run the CPU oracle, compare."""
from __future__ import annotations

import numpy as np
import torch

from das_b200 import synth
from das_b200.head import DecodePlan
from oracle import das_oracle as O


def make_case(cfg: synth.HeadConfig, batch, h, w, seed=1234, peaks=16, scales=(1.0, 1.0, 1.0, 1.0), smooth=9,
              identity_rt=False, coherent=0):
    levels = synth.make_levels(cfg, batch, h, w, seed=seed, peaks=peaks, scales=scales, smooth=smooth,
                               coherent=coherent)
    layers = synth.make_layers(cfg, seed=seed + 1)
    metas = synth.make_metas(batch, h, w, stride=cfg.strides[0], seed=seed + 2, identity_rt=identity_rt)
    return dict(cfg=cfg, levels=levels, layers=layers, metas=metas, batch=batch)


def score_maps(levels):
    return [(lv["cls"].sigmoid() * lv["ctr"].sigmoid()).flatten(1) for lv in levels]


def ulp_gap(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """|a-b| in units of float32 ulps (positive finite floats)."""
    ia = a.contiguous().view(torch.int32).to(torch.int64)
    ib = b.contiguous().view(torch.int32).to(torch.int64)
    return (ia - ib).abs()


def rank_margin_ulps(levels, nms_pre, score_thr=0.0):
    """Smallest ulp gap at any decision boundary of the ranking: adjacent scores among the first
    nms_pre+1 ranks of every (image, level), and every score against score_thr."""
    worst = 1 << 40
    for sm in score_maps(levels):
        hw = sm.shape[1]
        if nms_pre > 0 and hw > nms_pre:
            top = sm.topk(min(nms_pre + 1, hw), dim=1)[0]
        else:
            top = sm.sort(dim=1, descending=True)[0]
        if top.shape[1] > 1:
            worst = min(worst, int(ulp_gap(top[:, :-1], top[:, 1:]).min()))
        if score_thr > 0:
            worst = min(worst, int(ulp_gap(top, torch.full_like(top, score_thr)).min()))
    return worst


def run_oracle(case, test_cfg, stable=False, peak_kernel=0):
    cfg = case["cfg"]
    return O.decode_full(case["levels"], case["layers"], case["metas"], cfg.as_dict(), test_cfg,
                         stable=stable, peak_kernel=peak_kernel)


def make_plan(case, test_cfg, refine=True, peak_kernel=0, device="cuda", refine_mode=None):
    cfg = case["cfg"]
    sizes = [tuple(lv["cls"].shape[-2:]) for lv in case["levels"]]
    plan = DecodePlan(num_joints=cfg.num_joints, root_idx=cfg.root_idx, depth_factor=cfg.depth_factor,
                      z_norm=cfg.z_norm, strides=cfg.strides, level_sizes=sizes, batch=case["batch"],
                      test_cfg=test_cfg, num_heads=cfg.num_heads, feat_channels=cfg.feat_channels,
                      num_layers=cfg.num_layers, refine=refine, peak_kernel=peak_kernel, device=device,
                      refine_mode=refine_mode)
    if refine:
        plan.set_weights(synth.layers_to(case["layers"], device))
    return plan


def run_gpu(case, test_cfg, refine=True, peak_kernel=0, pose_override=None, use_graph=True, device="cuda",
            refine_mode=None):
    """Decode on the GPU through the C-ABI plan. pose_override: per-level final pose maps for refine=False."""
    plan = make_plan(case, test_cfg, refine, peak_kernel, device, refine_mode)
    dl = synth.levels_to(case["levels"], device)
    levels = []
    for l, lv in enumerate(dl):
        pose = lv["pose_raw"] if pose_override is None else pose_override[l].to(device)
        levels.append(dict(cls=lv["cls"], ctr=lv["ctr"], pose=pose, feats=lv["feats"], scales=lv["scales"]))
    plan.bind(levels)
    plan.set_metas(case["metas"])
    plan.run(use_graph=use_graph)
    torch.cuda.synchronize()
    return plan, plan.results(case["metas"])


def slot_to_level_index(plan, slots_b, cand_index_b):
    """Map candidate slots of one image to (level, cell index)."""
    lib = plan.lib
    bounds = []
    s0 = 0
    for (h, w) in plan.level_sizes:
        n = lib.das_level_slots(h, w, plan.cfg.nms_pre)
        bounds.append((s0, s0 + n))
        s0 += n
    lv = []
    for s in slots_b.tolist():
        for l, (a, b) in enumerate(bounds):
            if a <= s < b:
                lv.append(l)
                break
    idx = cand_index_b[slots_b.long()].tolist() if len(slots_b) else []
    return lv, idx


def rel_err(a, b, floor=1.0):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if a.size == 0:
        return 0.0
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor)))
