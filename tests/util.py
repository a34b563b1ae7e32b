"""Shared helpers of the parity tests: build a synthetic case, run the CUDA path through the
C-ABI, run the CPU oracle, compare."""
from __future__ import annotations

import numpy as np
import torch

from das_b200 import synth
from das_b200.head import DecodePlan
from oracle import das_oracle as O


MIN_OKS_MARGIN = 1e-6      # |oks - nms_thr| of every suppression decision
MIN_SOFT_GAP_ULPS = 64     # soft NMS: best vs second-best rescored candidate at every pick


def make_case(cfg: synth.HeadConfig, batch, h, w, seed=1234, peaks=16, scales=(1.0, 1.0, 1.0, 1.0), smooth=9,
              identity_rt=False, coherent=0, tc=None, peak_kernel=0, device="cpu"):
    """Seeded synthetic case.  With `tc` (the test_cfg the case will be decoded with) the generator reject-samples
    until every rank / threshold boundary of that decode has a safe margin (synth.enforce_rank_margins)."""
    margin_for = None if tc is None else dict(nms_pre=tc.get("nms_pre", -1), score_thr=tc.get("score_thr", 0.0),
                                              peak_kernel=peak_kernel)
    levels = synth.make_levels(cfg, batch, h, w, seed=seed, peaks=peaks, scales=scales, smooth=smooth,
                               coherent=coherent, margin_for=margin_for, device=device)
    layers = synth.make_layers(cfg, seed=seed + 1)
    metas = synth.make_metas(batch, h, w, stride=cfg.strides[0], seed=seed + 2, identity_rt=identity_rt)
    return dict(cfg=cfg, levels=levels, layers=layers, metas=metas, batch=batch)


def ulp_gap(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """|a-b| in units of float32 ulps (positive finite floats)."""
    ia = a.contiguous().view(torch.int32).to(torch.int64)
    ib = b.contiguous().view(torch.int32).to(torch.int64)
    return (ia - ib).abs()


def rank_margin_ulps(levels, nms_pre, score_thr=0.0, peak_kernel=0):
    """Smallest ulp gap at any boundary that decides the decode's output: adjacent ranks among the candidates that
    survive score_thr (per level incl. the first cell left out, and across levels for the NMS order) and every
    candidate against score_thr (synth.rank_margin_ulps)."""
    return synth.rank_margin_ulps(levels, nms_pre, score_thr, peak_kernel)


def assert_margins(case_or_levels, tc, ref=None, peak_kernel=0):
    """Every parity case must be decidable: rank margins >= 16 ulp and, when the oracle result is given, every OKS
    suppression decision at least MIN_OKS_MARGIN away from nms_thr.  Without this a mismatch could be excused as a
    near-tie -- and a vacuous pass could hide behind one."""
    levels = case_or_levels["levels"] if isinstance(case_or_levels, dict) else case_or_levels
    m = rank_margin_ulps(levels, tc.get("nms_pre", -1), tc.get("score_thr", 0.0), peak_kernel)
    assert m >= synth.MIN_MARGIN_ULPS, f"case has a rank margin of only {m} ulp: regenerate it with margin_for"
    if ref is not None:
        for b, o in enumerate(ref):
            assert o["oks_margin"] >= MIN_OKS_MARGIN, f"image {b}: an OKS decision sits {o['oks_margin']:.2e} from nms_thr"
            if tc.get("nms_type", "hard") != "hard":
                assert o["soft_gap_ulps"] >= MIN_SOFT_GAP_ULPS, f"image {b}: soft-NMS pick decided by {o['soft_gap_ulps']} ulp"
    return m


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
            refine_mode=None, on_demand=None):
    """Decode on the GPU through the C-ABI plan. pose_override: per-level final pose maps for refine=False.
    on_demand: None = the plan's default, else das_plan_set_on_demand_sampling (num_layers > 1)."""
    plan = make_plan(case, test_cfg, refine, peak_kernel, device, refine_mode)
    if on_demand is not None:
        plan.set_on_demand_sampling(on_demand)
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


def run_oracle64(levels, layers, metas, cfg, test_cfg):
    """fp64 ARBITER: the same algorithm with the pose / feature maps and weights in float64 (scores stay fp32, so the
    candidates are the same).  It says how far the fp32 reference itself is from the exact answer -- its noise floor."""
    hc = cfg.as_dict()
    lv64 = [dict(lv, pose_raw=lv["pose_raw"].double(), feats=[f.double() for f in lv["feats"]]) for lv in levels]
    lay64 = [{k: v.double() for k, v in l.items()} for l in layers]
    pp64 = [O.head_eval_tail(lv["pose_raw"], lv["feats"], lay64, lv["scales"], num_joints=hc["num_joints"],
                             num_heads=hc["num_heads"], root_idx=hc["root_idx"], depth_factor=hc["depth_factor"],
                             z_norm=hc["z_norm"], stride=lv["stride"]) for lv in lv64]
    res = O.get_poses([lv["cls"] for lv in levels], [p.float() for p in pp64], [lv["ctr"] for lv in levels], metas, test_cfg,
                      [lv["stride"] for lv in levels], hc["num_joints"])
    for r, meta in zip(res, metas):
        cam = meta["cam"]
        r["poses_cam"], r["poses_world"] = O.backproject(r["poses"].numpy(), cam["K"], cam["R"], cam["t"], hc["root_idx"])
    return res


def run_full_size(cfg, batch, h, w, test_cfg, seed, peaks, chunk=8, refine_mode=None, arbiter=False):
    """A BASELINE-sized case generated on the device (the host generator would take minutes), decoded through the C-ABI,
    and the oracle run on host copies of the same bits, `chunk` images at a time (its dense refinement materialises
    ~0.2 GB of temporaries per image).  Returns (case, plan, gpu results, oracle results)."""
    case = make_case(cfg, batch, h, w, seed=seed, peaks=peaks, tc=test_cfg, device="cuda")
    plan, got = run_gpu(case, test_cfg, refine=True, refine_mode=refine_mode)
    ref, ref64 = [], []
    for b0 in range(0, batch, chunk):
        b1 = min(batch, b0 + chunk)
        sub = [dict(lv, cls=lv["cls"][b0:b1].cpu(), ctr=lv["ctr"][b0:b1].cpu(), pose_raw=lv["pose_raw"][b0:b1].cpu(),
                    feats=[f[b0:b1].cpu() for f in lv["feats"]]) for lv in case["levels"]]
        r, _ = O.decode_full(sub, case["layers"], case["metas"][b0:b1], cfg.as_dict(), test_cfg)
        ref += r
        if arbiter:
            ref64 += run_oracle64(sub, case["layers"], case["metas"][b0:b1], cfg, test_cfg)
    if arbiter:
        return case, plan, got, ref, ref64
    return case, plan, got, ref


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
