"""Generate tests/golden/*.npz from the REFERENCE's own code.  TEST INFRASTRUCTURE ONLY.

Runs in the build container only (needs /root/reference).  For every case it
  1. builds the seeded synthetic head outputs (das_b200/synth.py),
  2. runs the reference functions extracted verbatim by oracle/ref_extract.py,
  3. asserts the restatement in oracle/das_oracle.py reproduces them BIT-EXACTLY,
  4. stores the reference outputs, plus input checksums so a drifting generator is detected.

    python -m oracle.make_golden
"""
from __future__ import annotations

import dataclasses
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from das_b200 import synth  # noqa: E402
from oracle import das_oracle as O  # noqa: E402
from oracle import ref_extract as R  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

P = synth.PANOPTIC
SHIPPED_TEST_CFG = dict(nms_across_levels=False, nms_pre=1000, nms_post=100, nms_thr=0.9, score_thr=0.07)   # exp_panoptic.py:47-53
CASES = {
    # name: (head cfg, batch, H, W, seed, peaks, scales, test_cfg)
    "panoptic_small": (P, 2, 24, 40, 1234, 12, (1.0, 1.0, 1.0, 1.0),
                       dict(nms_pre=10, nms_post=10, nms_thr=0.9, score_thr=0.0)),
    "panoptic_cfg1_128x208": (P, 1, 128, 208, 1235, 16, (1.0, 1.0, 1.0, 1.0),
                              dict(nms_pre=10, nms_post=10, nms_thr=0.9, score_thr=0.0)),
    "panoptic_scales_thr": (P, 2, 32, 48, 1236, 16, (1.1, 0.9, 1.05, 0.95),
                            dict(nms_pre=40, nms_post=20, nms_thr=0.9, score_thr=0.07)),
    "pyramid4": (dataclasses.replace(P, strides=(8, 16, 32, 64)), 2, 32, 48, 1237, 16, (1.0, 1.0, 1.0, 1.0),
                 dict(nms_pre=100, nms_post=30, nms_thr=0.9, score_thr=0.05)),
    "mupots17_L1": (dataclasses.replace(synth.MUPOTS17, num_layers=1), 2, 24, 32, 1238, 16, (1.0, 1.0, 1.0, 1.0),
                    dict(nms_pre=20, nms_post=20, nms_thr=0.9, score_thr=0.0)),
    "mupots17_L3": (synth.MUPOTS17, 1, 24, 32, 1239, 16, (1.0, 1.0, 1.0, 1.0),
                    dict(nms_pre=20, nms_post=20, nms_thr=0.9, score_thr=0.0)),
    "coherent_nms": (P, 2, 32, 48, 1241, 12, (1.0, 1.0, 1.0, 1.0),
                     dict(nms_pre=30, nms_post=30, nms_thr=0.9, score_thr=0.0), dict(coherent=8)),
    "soft_nms": (P, 2, 32, 48, 1242, 12, (1.0, 1.0, 1.0, 1.0),
                 dict(nms_pre=30, nms_post=12, nms_thr=0.9, score_thr=0.0, nms_type="soft"), dict(coherent=8)),
    "reference_cfg_nms1000": (P, 1, 40, 72, 1240, 24, (1.0, 1.0, 1.0, 1.0),
                              dict(nms_across_levels=False, nms_pre=1000, nms_post=100, nms_thr=0.9, score_thr=0.07)),
    # SURVEY 8(c) edge cases: nms_post absent -> the NMS block is skipped (das_head.py:770-772); nothing above score_thr
    # (N == 0, :772); uv offsets scaled 6x so most sampling targets fall outside the map (zero padding, recursive_update.py:25)
    "no_nms_post": (P, 2, 24, 40, 1243, 12, (1.0, 1.0, 1.0, 1.0), dict(nms_pre=12, nms_thr=0.9, score_thr=0.02)),
    "none_above_thr": (P, 2, 24, 40, 1244, 12, (1.0, 1.0, 1.0, 1.0),
                       dict(nms_pre=10, nms_post=10, nms_thr=0.9, score_thr=0.9999)),
    "border_targets": (P, 2, 16, 20, 1245, 10, (1.0, 1.0, 6.0, 1.0),
                       dict(nms_pre=10, nms_post=10, nms_thr=0.9, score_thr=0.0)),
    # the two SHIPPED models with their shipped test_cfg, on the 4-level pyramid of their shipped test pipelines:
    # Panoptic J=15, L=1, 1152x640 input -> 80x144 stride-8 map (exp_panoptic.py:31-53,138-155); MuPoTS J=21, L=2, root 14,
    # 2048^2 frames at (1280, 768) -> 96x96 (exp_mupots.py:7,37-56,142-159)
    "shipped_panoptic_80x144": (dataclasses.replace(P, strides=(8, 16, 32, 64)), 1, 80, 144, 1246, 24, (1.0, 1.0, 1.0, 1.0),
                                SHIPPED_TEST_CFG),
    "shipped_mupots21_L2_96x96": (dataclasses.replace(synth.MUPOTS17, num_joints=21, num_layers=2, strides=(8, 16, 32, 64)),
                                  1, 96, 96, 1247, 24, (1.0, 1.0, 1.0, 1.0), SHIPPED_TEST_CFG),
}


def checksum(levels):
    return np.array([[float(lv[k].double().sum()) for k in ("cls", "ctr", "pose_raw")] +
                     [float(f.double().sum()) for f in lv["feats"]] for lv in levels], dtype=np.float64)


def build_case(name):
    cfg, b, h, w, seed, peaks, scales, tc = CASES[name][:8]
    extra = CASES[name][8] if len(CASES[name]) > 8 else {}
    # reject-sampled so that every rank / score_thr boundary of THIS decode has a safe margin
    margin_for = dict(nms_pre=tc.get("nms_pre", -1), score_thr=tc.get("score_thr", 0.0))
    levels = synth.make_levels(cfg, b, h, w, seed=seed, peaks=peaks, scales=scales, margin_for=margin_for, **extra)
    layers = synth.make_layers(cfg, seed=seed + 1)
    metas = synth.make_metas(b, h, w, stride=cfg.strides[0], seed=seed + 2)
    return cfg, levels, layers, metas, tc


def run_reference(cfg, levels, layers, metas, tc):
    hc = cfg.as_dict()
    pose_preds = [R.refined_pose_pred(lv, layers, hc) for lv in levels]
    head = R.head(cfg.num_joints, cfg.strides, tc)
    # the reference's get_poses writes into its pose_pred argument on pass-through levels (HW <= nms_pre:
    # das_head.py:714 is a view and :725/:732 assign into it), so hand it a copy
    res = head.get_poses([lv["cls"] for lv in levels], [p.clone() for p in pose_preds],
                         [lv["ctr"] for lv in levels], metas)
    for r, m in zip(res, metas):
        r["poses_cam"], r["poses_world"] = R.backproject(r["poses"].numpy(), m["cam"], cfg.root_idx)
    return res, pose_preds


def main(only=()):
    """python -m oracle.make_golden [case ...]: regenerate every golden file, or only the named ones."""
    assert R.available(), "the reference tree is required to (re)generate golden vectors"
    os.makedirs(OUT, exist_ok=True)
    unknown = set(only) - set(CASES)
    assert not unknown, f"unknown cases {sorted(unknown)}"
    for name in (only or CASES):
        cfg, levels, layers, metas, tc = build_case(name)
        ref, ref_pp = run_reference(cfg, levels, layers, metas, tc)
        ours, our_pp = O.decode_full(levels, layers, metas, cfg.as_dict(), tc)
        for a, b in zip(our_pp, ref_pp):
            assert torch.equal(a, b), f"{name}: refined pose_pred differs from the reference"
        margin = synth.rank_margin_ulps(levels, tc.get("nms_pre", -1), tc.get("score_thr", 0.0))
        oks_margin = min(o["oks_margin"] for o in ours)
        assert margin >= synth.MIN_MARGIN_ULPS, f"{name}: rank margin {margin} ulp"
        assert oks_margin >= 1e-6, f"{name}: an OKS decision sits {oks_margin:.2e} from nms_thr"
        if tc.get("nms_type", "hard") != "hard":
            assert min(o["soft_gap_ulps"] for o in ours) >= 64, f"{name}: soft-NMS pick order is fragile"
        blob = dict(checksum=checksum(levels), n_images=np.array(len(ref)), rank_margin_ulps=np.array(margin),
                    oks_margin=np.array(min(oks_margin, 1e30)))
        for i, (o, r) in enumerate(zip(ours, ref)):
            assert torch.equal(o["poses"], r["poses"]), f"{name}[{i}] poses"
            assert torch.equal(o["centers"], r["centers"]), f"{name}[{i}] centers"
            assert torch.equal(o["vis"], r["vis"]), f"{name}[{i}] vis"
            assert o["scores_list"] == r["scores"], f"{name}[{i}] scores"
            assert np.array_equal(o["poses_cam"], r["poses_cam"]) and np.array_equal(o["poses_world"], r["poses_world"])
            blob[f"poses_{i}"] = r["poses"].numpy()
            blob[f"centers_{i}"] = r["centers"].numpy()
            blob[f"scores_{i}"] = np.asarray(r["scores"], dtype=np.float32)
            blob[f"cam_{i}"] = r["poses_cam"]
            blob[f"world_{i}"] = r["poses_world"]
            # bookkeeping the reference does not return (which cell each person came from), from the
            # restatement that was just shown to be bit-identical
            blob[f"level_{i}"] = o["level"].numpy()
            blob[f"index_{i}"] = o["index"].numpy()
        # a few refined pose_pred samples pin the dense refinement itself
        for l, pp in enumerate(ref_pp):
            flat = pp.flatten()
            pick = torch.linspace(0, flat.numel() - 1, 257).long()
            blob[f"pose_pred_samples_{l}"] = flat[pick].numpy()
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **blob)
        print(f"{name}: {[len(r['scores']) for r in ref]} people, bit-exact restatement, saved")


if __name__ == "__main__":
    main(sys.argv[1:])
