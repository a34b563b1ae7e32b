"""GPU: the hard OKS-NMS decision `oks > nms_thr` (pose_nms.py:97) through das_nms_backproject called directly.

The kernel settles a pair with a float32 estimate when the estimate is further than 1e-3 from the threshold (on either
side) and runs the reference's float64 chain only inside that band.  These cases put pairs everywhere: a few 1e-6 .. 1e-3
from the threshold on both sides (inside the band), just outside it, near-duplicates (OKS ~ 1) and distinct people
(OKS ~ 0) -- the kept lists must equal the oracle's oks_nms image by image."""
import ctypes as C

import numpy as np
import pytest
import torch

from das_b200 import _lib
from oracle import das_oracle as O

pytestmark = pytest.mark.gpu

THR = 0.9
# distance of the pair's OKS from the threshold: inside the band, at its edge, far outside
DELTAS = [2e-6, 5e-6, 2e-5, 1e-4, 4e-4, 9e-4, 1.1e-3, 2e-3, 1e-2, 5e-2, 9.9e-2]


def _areas(poses):
    hi = poses[..., :2].max(1)[0]
    lo = poses[..., :2].min(1)[0]
    return (hi - lo).prod(-1).numpy()


def _oks(g, d, a_g, a_d):
    kp = lambda p: torch.cat([p[..., :2], torch.ones_like(p[..., :1])], -1).reshape(-1).numpy()
    return float(O.oks_to_head(kp(g), kp(d)[None], a_g, np.asarray([a_d], np.float32))[0])


def _shift_for(base, target):
    """x-shift of every joint that brings OKS(base, base + shift) as close to `target` as float32 poses allow (bisection on
    the oracle's own OKS; a uniform shift keeps the area)."""
    a = float(_areas(base[None])[0])
    lo, hi = 0.0, 200.0
    for _ in range(80):
        mid = 0.5 * (lo + hi)
        moved = base.clone()
        moved[:, 0] += np.float32(mid)
        if _oks(base, moved, a, float(_areas(moved[None])[0])) > target:
            lo = mid
        else:
            hi = mid
    return np.float32(0.5 * (lo + hi))


@pytest.mark.parametrize("J,CT", [(15, 4), (17, 4), (15, 64), (17, 80)], ids=["J15", "J17_coco_sigmas", "J15_K64_matrix", "J17_K80_greedy"])
def test_hard_nms_decisions_around_the_threshold(J, CT):
    g = torch.Generator().manual_seed(900 + J + CT)
    targets = [THR + s * d for d in DELTAS for s in (+1, -1)] + [0.999, 0.02]
    B = len(targets)
    poses = torch.zeros(B, CT, J, 3)
    scores = torch.zeros(B, CT)
    for b, tgt in enumerate(targets):
        base = torch.rand(J, 3, generator=g) * torch.tensor([150.0, 220.0, 30.0]) + torch.tensor([300.0, 200.0, 100.0])
        moved = base.clone()
        moved[:, 0] += _shift_for(base, tgt)
        poses[b, 0], poses[b, 1] = base, moved
        for c in range(2, CT):
            if c % 3 == 2:      # somebody else, far away
                poses[b, c] = torch.rand(J, 3, generator=g) * torch.tensor([150.0, 220.0, 30.0]) + torch.tensor([900.0 + 40 * c, 300.0, 100.0])
            else:               # another near-duplicate of an earlier candidate, at a random small shift
                poses[b, c] = poses[b, c - 2] + (torch.rand(1, generator=g) * 14.0) * torch.tensor([1.0, 0.3, 0.0])
        scores[b] = torch.sort(torch.rand(CT, generator=g) * 0.5 + 0.3, descending=True)[0]
    scores += torch.arange(CT).flip(0)[None] * 1e-4          # distinct, strictly descending in slot order
    centers = poses[:, :, 0].clone()

    # oracle: kept slots per image
    want, margins = [], []
    for b in range(B):
        kp = torch.cat([poses[b, ..., :2], torch.ones_like(poses[b, ..., :1])], -1).reshape(CT, -1).numpy()
        trace = {}
        keep = O.oks_nms(scores[b].numpy(), kp, _areas(poses[b]), THR, stable=True, trace=trace).tolist()
        want.append(keep[:CT])
        margins.append(trace.get("oks_margin", np.inf))
    assert min(margins) >= 1e-6, f"a decision sits closer than 1e-6 to the threshold ({min(margins):.2e}): regenerate"
    assert sum(m < 1e-3 for m in margins) >= 10, "the case does not put pairs inside the float64 band"
    assert any(len(w) < CT for w in want) and any(1 in w for w in want), "both outcomes must occur"

    lib = _lib.load()
    cfg = _lib.DecodeCfg(num_joints=J, root_idx=2, num_heads=4, feat_channels=256, num_layers=1, depth_factor=1.0, z_norm=1.0,
                         nms_pre=CT, nms_post=CT, nms_thr=THR, score_thr=0.0, peak_kernel=0, refine=0, dataset_depth_factor=1.0,
                         nms_soft=0)
    dev = torch.device("cuda")
    d_score, d_pose, d_center = scores.to(dev), poses.to(dev).contiguous(), centers.to(dev).contiguous()
    cam = torch.tensor([1, 0, 0, 0, 1, 0, 1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0], dtype=torch.float64).repeat(B, 1).to(dev)
    P = CT
    out = dict(out_count=torch.zeros(B, dtype=torch.int32, device=dev), out_score=torch.zeros(B, P, device=dev),
               out_slot=torch.zeros(B, P, dtype=torch.int32, device=dev), out_pose=torch.zeros(B, P, J, 3, device=dev),
               out_center=torch.zeros(B, P, 3, device=dev), out_cam=torch.zeros(B, P, J, 3, dtype=torch.float64, device=dev),
               out_world=torch.zeros(B, P, J, 3, dtype=torch.float64, device=dev))
    bufs = _lib.Buffers(**{k: v.data_ptr() for k, v in out.items()})
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(lib.das_nms_backproject(C.byref(cfg), B, CT, C.c_void_p(d_score.data_ptr()), C.c_void_p(d_pose.data_ptr()),
                                       C.c_void_p(d_center.data_ptr()), C.c_void_p(cam.data_ptr()), bufs, C.c_void_p(st)),
               "das_nms_backproject")
    torch.cuda.synchronize()
    counts = out["out_count"].cpu().tolist()
    slots = out["out_slot"].cpu()
    for b in range(B):
        got = slots[b, :counts[b]].tolist()
        assert got == want[b], (b, targets[b], margins[b], got, want[b])
        assert torch.equal(out["out_pose"][b, :counts[b]].cpu(), poses[b, want[b]])
