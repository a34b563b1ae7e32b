"""CPU: the parity cases are DECIDABLE.  The GPU tests compare person order bit-exactly with no escape hatch, which is only
meaningful when no ranking / threshold / OKS decision of a case is a near-tie (rank margins >= 16 ulp at
every boundary, |oks - nms_thr| >= 1e-6).  These tests run the generator's reject-sampling and the oracle on the host for
every named GPU case and assert exactly what the GPU tests assert before they compare."""
import dataclasses

import pytest
import torch

import util
from das_b200 import synth
from test_gpu_parity import CASES, P


@pytest.mark.parametrize("case_id,cfg,B,H,W,tc,kw", CASES, ids=[c[0] for c in CASES])
def test_named_cases_have_safe_margins(case_id, cfg, B, H, W, tc, kw):
    if H * W > 128 * 208:
        cfg = dataclasses.replace(cfg, feat_channels=128)      # keeps the host-side oracle run short; margins do not depend on C
    for seed in (1234, 99):
        case = util.make_case(cfg, B, H, W, seed=seed, tc=tc, **kw)
        ref, _ = util.run_oracle(case, tc)
        assert util.assert_margins(case, tc, ref) >= synth.MIN_MARGIN_ULPS


def test_reject_sampling_repairs_a_near_tie_at_every_kind_of_boundary():
    tc = dict(nms_pre=5, nms_post=5, nms_thr=0.9, score_thr=0.3)
    lv = synth.make_levels(P, 1, 12, 16, seed=3, with_feats=False, peaks=6)
    flat_cls, flat_ctr = lv[0]["cls"].view(-1), lv[0]["ctr"].view(-1)
    order = (flat_cls.sigmoid() * flat_ctr.sigmoid()).argsort(descending=True)
    # (a) ranks 2 and 3 tie exactly; (b) rank K+1 ties with rank K
    flat_cls[order[2]], flat_ctr[order[2]] = flat_cls[order[1]], flat_ctr[order[1]]
    flat_cls[order[5]], flat_ctr[order[5]] = flat_cls[order[4]], flat_ctr[order[4]]
    assert synth.rank_margin_ulps(lv, 5, 0.3) == 0
    n = synth.enforce_rank_margins(lv, nms_pre=5, score_thr=0.3, seed=11)
    assert n >= 2 and synth.rank_margin_ulps(lv, 5, 0.3) >= 4 * synth.MIN_MARGIN_ULPS
    # ties among cells that score_thr drops anyway do not count (they never reach the output)
    lv2 = synth.make_levels(P, 1, 12, 16, seed=4, with_feats=False, peaks=0)
    lv2[0]["cls"].fill_(-6.0)
    lv2[0]["ctr"].fill_(0.0)
    lv2[0]["cls"][0, 0, 3, 3] = 4.0
    assert synth.rank_margin_ulps(lv2, 5, 0.3) >= synth.MIN_MARGIN_ULPS        # background ties are below the threshold
    assert synth.rank_margin_ulps(lv2, 5, 0.0) == 0                            # ... but decide the output without one


def test_margin_sees_cross_level_order_and_peak_mask_ties():
    cfg = dataclasses.replace(P, strides=(8, 16))
    lv = synth.make_levels(cfg, 1, 16, 24, seed=5, with_feats=False, peaks=4)
    synth.enforce_rank_margins(lv, nms_pre=3, score_thr=0.0, seed=1)
    assert synth.rank_margin_ulps(lv, 3, 0.0) >= synth.MIN_MARGIN_ULPS
    # copy the best cell of level 0 into level 1: same score in two levels -> the NMS visiting order is a coin flip
    s0 = (lv[0]["cls"].sigmoid() * lv[0]["ctr"].sigmoid()).view(-1)
    i0 = int(s0.argmax())
    lv[1]["cls"].view(-1)[7] = lv[0]["cls"].view(-1)[i0]
    lv[1]["ctr"].view(-1)[7] = lv[0]["ctr"].view(-1)[i0]
    assert synth.rank_margin_ulps(lv, 3, 0.0) == 0
    synth.enforce_rank_margins(lv, nms_pre=3, score_thr=0.0, seed=2)
    assert synth.rank_margin_ulps(lv, 3, 0.0) >= synth.MIN_MARGIN_ULPS
    # peak mode: a plateau of two equal neighbouring maxima makes the 3x3 mask itself a near-tie
    lp = synth.make_levels(P, 1, 16, 24, seed=6, with_feats=False, peaks=4)
    s = (lp[0]["cls"].sigmoid() * lp[0]["ctr"].sigmoid()).view(-1)
    i = int(s.argmax())
    j = i + 1 if (i % 24) < 23 else i - 1
    lp[0]["cls"].view(-1)[j], lp[0]["ctr"].view(-1)[j] = lp[0]["cls"].view(-1)[i], lp[0]["ctr"].view(-1)[i]
    assert synth.rank_margin_ulps(lp, 3, 0.0, peak_kernel=3) == 0
    synth.enforce_rank_margins(lp, nms_pre=3, score_thr=0.0, peak_kernel=3, seed=3)
    assert synth.rank_margin_ulps(lp, 3, 0.0, peak_kernel=3) >= synth.MIN_MARGIN_ULPS


def test_generator_is_deterministic_with_reject_sampling():
    tc = dict(nms_pre=1000, score_thr=0.0)
    a = synth.make_levels(P, 1, 40, 72, seed=9, with_feats=False, peaks=24, margin_for=tc)
    b = synth.make_levels(P, 1, 40, 72, seed=9, with_feats=False, peaks=24, margin_for=tc)
    assert torch.equal(a[0]["ctr"], b[0]["ctr"]) and synth.rank_margin_ulps(a, 1000, 0.0) >= synth.MIN_MARGIN_ULPS
