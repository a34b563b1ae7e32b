"""CPU, build container only: the oracle restatement against the reference's own extracted code on seeded random
configurations beyond the committed golden cases -- odd map sizes, pyramids with pass-through levels, J = 15 / 17 / 21,
1-3 refinement layers, per-level Scale values, thresholds on and off, hard and soft NMS.  Bit-exact, like make_golden."""
import dataclasses
import random

import numpy as np
import pytest
import torch

from das_b200 import synth
from oracle import das_oracle as O
from oracle import make_golden as G
from oracle import ref_extract as R

pytestmark = pytest.mark.skipif(not R.available(), reason="reference tree not mounted (GPU box)")


def _draw(seed):
    rnd = random.Random(seed)
    J, root = rnd.choice([(15, 2), (17, 14), (21, 14)])
    L = rnd.choice([1, 1, 2, 3])
    n_levels = rnd.choice([1, 1, 2, 4])
    strides = (8, 16, 32, 64)[:n_levels]
    cfg = synth.HeadConfig(num_joints=J, root_idx=root, depth_factor=rnd.choice([1.0, 20.0]), z_norm=50.0, num_layers=L,
                           strides=strides)
    # level 0 map; coarser levels halve it (rounded up by make_levels), so odd sizes stay odd down the pyramid
    h, w = rnd.randint(9, 30), rnd.randint(11, 38)
    tc = dict(nms_pre=rnd.choice([5, 12, 40, 1000]), nms_thr=rnd.choice([0.9, 0.8, 0.5]),
              score_thr=rnd.choice([0.0, 0.0, 0.03, 0.07]))
    if rnd.random() < 0.8:
        tc["nms_post"] = rnd.choice([5, 10, 100])
    if rnd.random() < 0.25:
        tc["nms_type"] = "soft"
    scales = tuple(round(rnd.uniform(0.85, 1.15), 3) for _ in range(4))
    extra = dict(coherent=8) if rnd.random() < 0.3 else {}
    return cfg, rnd.choice([1, 2, 3]), h, w, scales, tc, extra


@pytest.mark.parametrize("seed", range(9100, 9112))
def test_oracle_equals_extracted_reference_on_random_configs(seed):
    cfg, b, h, w, scales, tc, extra = _draw(seed)
    margin_for = dict(nms_pre=tc.get("nms_pre", -1), score_thr=tc.get("score_thr", 0.0))
    levels = synth.make_levels(cfg, b, h, w, seed=seed, peaks=10, scales=scales, margin_for=margin_for, **extra)
    layers = synth.make_layers(cfg, seed=seed + 1)
    metas = synth.make_metas(b, h, w, stride=cfg.strides[0], seed=seed + 2)
    ref, ref_pp = G.run_reference(cfg, levels, layers, metas, tc)
    ours, our_pp = O.decode_full(levels, layers, metas, cfg.as_dict(), tc)
    for a, r in zip(our_pp, ref_pp):
        assert torch.equal(a, r), "refined pose_pred differs from the reference"
    assert len(ours) == len(ref) == b
    for o, r in zip(ours, ref):
        assert torch.equal(o["poses"], r["poses"]) and torch.equal(o["centers"], r["centers"]) and torch.equal(o["vis"], r["vis"])
        assert o["scores_list"] == r["scores"]
        assert np.array_equal(o["poses_cam"], r["poses_cam"]) and np.array_equal(o["poses_world"], r["poses_world"])
