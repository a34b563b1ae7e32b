"""CPU: the oracle restatement against the golden vectors produced by the reference's own code
(oracle/make_golden.py), and -- when the reference tree is present -- against that code directly."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import das_oracle as O
from oracle import make_golden as G
from oracle import ref_extract as R

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
MODEL_FIXTURES = {"mspn_small", "das_head_small"}      # tests/test_model.py (oracle/make_model_golden.py, make_state_keys.py)
NAMES = sorted(n for n in (os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN, "*.npz")))
               if n not in MODEL_FIXTURES)


def test_golden_files_cover_all_cases():
    assert set(NAMES) == set(G.CASES), "tests/golden is out of date: run python -m oracle.make_golden"


@pytest.mark.parametrize("name", NAMES)
def test_oracle_reproduces_reference_golden(name):
    gold = np.load(os.path.join(GOLDEN, name + ".npz"))
    cfg, levels, layers, metas, tc = G.build_case(name)
    # the seeded generator must produce the very inputs the reference saw
    np.testing.assert_allclose(G.checksum(levels), gold["checksum"], rtol=1e-9, atol=1e-6)
    res, pps = O.decode_full(levels, layers, metas, cfg.as_dict(), tc)
    assert len(res) == int(gold["n_images"])
    for i, r in enumerate(res):
        # person order and the cells they came from: exact
        assert r["index"].tolist() == gold[f"index_{i}"].tolist()
        assert r["level"].tolist() == gold[f"level_{i}"].tolist()
        # values: identical code on identical torch => normally bit-equal; another CPU model may differ in the
        # last bits of vectorised sigmoid / conv, hence a tight tolerance rather than array_equal
        np.testing.assert_allclose(r["poses"].numpy(), gold[f"poses_{i}"], rtol=2e-5, atol=2e-4)
        np.testing.assert_allclose(r["centers"].numpy(), gold[f"centers_{i}"], rtol=2e-5, atol=2e-4)
        np.testing.assert_allclose(np.asarray(r["scores_list"], dtype=np.float32), gold[f"scores_{i}"], rtol=2e-6)
        np.testing.assert_allclose(r["poses_cam"], gold[f"cam_{i}"], rtol=2e-5, atol=2e-3)
        np.testing.assert_allclose(r["poses_world"], gold[f"world_{i}"], rtol=2e-5, atol=2e-3)
    for l, pp in enumerate(pps):
        flat = pp.flatten()
        pick = torch.linspace(0, flat.numel() - 1, 257).long()
        np.testing.assert_allclose(flat[pick].numpy(), gold[f"pose_pred_samples_{l}"], rtol=2e-5, atol=2e-4)


@pytest.mark.skipif(not R.available(), reason="reference tree not mounted (GPU box)")
def test_oracle_bit_exact_against_extracted_reference():
    name = "panoptic_scales_thr"
    cfg, levels, layers, metas, tc = G.build_case(name)
    ref, ref_pp = G.run_reference(cfg, levels, layers, metas, tc)
    ours, our_pp = O.decode_full(levels, layers, metas, cfg.as_dict(), tc)
    for a, b in zip(our_pp, ref_pp):
        assert torch.equal(a, b)
    for o, r in zip(ours, ref):
        assert torch.equal(o["poses"], r["poses"]) and torch.equal(o["centers"], r["centers"])
        assert o["scores_list"] == r["scores"]
        assert np.array_equal(o["poses_cam"], r["poses_cam"])


def test_oks_nms_known_answer():
    """Literal known-answer test in the style of the reference's tests/test_utils/test_nms.py."""
    J = 15
    base = np.zeros((4, J, 3), dtype=np.float32)
    rng = np.random.RandomState(0)
    base[0, :, :2] = rng.rand(J, 2) * 100
    base[1, :, :2] = base[0, :, :2] + 0.5          # near duplicate of 0 -> suppressed
    base[2, :, :2] = base[0, :, :2] + 60.0         # far away -> kept
    base[3, :, :2] = base[2, :, :2] + 0.25         # near duplicate of 2 -> suppressed
    base[..., 2] = 1
    scores = np.array([0.9, 0.8, 0.7, 0.95], dtype=np.float32)
    areas = np.array([(b[:, 0].max() - b[:, 0].min()) * (b[:, 1].max() - b[:, 1].min()) for b in base], dtype=np.float32)
    keep = O.oks_nms(scores, base.reshape(4, -1), areas, 0.9)
    assert keep.tolist() == [3, 0]


def test_oks_sigma_table_switches_at_17_joints():
    g15 = np.zeros(45, dtype=np.float32)
    d15 = np.zeros((1, 45), dtype=np.float32)
    d15[0, 0::3] = 10.0
    g17 = np.zeros(51, dtype=np.float32)
    d17 = np.zeros((1, 51), dtype=np.float32)
    d17[0, 0::3] = 10.0
    a = np.float32(1e4)
    o15 = O.oks_to_head(g15, d15, a, np.array([a]))[0]
    o17 = O.oks_to_head(g17, d17, a, np.array([a]))[0]
    assert abs(o15 - np.exp(-100 / 0.0256 / (1e4 + np.spacing(1)) / 2)) < 1e-6
    assert o17 != pytest.approx(o15, abs=1e-3)     # COCO table, pose_nms.py:66-70


def test_stable_variant_breaks_ties_towards_lower_index():
    cls = torch.full((1, 1, 4, 6), 0.3)
    ctr = torch.full((1, 1, 4, 6), -0.2)
    pose = torch.zeros(1, 3 + 6 * 15, 4, 6)
    meta = [dict(scale_factor=np.array([1, 1, 1, 1], dtype=np.float32), filename="t")]
    r = O.get_poses([cls], [pose], [ctr], meta, dict(nms_pre=5, nms_post=-1), [8], 15, stable=True)[0]
    assert r["index"].tolist() == [0, 1, 2, 3, 4]


def test_backproject_identity_camera():
    poses = np.zeros((1, 15, 3), dtype=np.float32)
    poses[0, :, 0] = 10
    poses[0, :, 1] = 20
    poses[0, :, 2] = 2.0
    K = np.array([[2.0, 0, 0], [0, 2.0, 0], [0, 0, 1]])
    cam, world = O.backproject(poses, K, np.eye(3), np.zeros(3), root_idx=2)
    # nd = 2, Z = 2*2 + 0 = 4, x = 10/2*4 = 20
    assert np.allclose(cam[0, 0], [20.0, 40.0, 4.0]) and np.allclose(world, cam)
