"""CPU: evaluator-side formats against the reference's own evaluator methods, extracted
with ast when the reference tree is mounted, plus literal known answers."""
import ast
import json
import os

import numpy as np
import pytest
import torch

from das_b200 import evalfmt
from oracle import ref_extract as R


def _results(rng, n_img=3, J=15):
    out = []
    for i in range(n_img):
        n = [2, 0, 3][i % 3]
        poses = torch.tensor(rng.rand(n, J, 3) * 100, dtype=torch.float32)
        out.append(dict(poses=poses, vis=torch.ones(n, J), centers=torch.zeros(n, 3),
                        image_paths=[f"/data/x/img_{i}.jpg"], scores=rng.rand(n).tolist()))
    return out


def test_records_known_answer(tmp_path):
    J = 2
    res = [dict(poses=torch.tensor([[[1., 2., 3.], [5., 1., 4.]]]), image_paths=["a/b/img.jpg"], scores=[0.5])]
    rec = evalfmt.keypoint_records(res, {"img.jpg": 7}, J)
    assert rec == [dict(image_id=7, category_id=1, keypoints=[1., 2., 3., 5., 1., 4.], score=0.5, bbox=[1., 1., 4., 1.])]
    f = evalfmt.write_result_keypoints(rec, str(tmp_path / "o" / "result_keypoints.json"))
    assert json.load(open(f)) == rec


@pytest.mark.skipif(not R.available(), reason="reference tree not mounted (GPU box)")
def test_records_and_mpjpe_match_reference_methods(tmp_path):
    src = open(os.path.join(R.REF, "mmdet3d/datasets/cmupanoptic_mono_dataset.py")).read()
    fns = {}
    for node in ast.parse(src).body:
        if isinstance(node, ast.ClassDef) and node.name == "CMUPanopticDataset":
            for f in node.body:
                if isinstance(f, ast.FunctionDef) and f.name in ("_coco_keypoint_results_one_category_kernel", "vectorize_distance", "mse"):
                    f.decorator_list = []
                    fns[f.name] = f
    ns = dict(np=np)
    for f in fns.values():
        exec(ast.unparse(f), ns)

    class Stub:
        num_joints = 15
    stub = Stub()
    rng = np.random.RandomState(1)
    results = _results(rng)
    name2id = {f"img_{i}.jpg": 100 + i for i in range(3)}
    # the reference's evaluate() builds these dicts (:283-303) before calling the kernel
    packs = []
    for r in results:
        packs.append([dict(keypoints=k[:, 0:3], score=s, image_id=name2id[os.path.basename(r["image_paths"][0])])
                      for k, s in zip(r["poses"].numpy(), r["scores"])])
    want = ns["_coco_keypoint_results_one_category_kernel"](stub, dict(cat_id=1, keypoints=packs))
    got = evalfmt.keypoint_records(results, name2id, 15)
    assert got == want
    # matching + per-joint error
    pred = rng.rand(4, 15, 3) * 50
    gt = rng.rand(3, 15, 3) * 50
    vis = (rng.rand(3, 15) > 0.2).astype(np.float64)
    idx = ns["vectorize_distance"](stub, pred, gt, vis)
    assert evalfmt.match_to_ground_truth(pred, gt, vis).tolist() == idx.tolist()
    p0, g0 = pred - pred[:, [2]], gt - gt[:, [2]]
    idx0 = ns["vectorize_distance"](stub, p0, g0, vis)
    ref_val = ns["mse"](stub, p0[idx0], g0, vis).mean() * 10
    assert abs(evalfmt.mpjpe([pred], [gt], [vis], root_idx=2) - ref_val) < 1e-9


def test_mpjpe_is_zero_for_perfect_predictions_and_skips_empty_gt():
    rng = np.random.RandomState(3)
    gt = rng.rand(2, 15, 3) * 100
    vis = np.ones((2, 15))
    shifted = gt[::-1] + 5.0                      # order and a global shift do not matter after root alignment
    assert evalfmt.mpjpe([shifted, gt], [gt, np.zeros((0, 15, 3))], [vis, np.zeros((0, 15))], root_idx=2) < 1e-9
    with pytest.raises(ValueError):
        evalfmt.mpjpe([np.zeros((0, 15, 3))], [gt], [vis], root_idx=2)


def _mupots_scene(rng, n_gt, n_pred, noise=20.0):
    """Synthetic MuPoTS-like frame: people 3-5 m from the camera, 17 joints within ~0.8 m of the pelvis (mm)."""
    gt = []
    for _ in range(n_gt):
        root = np.array([[rng.uniform(-1500, 1500)], [rng.uniform(-500, 500)], [rng.uniform(3000, 5000)]])
        g = root + rng.randn(3, 17) * 250.0
        g[:, 14:15] = root
        gt.append(g)
    pred = []
    for k in range(n_pred):
        base = gt[k % n_gt] if n_gt else rng.randn(3, 17) * 300 + np.array([[0.], [0.], [4000.]])
        p = base + rng.randn(3, 17) * noise + (0 if k < n_gt else 900.0)
        pred.append(p.T)
    return gt, np.array(pred).reshape(n_pred, 17, 3)


def test_pck_known_answer():
    # two people, one joint group check: errors 100 mm and 160 mm on every joint -> PCK@150 = 0.5 everywhere
    errs = [[np.full(17, 100.0), np.full(17, 160.0)]]
    curves, pck, auc = evalfmt.pck_tables(errs)
    assert all(abs(v - 0.5) < 1e-7 for v in pck[0]) and len(pck[0]) == 9 and len(auc[0]) == 8
    assert curves[0][-1][20] == 0.0 and curves[0][-1][21] == 0.5 and curves[0][-1][33] == 1.0   # thresholds 100, 105, 165 mm
    # a perfect prediction scores 100, a frame without predictions (zeros) scores 0 in 'all' mode
    rng = np.random.RandomState(0)
    gt, _ = _mupots_scene(rng, 2, 2)
    perfect = np.array([g.T for g in gt])
    seqs = [[dict(filename="TS1/img_000000.jpg", gt=gt)]]
    assert evalfmt.mupots_pck({"TS1/img_000000.jpg": perfect}, seqs)["PCK_MEAN"] == 100.0
    assert evalfmt.mupots_pck({"TS1/img_000000.jpg": np.zeros((1, 17, 3))}, seqs)["PCK_MEAN"] == 0.0
    # name2pred takes the decode's world-space joints and pads empty images with one all-zero pose
    res = [dict(image_paths=["/d/TS1/img_000000.jpg"], poses_world=torch.zeros(0, 21, 3, dtype=torch.float64)),
           dict(image_paths=["/d/TS1/img_000001.jpg"], poses_world=torch.ones(2, 21, 3, dtype=torch.float64))]
    n2p = evalfmt.mupots_name2pred(res, 17, data_root="/d")
    assert n2p["TS1/img_000000.jpg"].shape == (1, 17, 3) and n2p["TS1/img_000001.jpg"].shape == (2, 17, 3)


@pytest.mark.skipif(not R.available(), reason="reference tree not mounted (GPU box)")
def test_mupots_evaluator_matches_reference_functions():
    """The reference's own match / procrustes / norm_by_bone_length / PCK functions and its per-sequence loop
    (mupots_3dhp.py:389-682), executed from source with the .mat loaders replaced by synthetic annotations."""
    src = open(os.path.join(R.REF, "mmdet3d/datasets/mupots_3dhp.py")).read()
    keep = ("mpii_joint_groups", "mpii_get_joints", "mean", "mpii_compute_3d_pck", "calculate_multiperson_errors",
            "norm_by_bone_length", "procrustes", "match", "eval_mupots_abs")
    ns = dict(np=np, os=os)
    for node in ast.parse(src).body:
        if isinstance(node, ast.FunctionDef) and node.name in keep:
            exec(ast.unparse(node), ns)
    rng = np.random.RandomState(3)
    # one sequence: frames with 0..3 annotated people, missed detections, spurious detections, a zero-depth prediction
    frames, annots_by_person, occ = [], [[] for _ in range(3)], []
    name2pred = {}
    # (the reference itself raises LinAlgError on a frame with people but no usable prediction, so none is included)
    for i, (n_gt, n_pred) in enumerate([(2, 2), (3, 2), (0, 1), (1, 3), (2, 1), (3, 4)]):
        gt, pred = _mupots_scene(rng, n_gt, max(n_pred, 0), noise=40.0)
        if i == 3:
            pred[2, 14, 2] = 0.0                       # dropped by the reference (:618-620)
        fname = "TS1/img_%06d.jpg" % i
        name2pred[fname] = pred
        frames.append(dict(filename=fname, gt=gt))
        for k in range(3):
            valid = k < n_gt
            annots_by_person[k].append(dict(is_valid=1 if valid else 0, annot3=gt[k] if valid else np.zeros((3, 17)),
                                            annot2=np.zeros((2, 17))))
        occ.append([np.zeros((1, 17)) for _ in range(3)])
    ns["load_annot"] = lambda fname: annots_by_person
    ns["load_occ"] = lambda fname: occ
    for mode in ("all", "matched"):
        res = {}
        want = ns["eval_mupots_abs"](0, "unused", {k: v.copy() for k, v in name2pred.items()}, res, eval_mode=mode)
        want_abs = res[0]["sequencewise_per_joint_error_abs"]
        got_rel, got_abs = [], []
        for fr in frames:
            if len(fr["gt"]) == 0:
                continue
            e = evalfmt.mupots_frame_errors(fr["gt"], name2pred[fr["filename"]], eval_all=(mode == "all"))
            got_rel += e["rel"]
            got_abs += e["abs"]
        assert len(got_rel) == len(want[0]) > 0
        np.testing.assert_allclose(np.array(got_rel), np.array(want[0]), rtol=1e-9, atol=1e-6)
        np.testing.assert_allclose(np.array(got_abs), np.array(want_abs[0]), rtol=1e-9, atol=1e-6)
        curves, pck, auc = ns["calculate_multiperson_errors"](want)
        c2, p2, a2 = evalfmt.pck_tables([got_rel])
        np.testing.assert_allclose(np.array(p2), np.array(pck), rtol=0, atol=1e-7)
        np.testing.assert_allclose(np.array(c2), np.array(curves), rtol=0, atol=1e-7)
        np.testing.assert_allclose(np.array(a2), np.array(auc), rtol=0, atol=1e-7)
    # building blocks on their own
    g, p = _mupots_scene(rng, 1, 1, noise=60.0)
    o1 = ns["mpii_get_joints"]("relavant")[1]
    trav = [t - 1 for t in [15, 16, 2, 1, 17, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14]][1:]
    np.testing.assert_allclose(evalfmt.bone_length_normalise(p[0].T, g[0]), ns["norm_by_bone_length"](p[0].T, g[0], o1, trav), rtol=1e-12)
    np.testing.assert_allclose(evalfmt.procrustes_align(p[0].T, g[0]), ns["procrustes"](p[0].T.copy(), g[0].copy()), rtol=1e-9, atol=1e-9)
