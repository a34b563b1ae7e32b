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
