"""Evaluator-side formats of the step AFTER the decode.

Host-side NumPy, mirroring what ``CMUPanopticDataset.evaluate`` does with the head's result list
(reference: mmdet3d/datasets/cmupanoptic_mono_dataset.py:267-359 record writing, :361-424 MPJPE), so results
decoded on the GPU can be scored and exchanged in the reference's own file format:

* ``keypoint_records``      -> the COCO-style ``result_keypoints.json`` records
                               {image_id, category_id, keypoints[3J], score, bbox} (:339-357)
* ``write_result_keypoints`` -> json.dump(..., sort_keys=True, indent=4) like :326-327
* ``mpjpe``                 -> root-aligned MPJPE with nearest-ground-truth matching (:361-370, :406-421); it takes
                               WORLD/CAMERA-space joints, i.e. the ``poses_world`` / ``poses_cam`` the CUDA decode
                               already back-projected on the device (the reference does that step on the host after a
                               JSON round trip, :391-402).
"""
from __future__ import annotations

import json
import os
from typing import Iterable, List, Mapping, Sequence

import numpy as np


def _np(x):
    return x.detach().cpu().numpy() if hasattr(x, "detach") else np.asarray(x)


def keypoint_records(results: Sequence[Mapping], name2id: Mapping[str, int], num_joints: int,
                     category_id: int = 1) -> List[dict]:
    """One record per decoded person, images in result order (cmupanoptic_mono_dataset.py:283-357).

    ``results`` is the list ``get_poses`` returns; ``name2id`` maps ``os.path.basename(image_path)`` to the
    dataset's image id (the reference's ``self.name2id``)."""
    out = []
    for r in results:
        poses = _np(r["poses"]).reshape(-1, num_joints, 3)
        if len(poses) == 0:
            continue
        image_id = name2id[os.path.basename(r["image_paths"][0])]
        for kpt, score in zip(poses, r["scores"]):
            lt = np.amin(kpt, axis=0)
            rb = np.amax(kpt, axis=0)
            out.append({
                "image_id": image_id,
                "category_id": category_id,
                "keypoints": kpt.reshape(num_joints * 3).tolist(),
                "score": float(score),
                "bbox": np.array([lt[0], lt[1], rb[0] - lt[0], rb[1] - lt[1]]).tolist(),
            })
    return out


def write_result_keypoints(records: Iterable[dict], res_file: str) -> str:
    os.makedirs(os.path.dirname(os.path.abspath(res_file)), exist_ok=True)
    with open(res_file, "w") as f:
        json.dump(list(records), f, sort_keys=True, indent=4)
    return res_file


def match_to_ground_truth(pred: np.ndarray, gt: np.ndarray, vis: np.ndarray) -> np.ndarray:
    """For every ground-truth person the index of the closest prediction (mean visible-joint distance);
    cmupanoptic_mono_dataset.py:361-366."""
    d = np.sqrt(((gt[:, None] - pred[None]) ** 2).sum(axis=-1))
    d = d * vis[:, None]
    return d.mean(-1).argmin(1)


def mpjpe(pred_per_image: Sequence, gt_per_image: Sequence, vis_per_image: Sequence, root_idx: int,
          mean_pose: np.ndarray = None, unit_scale: float = 10.0) -> float:
    """Root-aligned MPJPE averaged over ground-truth people (cmupanoptic_mono_dataset.py:403-424).

    pred/gt: per image [N,J,3] / [M,J,3] in the same metric space (Panoptic: cm, ``unit_scale`` 10 -> mm);
    vis: [M,J].  Images without ground truth are skipped; images without predictions use ``mean_pose``."""
    total, count = 0.0, 0
    for pred, gt, vis in zip(pred_per_image, gt_per_image, vis_per_image):
        pred, gt, vis = _np(pred).astype(np.float64), _np(gt).astype(np.float64), _np(vis).astype(np.float64)
        if len(gt) == 0:
            continue
        pred = pred - pred[:, [root_idx]] if len(pred) else pred
        if len(pred) == 0:
            if mean_pose is None:
                raise ValueError("an image has no prediction and no mean_pose fallback was given")
            pred = np.asarray(mean_pose, dtype=np.float64)[None]
        gt = gt - gt[:, [root_idx]]
        idx = match_to_ground_truth(pred, gt, vis)
        sel = pred[idx]
        jpe = np.sqrt(((sel[vis > 0] - gt[vis > 0]) ** 2).sum(axis=-1))
        if len(jpe) > 0:
            total += jpe.mean() * unit_scale * len(gt)
            count += len(gt)
    return total / max(count, 1)
