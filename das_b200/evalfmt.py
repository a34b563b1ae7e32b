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
* ``mupots_name2pred`` / ``mupots_frame_errors`` / ``pck_tables`` / ``mupots_pck``
                            -> the MuPoTS-3D 3DPCK protocol of mmdet3d/datasets/mupots_3dhp.py:298-350, 389-682
                               (matching, depth-ratio rescaling, bone-length normalisation, Procrustes, PCK@150 mm / AUC),
                               checked against the reference's own functions executed from source.
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


# ---------------------------------------------------------------------------------------------------------------
# MuPoTS-3D: 3DPCK of camera-space poses (reference: mmdet3d/datasets/mupots_3dhp.py:298-350 result handling,
# :389-474 joint tables + PCK, :480-566 bone-length normalisation / Procrustes / matching, :569-682 per-sequence loop).
# Written from the formulas; poses are [3, 17] arrays (x, y, z rows) in millimetres, joint 14 is the pelvis.
# ---------------------------------------------------------------------------------------------------------------
MUPOTS_ROOT = 14
# parent of each of the 17 joints (mupots_3dhp.py:421, 1-based in the reference) and the root-first traversal (:577)
MUPOTS_PARENT = np.array([2, 16, 2, 3, 4, 2, 6, 7, 15, 9, 10, 15, 12, 13, 15, 15, 2]) - 1
MUPOTS_TRAVERSAL = np.array([15, 16, 2, 1, 17, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14]) - 1
# PCK joint groups (:389-403): head, neck, shoulders, elbows, wrists, hips, knees, ankles
MUPOTS_GROUPS = (("Head", (0,)), ("Neck", (1,)), ("Shou", (2, 5)), ("Elbow", (3, 6)), ("Wrist", (4, 7)),
                 ("Hip", (8, 11)), ("Knee", (9, 12)), ("Ankle", (10, 13)))
_MUPOTS_ALL = tuple(j for _, g in MUPOTS_GROUPS for j in g)


def mupots_name2pred(results: Sequence[Mapping], num_joints: int = 17, data_root: str = "") -> dict:
    """file name -> [N, num_joints, 3] camera-space joints (mm) for the MuPoTS evaluator (mupots_3dhp.py:298-327).
    Takes the ``poses_world`` the decode already produced on the device (the reference keeps pixel2world's last return
    value, :324-326); an image without people maps to zeros [1, J, 3] exactly like the reference (:311-312)."""
    root = data_root if (not data_root or data_root.endswith("/")) else data_root + "/"
    out = {}
    for r in results:
        name = r["image_paths"][0]
        if root and name.startswith(root):
            name = name[len(root):]
        src = _np(r["poses_world"] if "poses_world" in r else r["poses_cam"])      # MuPoTS: R = I, t = 0, so both agree
        pts = src.reshape(-1, src.shape[-2], 3)[:, :num_joints]
        out[name] = pts.astype(np.float64) if len(pts) else np.zeros((1, num_joints, 3))
    return out


def bone_length_normalise(pred: np.ndarray, gt: np.ndarray) -> np.ndarray:
    """Re-scale the bones of `pred` [3,17] to the lengths of the same bones in `gt`, walking the traversal order
    (:480-489; the bone DIRECTIONS are the original prediction's, the start point is the already re-scaled parent).
    Reference quirk kept on purpose: the i-th traversed joint is paired with `o1[i]`, not with its own parent
    `o1[joint]` (:483-486) -- the numbers the reference reports depend on it, so a drop-in scorer must do the same."""
    out = pred.copy()
    for i, j in enumerate(MUPOTS_TRAVERSAL[1:]):
        par = MUPOTS_PARENT[i]
        vec = pred[:, j] - pred[:, par]
        out[:, j] = out[:, par] + vec * np.linalg.norm(gt[:, j] - gt[:, par]) / np.linalg.norm(vec)
    return out


def procrustes_align(pred: np.ndarray, gt: np.ndarray) -> np.ndarray:
    """Similarity transform (rotation without reflection, scale, translation) of `pred` [3,17] onto `gt` (:492-528)."""
    X, Y = gt.T, pred.T
    mx, my = X.mean(0, keepdims=True), Y.mean(0, keepdims=True)
    X0, Y0 = X - mx, Y - my
    nx, ny = np.sqrt((X0 ** 2).sum()), np.sqrt((Y0 ** 2).sum())
    X0, Y0 = X0 / nx, Y0 / ny
    U, s, Vt = np.linalg.svd(X0.T @ Y0)
    V = Vt.T
    R = V @ U.T
    sign = np.sign(np.linalg.det(R))
    V[:, -1] *= sign
    s[-1] *= sign
    R = V @ U.T
    a = s.sum() * nx / ny
    t = mx - a * (my @ R)
    return (a * (Y @ R) + t).T


def _mupots_canonical(pred_abs: np.ndarray, gt_abs: np.ndarray):
    """Root-relative prediction with x, y rescaled by the depth ratio gt/pred of the roots (:638-641, 647-648)."""
    p = pred_abs - pred_abs[:, MUPOTS_ROOT:MUPOTS_ROOT + 1]
    p[:2] = p[:2] * (gt_abs[2, MUPOTS_ROOT] / pred_abs[2, MUPOTS_ROOT])
    return p


def mupots_match(gt: Sequence[np.ndarray], pred: np.ndarray, threshold: float = 250.0):
    """For every ground-truth pose [3,17] the index of the closest prediction ([N,3,17]) after depth-ratio rescaling and
    bone-length normalisation, by mean root-relative joint distance (and by absolute distance), -1 above `threshold`
    (:531-566; float32 like the reference)."""
    p2 = np.float32(pred)
    roots = p2[:, :, MUPOTS_ROOT:MUPOTS_ROOT + 1]
    rel = p2 - roots
    matches, matches_abs = [], []
    for g in gt:
        g32 = np.float32(g)
        g_root = g32[:, MUPOTS_ROOT:MUPOTS_ROOT + 1]
        g_rel = g32 - g_root
        d_rel, d_abs = [], []
        for j in range(len(rel)):
            p = rel[j].copy()
            p[:2] *= g_root[[2]] / roots[j, [2]]
            p = bone_length_normalise(p, g_rel)
            d_rel.append(np.sqrt(((p - g_rel) ** 2).sum(0)).mean())
            d_abs.append(np.sqrt(((p + roots[j] - g_rel - g_root) ** 2).sum(0)).mean())
        d_rel, d_abs = np.float32(d_rel), np.float32(d_abs)
        matches.append(-1 if d_rel.min() > threshold else int(np.argmin(d_rel)))
        matches_abs.append(-1 if d_abs.min() > threshold else int(np.argmin(d_abs)))
    return matches, matches_abs


def mupots_frame_errors(gt: Sequence[np.ndarray], pred: np.ndarray, eval_all: bool = True) -> dict:
    """Per-joint errors [17] of every annotated person of ONE frame (:625-673).  `pred` is [N,17,3] as produced by
    mupots_name2pred; predictions whose root depth is 0 are dropped (:618-620).  Unmatched people count with an error
    of 1e5 per joint when `eval_all` (the reference's eval_mode='all'), else they are skipped."""
    p = np.asarray(pred, dtype=np.float64).transpose(0, 2, 1)
    p = p[p[:, 2, MUPOTS_ROOT] != 0]
    # No usable prediction: everybody is undetected.  (The reference substitutes one all-zero pose here, :621-622, whose
    # zero root depth turns the depth ratio into inf and makes its own Procrustes step raise LinAlgError.)
    matches = mupots_match(gt, p)[0] if len(p) else [-1] * len(gt)
    rel, aligned, absolute, abs_aligned, undetected = [], [], [], [], 0
    for g_abs, m in zip(gt, matches):
        g_abs = np.asarray(g_abs, dtype=np.float64)
        g_rel = g_abs - g_abs[:, MUPOTS_ROOT:MUPOTS_ROOT + 1]
        if m != -1:
            p_abs = p[m]
            root = p_abs[:, MUPOTS_ROOT:MUPOTS_ROOT + 1]
            canon = _mupots_canonical(p_abs, g_abs)
            p_align = procrustes_align(canon, g_rel)
            p_rel = bone_length_normalise(canon, g_rel)
            p_abs_n = p_rel + root
            p_abs_align = p_align - p_align[:, MUPOTS_ROOT:MUPOTS_ROOT + 1] + root
        else:
            undetected += 1
            if not eval_all:
                continue
            p_rel = p_abs_n = p_align = p_abs_align = 100000.0 * np.ones_like(g_rel)
        rel.append(np.sqrt(((p_rel - g_rel) ** 2).sum(0)))
        aligned.append(np.sqrt(((p_align - g_rel) ** 2).sum(0)))
        absolute.append(np.sqrt(((p_abs_n - g_abs) ** 2).sum(0)))
        abs_aligned.append(np.sqrt(((p_abs_align - g_abs) ** 2).sum(0)))
    return dict(rel=rel, aligned=aligned, abs=absolute, abs_aligned=abs_aligned, undetected=undetected)


def pck_tables(seq_err: Sequence[Sequence[np.ndarray]], pck_thresh: float = 150.0):
    """Per sequence: PCK curves over thresholds 0, 5, ..., 195 mm for the 8 joint groups + all joints, PCK@150 mm for the
    same 9 entries, AUC (mean of the curve) for the 8 groups (:436-473)."""
    thresh = np.arange(0, 200, 5)
    curves, pcks, aucs = [], [], []
    for errs in seq_err:
        err = np.array(errs).astype(np.float32)
        n = len(err)
        sel = [list(g) for _, g in MUPOTS_GROUPS] + [list(_MUPOTS_ALL)]
        curve = [[float(np.float32(err[:, g] < t).sum() / len(g) / n) for t in thresh] for g in sel]
        curves.append(curve)
        pcks.append([float(np.float32(err[:, g] < pck_thresh).sum() / len(g) / n) for g in sel])
        aucs.append([sum(c) / len(c) for c in curve[:-1]])
    return curves, pcks, aucs


def mupots_pck(name2pred: Mapping[str, np.ndarray], sequences: Sequence[Sequence[Mapping]], eval_all: bool = True) -> dict:
    """`sequences[s]` = frames of test sequence s, each ``dict(filename=..., gt=[[3,17] per valid person])``; returns the
    reference's two headline numbers (:330-349): PCK_MEAN (root-relative) and PCK_MEAN_ABS (absolute), in percent."""
    rel_seq, abs_seq = [], []
    for frames in sequences:
        rel, ab = [], []
        for fr in frames:
            if len(fr["gt"]) == 0:
                continue
            e = mupots_frame_errors(fr["gt"], name2pred[fr["filename"]], eval_all)
            rel += e["rel"]
            ab += e["abs"]
        rel_seq.append(rel)
        abs_seq.append(ab)
    _, pck, _ = pck_tables(rel_seq)
    _, pck_abs, _ = pck_tables(abs_seq)
    return {"PCK_MEAN": 100.0 * sum(p[-1] for p in pck) / len(pck),
            "PCK_MEAN_ABS": 100.0 * sum(p[-1] for p in pck_abs) / len(pck_abs)}


def load_mupots_sequence(annot_mat: str, ts: int) -> List[dict]:
    """Frames of MuPoTS test sequence `ts` (0-based) from its ``annot.mat`` (the layout mupots_3dhp.py:353-374 parses);
    needs scipy and the dataset files, neither of which ships with this repository."""
    import scipy.io as sio
    data = sio.loadmat(annot_mat)["annotations"]
    frames = []
    for i in range(data.shape[0]):
        gt = [data[i, k]["annot3"][0, 0] for k in range(data.shape[1]) if data[i, k]["isValidFrame"][0, 0][0, 0] == 1]
        frames.append(dict(filename="TS%d/img_%06d.jpg" % (ts + 1, i), gt=gt))
    return frames
