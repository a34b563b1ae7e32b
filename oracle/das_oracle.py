"""CPU oracle for the DAS dense-head inference decode.  TEST INFRASTRUCTURE ONLY.

This file is a from-scratch restatement (torch-CPU fp32 + NumPy fp64) of the
algorithm the reference executes on this path.  It is the *checker* for the
CUDA path in ``das_b200/``; nothing under ``das_b200/`` may import it.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs use it.

Parity pin: the reference ships no test, fixture or golden vector for this
path.  The restatement is pinned instead against the
reference's own functions executed verbatim in the build container
(``oracle/ref_extract.py`` AST-extracts them from ``/root/reference``;
``oracle/make_golden.py`` asserts bit-equality of this file against them and
writes ``tests/golden/*.npz`` from the REFERENCE outputs).

Reference lines restated (all relative to /root/reference):
  * point grid ............ mmdet3d/models/pose_heads/das_head.py:269-279,
                            anchor_free_mono3d_pose_head.py:251-283
  * head eval tail ........ das_head.py:237-262
  * gated blend ........... recursive_update.py:186-197 (NextLevelOffset.forward,
                            minus the DCN feature update at :188 which is
                            outside the decode boundary)
  * progressive sampling .. recursive_update.py:9-31, 34-82
  * branch loop ........... recursive_update.py:250-255
  * decode ................ das_head.py:653-796
  * OKS-NMS ............... mmdet3d/core/post_processing/pose_nms.py:51-126
  * depth de-norm ......... mmdet3d/datasets/cmupanoptic_mono_dataset.py:391-401
  * back-projection ....... mytools/vis_3d.py:16-26
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

COCO17_SIGMAS = np.array([.26, .25, .25, .35, .35, .79, .79, .72, .72, .62, .62,
                          1.07, 1.07, .87, .87, .89, .89]) / 10.0


# --------------------------------------------------------------------------
# point grid
# --------------------------------------------------------------------------
def point_grid(h: int, w: int, stride: int, dtype=torch.float32, device=None) -> torch.Tensor:
    """[H*W, 2] image-space anchor of every cell, raster order i = y*W + x.

    das_head.py:276-278: (x*stride, y*stride) + stride // 2.
    """
    ys, xs = torch.meshgrid(torch.arange(h, dtype=dtype, device=device), torch.arange(w, dtype=dtype, device=device), indexing="ij")
    return torch.stack((xs.reshape(-1) * stride, ys.reshape(-1) * stride), dim=-1) + stride // 2


def cell_centres(h: int, w: int, dtype=torch.float32, device=None) -> torch.Tensor:
    """[2, H, W] feature-space cell centres (x+0.5, y+0.5); recursive_update.py:211-218."""
    ys, xs = torch.meshgrid(torch.arange(h, dtype=dtype, device=device), torch.arange(w, dtype=dtype, device=device), indexing="ij")
    return torch.stack((xs, ys), dim=0) + 0.5


# --------------------------------------------------------------------------
# progressive refinement (dense, as the reference runs it)
# --------------------------------------------------------------------------
def _f32_unless_f64(t: torch.Tensor) -> torch.Tensor:
    """The reference forces ``.float()`` before grid_sample (recursive_update.py:25,56).  Feeding float64
    maps keeps float64 instead: that is the fp64 ARBITER mode the noise-floor tests use."""
    return t if t.dtype == torch.float64 else t.float()


def project_1x1(feat: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor) -> torch.Tensor:
    """nn.Conv2d(C, O, 1) forward: weight [O, C], bias [O]."""
    return F.conv2d(feat, weight[:, :, None, None], bias)


def gated_blend(feat, offset, layer):
    """recursive_update.py:190-195 -> (offset', sampling_offset, sampling_conf)."""
    samp_off = project_1x1(feat, layer["so_w"], layer["so_b"])
    samp_conf = project_1x1(feat, layer["sc_w"], layer["sc_b"])
    gate = project_1x1(feat, layer["uw_w"], layer["uw_b"]).sigmoid()
    value = project_1x1(feat, layer["uv_w"], layer["uv_b"])
    offset = (1 - gate) * offset + gate * value
    return offset, samp_off, samp_conf


def progressive_sample(uvd, samp_off, joint_conf, num_joints: int, num_heads: int, dim: int = 3):
    """Dense progressive sampling step; recursive_update.py:34-82 + 9-31.

    uvd        [B, J*dim, H, W]   blended offsets (feature px, normalised depth)
    samp_off   [B, J*nh*2, H, W]  per-head sampling offsets
    joint_conf [B, J*dim, H, W]   per-dim confidence logits
    returns    [B, J, dim, H, W]
    """
    b = uvd.shape[0]
    h, w = uvd.shape[-2:]
    bj = b * num_joints
    pts = cell_centres(h, w, uvd.dtype, uvd.device)
    wh = uvd.new_tensor([w, h]).view(1, 2, 1, 1)

    uvd = uvd.view(bj, dim, h, w)
    to_target = uvd[:, :2]
    # heads anchored at the current joint estimate ("from target")
    tgt = ((pts + to_target) / wh).permute(0, 2, 3, 1)
    so_map = samp_off.view(bj, num_heads * 2, h, w)
    from_target = F.grid_sample(_f32_unless_f64(so_map), 2 * tgt - 1, mode="bilinear",
                                padding_mode="zeros", align_corners=False)
    from_target = from_target.view(bj, num_heads, 2, h, w) + to_target[:, None]
    # heads anchored at the source cell ("from source")
    from_source = samp_off.view(bj, num_heads, 2, h, w)
    heads = torch.cat([from_target, from_source], dim=1)            # [BJ, 2nh, 2, H, W]
    nh2 = 2 * num_heads
    heads = heads.view(bj * nh2, 2, h, w)
    loc = ((pts + heads) / wh).permute(0, 2, 3, 1)

    conf = joint_conf.view(bj, dim, h, w).repeat_interleave(nh2, dim=0)
    off = uvd.repeat_interleave(nh2, dim=0)
    if dim == 3:
        diff = torch.cat([heads, heads.new_zeros([heads.size(0), 1, h, w])], dim=1)
    else:
        diff = heads
    stacked = torch.cat([off, conf], dim=1)
    sampled = F.grid_sample(_f32_unless_f64(stacked), 2 * loc - 1, mode="bilinear",
                            padding_mode="zeros", align_corners=False)
    s_off, s_conf = torch.split(sampled, [dim, dim], dim=1)
    per_head = s_off + diff
    wgt = s_conf.reshape(bj, nh2, dim, h, w).softmax(dim=1)
    out = (per_head.reshape(bj, nh2, dim, h, w) * wgt).sum(1)
    return out.view(b, num_joints, dim, h, w)


def refine_branch(feats, uvd, layers, num_joints: int, num_heads: int, dim: int = 3):
    """recursive_update.py:250-255 with the conv feature updates factored out.

    ``feats[k]`` is the feature map layer k projects from (= the reference's
    ``feat + update_feat_conv(feat)`` at recursive_update.py:188; the 3x3
    deformable conv producing it is upstream of the decode boundary).
    """
    b, _, h, w = uvd.shape
    for feat, layer in zip(feats, layers):
        # the reference feature maps are plain NCHW-contiguous tensors (its .view()s require it)
        uvd, so, sc = gated_blend(feat.contiguous(), uvd, layer)
        uvd = progressive_sample(uvd, so, sc, num_joints, num_heads, dim).reshape(b, num_joints * dim, h, w)
    return uvd


def head_eval_tail(pose_raw, feats, layers, scales, *, num_joints, num_heads, root_idx,
                   depth_factor, z_norm, stride):
    """das_head.py:237-262 (eval branch).  Returns the final pose_pred [B, 3+6J, H, W].

    ``scales`` = (scale_offset, scale_depth, scale_uv, scale_d) learnable scalars
    of this level (mmcv ``Scale`` = multiply by an fp32 scalar parameter).
    """
    j3 = 3 * num_joints
    s_off, s_depth, s_uv, s_d = [torch.tensor(float(s), dtype=torch.float32) for s in scales]
    pose = pose_raw.clone()
    pose[:, :2] = pose_raw[:, :2] * s_off
    pose[:, 2] = pose_raw[:, 2] * s_depth
    pose[:, 3:3 + j3:3] = pose_raw[:, 3:3 + j3:3] * s_uv
    pose[:, 4:3 + j3:3] = pose_raw[:, 4:3 + j3:3] * s_uv
    pose[:, 5:3 + j3:3] = pose_raw[:, 5:3 + j3:3] * s_d
    pose[:, 3 + root_idx * 3 + 2] = 0
    pose[:, 3 + j3 + root_idx * 3 + 2] = 1
    ref = refine_branch(feats, pose[:, 3:3 + j3].clone(), layers, num_joints, num_heads)
    ref[:, root_idx * 3 + 2] = 0
    pose[:, 3:3 + j3] = ref
    pose[:, 2] /= depth_factor
    pose[:, 3 + root_idx * 3 + 2] = 0
    pose[:, 3:3 + j3:3] *= stride
    pose[:, 4:3 + j3:3] *= stride
    pose[:, 5:3 + j3:3] *= z_norm
    return pose


# --------------------------------------------------------------------------
# OKS-NMS (NumPy, float64 math on float32 inputs; pose_nms.py:51-126)
# --------------------------------------------------------------------------
def oks_to_head(g, d, a_g, a_d):
    """OKS of every row of ``d`` against ``g`` (flattened x,y,v triples)."""
    nj = len(g) // 3
    sig = COCO17_SIGMAS if nj == 17 else np.ones(nj, dtype=np.float64) * 0.08
    var = (sig * 2) ** 2
    out = np.zeros(len(d), dtype=np.float32)
    for n in range(len(d)):
        dx = d[n, 0::3] - g[0::3]
        dy = d[n, 1::3] - g[1::3]
        e = (dx ** 2 + dy ** 2) / var / ((a_g + a_d[n]) / 2 + np.spacing(1)) / 2
        out[n] = np.sum(np.exp(-e)) / len(e) if len(e) != 0 else 0.0
    return out


def oks_nms(scores, kpts, areas, thr, stable: bool = False, trace: dict | None = None):
    """Greedy OKS suppression; returns kept indices in pick order.

    ``stable=False`` reproduces the reference ordering (``argsort()[::-1]``);
    ``stable=True`` is the tie rule of this repo: equal scores -> lower index first.
    ``trace`` (test bookkeeping, not in the reference): receives 'oks_margin' = the smallest |oks - thr| of any
    suppression decision taken, so a test can tell a robust decision from a coin flip.
    """
    if len(scores) == 0:
        return np.zeros(0, dtype=np.int64)
    if stable:
        order = np.argsort(-scores.astype(np.float64), kind="stable")
    else:
        order = scores.argsort()[::-1]
    keep = []
    while len(order) > 0:
        i = order[0]
        keep.append(i)
        ovr = oks_to_head(kpts[i], kpts[order[1:]], areas[i], areas[order[1:]])
        if trace is not None and len(ovr):
            trace["oks_margin"] = min(trace.get("oks_margin", np.inf), float(np.abs(ovr.astype(np.float64) - thr).min()))
        order = order[np.where(ovr <= thr)[0] + 1]
    return np.array(keep)


def soft_oks_nms(scores, kpts, areas, thr, max_dets, stable: bool = False, trace: dict | None = None):
    """Gaussian soft OKS-NMS (pose_nms.py:129-194): nothing is removed; after every pick the remaining
    scores are multiplied by exp(-oks^2 / thr) (float32) and the best one is taken next.
    ``trace`` receives 'soft_gap_ulps' = the smallest float32-ulp gap between the best and the second-best rescored
    candidate at any pick (how robust the pick order is)."""
    if len(scores) == 0:
        return np.zeros(0, dtype=np.int64)
    order = np.argsort(-scores.astype(np.float64), kind="stable") if stable else scores.argsort()[::-1]
    cur = scores[order]
    keep = []
    while len(order) > 0 and len(keep) < max_dets:
        i = order[0]
        ovr = oks_to_head(kpts[i], kpts[order[1:]], areas[i], areas[order[1:]])
        order = order[1:]
        cur = cur[1:] * np.exp(-ovr ** 2 / thr)
        tmp = np.argsort(-cur.astype(np.float64), kind="stable") if stable else cur.argsort()[::-1]
        order, cur = order[tmp], cur[tmp]
        if trace is not None and len(cur) > 1 and len(keep) + 1 < max_dets:
            a = np.ascontiguousarray(cur[:2], dtype=np.float32).view(np.int32).astype(np.int64)
            trace["soft_gap_ulps"] = min(trace.get("soft_gap_ulps", 1 << 40), int(abs(a[0] - a[1])))
        keep.append(i)
    return np.array(keep)


# --------------------------------------------------------------------------
# decode (das_head.py:653-796)
# --------------------------------------------------------------------------
def decode_image(cls_l, pose_l, ctr_l, strides, scale_factor, cfg, num_joints,
                 stable: bool = False, peak_kernel: int = 0):
    """One image.  cls_l/ctr_l: list of [1,H,W]; pose_l: list of [3+6J,H,W].

    Returns dict(scores [N], poses [N,J,3], vis [N,J], centers [N,3],
    cand_level [N], cand_index [N]) in final (NMS pick) order, plus the
    pre-threshold candidate list under 'cand_*' keys for margin checks.
    """
    j3 = 3 * num_joints
    nms_pre = cfg.get("nms_pre", -1)
    all_c, all_p, all_s, all_lvl, all_idx = [], [], [], [], []
    for lvl, (cls, pose, ctr, stride) in enumerate(zip(cls_l, pose_l, ctr_l, strides)):
        h, w = cls.shape[-2:]
        pts = point_grid(h, w, stride, pose.dtype, pose.device)
        sc = cls.permute(1, 2, 0).reshape(-1, 1).sigmoid()
        ct = ctr.permute(1, 2, 0).reshape(-1).sigmoid()
        pp = pose.permute(1, 2, 0).reshape(-1, pose.shape[0])
        idx = torch.arange(h * w, device=pose.device)
        if nms_pre > 0 and sc.shape[0] > nms_pre:
            rank = (sc * ct[:, None]).max(dim=1)[0]
            if peak_kernel and peak_kernel > 1:
                # north-star option (not in the reference): keep only cells that equal the
                # max of their k x k neighbourhood (out-of-map neighbours ignored).
                m = rank.view(1, 1, h, w)
                pooled = F.max_pool2d(m, peak_kernel, 1, peak_kernel // 2)
                rank = torch.where(m == pooled, m, torch.zeros_like(m)).view(-1)
            if stable:
                idx = torch.sort(rank, descending=True, stable=True)[1][:nms_pre]
            else:
                idx = rank.topk(nms_pre)[1]
            pts, pp, sc, ct = pts[idx], pp[idx], sc[idx], ct[idx]
        pp = pp.clone()
        pp[:, :2] = pts - pp[:, :2]
        centre = pp[:, :3].clone()
        joints = pp[:, 3:3 + j3].reshape(-1, num_joints, 3)
        root = centre[:, None].clone()
        root[:, 0, :2] = pts
        scale = joints.new_tensor(np.asarray(scale_factor[:2], dtype=np.float32))
        q = torch.sqrt(scale.prod())
        root[..., 2] *= q
        centre[..., 2] *= q
        joints = joints + root
        joints[..., :2] = joints[..., :2] / scale
        centre[:, :2] = centre[:, :2] / scale
        all_c.append(centre)
        all_p.append(joints)
        all_s.append(sc[:, 0] * ct)
        all_lvl.append(torch.full((len(idx),), lvl, dtype=torch.int64, device=pose.device))
        all_idx.append(idx.to(torch.int64))
    centres, poses = torch.cat(all_c), torch.cat(all_p)
    scores, lvls, idxs = torch.cat(all_s), torch.cat(all_lvl), torch.cat(all_idx)
    cand = dict(cand_scores=scores.clone(), cand_level=lvls.clone(), cand_index=idxs.clone(),
                cand_poses=poses.clone())
    thr = cfg.get("score_thr", 0.)
    if thr > 0:
        ok = scores > thr
        scores, poses, centres, lvls, idxs = scores[ok], poses[ok], centres[ok], lvls[ok], idxs[ok]
    nms_post = cfg.get("nms_post", -1)
    trace = {}
    if nms_post > 0 and len(scores) > 0:
        hi = poses[..., :2].max(1)[0]
        lo = poses[..., :2].min(1)[0]
        areas = (hi - lo).prod(-1).cpu().numpy()
        kp = torch.cat([poses[..., :2], torch.ones_like(poses[..., :1])], -1).reshape(len(poses), -1).cpu().numpy()
        if cfg.get("nms_type", "hard") == "hard":
            keep = oks_nms(scores.cpu().numpy(), kp, areas, cfg.get("nms_thr", 0.9), stable=stable, trace=trace).tolist()
            keep = keep[:cfg.get("nms_post", 100)]
        else:                                           # das_head.py:789-790
            keep = soft_oks_nms(scores.cpu().numpy(), kp, areas, cfg.get("nms_thr", 0.9), cfg.get("nms_post", 100),
                                stable=stable, trace=trace).tolist()
        scores, poses, centres, lvls, idxs = scores[keep], poses[keep], centres[keep], lvls[keep], idxs[keep]
    out = dict(scores=scores, poses=poses, vis=torch.ones(poses.shape[:2], device=poses.device), centers=centres,
               level=lvls, index=idxs)
    out.update(cand)
    out["oks_margin"] = trace.get("oks_margin", float("inf"))          # test bookkeeping (see oks_nms)
    out["soft_gap_ulps"] = trace.get("soft_gap_ulps", 1 << 40)
    return out


def get_poses(cls_scores, pose_preds, centernesses, img_metas, cfg, strides, num_joints,
              stable: bool = False, peak_kernel: int = 0):
    """Batch decode with the reference's return structure (das_head.py:680-687) plus
    'level'/'index' bookkeeping used by the parity tests."""
    res = []
    for b, meta in enumerate(img_metas):
        r = decode_image([c[b] for c in cls_scores], [p[b] for p in pose_preds],
                         [c[b] for c in centernesses], strides, meta["scale_factor"], cfg,
                         num_joints, stable=stable, peak_kernel=peak_kernel)
        r["image_paths"] = [meta.get("filename", "")]
        r["scores_list"] = r["scores"].cpu().numpy().tolist()
        res.append(r)
    return res


# --------------------------------------------------------------------------
# back-projection (float64, host; vis_3d.py:16-26 + cmupanoptic...:391-401)
# --------------------------------------------------------------------------
def backproject(poses, K, R, t, root_idx: int, dataset_depth_factor: float = 1.0):
    """poses [N,J,3] (image px, normalised depth) -> (cam [N,J,3], world [N,J,3]) float64."""
    p = np.asarray(poses, dtype=np.float64).copy()
    K = np.asarray(K, dtype=np.float64)
    R = np.asarray(R, dtype=np.float64)
    t = np.asarray(t, dtype=np.float64).reshape(3, 1)
    if p.shape[0] == 0:
        return p.copy(), p.copy()
    nd = np.sqrt(K[0, 0] * K[1, 1])
    zr = p[:, [root_idx], 2]
    dz = p[..., 2] - zr
    p[..., 2] = zr * nd + dz
    p[..., 2] *= dataset_depth_factor
    X = p.reshape(-1, 3).T.copy()
    X[0, :] = X[0, :] - K[0, 2]
    X[1, :] = X[1, :] - K[1, 2]
    X[:2] = np.dot(np.linalg.inv(K[:2, :2]), X[:2])
    X[0:2, :] = X[0:2, :] * X[2, :]
    cam = X.copy()
    world = np.dot(np.linalg.inv(R), (X - t))
    return cam.T.reshape(p.shape), world.T.reshape(p.shape)


# --------------------------------------------------------------------------
# whole path: raw head outputs -> final pose lists (what the CUDA path is checked against)
# --------------------------------------------------------------------------
def decode_full(levels, layers, img_metas, head_cfg, test_cfg, stable: bool = False,
                peak_kernel: int = 0):
    """levels: list of dict(cls, ctr, pose_raw, feats=[L x [B,C,H,W]], stride, scales).

    Runs the reference order: dense refinement + eval tail on every level, then decode,
    then back-projection with img_metas[i]['cam'].
    """
    J = head_cfg["num_joints"]
    pose_preds = [head_eval_tail(lv["pose_raw"], lv["feats"], layers, lv["scales"],
                                 num_joints=J, num_heads=head_cfg["num_heads"],
                                 root_idx=head_cfg["root_idx"], depth_factor=head_cfg["depth_factor"],
                                 z_norm=head_cfg["z_norm"], stride=lv["stride"]) for lv in levels]
    res = get_poses([lv["cls"] for lv in levels], pose_preds, [lv["ctr"] for lv in levels],
                    img_metas, test_cfg, [lv["stride"] for lv in levels], J,
                    stable=stable, peak_kernel=peak_kernel)
    for r, meta in zip(res, img_metas):
        cam = meta.get("cam")
        if cam is not None:
            r["poses_cam"], r["poses_world"] = backproject(r["poses"].cpu().numpy(), cam["K"], cam["R"], cam["t"],
                                                           head_cfg["root_idx"])
    return res, pose_preds
