"""Run the UNMODIFIED reference hot-path functions in the build container.  TEST INFRASTRUCTURE ONLY.

The reference package cannot be imported (mmdet3d/__init__.py:2-5 needs mmcv/mmdet/mmseg, none
installed, no network), but the decode functions themselves only need torch + numpy.  This module
pulls their source out of ``/root/reference`` with ``ast`` at run time and ``exec``s it, so the
golden vectors under ``tests/golden/`` come from the reference's own code, not from our
restatement.  Nothing is copied into this repository.

``/root/reference`` does not exist on the GPU box: this module is only used by
``oracle/make_golden.py`` and by CPU tests that skip when the tree is absent.
"""
from __future__ import annotations

import ast
import importlib.util
import os

import numpy as np
import torch
import torch.nn.functional as F

REF = os.environ.get("DAS_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF, "mmdet3d/models/pose_heads/das_head.py"))


def _load_by_path(rel, name):
    if not hasattr(np, "float"):
        np.float = float              # pose_nms.py:72 uses the alias removed in NumPy 1.24
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _class_methods(rel, cls, names):
    tree = ast.parse(open(os.path.join(REF, rel)).read())
    out = {}
    for node in tree.body:
        if isinstance(node, ast.ClassDef) and node.name == cls:
            for fn in node.body:
                if isinstance(fn, ast.FunctionDef) and fn.name in names:
                    fn.decorator_list = []
                    out[fn.name] = fn
    return out


def _module_functions(rel, names):
    tree = ast.parse(open(os.path.join(REF, rel)).read())
    return {n.name: n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in names}


_cache = {}


def pose_nms():
    if "nms" not in _cache:
        _cache["nms"] = _load_by_path("mmdet3d/core/post_processing/pose_nms.py", "ref_pose_nms")
    return _cache["nms"]


def vis_3d():
    if "vis" not in _cache:
        _cache["vis"] = _load_by_path("mytools/vis_3d.py", "ref_vis_3d")
    return _cache["vis"]


def head(num_joints, strides, test_cfg):
    """A stub object carrying the reference's get_poses/_get_poses_single/get_points methods."""
    nms = pose_nms()
    ns = dict(torch=torch, np=np, oks_nms=nms.oks_nms, soft_oks_nms=nms.soft_oks_nms, INF=1e8)

    class Base:
        pass

    for k, fn in _class_methods("mmdet3d/models/pose_heads/anchor_free_mono3d_pose_head.py",
                                "AnchorFreeMono3DPoseHead", ["get_points", "_get_points_single"]).items():
        exec(ast.unparse(fn), ns)
        setattr(Base, k, ns[k])

    class Head(Base):
        pass

    ns["Head"] = Head
    for k, fn in _class_methods("mmdet3d/models/pose_heads/das_head.py", "DASHead",
                                ["get_poses", "_get_poses_single", "_get_points_single"]).items():
        exec(ast.unparse(fn).replace("super()", "super(Head, self)"), ns)
        setattr(Head, k, ns[k])
    h = Head()
    h.training = False
    h.num_joints = num_joints
    h.cls_out_channels = 1
    h.group_reg_dims = [2, 1, 3 * num_joints, 3 * num_joints]
    h.strides = list(strides)
    h.test_cfg = dict(test_cfg)
    return h


def offset_sample_fn():
    if "os" not in _cache:
        ns = dict(torch=torch, F=F)
        for k, fn in _module_functions("mmdet3d/models/pose_heads/recursive_update.py",
                                       ["offset_sample", "offset_sample_core"]).items():
            exec(ast.unparse(fn), ns)
        _cache["os"] = ns["offset_sample"]
    return _cache["os"]


def refined_pose_pred(level, layers, head_cfg):
    """Reference refinement + eval tail for one level.

    The sampling (recursive_update.py:9-82) is the reference's own code; the mmcv-bound
    lines around it (Scale modules das_head.py:237-250, 1x1 convs + gate recursive_update.py:190-195,
    eval tail das_head.py:256-262) are driven here with plain torch modules exactly as the
    SURVEY appendix describes.
    """
    J, nh, root = head_cfg["num_joints"], head_cfg["num_heads"], head_cfg["root_idx"]
    osamp = offset_sample_fn()
    pose_pred = level["pose_raw"].clone()
    s_off, s_depth, s_uv, s_d = [torch.tensor(float(s)) for s in level["scales"]]
    clone = pose_pred.clone()
    pose_pred[:, :2] = (clone[:, :2] * s_off).float()
    pose_pred[:, 2] = (clone[:, 2] * s_depth).float()
    cuvd = clone[:, 3:3 + J * 3]
    uvd = pose_pred[:, 3:3 + J * 3]
    uvd[:, 0::3] = cuvd[:, 0::3] * s_uv
    uvd[:, 1::3] = cuvd[:, 1::3] * s_uv
    uvd[:, 2::3] = cuvd[:, 2::3] * s_d
    pose_pred[:, 3 + root * 3 + 2] = 0
    pose_pred[:, 3 + J * 3 + root * 3 + 2] = 1
    offset = pose_pred[:, 3:3 + J * 3].clone()
    b, _, h, w = offset.shape
    ys, xs = torch.meshgrid(torch.arange(h, dtype=offset.dtype), torch.arange(w, dtype=offset.dtype), indexing="ij")
    pts = torch.stack((xs, ys), dim=0) + 0.5
    for feat, lw in zip(level["feats"], layers):
        feat = feat.contiguous()
        convs = {}
        for k in ("so", "sc", "uw", "uv"):
            wt = lw[k + "_w"]
            c = torch.nn.Conv2d(wt.shape[1], wt.shape[0], 1)
            c.weight.data.copy_(wt[:, :, None, None])
            c.bias.data.copy_(lw[k + "_b"])
            convs[k] = c
        with torch.no_grad():
            so = convs["so"](feat)
            sc = convs["sc"](feat)
            gate = convs["uw"](feat).sigmoid()
            nxt = convs["uv"](feat)
            offset = (1 - gate) * offset + gate * nxt
            new, _ = osamp(offset, so, sc, (b, J, nh, 3), pts)
        offset = new.reshape(b, J * 3, h, w)
    ref_uvd = offset
    ref_uvd[:, root * 3 + 2] = 0
    pose_pred[:, 3:3 + J * 3] = ref_uvd
    pose_pred[:, 2] /= head_cfg["depth_factor"]
    pose_pred[:, 3 + root * 3 + 2] = 0
    pose_pred[:, 3:3 + J * 3:3] *= level["stride"]
    pose_pred[:, 4:3 + J * 3:3] *= level["stride"]
    pose_pred[:, 5:3 + J * 3:3] *= head_cfg["z_norm"]
    return pose_pred


def backproject(poses, cam, root_idx, dataset_depth_factor=1.0):
    """cmupanoptic_mono_dataset.py:391-402 de-norm (restated: it lives inside a dataset method that
    needs pycocotools) followed by the reference's own pixel2world."""
    pred_img = np.asarray(poses, dtype=np.float64).copy()
    if pred_img.shape[0] == 0:
        return pred_img, pred_img.copy()
    K = np.array(cam["K"], dtype=np.float64)
    norm_depth = np.sqrt(K[0, 0] * K[1, 1])
    root_depth = pred_img[:, [root_idx], 2]
    dz = pred_img[..., 2] - root_depth
    pred_img[..., 2] = root_depth * norm_depth + dz
    pred_img[..., 2] *= dataset_depth_factor
    x1, x2, x3 = vis_3d().pixel2world(pred_img.reshape(-1, 3).T, K, np.array(cam["R"], dtype=np.float64),
                                      np.array(cam["t"], dtype=np.float64).reshape(3, 1))
    return x2.T.reshape(pred_img.shape), x3.T.reshape(pred_img.shape)
