"""TEST INFRASTRUCTURE: the state_dict keys/shapes of the reference's Panoptic model (exp_panoptic.py), so that
das_b200.model.DASNet.load_reference_state_dict can be checked for full coverage without mmcv/mmdet.

  backbone.*   the reference's MSPN2 source, executed under the shims of oracle/make_model_golden.py
  bbox_head.*  the reference's DASHead / AnchorFreeMono3DPoseHead / RecursiveUpdateBranch / RealNVP sources, executed
               under shims that restate the parameter layout of the mmcv pieces they use (ConvModule, DCNv2 pack, Scale)
  neck.*       mmdet 2.14 FPN is not in the reference tree: keys written out from its published module layout
               (lateral_convs.i / fpn_convs.i ConvModules, conv + bn when norm_cfg is given)
Run in the build container only (needs /root/reference):  python oracle/make_state_keys.py
Writes tests/golden/reference_state_keys.json."""
import json
import os
import re
import sys

import torch
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import make_model_golden as G  # noqa: E402

REF = G.REF
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "reference_state_keys.json")


class DCNv2Pack(nn.Module):
    """Parameter layout of mmcv's ModulatedDeformConv2dPack: weight, optional bias, conv_offset (27 channels for 3x3)."""

    def __init__(self, cin, cout, k, padding, bias):
        super().__init__()
        self.weight = nn.Parameter(torch.zeros(cout, cin, k, k))
        self.bias = nn.Parameter(torch.zeros(cout)) if bias else None
        self.conv_offset = nn.Conv2d(cin, 3 * k * k, k, padding=padding)


class ConvModule(nn.Module):
    def __init__(self, cin, cout, kernel_size, stride=1, padding=0, conv_cfg=None, norm_cfg=None, act_cfg="default",
                 inplace=True, bias="auto"):
        super().__init__()
        with_bias = (norm_cfg is None) if bias == "auto" else bias
        if conv_cfg is not None and conv_cfg.get("type") == "DCNv2":
            self.conv = DCNv2Pack(cin, cout, kernel_size, padding, with_bias)
        else:
            self.conv = nn.Conv2d(cin, cout, kernel_size, stride, padding, bias=with_bias)
        if norm_cfg is not None:
            if norm_cfg["type"] == "GN":
                self.gn = nn.GroupNorm(norm_cfg["num_groups"], cout)
            else:
                self.bn = nn.BatchNorm2d(cout)


class Scale(nn.Module):
    def __init__(self, scale=1.0):
        super().__init__()
        self.scale = nn.Parameter(torch.tensor(scale, dtype=torch.float))


class _Base(nn.Module):
    def __init__(self, init_cfg=None):
        super().__init__()


def _strip_imports(src):
    src = re.sub(r"from mmcv\.cnn import \([^)]*\)", "", src, flags=re.S)
    return re.sub(r"^from (mmcv|mmdet|mmdet3d|\.)[^\n]*$", "", src, flags=re.M)


def reference_head():
    ns = dict(ConvModule=ConvModule, Scale=Scale, force_fp32=lambda **kw: (lambda f: f), multi_apply=None,
              HEADS=G._Registry(), build_loss=lambda cfg: nn.Module(), bias_init_with_prob=None, normal_init=None,
              oks_nms=None, soft_oks_nms=None, BaseMono3DDensePoseHead=_Base, __name__="ref_head")
    for rel in ("real_nvp.py", "recursive_update.py", "anchor_free_mono3d_pose_head.py", "das_head.py"):
        path = os.path.join(REF, "mmdet3d/models/pose_heads", rel)
        exec(compile(_strip_imports(open(path).read()), path, "exec"), ns)
    # configs/_base_/models/das.py:24-51 merged with configs/das/exp_panoptic.py:31-44
    return ns["DASHead"](num_classes=1, in_channels=256, stacked_convs=2, feat_channels=256, strides=[8, 16, 32, 64],
                         center_sample_radius=1.5, num_joints=15, cls_branch=(256,),
                         reg_branch=((256,), (256,), (256,), (256,)), centerness_on_reg=True, conv_bias=True,
                         dcn_on_last_conv=True,
                         recursive_update=dict(prev_loss=True, num_heads=4, in_channels=256, feat_channels=256,
                                               num_layers=1, dim=3, num_joints=15),
                         regress_ranges=((-1, 80), (80, 160), (160, 320), (320, 1e8)), depth_factor=20, z_norm=50,
                         root_idx=2)


def fpn_keys(n_lateral=3, n_out=4, ch=256):
    keys = {}
    for name, n, k in (("lateral_convs", n_lateral, 1), ("fpn_convs", n_out, 3)):
        for i in range(n):
            keys[f"neck.{name}.{i}.conv.weight"] = [ch, ch, k, k]
            for p in ("weight", "bias", "running_mean", "running_var"):
                keys[f"neck.{name}.{i}.bn.{p}"] = [ch]
            keys[f"neck.{name}.{i}.bn.num_batches_tracked"] = []
    return keys


def main():
    MSPN2 = G.load_reference_mspn()
    bb = MSPN2(unit_channels=256, num_stages=2, num_units=4, num_blocks=[3, 4, 6, 3], norm_cfg=dict(type="SyncBN"))
    keys = {"backbone." + k: list(v.shape) for k, v in bb.state_dict().items()}
    keys.update(fpn_keys())
    keys.update({"bbox_head." + k: list(v.shape) for k, v in reference_head().state_dict().items()})
    with open(OUT, "w") as f:
        json.dump(keys, f, indent=0, sort_keys=True)
    print("wrote", os.path.normpath(OUT), len(keys), "keys;",
          sum(1 for k in keys if k.startswith("bbox_head.")), "head,", sum(1 for k in keys if k.startswith("backbone.")), "backbone")


if __name__ == "__main__":
    main()
