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
        self.padding = padding

    def forward(self, x):
        """mmcv's documented pack forward: one conv predicts (o1, o2, mask); offset = cat(o1, o2); mask = sigmoid(mask).
        The deformable convolution itself is torchvision's here (mmcv's CUDA op is absent) -- so this pins the WIRING of
        the towers, not the DCNv2 kernel."""
        from torchvision.ops import deform_conv2d
        o1, o2, mask = torch.chunk(self.conv_offset(x), 3, dim=1)
        return deform_conv2d(x, torch.cat((o1, o2), dim=1), self.weight, self.bias, padding=self.padding, mask=torch.sigmoid(mask))


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
        self.relu = act_cfg is not None

    def forward(self, x):
        x = self.conv(x)
        x = self.gn(x) if hasattr(self, "gn") else self.bn(x) if hasattr(self, "bn") else x
        return torch.relu(x) if self.relu else x


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


def make_head_golden():
    """Conv part of the reference head's forward (das_head.py:180-230, recursive_update.py:186-188, 250-252) on one small
    feature map, with weights from the shared synthetic recipe -> tests/golden/das_head_small.npz."""
    import numpy as np
    from model_fixture import synthetic_state
    head = reference_head().eval()
    sd = head.state_dict()
    keys = [k for k in sd if not k.startswith(("flow", "loss_"))]
    full = synthetic_state(["bbox_head." + k for k in keys], [tuple(sd[k].shape) for k in keys])
    head.load_state_dict({k: full["bbox_head." + k] for k in keys}, strict=False)
    g = torch.Generator().manual_seed(91)
    x = torch.randn(2, 256, 12, 16, generator=g)
    with torch.no_grad():
        cls, pose, cls_feat, reg_feat, pose_feat = head._forward_single(x)
        ctr = head._forward_centerness(cls_feat, reg_feat)
        f0 = head.recursive_update_branch.reduction(pose_feat)
        f1 = f0 + head.recursive_update_branch.layer_0.next_level_offset._update_feat(f0)
    out = os.path.join(os.path.dirname(OUT), "das_head_small.npz")
    np.savez_compressed(out, cls=cls.numpy(), pose=pose.numpy(), ctr=ctr.numpy(), feat=f1.numpy(),
                        keys=np.array(["bbox_head." + k for k in keys]),
                        shapes=np.array([",".join(map(str, sd[k].shape)) for k in keys]))
    print("wrote", os.path.normpath(out), os.path.getsize(out) // 1024, "KiB;", float(cls.abs().mean()), float(pose.abs().mean()),
          float(f1.abs().mean()))


def main():
    make_head_golden()
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
