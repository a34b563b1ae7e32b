"""TEST INFRASTRUCTURE: pins das_b200.model.MSPNBackbone against the reference's own MSPN2 source.

The reference module (mmdet3d/models/backbones/mspn_mmpose.py) imports mmcv / mmdet, which are not installed, so its
source is executed here with those imports replaced by ~40 lines of shims that restate the documented behaviour of the
few mmcv helpers it uses (ConvModule = conv -> norm -> ReLU with `bias='auto'`; build_norm_layer naming `bn<postfix>`;
SyncBN == BatchNorm2d in eval).  Everything else -- the network structure and forward -- is the reference's code.
Run in the build container only (needs /root/reference):  python oracle/make_model_golden.py
Writes tests/golden/mspn_small.npz: state_dict key/shape list, the input recipe, and the four reference output maps."""
import os
import re
import sys

import numpy as np
import torch
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from model_fixture import synthetic_image, synthetic_state  # noqa: E402

REF = os.environ.get("DAS_REFERENCE", "/root/reference")
SRC = os.path.join(REF, "mmdet3d/models/backbones/mspn_mmpose.py")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "mspn_small.npz")


def _norm(cfg, n):
    assert cfg["type"] in ("BN", "SyncBN"), cfg
    return nn.BatchNorm2d(n)


class ConvModule(nn.Module):
    """mmcv.cnn.ConvModule for the arguments mspn_mmpose.py passes: order (conv, norm, act), ReLU unless act_cfg=None,
    conv bias only without a norm."""

    def __init__(self, cin, cout, kernel_size, stride=1, padding=0, norm_cfg=None, act_cfg="default", inplace=True,
                 bias="auto", conv_cfg=None):
        super().__init__()
        assert conv_cfg is None
        self.conv = nn.Conv2d(cin, cout, kernel_size, stride, padding, bias=(norm_cfg is None) if bias == "auto" else bias)
        self.has_norm = norm_cfg is not None
        if self.has_norm:
            self.bn = _norm(norm_cfg, cout)
        self.activate = None if act_cfg is None else nn.ReLU(inplace=inplace)

    def forward(self, x):
        x = self.conv(x)
        if self.has_norm:
            x = self.bn(x)
        return x if self.activate is None else self.activate(x)


def build_norm_layer(cfg, n, postfix=""):
    return "bn" + str(postfix), _norm(cfg, n)


def build_conv_layer(cfg, *a, **kw):
    assert cfg is None
    return nn.Conv2d(*a, **kw)


class _Registry:
    def register_module(self):
        return lambda cls: cls


def load_reference_mspn():
    src = open(SRC).read()
    # drop the imports of absent packages (mmcv, mmdet, the package-relative registry); keep everything else
    src = re.sub(r"from mmcv\.cnn import \([^)]*\)", "", src, flags=re.S)
    src = re.sub(r"^from (mmcv|mmdet|\.\.)[^\n]*$", "", src, flags=re.M)
    ns = dict(ConvModule=ConvModule, MaxPool2d=nn.MaxPool2d, build_conv_layer=build_conv_layer,
              build_norm_layer=build_norm_layer, constant_init=None, kaiming_init=None, normal_init=None,
              BACKBONES=_Registry(), get_root_logger=None, _load_checkpoint=None, load_state_dict=None,
              load_checkpoint=None, __name__="ref_mspn")
    exec(compile(src, SRC, "exec"), ns)
    return ns["MSPN2"]


CFG = dict(unit_channels=64, num_stages=2, num_units=4, num_blocks=[2, 1, 2, 1])
IMG = dict(batch=2, h=64, w=96, seed=77)


def main():
    torch.manual_seed(0)
    torch.set_num_threads(max(1, os.cpu_count() // 2))
    MSPN2 = load_reference_mspn()
    net = MSPN2(norm_cfg=dict(type="SyncBN"), **CFG).eval()
    sd = net.state_dict()
    keys = [k for k in sd if not k.endswith("num_batches_tracked")]
    shapes = [tuple(sd[k].shape) for k in keys]
    net.load_state_dict(synthetic_state(keys, shapes), strict=False)
    img = synthetic_image(**IMG)
    with torch.no_grad():
        outs = net(img)
    assert len(outs) == 4
    blob = dict(keys=np.array(keys), shapes=np.array([",".join(map(str, s)) for s in shapes]),
                cfg=np.array(repr(CFG)), img=np.array(repr(IMG)))
    for i, o in enumerate(outs):
        blob[f"out{i}"] = o.numpy().astype(np.float32)
        print(i, tuple(o.shape), float(o.abs().mean()), float(o.abs().max()))
    np.savez_compressed(OUT, **blob)
    print("wrote", os.path.normpath(OUT), os.path.getsize(OUT) // 1024, "KiB;", len(keys), "tensors,",
          sum(int(np.prod(s)) for s in shapes) / 1e6, "M values")


if __name__ == "__main__":
    main()
