"""Image -> decode inputs: the step *before* the hot path (BASELINE config #5).

Written from scratch for inference on one B200 per process: PyTorch/cuDNN convolutions in channels-last storage,
BatchNorm folded into the preceding convolution, optional bf16 autocast for the backbone/neck/towers, and -- the
point of the exercise -- the head emits exactly what the CUDA decode wants (raw fp32 predictor maps + NHWC refinement
features), so the reference's dense refinement (recursive_update.py:34-82, ~77 MB of temporaries per image) never runs.

What it mirrors (behaviour, not code):
  * MSPNBackbone      <- MSPN2 (mmdet3d/models/backbones/mspn_mmpose.py:17-667; config exp_panoptic.py:13-23)
  * FPNNeck           <- mmdet 2.14 FPN as configured in configs/_base_/models/das.py:16-23 + exp_panoptic.py:24-30
                         (not in the reference tree; restated from the published algorithm)
  * DASTowers         <- DASHead conv stack (das_head.py:103-230, anchor_free_mono3d_pose_head.py:100-249) and
                         RecursiveUpdateBranch's conv part (recursive_update.py:171-180, 243-255)
  * DASNet            <- DAS.extract_feat + bbox_head forward (detectors/das.py:34-39)

Parity status: MSPNBackbone is pinned against the reference's own MSPN2 source executed under mmcv shims
(oracle/make_model_golden.py -> tests/golden/mspn_small.npz).  DASTowers is pinned against the reference head's own
forward code (das_head.py:180-230 + the RecursiveUpdateBranch conv part) executed under shims
(oracle/make_state_keys.py -> tests/golden/das_head_small.npz), with ONE substitution: mmcv's DCNv2 CUDA op is absent, so
both sides use torchvision.ops.deform_conv2d with mmcv's (o1, o2, mask) channel split -- the wiring, GroupNorm / bias
conventions and the key map are pinned, the DCNv2 kernel's offset ordering vs mmcv is not.  FPN (mmdet, not in the tree)
is restated from the published algorithm and checked against torchvision's FeaturePyramidNetwork.
`load_reference_state_dict` maps the reference's checkpoint keys onto these modules (all 1200 keys of the Panoptic
model, tests/golden/reference_state_keys.json).
"""
from __future__ import annotations

import re
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F


class ConvUnit(nn.Module):
    """conv -> optional norm ('bn' | 'gn') -> optional ReLU.  Without an explicit `bias` the convolution carries one
    only when there is no norm (the ConvModule 'auto' rule the reference relies on)."""

    def __init__(self, cin: int, cout: int, k: int, stride: int = 1, norm: Optional[str] = "bn", act: bool = True,
                 bias: Optional[bool] = None):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, k, stride, k // 2, bias=(norm is None) if bias is None else bias)
        self.norm = nn.BatchNorm2d(cout) if norm == "bn" else nn.GroupNorm(32, cout) if norm == "gn" else None
        self.act = act
        self.fused = False          # set by DASNet.prepare_inference on CUDA

    def forward(self, x):
        x = x.to(self.conv.weight.dtype)
        if self.fused and self.act and self.norm is None and x.is_cuda:
            c = self.conv       # cuDNN's fused conv + bias + ReLU: no separate bias-add / ReLU passes over the map
            return torch.cudnn_convolution_relu(x, c.weight, c.bias, c.stride, c.padding, c.dilation, 1)
        x = self.conv(x)
        if self.norm is not None:
            x = self.norm(x)
        return F.relu(x, inplace=True) if self.act else x

    @torch.no_grad()
    def fold_batchnorm(self):
        """Inference: absorb an eval-mode BatchNorm into the convolution (what tools/misc/fuse_conv_bn.py:9-22 does)."""
        bn = self.norm
        if not isinstance(bn, nn.BatchNorm2d):
            return
        g = bn.weight / torch.sqrt(bn.running_var + bn.eps)
        w = self.conv.weight * g.view(-1, 1, 1, 1)
        b = bn.bias - bn.running_mean * g
        if self.conv.bias is not None:
            b = b + self.conv.bias * g
        fused = nn.Conv2d(self.conv.in_channels, self.conv.out_channels, self.conv.kernel_size, self.conv.stride,
                          self.conv.padding, bias=True).to(w.device, w.dtype)
        fused.weight.copy_(w)
        fused.bias.copy_(b)
        self.conv, self.norm = fused, None


class DeformUnit(nn.Module):
    """3x3 modulated deformable conv (DCNv2, one deformable group) -> GroupNorm(32) -> ReLU.
    The offset/mask predictor is a plain 3x3 conv with 27 outputs, zero-initialised so a fresh unit is an ordinary
    convolution (mmcv's ModulatedDeformConv2dPack convention).  Runs in fp32 like the reference (`force_fp32`)."""

    def __init__(self, cin: int, cout: int, bias: bool):
        super().__init__()
        self.offset_mask = nn.Conv2d(cin, 27, 3, 1, 1)
        nn.init.zeros_(self.offset_mask.weight)
        nn.init.zeros_(self.offset_mask.bias)
        self.weight = nn.Parameter(torch.empty(cout, cin, 3, 3))
        nn.init.kaiming_uniform_(self.weight, a=5 ** 0.5)
        self.bias = nn.Parameter(torch.zeros(cout)) if bias else None
        self.norm = nn.GroupNorm(32, cout)
        self.tf32 = False           # set by DASNet.prepare_inference: the column GEMM on TF32 tensor cores

    def forward(self, x):
        from torchvision.ops import deform_conv2d
        x = x.float()
        om = self.offset_mask(x)
        first, second, mask = torch.chunk(om, 3, dim=1)
        was = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = was or self.tf32
        try:
            y = deform_conv2d(x, torch.cat((first, second), 1), self.weight, self.bias, padding=1, mask=torch.sigmoid(mask))
        finally:
            torch.backends.cuda.matmul.allow_tf32 = was
        return F.relu(self.norm(y), inplace=True)


# ------------------------------------------------------------------------------------------ backbone

class Bottleneck(nn.Module):
    """1x1 -> 3x3 (carries the stride) -> 1x1 residual block, output = 4 x mid channels."""

    def __init__(self, cin: int, mid: int, stride: int = 1):
        super().__init__()
        cout = 4 * mid
        self.reduce = ConvUnit(cin, mid, 1)
        self.spatial = ConvUnit(mid, mid, 3, stride)
        self.expand = ConvUnit(mid, cout, 1, act=False)
        self.shortcut = ConvUnit(cin, cout, 1, stride, act=False) if (stride != 1 or cin != cout) else None
        self.tail_bias = None       # expand bias (+ shortcut bias) once BatchNorm is folded (prepare_inference)

    def forward(self, x):
        h = self.spatial(self.reduce(x))
        if self.tail_bias is not None and h.is_cuda:
            # relu(expand(h) + identity) as ONE cuDNN call: conv + residual add + bias + ReLU
            e = self.expand.conv
            x = x.to(e.weight.dtype)
            z = x if self.shortcut is None else F.conv2d(x, self.shortcut.conv.weight, None, self.shortcut.conv.stride)
            return torch.cudnn_convolution_add_relu(h, e.weight, z, 1.0, self.tail_bias, e.stride, e.padding, e.dilation, 1)
        y = self.expand(h)
        y = y + (x.to(y.dtype) if self.shortcut is None else self.shortcut(x))
        return F.relu(y, inplace=True)


class UpUnit(nn.Module):
    """One rung of the hourglass decoder: 1x1 on the encoder map, plus the 1x1 of the bilinearly (align_corners=True)
    upsampled coarser rung, ReLU.  Optionally emits the two skips for the next stage's encoder and, on the finest
    rung, the 64-channel map the next stage starts from."""

    def __init__(self, rung: int, n_rungs: int, cin: int, width: int, feeds_next: bool, stem: int):
        super().__init__()
        self.lateral = ConvUnit(cin, width, 1, act=False)
        self.from_coarse = ConvUnit(width, width, 1, act=False) if rung > 0 else None
        self.skip_enc = ConvUnit(cin, cin, 1) if feeds_next else None
        self.skip_dec = ConvUnit(width, cin, 1) if feeds_next else None
        self.to_next = ConvUnit(width, stem, 1) if (feeds_next and rung == n_rungs - 1) else None
        self.merge_bias = None      # lateral bias (+ from_coarse bias) once BatchNorm is folded (prepare_inference)

    def forward(self, enc, coarse):
        if self.merge_bias is not None and enc.is_cuda:
            l = self.lateral.conv
            enc = enc.to(l.weight.dtype)
            if self.from_coarse is None:
                y = torch.cudnn_convolution_relu(enc, l.weight, self.merge_bias, l.stride, l.padding, l.dilation, 1)
            else:
                up = F.interpolate(coarse, size=enc.shape[-2:], mode="bilinear", align_corners=True)
                z = F.conv2d(up.to(l.weight.dtype), self.from_coarse.conv.weight, None)
                y = torch.cudnn_convolution_add_relu(enc, l.weight, z, 1.0, self.merge_bias, l.stride, l.padding, l.dilation, 1)
        else:
            y = self.lateral(enc)
            if self.from_coarse is not None:
                up = F.interpolate(coarse, size=enc.shape[-2:], mode="bilinear", align_corners=True)
                y = y + self.from_coarse(up)
            y = F.relu(y, inplace=True)
        s1 = self.skip_enc(enc) if self.skip_enc is not None else None
        s2 = self.skip_dec(y) if self.skip_dec is not None else None
        nxt = self.to_next(y) if self.to_next is not None else None
        return y, s1, s2, nxt


class Hourglass(nn.Module):
    """One MSPN stage: a ResNet-style encoder (n_rungs groups of bottlenecks, mid channels stem * 2^u, stride 2 from
    the second group on) and the UpUnit decoder.  A stage after the first adds the previous stage's two skips to
    every encoder output."""

    def __init__(self, blocks: Sequence[int], width: int, takes_skips: bool, feeds_next: bool, stem: int):
        super().__init__()
        self.takes_skips = takes_skips
        n = len(blocks)
        self.encoder = nn.ModuleList()
        cin = stem
        for u, nb in enumerate(blocks):
            mid = stem << u
            group = [Bottleneck(cin, mid, 1 if u == 0 else 2)]
            cin = 4 * mid
            group += [Bottleneck(cin, mid) for _ in range(nb - 1)]
            self.encoder.append(nn.Sequential(*group))
        # decoder rung r works on encoder group n-1-r (coarsest first)
        self.decoder = nn.ModuleList([UpUnit(r, n, 4 * (stem << (n - 1 - r)), width, feeds_next, stem) for r in range(n)])

    def forward(self, x, skips):
        enc = []
        for u, group in enumerate(self.encoder):
            x = group(x)
            if self.takes_skips:
                x = x + skips[0][u] + skips[1][u]
            enc.append(x)
        outs, s1, s2, nxt, coarse = [], [], [], None, None
        for r, unit in enumerate(self.decoder):
            coarse, a, b, t = unit(enc[len(enc) - 1 - r], coarse)
            outs.append(coarse)
            s1.append(a)
            s2.append(b)
            nxt = t if t is not None else nxt
        return outs, (s1[::-1], s2[::-1]), nxt


class MSPNBackbone(nn.Module):
    """Multi-stage pose network: 7x7/2 stem + 3x3/2 max-pool, then `num_stages` hourglasses chained through a
    64-channel cross map and per-rung skips.  Returns the LAST stage's decoder maps, finest first (strides 4, 8, 16, 32
    for four rungs), `unit_channels` channels each (mspn_mmpose.py:646-654)."""

    def __init__(self, unit_channels: int = 256, num_stages: int = 2, num_blocks: Sequence[int] = (3, 4, 6, 3),
                 stem_channels: int = 64):
        super().__init__()
        assert num_stages >= 1 and len(num_blocks) >= 2
        self.stem = ConvUnit(3, stem_channels, 7, 2)
        self.stages = nn.ModuleList([
            Hourglass(num_blocks, unit_channels, takes_skips=i > 0, feeds_next=i < num_stages - 1, stem=stem_channels)
            for i in range(num_stages)])

    def forward(self, img):
        x = F.max_pool2d(self.stem(img), 3, 2, 1)
        skips, outs = None, None
        for stage in self.stages:
            outs, skips, x = stage(x, skips)
        return outs[::-1]


# ------------------------------------------------------------------------------------------ neck

class FPNNeck(nn.Module):
    """Feature pyramid over backbone maps start_level.. : 1x1 laterals, nearest-neighbour top-down sums, a 3x3 on every
    merged map, then extra stride-2 3x3 convs fed from the last OUTPUT ('on_output') until num_outs maps exist, with a
    ReLU before every extra conv except the first (relu_before_extra_convs).  With a norm, convs have no bias and there
    is no activation (configs/_base_/models/das.py:16-23, exp_panoptic.py:24-30: SyncBN, num_outs=4, start_level=1)."""

    def __init__(self, in_channels: Sequence[int] = (256, 256, 256, 256), out_channels: int = 256, start_level: int = 1,
                 num_outs: int = 4, norm: Optional[str] = "bn", relu_before_extra_convs: bool = True):
        super().__init__()
        self.start_level = start_level
        self.relu_before_extra = relu_before_extra_convs
        used = list(in_channels)[start_level:]
        assert num_outs >= len(used)
        self.lateral = nn.ModuleList([ConvUnit(c, out_channels, 1, norm=norm, act=False) for c in used])
        self.smooth = nn.ModuleList([ConvUnit(out_channels, out_channels, 3, norm=norm, act=False) for _ in used])
        self.extra = nn.ModuleList([ConvUnit(out_channels, out_channels, 3, 2, norm=norm, act=False)
                                    for _ in range(num_outs - len(used))])

    def forward(self, feats):
        lat = [l(f) for l, f in zip(self.lateral, feats[self.start_level:])]
        for i in range(len(lat) - 1, 0, -1):
            lat[i - 1] = lat[i - 1] + F.interpolate(lat[i], size=lat[i - 1].shape[-2:], mode="nearest")
        outs = [s(x) for s, x in zip(self.smooth, lat)]
        for i, conv in enumerate(self.extra):
            src = outs[-1]
            outs.append(conv(F.relu(src) if (self.relu_before_extra and i > 0) else src))
        return outs


# ------------------------------------------------------------------------------------------ head conv stack

def _tower(cin: int, width: int, depth: int, deform_last: bool) -> nn.Sequential:
    units = []
    for i in range(depth):
        c = cin if i == 0 else width
        if deform_last and i == depth - 1:
            units.append(DeformUnit(c, width, bias=True))
        else:
            units.append(ConvUnit(c, width, 3, norm="gn", bias=True))
    return nn.Sequential(*units)


class RefineLayerConvs(nn.Module):
    """Conv part of one refinement layer (recursive_update.py:171-197): F <- F + ReLU(GN(DCNv2(F))) and the four 1x1
    projections, which are NOT applied here -- their weights are handed to the decode kernels."""

    def __init__(self, channels: int, num_joints: int, num_heads: int):
        super().__init__()
        self.update = DeformUnit(channels, channels, bias=False)
        self.sampling_offset = nn.Conv2d(channels, num_joints * num_heads * 2, 1)
        self.sampling_conf = nn.Conv2d(channels, num_joints * 3, 1)
        self.update_weight = nn.Conv2d(channels, num_joints * 3, 1)
        self.update_offset_value = nn.Conv2d(channels, num_joints * 3, 1)
        nn.init.normal_(self.sampling_offset.weight, 0.0, 1e-2)
        nn.init.zeros_(self.sampling_offset.bias)

    def projection_weights(self) -> Dict[str, torch.Tensor]:
        out = {}
        for tag, m in (("so", self.sampling_offset), ("sc", self.sampling_conf), ("uw", self.update_weight),
                       ("uv", self.update_offset_value)):
            out[tag + "_w"] = m.weight.detach().reshape(m.out_channels, m.in_channels)
            out[tag + "_b"] = m.bias.detach()
        return out


class DASTowers(nn.Module):
    """The head's convolutions, shared by all levels: three towers (centre / regression / pose) of `stacked_convs`
    3x3+GN+ReLU units, the last one deformable; one 3x3+GN+ReLU 'prev' unit and a 1x1 predictor per output group
    (centre logit; root offset 2; root depth 1; uvd 3J; sigma 3J); centerness from the regression tower through a
    64-channel unit; the refinement branch's 1x1+GN+ReLU reduction and per-layer deformable feature update.
    forward(x) -> cls [B,1,H,W], raw pose [B,3+6J,H,W] (no Scale, no tail: the decode applies them), centerness
    [B,1,H,W], feats = [F_1 .. F_L] channels-last fp32."""

    def __init__(self, num_joints: int = 15, in_channels: int = 256, feat_channels: int = 256, stacked_convs: int = 2,
                 branch_channels: int = 256, centerness_channels: int = 64, refine_channels: int = 256,
                 num_layers: int = 1, num_heads: int = 4, dcn_on_last_conv: bool = True, with_sigma: bool = False):
        super().__init__()
        J = num_joints
        self.num_joints, self.with_sigma = J, with_sigma
        self.cls_tower = _tower(in_channels, feat_channels, stacked_convs, dcn_on_last_conv)
        self.reg_tower = _tower(in_channels, feat_channels, stacked_convs, dcn_on_last_conv)
        self.pose_tower = _tower(in_channels, feat_channels, stacked_convs, dcn_on_last_conv)

        def branch(width, out):
            return nn.Sequential(ConvUnit(feat_channels, width, 3, norm="gn", bias=True), nn.Conv2d(width, out, 1))
        self.cls_out = branch(branch_channels, 1)
        self.offset_out = branch(branch_channels, 2)
        self.depth_out = branch(branch_channels, 1)
        self.uvd_out = branch(branch_channels, 3 * J)
        self.sigma_out = branch(branch_channels, 3 * J)
        self.centerness_out = branch(centerness_channels, 1)
        self.reduction = ConvUnit(feat_channels, refine_channels, 1, norm="gn")
        self.layers = nn.ModuleList([RefineLayerConvs(refine_channels, J, num_heads) for _ in range(num_layers)])
        for m in (self.cls_out, self.offset_out, self.depth_out, self.uvd_out, self.sigma_out, self.centerness_out):
            nn.init.normal_(m[1].weight, 0.0, 0.01)
            nn.init.zeros_(m[1].bias)
        nn.init.constant_(self.cls_out[1].bias, -4.59511985)      # bias_prob 0.01 (das_head.py:94-99)

    def forward(self, x):
        cls_feat, reg_feat, pose_feat = self.cls_tower(x), self.reg_tower(x), self.pose_tower(x)

        def predict(branch, feat):                # 3x3+GN+ReLU unit in the network dtype, the 1x1 predictor always in fp32
            return branch[1](branch[0](feat).float())
        cls = predict(self.cls_out, cls_feat)
        ctr = predict(self.centerness_out, reg_feat)
        off, dep, uvd = predict(self.offset_out, reg_feat), predict(self.depth_out, reg_feat), predict(self.uvd_out, pose_feat)
        if self.with_sigma:
            sig = predict(self.sigma_out, pose_feat)
        else:                                     # sigma is never read at test time (das_head.py:732)
            sig = uvd.new_zeros(uvd.shape)
        pose = torch.cat((off, dep, uvd, sig), 1).contiguous()
        f = self.reduction(pose_feat)
        feats = []
        for layer in self.layers:
            f = f.float() + layer.update(f)
            feats.append(f.contiguous(memory_format=torch.channels_last))
        return cls.contiguous(), pose, ctr.contiguous(), feats


# ------------------------------------------------------------------------------------------ whole network

class DASNet(nn.Module):
    """backbone -> neck -> towers; `forward(img)` returns `(cls_scores, raw_pose_preds, centernesses, refine_feats)`,
    each a list over levels, i.e. the extended `DASHeadB200.get_poses` call (`head.get_poses(*outs, img_metas)`)."""

    def __init__(self, num_joints: int = 15, strides: Sequence[int] = (8, 16, 32, 64),
                 backbone: Optional[dict] = None, fpn_channels: int = 256, num_layers: int = 1, num_heads: int = 4,
                 stacked_convs: int = 2, with_sigma: bool = False):
        super().__init__()
        bb = dict(unit_channels=256, num_stages=2, num_blocks=(3, 4, 6, 3))
        bb.update(backbone or {})
        self.strides = list(strides)
        self.backbone = MSPNBackbone(**bb)
        n_maps = len(bb["num_blocks"])
        self.neck = FPNNeck([bb["unit_channels"]] * n_maps, fpn_channels, start_level=1, num_outs=len(self.strides))
        self.towers = DASTowers(num_joints, fpn_channels, fpn_channels, stacked_convs, fpn_channels, 64, fpn_channels,
                                num_layers, num_heads, with_sigma=with_sigma)
        # learnable per-level Scale factors for (offset, depth, uv, d), init 1 (das_head.py:171-173)
        self.scales = nn.Parameter(torch.ones(len(self.strides), 4))
        self.compute_dtype: Optional[torch.dtype] = None

    @torch.no_grad()
    def prepare_inference(self, dtype: Optional[torch.dtype] = None):
        """eval mode, BatchNorm folded, channels-last weights, cuDNN fused conv+bias(+residual)+ReLU calls, DCNv2 column
        GEMMs on TF32 tensor cores; `dtype=torch.bfloat16` stores and runs the plain convolutions in bf16 (the 1x1
        predictors, DCNv2 and the refinement features stay fp32)."""
        self.eval()
        for m in self.modules():
            if isinstance(m, ConvUnit):
                m.fold_batchnorm()
        self.to(memory_format=torch.channels_last)
        self.compute_dtype = dtype
        for m in self.modules():
            if isinstance(m, ConvUnit):
                m.fused = True
                if dtype is not None:
                    m.to(dtype)                   # conv (+ GroupNorm affine) in the network dtype; predictors / DCNv2 stay fp32
            elif isinstance(m, DeformUnit):
                m.tf32 = True
        for m in self.modules():
            if isinstance(m, Bottleneck):
                b = m.expand.conv.bias
                m.tail_bias = (b if m.shortcut is None else b + m.shortcut.conv.bias).detach().clone()
            elif isinstance(m, UpUnit):
                b = m.lateral.conv.bias
                m.merge_bias = (b if m.from_coarse is None else b + m.from_coarse.conv.bias).detach().clone()
        return self

    def level_scales(self) -> List[Tuple[float, float, float, float]]:
        return [tuple(float(v) for v in row) for row in self.scales.detach().cpu()]

    def refine_weights(self) -> List[Dict[str, torch.Tensor]]:
        """Per-layer 1x1 projection weights in the form DASHeadB200.load_refine_weights takes."""
        return [l.projection_weights() for l in self.towers.layers]

    def forward(self, img):
        img = img.contiguous(memory_format=torch.channels_last)
        pyramid = self.neck(self.backbone(img))
        per_level = [self.towers(x) for x in pyramid]
        cls, pose, ctr, feats = (list(t) for t in zip(*per_level))
        return cls, pose, ctr, feats

    # -------------------------------------------------------------------------------- reference checkpoints
    def load_reference_state_dict(self, state: Dict[str, torch.Tensor], strict: bool = True):
        """Load a checkpoint written by the reference (keys `backbone.* / neck.* / bbox_head.*`, optionally under
        `module.` or in a `state_dict` entry).  Must be called before `prepare_inference` (BatchNorm still present)."""
        if "state_dict" in state:
            state = state["state_dict"]
        mine = self.state_dict()
        loaded, unknown = set(), []
        for key, val in state.items():
            k = key[7:] if key.startswith("module.") else key
            if k.endswith("num_batches_tracked") or k.startswith(_TRAINING_ONLY):
                continue
            tgt = reference_key_to_local(k, n_lateral=len(self.neck.lateral))
            if tgt is not None and tgt not in mine:      # a DCNv2 pack keeps its own weight/bias where a plain unit has .conv
                alt = tgt.replace(".conv.weight", ".weight").replace(".conv.bias", ".bias")
                tgt = alt if alt in mine else tgt
            if tgt == "scales":
                m = re.match(r"bbox_head\.scales\.(\d+)\.(\d+)\.scale", k)
                self.scales.data[int(m.group(1)), int(m.group(2))] = float(val)
                loaded.add("scales")
                continue
            if tgt is None or tgt not in mine:
                unknown.append(key)
                continue
            assert mine[tgt].shape == val.shape, (key, tgt, tuple(mine[tgt].shape), tuple(val.shape))
            mine[tgt].copy_(val)
            loaded.add(tgt)
        missing = [k for k in mine if k not in loaded and not k.endswith("num_batches_tracked")]
        if strict and (unknown or missing):
            raise KeyError(f"unmapped checkpoint keys {unknown[:5]} (+{max(0, len(unknown) - 5)}), "
                           f"uninitialised parameters {missing[:5]} (+{max(0, len(missing) - 5)})")
        return missing, unknown


# training-only state of the reference head: the RLE loss's normalising flows (das_head.py:94-98) and loss modules
_TRAINING_ONLY = ("bbox_head.flow2d", "bbox_head.flow3d", "bbox_head.loss_")
_UNIT = {"conv": "conv", "bn": "norm", "gn": "norm"}
_BLOCK = {"conv1": "reduce.conv", "bn1": "reduce.norm", "conv2": "spatial.conv", "bn2": "spatial.norm",
          "conv3": "expand.conv", "bn3": "expand.norm"}
_UP = {"in_skip": "lateral", "up_conv": "from_coarse", "out_skip1": "skip_enc", "out_skip2": "skip_dec",
       "cross_conv": "to_next"}
_BRANCH = {"conv_cls_prev.0": "cls_out.0", "conv_cls": "cls_out.1",
           "conv_reg_prevs.0.0": "offset_out.0", "conv_regs.0": "offset_out.1",
           "conv_reg_prevs.1.0": "depth_out.0", "conv_regs.1": "depth_out.1",
           "conv_pose_prevs.0.0": "uvd_out.0", "conv_poses.0": "uvd_out.1",
           "conv_pose_prevs.1.0": "sigma_out.0", "conv_poses.1": "sigma_out.1",
           "conv_centerness_prev.0": "centerness_out.0", "conv_centerness": "centerness_out.1"}
_TOWER = {"cls_convs": "cls_tower", "reg_convs": "reg_tower", "pose_convs": "pose_tower"}


def _unit_tail(rest: str) -> Optional[str]:
    """'<conv|bn|gn>.<param>' of a reference ConvModule -> the ConvUnit / DeformUnit parameter path."""
    parts = rest.split(".")
    if parts[0] == "conv" and len(parts) == 3 and parts[1] == "conv_offset":      # DCNv2 pack: offset/mask predictor
        return "offset_mask." + parts[2]
    if parts[0] in _UNIT and len(parts) == 2:
        return _UNIT[parts[0]] + "." + parts[1]
    return None


def reference_key_to_local(k: str, n_lateral: int = 3) -> Optional[str]:
    """Reference state_dict key -> this module's key ('scales' for the Scale scalars, None if unknown).
    Reference names: mspn_mmpose.py:228-262 (layer{u}), :322-379 (up-unit convs), :404-426 (up{r}), :552-566 (top),
    :607-627 (multi_stage_mspn); das_head.py:103-175; anchor_free_mono3d_pose_head.py:100-198;
    recursive_update.py:166-180, 243-249."""
    m = re.match(r"backbone\.top\.top\.0\.(.+)", k)
    if m:
        t = _unit_tail(m.group(1))
        return t and "backbone.stem." + t
    m = re.match(r"backbone\.multi_stage_mspn\.(\d+)\.downsample\.layer(\d+)\.(\d+)\.(.+)", k)
    if m:
        s, u, b, rest = int(m.group(1)), int(m.group(2)) - 1, int(m.group(3)), m.group(4)
        head, _, param = rest.rpartition(".")
        if head in _BLOCK:
            tail = _BLOCK[head] + "." + param
        elif head.startswith("downsample."):
            t = _unit_tail(rest[len("downsample."):])
            if t is None:
                return None
            tail = "shortcut." + t
        else:
            return None
        return f"backbone.stages.{s}.encoder.{u}.{b}.{tail}"
    m = re.match(r"backbone\.multi_stage_mspn\.(\d+)\.upsample\.up(\d+)\.(\w+)\.(.+)", k)
    if m and m.group(3) in _UP:
        t = _unit_tail(m.group(4))
        return t and f"backbone.stages.{int(m.group(1))}.decoder.{int(m.group(2)) - 1}.{_UP[m.group(3)]}.{t}"
    m = re.match(r"neck\.(lateral_convs|fpn_convs)\.(\d+)\.(.+)", k)
    if m:
        t = _unit_tail(m.group(3))
        if t is None:
            return None
        i = int(m.group(2))
        if m.group(1) == "lateral_convs":
            return f"neck.lateral.{i}.{t}"
        return ("neck.smooth.%d.%s" % (i, t)) if i < n_lateral else ("neck.extra.%d.%s" % (i - n_lateral, t))
    if k.startswith("bbox_head.scales."):
        return "scales"
    m = re.match(r"bbox_head\.(cls_convs|reg_convs|pose_convs)\.(\d+)\.(.+)", k)
    if m:
        rest = m.group(3)
        base = f"towers.{_TOWER[m.group(1)]}.{int(m.group(2))}."
        t = _unit_tail(rest)
        return t and base + t
    m = re.match(r"bbox_head\.recursive_update_branch\.reduction\.(.+)", k)
    if m:
        t = _unit_tail(m.group(1))
        return t and "towers.reduction." + t
    m = re.match(r"bbox_head\.recursive_update_branch\.layer_(\d+)\.next_level_offset\.(.+)", k)
    if m:
        base, rest = f"towers.layers.{int(m.group(1))}.", m.group(2)
        if rest.startswith("update_feat_conv."):
            r = rest[len("update_feat_conv."):]
            t = _unit_tail(r)
            return t and base + "update." + t
        return base + rest
    for ref, loc in _BRANCH.items():
        pre = "bbox_head." + ref + "."
        if k.startswith(pre):
            rest = k[len(pre):]
            if loc.endswith(".1"):
                return f"towers.{loc}.{rest}"
            t = _unit_tail(rest)
            return t and f"towers.{loc}.{t}"
    return None
