"""Deterministic synthetic DAS head outputs (the decode path's inputs).

Shapes and value distributions follow SURVEY.md section 8(d): one entry per pyramid
level holding the raw predictor outputs the reference head produces right before its
eval tail (``das_head.py:232-236``) plus the refinement feature map(s) each
``RecursiveUpdateLayer`` projects from (``recursive_update.py:188`` output).

Everything is generated with a seeded ``torch.Generator`` on the requested device
(CPU for parity tests so the oracle sees the very same bits; CUDA for full-size
benchmarks where 1.7 GB of features per batch would take too long on the host).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Sequence

import numpy as np
import torch
import torch.nn.functional as F


@dataclass
class HeadConfig:
    """Keys mirror the reference head config (configs/_base_/models/das.py:24-51,
    configs/das/exp_panoptic.py:31-44)."""
    num_joints: int = 15
    root_idx: int = 2
    depth_factor: float = 20.0
    z_norm: float = 50.0
    strides: Sequence[int] = (8,)
    num_heads: int = 4
    feat_channels: int = 256
    num_layers: int = 1
    dim: int = 3

    def as_dict(self):
        return dict(num_joints=self.num_joints, root_idx=self.root_idx, depth_factor=self.depth_factor,
                    z_norm=self.z_norm, strides=list(self.strides), num_heads=self.num_heads,
                    feat_channels=self.feat_channels, num_layers=self.num_layers, dim=self.dim)


PANOPTIC = HeadConfig(num_joints=15, root_idx=2, depth_factor=20.0, z_norm=50.0)
MUPOTS17 = HeadConfig(num_joints=17, root_idx=14, depth_factor=1.0, z_norm=50.0, num_layers=3)


def _gen(seed: int, device) -> torch.Generator:
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    return g


def make_layers(cfg: HeadConfig, seed: int = 7, device="cpu", so_std: float = 0.01) -> List[dict]:
    """1x1 projection weights of every refinement layer, named after the reference
    modules (recursive_update.py:171-180): so = sampling_offset [J*nh*2, C],
    sc = sampling_conf [3J, C], uw = update_weight [3J, C], uv = update_offset_value [3J, C].
    so_std: std of the sampling-offset weights; 0.01 is the reference's init (recursive_update.py:174-175, sampled
    offsets of ~0.2 px), the "spread" bench workload uses 0.3 (offsets of ~5 px, like a trained model's)."""
    g = _gen(seed, device)
    J, nh, C = cfg.num_joints, cfg.num_heads, cfg.feat_channels
    bnd = 1.0 / math.sqrt(C)
    layers = []
    for _ in range(cfg.num_layers):
        def w(o, std):
            return torch.randn(o, C, generator=g, device=device) * std

        def b(o):
            return (torch.rand(o, generator=g, device=device) * 2 - 1) * bnd
        layers.append(dict(so_w=w(J * nh * 2, so_std), so_b=b(J * nh * 2),
                           sc_w=w(3 * J, 0.02), sc_b=b(3 * J),
                           uw_w=w(3 * J, 0.02), uw_b=b(3 * J),
                           uv_w=w(3 * J, 0.02), uv_b=b(3 * J)))
    return layers


def _smooth(x: torch.Tensor, k: int = 9) -> torch.Tensor:
    """k x k box filter, rescaled so the per-pixel std stays close to the input's."""
    if k <= 1:
        return x
    return F.avg_pool2d(x, k, 1, k // 2, count_include_pad=False) * float(k)


def make_level(cfg: HeadConfig, batch: int, h: int, w: int, stride: int, seed: int, device="cpu",
               peaks: int = 16, smooth: int = 9, scales=(1.0, 1.0, 1.0, 1.0), channels_last: bool = True,
               with_feats: bool = True, coherent: int = 0, uv_scale: float = 4.0) -> dict:
    g = _gen(seed, device)
    J, C = cfg.num_joints, cfg.feat_channels

    def randn(*s):
        return torch.randn(*s, generator=g, device=device)

    def rand(*s):
        return torch.rand(*s, generator=g, device=device)

    cls = randn(batch, 1, h, w) - 6.0
    if peaks > 0:
        pos_y = torch.randint(0, h, (batch, peaks), generator=g, device=device)
        pos_x = torch.randint(0, w, (batch, peaks), generator=g, device=device)
        amp = rand(batch, peaks) * 6.0 - 1.0
        flat = cls.view(batch, -1)
        for dy in (-1, 0, 1):
            for dx in (-1, 0, 1):
                yy = (pos_y + dy).clamp(0, h - 1)
                xx = (pos_x + dx).clamp(0, w - 1)
                val = amp - 1.5 * float(dx * dx + dy * dy)
                flat.scatter_reduce_(1, yy * w + xx, val, "amax", include_self=True)
    ctr = randn(batch, 1, h, w) + 1.0

    pose = torch.empty(batch, 3 + 6 * J, h, w, device=device)
    pose[:, 0:2] = _smooth(randn(batch, 2, h, w) * 0.5, smooth)
    pose[:, 2:3] = (0.05 + 0.45 * rand(batch, 1, h, w)) * cfg.depth_factor
    if smooth > 1:
        pose[:, 2:3] = F.avg_pool2d(pose[:, 2:3], smooth, 1, smooth // 2, count_include_pad=False)
    uvd = _smooth(randn(batch, 3 * J, h, w), smooth)
    uvd[:, 0::3] *= uv_scale
    uvd[:, 1::3] *= uv_scale
    if coherent > 0:
        # person-like fields: every cell of a coherent x coherent block points at the same joints (u = target - x),
        # so neighbouring candidates decode to near-identical poses and OKS-NMS has something to suppress
        blk = coherent
        ys = torch.arange(h, device=device, dtype=torch.float32).view(1, 1, h, 1)
        xs = torch.arange(w, device=device, dtype=torch.float32).view(1, 1, 1, w)
        bh, bw = -(-h // blk), -(-w // blk)
        per_block = randn(batch, 3 * J, bh, bw)
        per_block = per_block.repeat_interleave(blk, 2)[:, :, :h].repeat_interleave(blk, 3)[:, :, :, :w]
        uvd = 0.02 * uvd + per_block
        uvd[:, 0::3] = uvd[:, 0::3] * 4.0 + (torch.floor(xs / blk) * blk + blk / 2 - xs)
        uvd[:, 1::3] = uvd[:, 1::3] * 4.0 + (torch.floor(ys / blk) * blk + blk / 2 - ys)
    pose[:, 3:3 + 3 * J] = uvd
    pose[:, 3 + 3 * J:] = randn(batch, 3 * J, h, w)

    lvl = dict(cls=cls, ctr=ctr, pose_raw=pose, stride=int(stride), scales=tuple(float(s) for s in scales),
               feats=[])
    if with_feats:
        for _ in range(cfg.num_layers):
            f = randn(batch, h, w, C).permute(0, 3, 1, 2)      # NHWC storage, NCHW logical view
            if not channels_last:
                f = f.contiguous()
            lvl["feats"].append(f)
    return lvl


def make_levels(cfg: HeadConfig, batch: int, h: int, w: int, seed: int, device="cpu", margin_for=None, **kw) -> List[dict]:
    """One dict per stride in ``cfg.strides``; level l has size ceil(h / 2^l) x ceil(w / 2^l).

    ``margin_for`` = dict(nms_pre=..., score_thr=..., peak_kernel=...) reject-samples the centerness logits until
    every decision boundary of that decode has a rank margin of at least MIN_MARGIN_ULPS."""
    out = []
    for l, s in enumerate(cfg.strides):
        hl, wl = -(-h // (1 << l)), -(-w // (1 << l))
        out.append(make_level(cfg, batch, hl, wl, s, seed + 101 * l, device, **kw))
    if margin_for is not None:
        enforce_rank_margins(out, seed=seed + 7919, **margin_for)
    return out


# ---- rank margins ("reject-sample to guarantee rank margins >= 16 ulp at every boundary <= K+1
# and for |score - score_thr|") ------------------------------------------------------------------------------
MIN_MARGIN_ULPS = 16


def _ulps(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """|a - b| in float32 ulps (non-negative finite floats order like their bit patterns)."""
    return (a.contiguous().view(torch.int32).to(torch.int64) - b.contiguous().view(torch.int32).to(torch.int64)).abs()


def _level_candidates(lv: dict, nms_pre: int, peak_kernel: int):
    """(ranked score [B,R], cell index [B,R], n_cand) of one level: the first n_cand columns are the cells the decode
    hands on (das_head.py:716-723), column n_cand (if present) is the best cell that is left out."""
    sc = (lv["cls"].sigmoid() * lv["ctr"].sigmoid())
    b, _, h, w = sc.shape
    rank = sc
    if peak_kernel and peak_kernel > 1 and nms_pre > 0 and h * w > nms_pre:
        pooled = F.max_pool2d(sc, peak_kernel, 1, peak_kernel // 2)
        rank = torch.where(sc == pooled, sc, torch.zeros_like(sc))
    rank = rank.flatten(1)
    hw = h * w
    if nms_pre > 0 and hw > nms_pre:
        v, i = rank.topk(min(nms_pre + 1, hw), dim=1)
        return v, i, nms_pre
    v, i = rank.sort(dim=1, descending=True)
    return v, i, hw


def _margin_violations(levels: List[dict], nms_pre: int, score_thr: float, peak_kernel: int, min_ulps: int):
    """Smallest ulp gap at any boundary that decides the decode's output, plus per level the (image, cell) pairs whose
    score sits too close to such a boundary.  Boundaries (das_head.py:716-723, 763-769, pose_nms.py:97):
      * adjacent ranks among a level's candidates and the first cell left out, whenever the better one survives
        score_thr (order inside a level, and which cell makes the cut),
      * every candidate and the first cell left out against score_thr,
      * adjacent scores of the surviving candidates of ALL levels together (OKS-NMS visits them by score),
      * peak mode: every cell that could rank among the candidates against its best 3x3 neighbour."""
    worst = 1 << 40
    bad = []
    per_level = []
    thr = float(score_thr)
    for lv in levels:
        v, i, n = _level_candidates(lv, nms_pre, peak_kernel)
        per_level.append((v, i, n))
        mask = torch.zeros_like(v, dtype=torch.bool)
        if v.shape[1] > 1:
            gap = _ulps(v[:, :-1], v[:, 1:])
            rel = v[:, :-1] > thr if thr > 0 else torch.ones_like(gap, dtype=torch.bool)
            rel = rel & (v[:, :-1] > 0)
            if rel.any():
                worst = min(worst, int(gap[rel].min()))
            close = rel & (gap < min_ulps)
            mask[:, 1:] |= close
        if thr > 0:
            g = _ulps(v, torch.full_like(v, thr))
            worst = min(worst, int(g.min()))
            mask |= g < min_ulps
        bad.append((len(bad), mask, i))
    # cross-level order of the survivors (only matters when several levels feed one NMS)
    if len(levels) > 1:
        vs = torch.cat([v[:, :n] for v, _, n in per_level], dim=1)
        where = torch.cat([torch.full((n,), l, dtype=torch.int64) for l, (_, _, n) in enumerate(per_level)]).to(vs.device)
        col = torch.cat([torch.arange(n) for _, _, n in per_level]).to(vs.device)
        sv, si = vs.sort(dim=1, descending=True)
        gap = _ulps(sv[:, :-1], sv[:, 1:])
        rel = (sv[:, :-1] > thr) if thr > 0 else torch.ones_like(gap, dtype=torch.bool)
        if rel.any():
            worst = min(worst, int(gap[rel].min()))
        close = rel & (gap < min_ulps)
        if close.any():
            bi, pi = close.nonzero(as_tuple=True)
            lower = si[bi, pi + 1]
            for l in range(len(levels)):
                sel = where[lower] == l
                if sel.any():
                    bad[l][1][bi[sel], col[lower[sel]]] = True
    if peak_kernel and peak_kernel > 1:
        for l, lv in enumerate(levels):
            v, i, n = per_level[l]
            sc = (lv["cls"].sigmoid() * lv["ctr"].sigmoid())
            b, _, h, w = sc.shape
            if not (nms_pre > 0 and h * w > nms_pre):
                continue
            pad = F.pad(sc, (1, 1, 1, 1), value=0.0)
            nb = torch.stack([pad[:, :, dy:dy + h, dx:dx + w] for dy in range(3) for dx in range(3) if (dy, dx) != (1, 1)]).amax(0)
            cut = v[:, min(n, v.shape[1] - 1)].view(b, 1, 1, 1)
            rel = (sc >= cut) & (sc > 0)
            gap = _ulps(sc, nb)
            if rel.any():
                worst = min(worst, int(gap[rel].min()))
            close = (rel & (gap < min_ulps)).flatten(1)
            if close.any():
                bi, ci = close.nonzero(as_tuple=True)
                extra_mask = torch.zeros((b, h * w), dtype=torch.bool, device=sc.device)
                extra_mask[bi, ci] = True
                bad.append((l, extra_mask, torch.arange(h * w, device=sc.device).expand(b, -1)))
    return worst, bad


def rank_margin_ulps(levels: List[dict], nms_pre: int, score_thr: float = 0.0, peak_kernel: int = 0) -> int:
    """Smallest float32-ulp gap at any boundary that decides which candidates the decode outputs and in which order
    (see _margin_violations).  The CPU reference's own sigmoid is only ~2 ulp accurate, so bit-exact index parity is
    only defined when this is comfortably above that; the parity tests assert >= MIN_MARGIN_ULPS."""
    return _margin_violations(levels, int(nms_pre), float(score_thr or 0.0), int(peak_kernel or 0), MIN_MARGIN_ULPS)[0]


def enforce_rank_margins(levels: List[dict], nms_pre: int = -1, score_thr: float = 0.0, peak_kernel: int = 0,
                         seed: int = 0, min_ulps: int = 4 * MIN_MARGIN_ULPS, max_iter: int = 200) -> int:
    """Reject-sampling step of the generator: re-draw the centerness logit of every cell whose score sits within
    `min_ulps` of a decision boundary until none is left (in place; deterministic for a given seed).  The default
    leaves 4x the margin the tests require.  Returns the number of re-drawn cells."""
    dev = levels[0]["ctr"].device
    g = _gen(seed, dev)
    redrawn = 0
    for _ in range(max_iter):
        worst, bad = _margin_violations(levels, int(nms_pre), float(score_thr or 0.0), int(peak_kernel or 0), min_ulps)
        n_bad = 0
        for l, mask, idx in bad:
            if not mask.any():
                continue
            bi, ci = mask.nonzero(as_tuple=True)
            cells = idx[bi, ci]
            flat = levels[l]["ctr"].view(levels[l]["ctr"].shape[0], -1)
            flat[bi, cells] = torch.randn(len(bi), generator=g, device=dev) + 1.0
            n_bad += len(bi)
        redrawn += n_bad
        if n_bad == 0:
            return redrawn
    raise RuntimeError(f"enforce_rank_margins: still {n_bad} cells within {min_ulps} ulp of a boundary after {max_iter} rounds")


def make_metas(batch: int, h: int, w: int, stride: int = 8, seed: int = 3, identity_rt: bool = False) -> List[dict]:
    """img_metas entries with the keys the path reads: 'scale_factor' (np.float32[4], das_head.py:698),
    'filename' (:684) and 'cam' {K,R,t} (formating.py:140)."""
    rng = np.random.RandomState(seed)
    metas = []
    for b in range(batch):
        sx, sy = 0.6 + 0.01 * rng.rand(), 0.6 + 0.01 * rng.rand()
        fx, fy = 1400.0 + 50 * rng.rand(), 1410.0 + 50 * rng.rand()
        K = np.array([[fx, 0.3 * rng.rand(), w * stride / sx / 2 + rng.rand()],
                      [0.0, fy, h * stride / sy / 2 + rng.rand()],
                      [0.0, 0.0, 1.0]])
        if identity_rt:
            R, t = np.eye(3), np.zeros((3, 1))
        else:
            a, bb, c = 0.3 * rng.rand(3)
            Rx = np.array([[1, 0, 0], [0, math.cos(a), -math.sin(a)], [0, math.sin(a), math.cos(a)]])
            Ry = np.array([[math.cos(bb), 0, math.sin(bb)], [0, 1, 0], [-math.sin(bb), 0, math.cos(bb)]])
            Rz = np.array([[math.cos(c), -math.sin(c), 0], [math.sin(c), math.cos(c), 0], [0, 0, 1]])
            R = Rz @ Ry @ Rx
            t = np.array([[10.0 * rng.rand()], [-120.0 + 5 * rng.rand()], [300.0 + 20 * rng.rand()]])
        metas.append(dict(scale_factor=np.array([sx, sy, sx, sy], dtype=np.float32),
                          filename=f"synthetic_{b:05d}.jpg", cam=dict(K=K, R=R, t=t)))
    return metas


def levels_to(levels: List[dict], device) -> List[dict]:
    out = []
    for lv in levels:
        d = dict(lv)
        for k in ("cls", "ctr", "pose_raw"):
            d[k] = lv[k].to(device)
        d["feats"] = [f.to(device) for f in lv["feats"]]     # preserves the NHWC strides
        out.append(d)
    return out


def layers_to(layers: List[dict], device) -> List[dict]:
    return [{k: v.to(device) for k, v in l.items()} for l in layers]
