"""Host-side mirror of the reference dense-head decode interface, backed by the CUDA C-ABI.

``DASHeadB200`` keeps the reference plugin contract of ``DASHead`` for the inference decode
(reference: mmdet3d/models/pose_heads/das_head.py:653-688 ``get_poses``; config keys from
configs/_base_/models/das.py:24-51 and configs/das/exp_panoptic.py:31-53):

* same constructor keys for the path (``num_joints, strides, depth_factor, z_norm, root_idx,
  recursive_update{num_heads, feat_channels, num_layers, dim}``, ``test_cfg{nms_pre, nms_post,
  nms_thr, score_thr, nms_type}``); unrelated keys (losses, conv towers) are accepted and ignored;
* ``get_poses(cls_scores, pose_preds, centernesses, img_metas, cfg=None, rescale=None)`` returns
  one dict per image with ``poses [N,J,3]``, ``vis [N,J]``, ``centers [N,3]``, ``image_paths`` and
  ``scores`` (python list), plus ``poses_cam`` / ``poses_world`` (float64) -- the evaluator-side
  back-projection (cmupanoptic_mono_dataset.py:391-402) moved onto the device;
* the same python ``assert`` error behaviour on list-length / shape mismatch (das_head.py:660,701,711).

When ``get_poses`` additionally receives the refinement feature maps (``refine_feats``), the pose
maps are taken as RAW predictor outputs and the progressive refinement + eval tail
(das_head.py:237-262) run sparsely at the selected centres on the GPU (SURVEY.md 8.0 divergence B).

No CPU path exists here: everything raises if the CUDA library or device is missing.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import Buffers, DecodeCfg, Levels


class _DevArray:
    """Expose plan-owned device memory to torch through __cuda_array_interface__ (zero copy)."""

    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = dict(shape=tuple(int(s) for s in shape), typestr=typestr,
                                             data=(int(ptr), False), version=2, strides=None)


def _as_tensor(ptr, shape, typestr, device):
    if int(np.prod(shape)) == 0:
        dt = {"<f4": torch.float32, "<i4": torch.int32, "<f8": torch.float64}[typestr]
        return torch.empty(tuple(shape), dtype=dt, device=device)
    return torch.as_tensor(_DevArray(ptr, shape, typestr), device=device)


def _stream_ptr(device) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _check_f32(t: torch.Tensor, what: str):
    if t.dtype != torch.float32:
        raise TypeError(f"{what}: expected float32 at the C ABI, got {t.dtype}")
    if not t.is_cuda:
        raise RuntimeError(f"{what}: the decode path only runs on CUDA tensors")


_IN_DTYPES = {torch.float32: _lib.DTYPE_F32, torch.float16: _lib.DTYPE_F16, torch.bfloat16: _lib.DTYPE_BF16}


def _to_f32(t: torch.Tensor) -> torch.Tensor:
    """Refinement feature maps are fp32 at the C ABI (inside the reference's force_fp32 region they are fp32 too)."""
    return t if t.dtype == torch.float32 else t.float()


def _head_maps(cls_l, ctr_l, pose_l):
    """The three head-output maps of every level in ONE common element type the kernels read natively: fp32, or -- the
    reference's shipped fp16 mode (das_head.py:180,218 out_fp16=True, exp_panoptic.py:222) -- fp16 / bf16 as they are
    (half the scan and gather bytes; the arithmetic is fp32 either way, so results equal the up-cast path bit for bit).
    Mixed or other dtypes are up-cast to fp32."""
    dts = {t.dtype for t in list(cls_l) + list(ctr_l) + list(pose_l)}
    if len(dts) == 1 and next(iter(dts)) in _IN_DTYPES:
        return cls_l, ctr_l, pose_l
    f = lambda ts: [t.float() for t in ts]
    return f(cls_l), f(ctr_l), f(pose_l)


_BLOCK_SPEC = (("out_count", torch.int32, 0), ("out_score", torch.float32, 1), ("out_slot", torch.int32, 1),
               ("out_pose", torch.float32, 2), ("out_center", torch.float32, 3), ("out_cam", torch.float64, 2),
               ("out_world", torch.float64, 2))


def block_layout(B: int, P: int, J: int):
    """(name, dtype, shape, byte offset) of every out_* buffer inside the packed output block, and its size.
    Mirrors das_plan_create: [count | score | slot | pose | center | cam | world], each 256-byte aligned."""
    shapes = {0: (B,), 1: (B, P), 2: (B, P, J, 3), 3: (B, P, 3)}
    out, off = [], 0
    for name, dt, kind in _BLOCK_SPEC:
        shape = shapes[kind]
        n = int(np.prod(shape)) * torch.empty((), dtype=dt).element_size()
        out.append((name, dt, shape, off))
        off = (off + n + 255) & ~255
    return out, off


def block_views(block: torch.Tensor, B: int, P: int, J: int) -> Dict[str, torch.Tensor]:
    lay, total = block_layout(B, P, J)
    assert block.dtype == torch.uint8 and block.numel() == total, (block.dtype, block.numel(), total)
    views = {}
    for name, dt, shape, off in lay:
        n = int(np.prod(shape)) * torch.empty((), dtype=dt).element_size()
        views[name] = block[off:off + n].view(dt).view(shape)
    return views


class DecodePlan:
    """One (batch, level shapes, config) instance of the C ``das_plan``."""

    def __init__(self, *, num_joints, root_idx, depth_factor, z_norm, strides, level_sizes, batch,
                 test_cfg, num_heads=4, feat_channels=256, num_layers=1, refine=True, peak_kernel=0,
                 dataset_depth_factor=1.0, device="cuda", refine_mode=None):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise RuntimeError("das_b200 needs a CUDA device; there is no CPU path")
        self.device = torch.device(device)
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        nms_type = test_cfg.get("nms_type", "hard")       # anything but 'hard' selects soft_oks_nms (das_head.py:784-790)
        nms_post = test_cfg.get("nms_post", -1)
        self.cfg = DecodeCfg(num_joints=num_joints, root_idx=root_idx, num_heads=num_heads,
                             feat_channels=feat_channels, num_layers=num_layers,
                             depth_factor=float(depth_factor), z_norm=float(z_norm),
                             nms_pre=int(test_cfg.get("nms_pre", -1)),
                             # das_head.py:770,785: absent/<=0 skips NMS; when present the cap re-reads it
                             nms_post=int(nms_post),
                             nms_thr=float(test_cfg.get("nms_thr", 0.9)),
                             score_thr=float(test_cfg.get("score_thr", 0.0)),
                             peak_kernel=int(peak_kernel), refine=int(bool(refine)),
                             dataset_depth_factor=float(dataset_depth_factor), nms_soft=int(nms_type != "hard"))
        self.batch = int(batch)
        self.strides = [int(s) for s in strides]
        self.level_sizes = [(int(h), int(w)) for h, w in level_sizes]
        assert len(self.strides) == len(self.level_sizes)
        shape = Levels()
        shape.n_levels = len(self.strides)
        shape.batch = self.batch
        for l, ((h, w), s) in enumerate(zip(self.level_sizes, self.strides)):
            shape.lv[l].H, shape.lv[l].W, shape.lv[l].stride = h, w, s
            shape.lv[l].scale_offset = shape.lv[l].scale_depth = shape.lv[l].scale_uv = shape.lv[l].scale_d = 1.0
        self._shape = shape
        self._plan = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.das_plan_create(C.byref(self.cfg), C.byref(shape), C.byref(self._plan)), "das_plan_create")
            bufs = Buffers()
            ct, p = C.c_int32(), C.c_int32()
            _lib.check(self.lib.das_plan_buffers(self._plan, C.byref(bufs), C.byref(ct), C.byref(p)), "das_plan_buffers")
            if refine_mode is not None:      # 0 fp32 SIMT, 1 tcgen05 3xTF32 (default when supported), 2 tcgen05 TF32
                _lib.check(self.lib.das_plan_set_refine_mode(self._plan, int(refine_mode)), "das_plan_set_refine_mode")
        self.cand_slots, self.out_slots = ct.value, p.value
        self.refine_mode = (int(refine_mode) if refine_mode is not None else
                            (1 if (refine and feat_channels == 256 and num_heads == 4) else 0))
        B, CT, P, J = self.batch, ct.value, p.value, num_joints
        dev = self.device
        self.t = dict(
            cand_score=_as_tensor(bufs.cand_score, (B, CT), "<f4", dev),
            cand_index=_as_tensor(bufs.cand_index, (B, CT), "<i4", dev),
            cand_pose=_as_tensor(bufs.cand_pose, (B, CT, J, 3), "<f4", dev),
            cand_center=_as_tensor(bufs.cand_center, (B, CT, 3), "<f4", dev),
            out_count=_as_tensor(bufs.out_count, (B,), "<i4", dev),
            out_score=_as_tensor(bufs.out_score, (B, P), "<f4", dev),
            out_slot=_as_tensor(bufs.out_slot, (B, P), "<i4", dev),
            out_pose=_as_tensor(bufs.out_pose, (B, P, J, 3), "<f4", dev),
            out_center=_as_tensor(bufs.out_center, (B, P, 3), "<f4", dev),
            out_cam=_as_tensor(bufs.out_cam, (B, P, J, 3), "<f8", dev),
            out_world=_as_tensor(bufs.out_world, (B, P, J, 3), "<f8", dev),
        )
        self._keep = []          # tensors whose storage the plan currently points at
        self._out_keep = None    # caller-owned output block (kept alive while the plan points at it)
        self._run = self.lib.das_plan_run

    def __del__(self):
        try:
            if getattr(self, "_plan", None) and self._plan.value:
                self.lib.das_plan_destroy(self._plan)
                self._plan = C.c_void_p()
        except Exception:
            pass

    # ---- inputs ------------------------------------------------------------------------------
    def set_weights(self, layers: Sequence[Dict[str, torch.Tensor]]):
        """layers[k]: dict(so_w, so_b, sc_w, sc_b, uw_w, uw_b, uv_w, uv_b) = the nn.Conv2d parameters of
        recursive_update_branch.layer_k.next_level_offset.{sampling_offset, sampling_conf, update_weight,
        update_offset_value} (weights [O,C] or [O,C,1,1])."""
        assert len(layers) == self.cfg.num_layers, (len(layers), self.cfg.num_layers)
        J, nh, Cc = self.cfg.num_joints, self.cfg.num_heads, self.cfg.feat_channels
        want = dict(so=J * nh * 2, sc=3 * J, uw=3 * J, uv=3 * J)
        with torch.cuda.device(self.device):
            for k, lw in enumerate(layers):
                args = []
                hold = []
                for name in ("so", "sc", "uw", "uv"):
                    w = lw[name + "_w"].detach().to(self.device, torch.float32).reshape(want[name], Cc).contiguous()
                    b = lw[name + "_b"].detach().to(self.device, torch.float32).reshape(want[name]).contiguous()
                    hold += [w, b]
                    args += [C.c_void_p(w.data_ptr()), C.c_void_p(b.data_ptr())]
                _lib.check(self.lib.das_plan_set_weights(self._plan, k, *args, _stream_ptr(self.device)), "das_plan_set_weights")
                torch.cuda.current_stream(self.device).synchronize()   # `hold` may be freed afterwards

    def _levels_struct(self, levels: Sequence[dict], host: bool = False) -> Levels:
        lv = Levels()
        lv.n_levels = len(levels)
        lv.batch = self.batch
        J = self.cfg.num_joints
        keep = []
        assert len(levels) == len(self.level_sizes), "number of levels differs from the plan"
        for l, d in enumerate(levels):
            h, w = self.level_sizes[l]
            cls, ctr, pose = d["cls"], d["ctr"], d["pose"]
            assert tuple(cls.shape) == (self.batch, 1, h, w), (tuple(cls.shape), (self.batch, 1, h, w))
            assert tuple(ctr.shape) == (self.batch, 1, h, w)
            assert tuple(pose.shape) == (self.batch, 3 + 6 * J, h, w), tuple(pose.shape)
            for name, t in (("cls", cls), ("ctr", ctr), ("pose", pose)):
                if t.dtype not in _IN_DTYPES:
                    raise TypeError(f"{name}: expected float32 / float16 / bfloat16 at the C ABI, got {t.dtype}")
                if not host and not t.is_cuda:
                    raise RuntimeError(f"{name}: the decode path only runs on CUDA tensors")
                if l == 0 and name == "cls":
                    lv.in_dtype = _IN_DTYPES[t.dtype]
                elif _IN_DTYPES[t.dtype] != lv.in_dtype:
                    raise TypeError(f"{name} of level {l} is {t.dtype}: cls / ctr / pose of every level must share one dtype")
            cls, ctr, pose = cls.contiguous(), ctr.contiguous(), pose.contiguous()
            keep += [cls, ctr, pose]
            lv.lv[l].cls, lv.lv[l].ctr, lv.lv[l].pose = cls.data_ptr(), ctr.data_ptr(), pose.data_ptr()
            lv.lv[l].H, lv.lv[l].W, lv.lv[l].stride = h, w, self.strides[l]
            sc = d.get("scales", (1.0, 1.0, 1.0, 1.0))
            lv.lv[l].scale_offset, lv.lv[l].scale_depth, lv.lv[l].scale_uv, lv.lv[l].scale_d = [float(s) for s in sc]
            if self.cfg.refine:
                feats = d["feats"]
                assert len(feats) == self.cfg.num_layers, "one refinement feature map per layer is required"
                for k, f in enumerate(feats):
                    assert tuple(f.shape) == (self.batch, self.cfg.feat_channels, h, w), tuple(f.shape)
                    if not host:
                        _check_f32(f, "feats")
                    # the kernels read NHWC; channels_last tensors are used in place, NCHW ones converted once
                    f = f.contiguous(memory_format=torch.channels_last)
                    keep.append(f)
                    lv.lv[l].feats[k] = f.data_ptr()
        self._keep_tmp = keep
        return lv

    def bind(self, levels: Sequence[dict]):
        lv = self._levels_struct(levels)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.das_plan_bind(self._plan, C.byref(lv), _stream_ptr(self.device)), "das_plan_bind")
        self._keep = self._keep_tmp

    @staticmethod
    def pack_metas(img_metas: Sequence[dict]):
        """(scale_xy [B,2] float32, cam [B,18] float64 = K[0,:3] K[1,:3] R t) from the reference's img_metas entries.
        One array construction per field instead of per-image slicing: this runs on the host inside every end-to-end
        call (64 images: 640 us -> ~80 us)."""
        B = len(img_metas)
        sxy = np.ascontiguousarray(np.array([m["scale_factor"] for m in img_metas], dtype=np.float32)[:, :2])
        cam = np.zeros((B, _lib.CAM_DOUBLES), dtype=np.float64)
        cams = [m.get("cam") for m in img_metas]
        if all(c is not None and "R" in c and "t" in c for c in cams):
            try:
                K = np.array([c["K"] for c in cams], dtype=np.float64)      # np.array over a list of equal-shape arrays
                cam[:, 0:3] = K[:, 0, :3]                                    # is the cheapest way to gather them
                cam[:, 3:6] = K[:, 1, :3]
                cam[:, 6:15] = np.array([c["R"] for c in cams], dtype=np.float64).reshape(B, 9)
                cam[:, 15:18] = np.array([c["t"] for c in cams], dtype=np.float64).reshape(B, 3)
                return sxy, cam
            except (ValueError, IndexError):     # images with differently shaped K (2x3 next to 3x3): general path
                pass
        for b, c in enumerate(cams):                 # mixed / partial camera entries: the general path
            if c is None:
                K, R, t = np.eye(3), np.eye(3), np.zeros(3)
            else:
                K = np.asarray(c["K"], dtype=np.float64)
                R = np.asarray(c.get("R", np.eye(3)), dtype=np.float64)
                t = np.asarray(c.get("t", np.zeros(3)), dtype=np.float64).reshape(3)
            cam[b, 0:3] = K[0, :3]
            cam[b, 3:6] = K[1, :3]
            cam[b, 6:15] = R.reshape(9)
            cam[b, 15:18] = t
        return sxy, cam

    def set_metas(self, img_metas: Sequence[dict]):
        assert len(img_metas) == self.batch
        sxy, cam = self.pack_metas(img_metas)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.das_plan_set_metas(self._plan, C.c_void_p(sxy.ctypes.data), C.c_void_p(cam.ctypes.data),
                                                   _stream_ptr(self.device)), "das_plan_set_metas")

    # ---- outputs: where the packed result block lives ------------------------------------------------------------
    def _refresh_views(self):
        bufs = Buffers()
        ct, p = C.c_int32(), C.c_int32()
        _lib.check(self.lib.das_plan_buffers(self._plan, C.byref(bufs), C.byref(ct), C.byref(p)), "das_plan_buffers")
        B, P, J, dev = self.batch, self.out_slots, self.cfg.num_joints, self.device
        self.t.update(out_count=_as_tensor(bufs.out_count, (B,), "<i4", dev), out_score=_as_tensor(bufs.out_score, (B, P), "<f4", dev),
                      out_slot=_as_tensor(bufs.out_slot, (B, P), "<i4", dev), out_pose=_as_tensor(bufs.out_pose, (B, P, J, 3), "<f4", dev),
                      out_center=_as_tensor(bufs.out_center, (B, P, 3), "<f4", dev), out_cam=_as_tensor(bufs.out_cam, (B, P, J, 3), "<f8", dev),
                      out_world=_as_tensor(bufs.out_world, (B, P, J, 3), "<f8", dev))

    @property
    def block_stride(self) -> int:
        """Bytes one plan occupies in a caller-owned buffer: the packed outputs plus the 256-byte sequence trailer."""
        return int(self.output_block().numel()) + 256

    def set_output_block(self, block: Optional[torch.Tensor]):
        """Make the plan write its results into `block` (uint8 CUDA tensor of >= block_stride bytes, 256-B aligned) --
        typically a slice of one staging / gathered buffer shared by several plans, so that collecting results needs no
        device-to-device copy.  None returns to the plan's own allocation."""
        with torch.cuda.device(self.device):
            if block is None:
                _lib.check(self.lib.das_plan_set_output_block(self._plan, None, 0), "das_plan_set_output_block")
            else:
                assert block.dtype == torch.uint8 and block.is_cuda and block.is_contiguous()
                _lib.check(self.lib.das_plan_set_output_block(self._plan, C.c_void_p(block.data_ptr()), block.numel()),
                           "das_plan_set_output_block")
        self._out_keep = block
        self._refresh_views()

    def set_output_ptr(self, ptr: int, nbytes: int, keep=None):
        """set_output_block for raw device memory (e.g. from das_ipc_alloc)."""
        with torch.cuda.device(self.device):
            _lib.check(self.lib.das_plan_set_output_block(self._plan, C.c_void_p(int(ptr)), int(nbytes)), "das_plan_set_output_block")
        self._out_keep = keep
        self._refresh_views()

    def set_peer_blocks(self, peer_ptrs: Sequence[int]):
        """Fused result all-gather: the NMS / back-projection kernel also stores every result value at the same offset of
        each address in `peer_ptrs` (this rank's slot in every peer's gathered buffer, mapped with das_ipc_open), over
        NVLink.  An empty list switches it off."""
        arr = (C.c_void_p * max(len(peer_ptrs), 1))(*[int(q) for q in peer_ptrs])
        with torch.cuda.device(self.device):
            _lib.check(self.lib.das_plan_set_peer_blocks(self._plan, len(peer_ptrs), arr), "das_plan_set_peer_blocks")

    # ---- run -----------------------------------------------------------------------------------
    def run(self, use_graph: bool = True, stage_events: bool = False, stream: Optional[int] = None):
        """Enqueue one decode on the current stream (or on the raw cudaStream_t `stream`): eager launches, a CUDA-graph
        replay, or a replay with event nodes at the stage boundaries (then stage_ms() after a synchronize)."""
        mode = 2 if stage_events else int(bool(use_graph))
        if stream is not None:             # lean path for tight loops: no torch stream / device context lookups
            st = self._run(self._plan, C.c_void_p(stream), mode)
            if st != 0:
                _lib.check(st, "das_plan_run")
            return
        with torch.cuda.device(self.device):
            _lib.check(self.lib.das_plan_run(self._plan, _stream_ptr(self.device), mode), "das_plan_run")

    def stage_ms(self) -> List[float]:
        """[score_topk, dense layers, refine phases 1-2 (tensor-core mode only), refine+assemble, nms+backproject]
        of the last stage_events run."""
        arr = (C.c_float * 5)()
        _lib.check(self.lib.das_plan_stage_ms(self._plan, arr), "das_plan_stage_ms")
        return [float(x) for x in arr]

    def output_block(self) -> torch.Tensor:
        """uint8 view of the single device block holding every out_* buffer (one NCCL all-gather
        moves a rank's whole result)."""
        ptr, n = C.c_void_p(), C.c_int64()
        _lib.check(self.lib.das_plan_output_block(self._plan, C.byref(ptr), C.byref(n)), "das_plan_output_block")
        return torch.as_tensor(_DevArray(ptr.value, (n.value,), "|u1"), device=self.device)

    def run_host(self, levels: Sequence[dict], img_metas: Sequence[dict], host_out: Dict[str, torch.Tensor], sync: bool = True):
        """End-to-end entry with HOST tensors (pinned for full PCIe speed): H2D of every input,
        decode, D2H of the packed outputs into ``host_out`` (see alloc_host_out), stream sync.
        sync=False (das_plan_run_host_async): everything is enqueued on the current stream and the call returns; the
        caller synchronises that stream before it reads ``host_out`` or reuses the inputs / this plan."""
        lv = self._levels_struct(levels, host=True)
        sxy, cam = self.pack_metas(img_metas)
        ob = Buffers()
        for k in ("out_count", "out_score", "out_slot", "out_pose", "out_center", "out_cam", "out_world"):
            if k in host_out:
                setattr(ob, k, host_out[k].data_ptr())
        with torch.cuda.device(self.device):
            fn = self.lib.das_plan_run_host if sync else self.lib.das_plan_run_host_async
            _lib.check(fn(self._plan, C.byref(lv), C.c_void_p(sxy.ctypes.data), C.c_void_p(cam.ctypes.data), ob,
                          _stream_ptr(self.device)), "das_plan_run_host")

    def alloc_host_out(self, pinned: bool = True, contiguous: bool = True) -> Dict[str, torch.Tensor]:
        """Host arrays for run_host's results.  contiguous=True: views into ONE host block with the layout of the device
        output block, which das_plan_run_host recognises and fills with a single D2H copy; False: one array per field."""
        if contiguous:
            block = torch.empty(int(self.output_block().numel()), dtype=torch.uint8, pin_memory=pinned)
            return self.views_of_block(block)
        return {k: torch.empty(v.shape, dtype=v.dtype, pin_memory=pinned)
                for k, v in self.t.items() if k.startswith("out_")}

    def views_of_block(self, block: torch.Tensor) -> Dict[str, torch.Tensor]:
        """Typed views into a copy of an output block (this rank's, or one gathered from a peer rank)."""
        return block_views(block, self.batch, self.out_slots, self.cfg.num_joints)

    def set_host_mode(self, zero_copy: bool, row_cache: bool = True):
        """run_host policy: False = bulk H2D of every map; True = copy only the logit planes and let the gather
        kernels read the (sparsely used) pose / feature maps in place from pinned host memory; `row_cache` adds the
        pass that copies every distinct feature row of the sampling phase to the device once."""
        mode = (2 if row_cache else 1) if zero_copy else 0
        _lib.check(self.lib.das_plan_set_host_mode(self._plan, mode), "das_plan_set_host_mode")

    def row_cache_stats(self):
        """(distinct feature rows fetched into the device row cache, candidates above score_thr) of the last host-mode
        run; synchronises with the device."""
        arr = (C.c_int32 * 2)()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.das_plan_row_cache_stats(self._plan, arr), "das_plan_row_cache_stats")
        return int(arr[0]), int(arr[1])

    def publish_wait(self, stream=None):
        """Make `stream` (a raw cudaStream_t value; default: the current stream) wait until the plan's last result
        publication to its peers has completed (no-op without peer blocks)."""
        sp = _stream_ptr(self.device) if stream is None else C.c_void_p(stream)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.das_plan_publish_wait(self._plan, sp), "das_plan_publish_wait")

    def set_pdl(self, mode: int):
        """Programmatic dependent launch along the kernel chain: 1 on (latency mode, one decode at a time), 0 off (throughput
        mode, several decodes in flight on different streams), -1 auto (default: on for decodes too small to fill the GPU)."""
        _lib.check(self.lib.das_plan_set_pdl(self._plan, int(mode)), "das_plan_set_pdl")

    def set_on_demand_sampling(self, on: bool):
        """num_layers > 1, tensor-core path: True (default) = layer L-2 only projects and the sparse last layer evaluates its
        progressive sampling at the cells it looks at; False = every dense layer samples its whole map.  Same results."""
        _lib.check(self.lib.das_plan_set_on_demand_sampling(self._plan, int(bool(on))), "das_plan_set_on_demand_sampling")

    def refine_stats(self):
        """(distinct (cell, joint) feature rows the gathered GEMM multiplied, candidates above score_thr, rows without the
        de-duplication) of the last tensor-core-mode run; synchronises with the device."""
        arr = (C.c_int64 * 3)()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.das_plan_refine_stats(self._plan, arr), "das_plan_refine_stats")
        return int(arr[0]), int(arr[1]), int(arr[2])

    @property
    def h2d_explicit_bytes(self) -> int:
        return int(self.lib.das_plan_h2d_explicit_bytes(self._plan))

    @property
    def kernel_launches(self) -> int:
        return int(self.lib.das_plan_kernel_launches(self._plan))

    @property
    def h2d_bytes(self) -> int:
        return int(self.lib.das_plan_h2d_bytes(self._plan))

    @property
    def d2h_bytes(self) -> int:
        return int(self.lib.das_plan_d2h_bytes(self._plan))

    # ---- results ------------------------------------------------------------------------------
    def results(self, img_metas: Sequence[dict], src: Optional[Dict[str, torch.Tensor]] = None) -> List[dict]:
        """Build the reference's list-of-dicts (das_head.py:680-687) with ONE small device->host copy
        (counts + scores); pose tensors stay on the device like the reference's."""
        t = src or self.t
        counts = t["out_count"].cpu().tolist() if t["out_count"].is_cuda else t["out_count"].tolist()
        scores = t["out_score"].cpu() if t["out_score"].is_cuda else t["out_score"]
        res = []
        for b, m in enumerate(img_metas):
            n = int(counts[b])
            poses = t["out_pose"][b, :n].clone()
            res.append(dict(poses=poses, vis=torch.ones(poses.shape[:2], dtype=poses.dtype, device=poses.device),
                            centers=t["out_center"][b, :n].clone(),
                            image_paths=[m.get("filename", "")],
                            scores=scores[b, :n].tolist(),
                            poses_cam=t["out_cam"][b, :n].clone(), poses_world=t["out_world"][b, :n].clone(),
                            slots=t["out_slot"][b, :n].clone()))
        return res


class DASHeadB200:
    """Drop-in for the inference decode of the reference ``DASHead`` (das_head.py:30-63, 653-796)."""

    def __init__(self, num_classes=1, in_channels=256, *, num_joints=15, strides=(8, 16, 32, 64, 128),
                 depth_factor=1, z_norm=1, root_idx=None, recursive_update=None, test_cfg=None,
                 train_cfg=None, peak_kernel=0, device="cuda", dataset_depth_factor=1.0, **unused):
        assert num_classes == 1, "the DAS head has one class (configs/_base_/models/das.py:26)"
        ru = dict(num_heads=4, feat_channels=256, num_layers=1, dim=3)
        ru.update(recursive_update or {})
        assert ru["dim"] == 3, "dim=3 is the only shipped value (configs/_base_/models/das.py:49)"
        self.num_classes = num_classes
        self.cls_out_channels = 1
        self.in_channels = in_channels
        self.num_joints = int(num_joints)
        self.strides = list(strides)
        self.depth_factor = depth_factor
        self.z_norm = z_norm
        self.root_idx = int(root_idx)
        self.recursive_update = ru
        self.test_cfg = dict(test_cfg or {})
        self.train_cfg = train_cfg
        self.peak_kernel = peak_kernel
        self.device = device
        self.dataset_depth_factor = float(dataset_depth_factor)     # cmupanoptic_mono_dataset.py:399 (dataset-side, default 1)
        self.group_reg_dims = [2, 1, 3 * self.num_joints, 3 * self.num_joints]
        self.training = False
        self._plans: Dict[tuple, DecodePlan] = {}
        self._layers = None
        self.scales = [(1.0, 1.0, 1.0, 1.0) for _ in self.strides]

    # refinement 1x1 weights (state_dict names in SURVEY.md section 5, checkpoint row)
    def load_refine_weights(self, layers: Sequence[Dict[str, torch.Tensor]]):
        self._layers = list(layers)
        for p in self._plans.values():
            if p.cfg.refine:
                p.set_weights(self._layers)

    def _plan(self, batch, sizes, cfg, refine) -> DecodePlan:
        key = (batch, tuple(sizes), tuple(sorted(cfg.items())), bool(refine))
        p = self._plans.get(key)
        if p is None:
            p = DecodePlan(num_joints=self.num_joints, root_idx=self.root_idx, depth_factor=self.depth_factor,
                           z_norm=self.z_norm, strides=self.strides[:len(sizes)], level_sizes=sizes, batch=batch,
                           test_cfg=cfg, num_heads=self.recursive_update["num_heads"],
                           feat_channels=self.recursive_update["feat_channels"],
                           num_layers=self.recursive_update["num_layers"], refine=refine,
                           peak_kernel=self.peak_kernel, device=self.device,
                           dataset_depth_factor=self.dataset_depth_factor)
            if refine:
                assert self._layers is not None, "call load_refine_weights() before a refining decode"
                p.set_weights(self._layers)
            self._plans[key] = p
        return p

    @staticmethod
    def _is_metas(x) -> bool:
        """img_metas is a list of per-image dicts (das_head.py:653-659); an empty list counts as one (empty batch)."""
        return isinstance(x, (list, tuple)) and all(isinstance(m, dict) for m in x)

    @classmethod
    def _split_rest(cls, rest, cfg, img_metas=None):
        """Resolve the positional tail of get_poses by CONTENT:
          reference form   (img_metas[, cfg[, rescale]])                  -- das_head.py:653-659
          extended form    (refine_feats, img_metas[, cfg[, rescale]])    -- refine_feats[level] = list of feature maps
        and the keyword form get_poses(..., img_metas=..., [refine_feats positional]).  Returns (refine_feats, img_metas, cfg)."""
        rest = tuple(rest)
        if img_metas is not None:                      # keyword img_metas: what is left can only be refine_feats
            if len(rest) > 1:
                raise TypeError("get_poses(): img_metas was passed by keyword; at most refine_feats may be positional")
            return (rest[0] if rest else None), img_metas, cfg
        if not rest:
            raise TypeError("get_poses() missing img_metas")
        second_is_metas = len(rest) >= 2 and cls._is_metas(rest[1])
        # an empty rest[0] followed by a non-empty list of dicts can only be (refine_feats of zero levels, img_metas)
        if cls._is_metas(rest[0]) and not (second_is_metas and len(rest[0]) == 0 and len(rest[1]) > 0):
            tail, feats, metas = rest[1:], None, rest[0]
        elif second_is_metas:
            tail, feats, metas = rest[2:], rest[0], rest[1]
        else:
            raise TypeError("get_poses(): expected (img_metas[, cfg[, rescale]]) or (refine_feats, img_metas[, cfg[, rescale]]); "
                            "img_metas must be a list of dicts")
        if len(tail) > 2:
            raise TypeError("get_poses(): too many positional arguments")
        if tail and tail[0] is not None and cfg is None:
            cfg = tail[0]                                # positional cfg like the reference allows
        return feats, metas, cfg

    def decode_to_device(self, cls_scores, pose_preds, centernesses, *rest, img_metas=None, cfg=None, rescale=None) -> DecodePlan:
        """get_poses without the device->host read: enqueues the decode on the current stream and returns the plan
        whose `output_block()` / `views_of_block()` hold the padded pose lists on the device (same arguments)."""
        refine_feats, img_metas, cfg = self._split_rest(rest, cfg, img_metas)
        assert len(cls_scores) == len(pose_preds) == len(centernesses)
        cfg = self.test_cfg if cfg is None else dict(cfg)
        num_levels = len(cls_scores)
        batch = len(img_metas)
        sizes = [tuple(int(s) for s in c.shape[-2:]) for c in cls_scores]
        refine = refine_feats is not None
        plan = self._plan(batch, sizes, cfg, refine)
        levels = []
        cls_l, ctr_l, pose_l = _head_maps([t.detach() for t in cls_scores], [t.detach() for t in centernesses],
                                          [t.detach() for t in pose_preds])
        for l in range(num_levels):
            assert cls_scores[l].shape[-2:] == pose_preds[l].shape[-2:]
            d = dict(cls=cls_l[l], ctr=ctr_l[l], pose=pose_l[l], scales=self.scales[l])
            if refine:
                d["feats"] = [_to_f32(f.detach()) for f in refine_feats[l]]
            levels.append(d)
        plan.bind(levels)
        plan.set_metas(img_metas)
        plan.run()
        return plan

    def get_poses(self, cls_scores, pose_preds, centernesses, *rest, img_metas=None, cfg=None, rescale=None):
        """Reference call: get_poses(cls_scores, pose_preds, centernesses, img_metas, cfg=None, rescale=None)
        (das_head.py:653-659; img_metas / cfg / rescale positional or by keyword) with pose_preds already refined +
        eval-tailed (das_head.py:264-267 outputs).
        Extended call: get_poses(cls_scores, raw_pose_preds, centernesses, refine_feats, img_metas, ...)
        where refine_feats[level] is the list of per-layer feature maps; refinement then runs on the GPU.
        Returns the reference's list of per-image dicts (das_head.py:680-687) plus 'poses_cam' / 'poses_world'
        (valid for the shipped dataset flags norm_depth=True, abs_dz=True: cmupanoptic_mono_dataset.py:391-401;
        `dataset_depth_factor` is that file's :399 factor)."""
        metas = self._split_rest(rest, cfg, img_metas)[1]
        plan = self.decode_to_device(cls_scores, pose_preds, centernesses, *rest, img_metas=img_metas, cfg=cfg, rescale=rescale)
        return plan.results(metas)

    def simple_test_decode(self, outs, img_metas, rescale=False):
        """What DAS.simple_test does after the head forward (detectors/das.py:34-39)."""
        return self.get_poses(*outs, img_metas, rescale=rescale)
