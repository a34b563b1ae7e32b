#!/usr/bin/env python
"""Benchmark of the DAS dense-head inference decode (BASELINE.json metric: decoded images/s).

  python bench.py --gpus N --steps K --warmup W              this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...    reference algorithm on the host cores
                                                              (oracle port; rank 0 only)

One "step" = one decode of one batch of synthetic Panoptic-shaped head outputs (BASELINE config #2:
B=64 per GPU, J=15, 128x208 stride-8 map, K=nms_pre=nms_post=10, one refinement layer, C=256).
Multi-GPU is weak scaling: every rank decodes its own B=64 shard, the only traffic is one NCCL
all-gather of the ranks' packed pose lists per step.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from das_b200 import synth  # noqa: E402

WORKLOADS = {
    # BASELINE config #2 (the metric's configuration) -- the default
    "panoptic": dict(name="panoptic_decode_B64_J15_128x208_K10_L1", batch=64, h=128, w=208, stride=8, K=10, head=synth.PANOPTIC),
    # config #2 with trained-model-like fields: sampling heads ~8 px apart, joints ~10 px from the centre, so the 32 rows
    # of an item are 32 DISTINCT rows and the sampling phase's bytes really come from HBM (honest traffic)
    "panoptic_spread": dict(name="panoptic_spread_decode_B64_J15_128x208_K10_L1_heads8px", batch=64, h=128, w=208, stride=8, K=10,
                            head=synth.PANOPTIC, so_std=0.3, uv_scale=16.0),
    # BASELINE config #3: MuPoTS-shaped, 17 joints, 3 refinement layers (2 dense + 1 sparse), K=20
    "mupots": dict(name="mupots_decode_B64_J17_128x208_K20_L3", batch=64, h=128, w=208, stride=8, K=20, head=synth.MUPOTS17),
    # BASELINE config #4: crowded scene, 256x416 map, K=64
    "crowded": dict(name="crowded_decode_B32_J15_256x416_K64_L1", batch=32, h=256, w=416, stride=8, K=64, head=synth.PANOPTIC),
    # BASELINE config #1: one image (latency-bound: judge ms_per_step, not the roofline fraction)
    "single": dict(name="single_image_decode_B1_J15_128x208_K10_L1", batch=1, h=128, w=208, stride=8, K=10, head=synth.PANOPTIC),
    # BASELINE config #5: images through the whole network (run_model); h, w are IMAGE sizes here
    "e2e_model": dict(name="e2e_model_B16_per_gpu_1024x1664_J15_K10_L1", batch=16, h=1024, w=1664, stride=8, K=10, head=synth.PANOPTIC),
}
for _k, _v in WORKLOADS.items():
    _v["key"] = _k
WORKLOAD = dict(WORKLOADS["panoptic"])
TEST_CFG = dict(nms_pre=10, nms_post=10, nms_thr=0.9, score_thr=0.0)
METRIC = "decoded_images_per_sec"
UNIT = "images/s"


# ------------------------------------------------------------------------------------------------
_JSON_FD = None


def guard_stdout():
    """stdout carries exactly ONE JSON line: libraries that print to fd 1 (NCCL's version banner under NCCL_DEBUG) are
    sent to stderr for the whole run, and emit() writes the result to the real stdout."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(obj):
    line = (json.dumps(obj) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(line.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, line)


def pin_rank_to_cores(local_rank, world):
    """One rank per GPU: give every rank its own slice of the host cores nearest to ITS GPU (NVML's CPU affinity of the
    device = its NUMA node / root complex), so that 8 ranks do not migrate across each other and the pinned host buffers
    a rank allocates afterwards come from that node (first-touch).  Returns a short description for the JSON line."""
    try:
        allowed = sorted(os.sched_getaffinity(0))
        near = None
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(int(os.environ.get("CUDA_VISIBLE_DEVICES", "").split(",")[local_rank])
                                                  if os.environ.get("CUDA_VISIBLE_DEVICES") else local_rank)
            words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
            near = [c for c in allowed if (words[c // 64] >> (c % 64)) & 1]
        except Exception:
            near = None
        pool = near if near else allowed
        per = max(1, len(pool) // max(1, world))
        mine = pool[(local_rank % max(1, len(pool) // per)) * per:][:per] or pool
        os.sched_setaffinity(0, set(mine))
        return dict(cores=[mine[0], mine[-1]], n=len(mine), near_gpu=bool(near))
    except Exception as e:       # affinity is an optimisation, never a reason to fail
        return dict(error=repr(e)[:100])


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


class ClockSampler:
    """nvidia-smi clocks/throttle reasons while the GPU is under load (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index, enabled: bool = True):
        """index: one GPU index or a list of them (one poller for all GPUs of the job: NVML queries from one poller per
        rank measurably slowed every rank's launches at N = 8)."""
        self.index = ",".join(str(i) for i in index) if isinstance(index, (list, tuple)) else str(index)
        self.rows, self.proc, self.enabled = [], None, enabled

    def start(self):
        if not self.enabled:
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", self.index, f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self):
        if self.proc is None:
            return
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()

    def summary(self, t0=None, t1=None):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"], samples=0)
        sm, smax, reasons = [], [], set()
        for ts, line in self.rows:
            if t0 is not None and not (t0 - 0.05 <= ts <= t1 + 0.15):
                continue
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["no samples"], samples=0)
        return dict(sm_mhz=statistics.median(sm), sm_max_mhz=max(smax), reasons=sorted(reasons), samples=len(sm))


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------------------------------------
def cpu_reference_impl():
    """The CPU arm: the REFERENCE's own functions when they can be run here (live tree, or the oracle/_ref bundle that
    oracle/make_ref.py extracted from it -- kind "reference"), else the restatement in oracle/das_oracle.py (kind "port")."""
    from oracle import ref_extract as R
    if R.available():
        from oracle import make_golden as G

        def step(levels, layers, metas, head_cfg):
            return G.run_reference(head_cfg, levels, layers, metas, TEST_CFG)
        return step, "reference", ("the reference's own get_poses/_get_poses_single/offset_sample/oks_nms/pixel2world (extracted from its "
                                   "sources by oracle/make_ref.py), driven by restated mmcv-bound glue (Scale, 1x1 convs, eval tail)")
    from oracle import das_oracle as O

    def step(levels, layers, metas, head_cfg):
        return O.decode_full(levels, layers, metas, head_cfg.as_dict(), TEST_CFG)
    return step, "port", "oracle port of the reference's torch/NumPy code (reference tree and oracle/_ref bundle absent)"


def make_cpu_sample(n_img, seed):
    w = WORKLOAD
    levels = synth.make_levels(w["head"], n_img, w["h"], w["w"], seed=seed, peaks=16, uv_scale=w.get("uv_scale", 4.0))
    metas = synth.make_metas(n_img, w["h"], w["w"], stride=w["stride"], seed=seed + 2)
    return levels, metas


def time_cpu_baseline(budget_s=15.0, chunk=4):
    """Reference algorithm on the host cores, bounded sample of the same workload."""
    torch.set_num_threads(os.cpu_count() or 1)
    step, kind, what = cpu_reference_impl()
    layers = synth.make_layers(WORKLOAD["head"], seed=1235, so_std=WORKLOAD.get("so_std", 0.01))
    levels, metas = make_cpu_sample(chunk, 99)
    step(levels, layers, metas, WORKLOAD["head"])          # warm-up
    n, t0 = 0, time.perf_counter()
    while True:
        step(levels, layers, metas, WORKLOAD["head"])
        n += chunk
        el = time.perf_counter() - t0
        if el >= budget_s or n >= 64:
            break
    out = dict(value=n / el, unit=UNIT, cores=torch.get_num_threads(), kind=kind,
               sample=f"{n} images of the {WORKLOAD['name']} workload in chunks of {chunk} "
                      f"(dense refinement + eval tail + get_poses + OKS-NMS + back-projection), {el:.1f} s", implementation=what)
    out["split_ms_per_image"] = cpu_split(levels, layers, metas, chunk)
    if torch.cuda.is_available():
        out["eager_gpu"] = time_eager_gpu(layers, chunk)
    return out


def cpu_split(levels, layers, metas, chunk):
    """where the host time goes -- dense refinement + eval tail / get_poses incl. OKS-NMS / back-projection."""
    from oracle import das_oracle as O
    hc = WORKLOAD["head"].as_dict()
    t0 = time.perf_counter()
    pps = [O.head_eval_tail(lv["pose_raw"], lv["feats"], layers, lv["scales"], num_joints=hc["num_joints"],
                            num_heads=hc["num_heads"], root_idx=hc["root_idx"], depth_factor=hc["depth_factor"],
                            z_norm=hc["z_norm"], stride=lv["stride"]) for lv in levels]
    t1 = time.perf_counter()
    res = O.get_poses([lv["cls"] for lv in levels], pps, [lv["ctr"] for lv in levels], metas, TEST_CFG,
                      [lv["stride"] for lv in levels], hc["num_joints"])
    t2 = time.perf_counter()
    for r, m in zip(res, metas):
        O.backproject(r["poses"].numpy(), m["cam"]["K"], m["cam"]["R"], m["cam"]["t"], hc["root_idx"])
    t3 = time.perf_counter()
    return dict(refinement_and_tail=(t1 - t0) / chunk * 1e3, get_poses_and_nms=(t2 - t1) / chunk * 1e3,
                backprojection=(t3 - t2) / chunk * 1e3)


def time_eager_gpu(layers, chunk, budget_s=6.0):
    """SURVEY 8(d) secondary baseline: the same reference-order algorithm (dense refinement in eager PyTorch, per-image
    python decode with host NumPy OKS-NMS and its .cpu() syncs) with the tensors on the B200."""
    from oracle import das_oracle as O
    dev = torch.device("cuda", torch.cuda.current_device())
    levels, metas = make_cpu_sample(chunk, 99)
    levels = synth.levels_to(levels, dev)
    layers = synth.layers_to(layers, dev)
    hc = WORKLOAD["head"].as_dict()
    O.decode_full(levels, layers, metas, hc, TEST_CFG)
    torch.cuda.synchronize()
    n, t0 = 0, time.perf_counter()
    while True:
        O.decode_full(levels, layers, metas, hc, TEST_CFG)
        torch.cuda.synchronize()
        n += chunk
        el = time.perf_counter() - t0
        if el >= budget_s or n >= 256:
            break
    return dict(value=n / el, unit=UNIT, kind="port, eager PyTorch on the GPU",
                sample=f"{n} images in chunks of {chunk}, {el:.1f} s")


def run_reference(args):
    rank, _, world = dist_env()
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)
    step, kind, what = cpu_reference_impl()
    head = WORKLOAD["head"]
    layers = synth.make_layers(head, seed=1235, so_std=WORKLOAD.get("so_std", 0.01))
    lv1, m1 = make_cpu_sample(1, 7)
    step(lv1, layers, m1, head)
    t = time.perf_counter()
    step(lv1, layers, m1, head)
    t_img = time.perf_counter() - t
    total = max(args.steps + args.warmup, 1)
    sample_b = int(max(1, min(16, (150.0 / total) / max(t_img, 1e-3))))   # <=16 images: ~3 GB of dense temporaries
    levels, metas = make_cpu_sample(sample_b, 1234)
    for _ in range(args.warmup):
        step(levels, layers, metas, head)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step(levels, layers, metas, head)
    el = time.perf_counter() - t0
    value = sample_b * args.steps / el
    cores = torch.get_num_threads()
    sample = f"{sample_b} images per step of the {WORKLOAD['name']} workload, torch CPU threads={cores}"
    emit({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": el / max(args.steps, 1) * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD["name"], "images_per_step": sample_b, "J": head.num_joints, "map": f"{WORKLOAD['h']}x{WORKLOAD['w']}",
                   "K": WORKLOAD["K"], "refine_layers": head.num_layers, "note": what + ", on the host cores"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    })


# ------------------------------------------------------------------------------------------------
STAGES = ("score_topk", "dense_layers", "refine_phase12", "refine_assemble", "nms_backproject")


def stage_bytes(w, head, mode):
    """SURVEY.md 8(d) algorithmic bytes of every stage of ONE batch (fp32): the per-unit figures times the units a launch
    processes.  scan: 2*4*H*W per image; dense layers: (L-1) * (C*4 + (3+3J)*4) * H*W; sparse refinement: feature rows of
    C*4 bytes per (centre, joint): 1 + 4 in phases 1-2, 32 in the sampling phase; NMS: CT*(3J+4)*4."""
    B, H, W, K, J, C, L = w["batch"], w["h"], w["w"], w["K"], head.num_joints, head.feat_channels, head.num_layers
    scan = B * 2 * 4 * H * W
    dense = (L - 1) * B * H * W * (C * 4 + (3 + 3 * J) * 4)
    p12 = B * K * J * 5 * C * 4
    p3 = B * K * J * 32 * C * 4
    nms = B * K * (3 * J + 4) * 4
    if mode == 0:
        return dict(score_topk=scan, dense_layers=dense, refine_phase12=0, refine_assemble=p12 + p3, nms_backproject=nms)
    return dict(score_topk=scan, dense_layers=dense, refine_phase12=p12, refine_assemble=p3, nms_backproject=nms)


KERNEL_OF_STAGE = {
    "score_topk": "score_topk_kernel", "dense_layers": "dense_project_tc_kernel (every dense layer) + dense_sample2_kernel (every dense layer but the last, whose sampling the sparse last layer evaluates on demand)",
    "refine_phase12": "refine_heads8_kernel (phases 1-2, 4 candidates per warp; warp-per-item kernel for small decodes)", "refine_assemble": None,
    "nms_backproject": "nms_backproject_kernel"}


class Runner:
    """One workload on one rank: `n_sets` rotating device-resident input sets (each far larger than L2), one plan per
    set, pipelined over `n_streams` CUDA streams; multi-rank result collection fused into the NMS kernel (P2P stores
    into every peer's gathered buffer over NVLink) or, as the fallback / comparison, NCCL all-gathers."""

    def __init__(self, wname, args, rank, world, dev):
        from das_b200.head import DecodePlan
        import ctypes as C
        self.C = C
        self.w = w = dict(WORKLOADS[wname])
        self.head = head = w["head"]
        self.rank, self.world, self.dev, self.args = rank, world, dev, args
        B = w["batch"]
        self.tc = dict(TEST_CFG, nms_pre=w["K"], nms_post=w["K"])
        self.n_sets = n_sets = max(1, args.input_sets if w["batch"] > 1 else min(args.input_sets, 4))
        layers = synth.make_layers(head, seed=1235, device=dev, so_std=w.get("so_std", 0.01))
        self.metas = synth.make_metas(B, w["h"], w["w"], stride=w["stride"], seed=1236 + rank)
        mode_env = int(os.environ["DAS_REFINE_MODE"]) if "DAS_REFINE_MODE" in os.environ else None
        self.collect = "none" if world == 1 else args.collect
        n_plans = n_sets * (2 if self.collect == "nccl" else 1)
        self.keep, self.plans = [], []
        for s in range(n_sets):
            self.keep.append(synth.make_levels(head, B, w["h"], w["w"], seed=1234 + 17 * s + 1000 * rank, device=dev,
                                               peaks=max(16, (3 * w["K"]) // 2), uv_scale=w.get("uv_scale", 4.0),
                                               margin_for=dict(nms_pre=w["K"], score_thr=0.0)))
        for i in range(n_plans):
            plan = DecodePlan(num_joints=head.num_joints, root_idx=head.root_idx, depth_factor=head.depth_factor,
                              z_norm=head.z_norm, strides=head.strides, level_sizes=[(w["h"], w["w"])], batch=B,
                              test_cfg=self.tc, num_heads=head.num_heads, feat_channels=head.feat_channels,
                              num_layers=head.num_layers, refine=True, device=dev, refine_mode=mode_env)
            plan.set_weights(layers)
            plan.bind([dict(cls=lv["cls"], ctr=lv["ctr"], pose=lv["pose_raw"], feats=lv["feats"], scales=lv["scales"])
                       for lv in self.keep[i % n_sets]])
            plan.set_metas(self.metas)
            self.plans.append(plan)
        # dense layers (num_layers > 1) fill the GPU by themselves: concurrent batches only thrash L2, so they run on one stream
        self.n_streams = 1 if head.num_layers > 1 else max(1, min(args.streams, n_sets))
        self.streams = ([torch.cuda.Stream(device=dev) for _ in range(self.n_streams)] if self.n_streams > 1
                        else [torch.cuda.current_stream(dev)])
        self.stream_ptrs = [st.cuda_stream for st in self.streams]
        self.comm_stream = None
        self.gathers = 0
        self._setup_collection()
        # every plan: one eager run (module load, attributes) and one run that captures its graph -- outside any timing
        for p in self.plans:
            p.run()
            p.run()
        torch.cuda.synchronize()

    # ---- result collection across ranks -------------------------------------------------------------------------
    def _setup_collection(self):
        import torch.distributed as dist
        world, rank, dev = self.world, self.rank, self.dev
        self.stride = self.plans[0].block_stride
        self.nb = int(self.plans[0].output_block().numel())
        if self.collect == "p2p":
            lib = self.plans[0].lib
            C = self.C
            nbytes = world * self.n_sets * self.stride
            ptr, handle = C.c_void_p(), C.create_string_buffer(64)
            ok = lib.das_ipc_alloc(nbytes, C.byref(ptr), handle) == 0
            flags = [None] * world
            dist.all_gather_object(flags, (ok, handle.raw if ok else b""))
            peers = {}
            if all(f[0] for f in flags):
                for r in range(world):
                    if r == rank:
                        continue
                    q = C.c_void_p()
                    if lib.das_ipc_open(flags[r][1], C.byref(q)) != 0:
                        ok = False
                        break
                    peers[r] = q.value
            oks = [None] * world
            dist.all_gather_object(oks, bool(ok) and all(f[0] for f in flags))
            if not all(oks):
                if rank == 0:
                    sys.stderr.write("bench.py: CUDA IPC / peer access unavailable, falling back to NCCL all-gathers\n")
                self.collect = "nccl"
                from das_b200.head import DecodePlan  # noqa: F401
                # the nccl path needs a second set of plans (double-buffered staging); build them by re-running setup
                extra = []
                for i in range(self.n_sets):
                    src = self.plans[i]
                    p2 = type(src)(num_joints=self.head.num_joints, root_idx=self.head.root_idx, depth_factor=self.head.depth_factor,
                                   z_norm=self.head.z_norm, strides=self.head.strides, level_sizes=[(self.w["h"], self.w["w"])],
                                   batch=self.w["batch"], test_cfg=self.tc, num_heads=self.head.num_heads,
                                   feat_channels=self.head.feat_channels, num_layers=self.head.num_layers, refine=True, device=dev)
                    p2.set_weights(synth.make_layers(self.head, seed=1235, device=dev, so_std=self.w.get("so_std", 0.01)))
                    p2.bind([dict(cls=lv["cls"], ctr=lv["ctr"], pose=lv["pose_raw"], feats=lv["feats"], scales=lv["scales"])
                             for lv in self.keep[i]])
                    p2.set_metas(self.metas)
                    extra.append(p2)
                self.plans += extra
            else:
                self.ipc_ptr, self.peer_ptrs = ptr.value, peers
                # gathered[r][s] = rank r's block of input set s; this rank writes its own row locally and into every peer
                for s, p in enumerate(self.plans):
                    off = (rank * self.n_sets + s) * self.stride
                    p.set_output_ptr(ptr.value + off, self.stride)
                    p.set_peer_blocks([peers[r] + off for r in sorted(peers)])
                self.gathered = torch.as_tensor(_DevBytes(ptr.value, nbytes), device=dev).view(world, self.n_sets, self.stride)
        if self.collect == "nccl":
            n_sets = self.n_sets
            self.staging = torch.zeros((2, n_sets, self.stride), dtype=torch.uint8, device=dev)
            self.gathered_nccl = torch.empty((2, world, n_sets * self.stride), dtype=torch.uint8, device=dev)
            for i, p in enumerate(self.plans):
                p.set_output_block(self.staging[i // n_sets, i % n_sets])
            self.comm_stream = torch.cuda.Stream(device=dev)
            self.decoded = [torch.cuda.Event() for _ in self.plans]
            self.gathered_ev = [None, None]

    def _flush(self, half, n):
        import torch.distributed as dist
        for k in range(n):
            self.comm_stream.wait_event(self.decoded[half * self.n_sets + k])
        with torch.cuda.stream(self.comm_stream):
            dist.all_gather_into_tensor(self.gathered_nccl[half].view(-1), self.staging[half].view(-1))
            ev = torch.cuda.Event()
            ev.record(self.comm_stream)
            self.gathered_ev[half] = ev
        self.gathers += 1

    def loop(self, steps):
        """Enqueue `steps` decodes round-robin over the plans / streams (no host synchronisation inside)."""
        plans, ptrs, n_plans, ns = self.plans, self.stream_ptrs, len(self.plans), self.n_streams
        if self.collect != "nccl":
            for i in range(steps):
                plans[i % n_plans].run(stream=ptrs[i % ns])       # p2p: the collective is inside the NMS kernel
            return
        n_sets = self.n_sets
        for i in range(steps):
            k = i % n_plans
            half, slot = divmod(k, n_sets)
            st = self.streams[i % ns]
            if slot == 0 and self.gathered_ev[half] is not None:
                for s2 in self.streams:
                    s2.wait_event(self.gathered_ev[half])        # this half's previous gather has read the staging buffer
            plans[k].run(stream=ptrs[i % ns])
            self.decoded[k].record(st)
            if slot == n_sets - 1:
                self._flush(half, n_sets)
        rem = steps % n_sets
        if rem:
            self._flush(((steps - 1) % n_plans) // n_sets, rem)

    def fork(self):
        cur = torch.cuda.current_stream(self.dev)
        if self.n_streams > 1:
            for st in self.streams:
                st.wait_stream(cur)
        if self.comm_stream is not None:
            self.comm_stream.wait_stream(cur)

    def join(self):
        cur = torch.cuda.current_stream(self.dev)
        if self.collect == "p2p":
            for p in self.plans:          # the timed region ends when every peer has every result
                p.publish_wait()
        if self.n_streams > 1:
            for st in self.streams:
                cur.wait_stream(st)
        if self.comm_stream is not None:
            cur.wait_stream(self.comm_stream)

    def barrier(self):
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def timed(self, steps, warmup, min_seconds):
        """W warm-up steps, then the timed region: exactly `steps` steps bracketed by barrier + synchronize, device-timed
        with CUDA events, max over ranks -- repeated until the region has lasted `min_seconds` (a 20-step region is
        1.7 ms: too short for the clock sampler and at the mercy of one scheduling hiccup); the MEDIAN repetition is
        reported.  Returns (ms per `steps` steps, list of all repetitions, wall-clock window, launches per repetition)."""
        import torch.distributed as dist
        self.fork()
        self.loop(max(warmup, 3))
        self.join()
        self.barrier()
        reps, t_wall0 = [], time.time()
        launches0 = sum(p.kernel_launches for p in self.plans)
        if self.args.profile_region:
            torch.cuda.profiler.start()
        while True:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            self.fork()
            self.loop(steps)
            self.join()
            e1.record()
            self.barrier()
            ms = e0.elapsed_time(e1)
            if self.world > 1:
                t = torch.tensor([ms, time.time() - t_wall0], device=self.dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms, elapsed = float(t[0]), float(t[1])
            else:
                elapsed = time.time() - t_wall0
            reps.append(ms)
            if elapsed >= min_seconds or len(reps) >= 2000 or self.args.profile_region:
                break
        if self.args.profile_region:
            torch.cuda.profiler.stop()
        t_wall1 = time.time()
        launches = (sum(p.kernel_launches for p in self.plans) - launches0) // len(reps)
        return statistics.median(reps), reps, (t_wall0, t_wall1), launches

    def stage_profile(self, n_prof):
        stage = np.zeros(5)
        n_sets = self.n_sets
        for i in range(3):
            self.plans[i % n_sets].run(stage_events=True)
        torch.cuda.synchronize()
        for i in range(n_prof):
            p = self.plans[i % n_sets]
            p.run(stage_events=True)
            torch.cuda.synchronize()
            stage += np.array(p.stage_ms())
        return stage / n_prof

    def serial_replay(self, n=200):
        """One decode at a time: graph replays back to back on ONE stream over the rotating input sets, with programmatic
        dependent launch along the kernel chain (das_plan_set_pdl(1), the latency mode).  ms per decode."""
        plans = self.plans[:self.n_sets]
        for p in plans:
            p.set_pdl(1)
        try:
            for i in range(2 * len(plans)):
                plans[i % len(plans)].run()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(n):
                plans[i % len(plans)].run()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / n
        finally:
            for p in plans:
                p.set_pdl(-1)

    def roofline(self, stage, n_prof):
        """Roofline of the dominant kernel.  Two byte counts are reported for the sparse refinement stages:
        * `frac` (`achieved`, `algorithmic_bytes_per_launch`): the task's definition -- SURVEY.md 8(d)'s per-unit figure (32
          feature rows of C*4 bytes per (centre, joint) in the sampling phase) x the units of one launch / time / measured HBM
          peak.  It can EXCEED 1: the sampling-phase GEMM multiplies every DISTINCT (cell, joint) row once (the 32 (head,
          corner) rows of an item that land on the same cell share one row), so the kernel no longer moves SURVEY's bytes;
        * `gathered_frac` (`gathered_bytes_per_launch`): the bytes the kernel REALLY gathers = distinct rows x C x 4, counted on
          the device in this run (das_plan_refine_stats) -- the bandwidth picture of the kernel as built;
        * `dram_frac`: ncu's DRAM bytes of the same kernels / time / peak (what came from HBM rather than L2).
        `survey_frac` repeats `frac` under the name the round-2 documents use."""
        w, head = self.w, self.head
        mode = self.plans[0].refine_mode
        sb = stage_bytes(w, head, mode)
        sg = dict(sb)                                   # ... with the sampling phase at its distinct gathered rows
        dedup = None
        if mode != 0 and head.num_layers >= 1:
            rows, n_valid, rows_nodedup = self.plans[0].refine_stats()
            if rows > 0:
                sg["refine_assemble"] = rows * head.feat_channels * 4
                dedup = dict(distinct_rows=rows, rows_without_dedup=rows_nodedup, valid_candidates=n_valid,
                             distinct_rows_per_item=rows / max(1, n_valid * head.num_joints))
        ms = dict(zip(STAGES, [float(x) for x in stage]))
        dom = max(STAGES, key=lambda k: ms[k])
        kernel = KERNEL_OF_STAGE[dom]
        if dom == "refine_assemble":
            kernel = ("refine_sparse_kernel (fp32 SIMT: phases 1-3)" if mode == 0 else
                      "refine_tc2_kernel (tcgen05 %s: gathered GEMM over the distinct sampled rows) + refine_finish_kernel"
                      % ("3xTF32" if mode == 1 else "TF32"))
        peak, peak_src = measured_peak()

        def gbs(nbytes, t_ms):
            return nbytes / (t_ms / 1e3) / 1e9

        achieved = gbs(sb[dom], ms[dom])
        traffic = ncu_traffic(w["key"], dom)
        path_bytes, path_gathered = sum(sb.values()), sum(sg.values())
        total_ms = float(stage.sum())
        out = dict(bound="hbm", kernel=kernel, stage=dom, achieved=achieved, peak=peak, unit="GB/s", frac=achieved / peak,
                   survey_frac=achieved / peak,
                   gathered_achieved=gbs(sg[dom], ms[dom]), gathered_frac=gbs(sg[dom], ms[dom]) / peak,
                   traffic=traffic, dram_frac=(gbs(traffic, ms[dom]) / peak) if traffic else None,
                   traffic_over_algorithmic=(traffic / sb[dom]) if traffic else None,
                   traffic_over_gathered=(traffic / sg[dom]) if traffic else None,
                   peak_source=peak_src, algorithmic_bytes_per_launch=sb[dom], gathered_bytes_per_launch=sg[dom],
                   kernel_ms=ms[dom], stage_ms=ms, stage_algorithmic_bytes=sb, stage_gathered_bytes=sg, dedup=dedup,
                   stage_frac={k: (gbs(sb[k], ms[k]) / peak if ms[k] > 0 else None) for k in STAGES},
                   stage_gathered_frac={k: (gbs(sg[k], ms[k]) / peak if ms[k] > 0 else None) for k in STAGES},
                   path=dict(algorithmic_bytes=path_bytes, gathered_bytes=path_gathered, ms=total_ms,
                             frac=gbs(path_bytes, total_ms) / peak, survey_frac=gbs(path_bytes, total_ms) / peak,
                             gathered_frac=gbs(path_gathered, total_ms) / peak,
                             note="whole decode over the summed single-stream stage times"),
                   how=f"CUDA-event nodes inside the replayed graph (single stream), mean of {n_prof} replays with a sync between them; "
                       "frac = SURVEY 8(d) algorithmic bytes (32 feature rows per (centre, joint) in the sampling phase) / time / peak "
                       "-- above 1 where the kernel multiplies each DISTINCT sampled row once instead of moving those bytes; "
                       "gathered_frac = the distinct rows it really gathers (device counter of the same run) / time / peak; "
                       "dram_frac = ncu dram__bytes of the same kernels / time / peak")
        return out

    def close(self):
        if getattr(self, "ipc_ptr", None):
            lib = self.plans[0].lib
            self.barrier()
            for q in self.peer_ptrs.values():
                lib.das_ipc_close(self.C.c_void_p(q))
            self.plans = []
            self.barrier()
            lib.das_ipc_free(self.C.c_void_p(self.ipc_ptr))
        self.plans, self.keep = [], []
        torch.cuda.empty_cache()


class _DevBytes:
    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = dict(shape=(int(n),), typestr="|u1", data=(int(ptr), False), version=2, strides=None)


def ncu_traffic(workload_key, stage):
    """dram bytes per launch of a stage's kernel from the committed ncu capture (profiles/dominant_kernel_traffic.json:
    {workload: {stage: bytes}}), or None."""
    p = os.path.join(ROOT, "profiles", "dominant_kernel_traffic.json")
    try:
        d = json.load(open(p))
        v = d.get(workload_key, {}).get(stage)
        return float(v) if v else None
    except Exception:
        return None


def run_b200(args):
    import torch.distributed as dist

    rank, local_rank, world = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the decode path has no CPU fallback")
    pinning = pin_rank_to_cores(local_rank, world) if world > 1 else None
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")    # NCCL's version / debug lines: keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=dev)
    w, head = WORKLOAD, WORKLOAD["head"]
    B = w["batch"]

    sampler = ClockSampler(list(range(world)) if world > 1 else local_rank, enabled=rank == 0)
    sampler.start()
    run = Runner(w["key"], args, rank, world, dev)
    ms, reps, (t_wall0, t_wall1), launches = run.timed(args.steps, args.warmup, args.min_seconds)
    value = world * B * args.steps / (ms / 1e3)

    collect_check = None
    if run.collect == "p2p":
        # the fused all-gather delivered what NCCL would have: compare every peer's row with an NCCL gather of the local blocks
        local = run.gathered[rank].contiguous()
        ref = torch.empty((world,) + tuple(local.shape), dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(ref.view(-1), local.view(-1))
        nb = run.nb
        same = bool(torch.equal(ref[:, :, :nb], run.gathered[:, :, :nb]))
        flag = torch.tensor([int(same)], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        collect_check = "ok" if int(flag) == 1 else "MISMATCH"

    n_prof = max(min(args.steps, 200), 5)
    stage = run.stage_profile(n_prof)
    roofline = run.roofline(stage, n_prof)
    step_ms = ms / args.steps
    ser_ms = run.serial_replay(200 if w["key"] != "mupots" else 10)
    roofline["path"]["serial_replay"] = dict(
        ms=ser_ms, frac=roofline["path"]["algorithmic_bytes"] / (ser_ms / 1e3) / 1e9 / roofline["peak"],
        survey_frac=roofline["path"]["algorithmic_bytes"] / (ser_ms / 1e3) / 1e9 / roofline["peak"],
        gathered_frac=roofline["path"]["gathered_bytes"] / (ser_ms / 1e3) / 1e9 / roofline["peak"],
        note="one decode at a time: graph replays back to back on ONE stream, programmatic dependent launch on "
             "(the stage times above carry event nodes between the kernels, which rules PDL out)")
    roofline["path"]["pipelined"] = dict(
        ms=step_ms, frac=roofline["path"]["algorithmic_bytes"] / (step_ms / 1e3) / 1e9 / roofline["peak"],
        survey_frac=roofline["path"]["algorithmic_bytes"] / (step_ms / 1e3) / 1e9 / roofline["peak"],
        gathered_frac=roofline["path"]["gathered_bytes"] / (step_ms / 1e3) / 1e9 / roofline["peak"],
        note="the same byte counts over the timed step (independent batches pipelined on %d streams)" % run.n_streams)

    # ---- end to end through the host-buffer C-ABI entry: pinned host inputs, H2D + decode + D2H ---
    e2e = None
    plans, metas = run.plans, run.metas
    J, K, C = head.num_joints, w["K"], head.feat_channels
    if not args.no_e2e:
        def pin_set(levels):
            lv0 = levels[0]
            return [dict(cls=lv0["cls"].cpu().pin_memory(), ctr=lv0["ctr"].cpu().pin_memory(),
                         pose=lv0["pose_raw"].cpu().pin_memory(),
                         feats=[f.permute(0, 2, 3, 1).cpu().pin_memory().permute(0, 3, 1, 2) for f in lv0["feats"]],
                         scales=lv0["scales"])]
        # two distinct pinned input sets alternate, so no call sees the host tensors of the call before it
        # one pinned set per call in flight on one GPU (depth 2 / 3 / 4: 72 k / 78 k / 80 k images/s); two per rank when several
        # ranks pin host memory at once (calls two apart then read the same, read-only, set)
        host_sets = [pin_set(run.keep[s]) for s in range(min(max(2, args.e2e_depth if world == 1 else 2), run.n_sets))]
        plan0 = plans[0]
        if run.collect == "p2p":
            plan0.set_peer_blocks([])
        host_out = plan0.alloc_host_out(pinned=True)
        n_e2e = max(min(args.steps, args.e2e_steps), 1)

        def n_calls(warm_s_per_call):
            # at least --e2e-steps calls and at least ~0.15 s of them (a 1-ms call timed over 10 calls is mostly noise)
            return int(max(n_e2e, min(400, math.ceil(0.15 / max(warm_s_per_call, 1e-5)))))

        def max_over_ranks(el):
            if world > 1:
                tel = torch.tensor([el], device=dev)
                dist.all_reduce(tel, op=dist.ReduceOp.MAX)
                el = float(tel.item())
            return el

        def time_host(zero_copy, row_cache=True):
            plan0.set_host_mode(zero_copy, row_cache)
            plan0.run_host(host_sets[0], metas, host_out)
            t0 = time.perf_counter()
            for i in range(2):
                plan0.run_host(host_sets[(i + 1) % len(host_sets)], metas, host_out)
            n = n_calls((time.perf_counter() - t0) / 2)
            run.barrier()
            t0 = time.perf_counter()
            for i in range(n):
                plan0.run_host(host_sets[i % len(host_sets)], metas, host_out)     # synchronises its stream before returning
            run.barrier()
            return max_over_ranks(time.perf_counter() - t0) / n, plan0.h2d_explicit_bytes, n

        def time_host_pipelined(depth):
            # `depth` plans on `depth` streams driven in turn through the asynchronous entry (das_plan_run_host_async): the
            # H2D copy of batch n+1 and the D2H of batch n-1 overlap the kernels of batch n, which read their rows in place
            # over the same PCIe link.  Every call still copies its inputs from pinned host memory and its results back; a
            # plan is reused only after the stream of its previous call has drained (its host_out is complete by then).
            pp = plans[:depth]
            outs = [q.alloc_host_out(pinned=True) for q in pp]
            sts = [torch.cuda.Stream(device=dev) for _ in pp]
            for q in pp:
                if run.collect == "p2p":
                    q.set_peer_blocks([])
                q.set_host_mode(True, True)

            def loop(n):
                for i in range(n):
                    k = i % depth
                    sts[k].synchronize()
                    with torch.cuda.stream(sts[k]):
                        pp[k].run_host(host_sets[k % len(host_sets)], metas, outs[k], sync=False)
                for st in sts:
                    st.synchronize()

            loop(depth)
            t0 = time.perf_counter()
            loop(2 * depth)
            n = n_calls((time.perf_counter() - t0) / (2 * depth))
            n -= n % depth
            run.barrier()
            t0 = time.perf_counter()
            loop(n)
            run.barrier()
            el = max_over_ranks(time.perf_counter() - t0) / n
            # same bits as the synchronous call on the same host set
            plan0.run_host(host_sets[0], metas, host_out)
            same = all(torch.equal(outs[0][k], host_out[k]) for k in host_out)
            return el, n, "ok" if same else "MISMATCH"

        spc_bulk, bytes_bulk, n_bulk = time_host(False)
        spc_nc, _, _ = time_host(True, False)
        spc_zc, bytes_zc, n_zc = time_host(True)
        depth = min(max(args.e2e_depth, 1), len(plans))
        pipe = time_host_pipelined(depth) if depth > 1 else None
        el_bulk, el_nc, el_zc = spc_bulk * n_e2e, spc_nc * n_e2e, spc_zc * n_e2e     # per --e2e-steps calls (the formulas below)
        rows, n_valid = plan0.row_cache_stats()
        row_bytes = (rows + n_valid) * C * 4
        # one 32-byte sector per scattered pose value: root depth / offsets per candidate, and (u, v, d) at the candidate's own
        # cell and at every DISTINCT sampled cell of every (candidate, joint) item (device counter of the same call);
        # cross-checked with ncu pcie__read_bytes per kernel: profiles/r02_e2e_probe_mode2.csv (30.7 MB in place + 6.8 MB copied)
        distinct_rows = plan0.refine_stats()[0]
        pose_bytes = (n_valid * 3 + (distinct_rows + n_valid * J) * 3) * 32 if head.num_layers == 1 else 0
        sparse_ub = B * K * J * 37 * C * 4 + (B * K * (3 + J * 33 * 3) * 32 if head.num_layers == 1 else 0)
        e2e_s = pipe[0] if pipe else spc_zc            # seconds per call
        e2e = dict(value=world * B / e2e_s, unit=UNIT, h2d_bytes_per_step=int(bytes_zc + row_bytes + pose_bytes),
                   d2h_bytes_per_step=plan0.d2h_bytes, steps=pipe[1] if pipe else n_zc, ms_per_step=e2e_s * 1e3,
                   pipeline_depth=depth if pipe else 1, pipelined_check=pipe[2] if pipe else None,
                   serial=dict(value=world * B / spc_zc, ms_per_step=spc_zc * 1e3, steps=n_zc,
                               note="one synchronous das_plan_run_host call at a time (what a single call costs)"),
                   h2d_explicit_bytes=int(bytes_zc), h2d_in_place_row_bytes=int(row_bytes),
                   h2d_in_place_pose_bytes_estimate=int(pose_bytes), h2d_in_place_bytes_upper_bound=int(sparse_ub),
                   h2d_how="explicit cudaMemcpyAsync bytes + distinct feature rows the device row cache fetched over PCIe (counted on "
                           "the device: %d rows + %d candidate rows of %d B) + one 32-B sector per pose value read in place (candidate cells and the "
                           "distinct sampled cells, counted on the device)" % (rows, n_valid, C * 4),
                   host_input_sets=len(host_sets),
                   api="das_plan_run_host_async on %d plans / streams in turn (stream sync before a plan is reused), host_mode 2: " % depth +
                       "pinned host inputs; logit planes H2D-copied, pose / feature maps read "
                       "in place over PCIe by the gather kernels (the decode touches ~5 % of them), every distinct row of the "
                       "sampling phase copied once into a device row cache -> graph replay -> D2H of the packed pose lists",
                   no_row_cache=dict(value=world * B * n_e2e / el_nc, ms_per_step=el_nc / n_e2e * 1e3,
                                     api="das_plan_run_host, host_mode 1: as above without the row-cache pass"),
                   bulk_copy=dict(value=world * B * n_e2e / el_bulk, ms_per_step=el_bulk / n_e2e * 1e3,
                                  h2d_bytes_per_step=int(bytes_bulk),
                                  api="das_plan_run_host, host_mode 0: every input map H2D-copied"))
        del host_sets
    clocks = sampler.summary(t_wall0, t_wall1)
    clocks["span"] = "timed region"
    in_bytes = plans[0].h2d_bytes
    n_sets, n_streams, collect = run.n_sets, run.n_streams, run.collect
    run.close()

    # ---- the other BASELINE configs as short device-timed sub-runs, so that they land in the driver's record --------
    extras = {}
    if not args.no_extra and w["key"] == "panoptic" and world == 1:
        for name in ("panoptic_spread", "single", "crowded", "mupots"):
            try:
                sub = Runner(name, args, rank, world, dev)
                steps = {"single": 400, "panoptic_spread": 200, "crowded": 100, "mupots": 12}[name]
                sms, sreps, _, _ = sub.timed(steps, 3, 0.15)
                sst = sub.stage_profile(5 if name == "mupots" else 20)
                sser = sub.serial_replay(10 if name == "mupots" else 100)
                sroof = sub.roofline(sst, 5 if name == "mupots" else 20)
                sw = sub.w
                extras[name] = dict(workload=sw["name"], value=world * sw["batch"] * steps / (sms / 1e3), unit=UNIT,
                                    ms_per_step=sms / steps, steps=steps, repetitions=len(sreps),
                                    serial_latency_ms=float(sst.sum()), serial_replay_ms=sser,
                                    roofline={k: sroof[k] for k in ("kernel", "stage", "frac", "gathered_frac", "dram_frac", "achieved", "peak",
                                                                    "traffic", "algorithmic_bytes_per_launch", "gathered_bytes_per_launch",
                                                                    "kernel_ms", "stage_ms", "stage_frac", "dedup")},
                                    path_frac=sroof["path"]["frac"], path_gathered_frac=sroof["path"]["gathered_frac"])
                sub.close()
                del sub
            except Exception as e:       # an extra must never take the headline line down with it
                extras[name] = dict(error=repr(e)[:300])
    sampler.stop()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = time_cpu_baseline(args.cpu_budget)

    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "timed_region_s": float(sum(reps)) / 1e3, "repetitions": len(reps),
            "repetition_ms": dict(median=ms, min=min(reps), max=max(reps)),
            "config": {"workload": w["name"], "images_per_gpu": B, "J": J, "map": f"{w['h']}x{w['w']}", "K": K,
                       "refine_layers": head.num_layers, "feat_channels": C, "test_cfg": dict(TEST_CFG),
                       "algorithm": "sparse last-layer refinement at the selected centres (SURVEY 8.0 divergence B)",
                       "l2": f"{n_sets} distinct input sets of {in_bytes / 1e9:.2f} GB rotated round-robin "
                             f"(each far larger than the 126 MB L2)",
                       "streams": n_streams,
                       "timing": f"{args.steps} steps per repetition, repeated until the timed region lasted >= {args.min_seconds} s; "
                                 "value and ms_per_step are the MEDIAN repetition (each one device-timed, max over ranks)",
                       "parallelism": (f"dp{world} batch-sharded, no traffic inside the path; result collection: "
                                       + ("fused into nms_backproject_kernel -- every rank stores its packed pose lists straight into each "
                                          "peer's gathered buffer over NVLink (IPC-mapped P2P stores), no NCCL kernel" if collect == "p2p"
                                          else f"one NCCL all-gather of the packed pose lists per {n_sets} steps on its own stream, plans write "
                                               "straight into the double-buffered staging block"))
                                      if world > 1 else "single GPU"},
            "clocks": clocks, "gpu_launches": int(launches) * world, "roofline": roofline,
        }
        if collect_check is not None:
            out["config"]["collect_check"] = collect_check
        if pinning is not None:
            out["config"]["host_cores_rank0"] = pinning
        if extras:
            out["extra_workloads"] = extras
        if e2e is not None:
            out["e2e"] = e2e
        if cpu is not None:
            out["cpu_baseline"] = cpu
        emit(out)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_model(args):
    """BASELINE config #5: image -> MSPN2 (2 stages, [3,4,6,3]) + FPN + head towers (PyTorch/cuDNN) -> CUDA decode.
    16 images of 1024x1664 per GPU, weak scaling; random-init weights, synthetic images."""
    import torch.distributed as dist
    from das_b200.head import DASHeadB200
    from das_b200.model import DASNet

    rank, local_rank, world = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the decode path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    w = WORKLOAD
    B, H, W, J, K = w["batch"], w["h"], w["w"], 15, w["K"]
    strides = (8, 16, 32, 64)
    dtype = None if os.environ.get("DAS_MODEL_DTYPE", "bf16") == "tf32" else torch.bfloat16
    sampler = ClockSampler(list(range(world)) if world > 1 else local_rank, enabled=rank == 0)
    sampler.start()
    torch.manual_seed(1238)
    net = DASNet(num_joints=J, strides=strides)
    with torch.no_grad():      # a random-init head is flat; widen the predictors so the decode sees separated peaks
        for branch, gain in ((net.towers.cls_out, 40.0), (net.towers.centerness_out, 20.0), (net.towers.uvd_out, 30.0)):
            branch[1].weight.mul_(gain)
    net = net.to(dev).prepare_inference(dtype)
    test_cfg = dict(nms_pre=K, nms_post=K, nms_thr=0.9, score_thr=0.0)
    head = DASHeadB200(1, 256, num_joints=J, strides=strides, depth_factor=20, z_norm=50, root_idx=2,
                       recursive_update=dict(num_heads=4, feat_channels=256, num_layers=1), test_cfg=test_cfg, device=dev)
    head.scales = net.level_scales()
    head.load_refine_weights(net.refine_weights())
    metas = synth.make_metas(B, H // 8, W // 8, stride=8, seed=1236 + rank)
    g = torch.Generator().manual_seed(1234 + rank)
    host_imgs = [torch.randn(B, 3, H, W, generator=g).pin_memory() for _ in range(2)]
    dev_imgs = [t.to(dev) for t in host_imgs]

    ev_model = [torch.cuda.Event(enable_timing=True) for _ in range(3)]

    def step(i, mark=False):
        with torch.no_grad():
            if mark:
                ev_model[0].record()
            outs = net(dev_imgs[i % 2])
            if mark:
                ev_model[1].record()
            plan = head.decode_to_device(*outs, metas)
            if mark:
                ev_model[2].record()
        return plan

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(max(args.warmup, 3)):
        plan = step(i)
    barrier()
    launches0 = plan.kernel_launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.time()
    e0.record()
    for i in range(args.steps):
        step(i)
    e1.record()
    barrier()
    t_wall1 = time.time()
    ms = e0.elapsed_time(e1)
    launches = plan.kernel_launches - launches0
    if world > 1:
        tms = torch.tensor([ms], device=dev)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms = float(tms.item())
    value = world * B * args.steps / (ms / 1e3)
    step(0, mark=True)
    torch.cuda.synchronize()
    model_ms, decode_ms = ev_model[0].elapsed_time(ev_model[1]), ev_model[1].elapsed_time(ev_model[2])

    # end to end through the public API: pinned host images -> H2D -> network -> decode -> list of result dicts on the host
    n_e2e = max(min(args.steps, args.e2e_steps), 1)
    res = None
    barrier()
    t0 = time.perf_counter()
    for i in range(n_e2e):
        with torch.no_grad():
            img = host_imgs[i % 2].to(dev, non_blocking=True)
            res = head.get_poses(*net(img), metas)
    barrier()
    el = time.perf_counter() - t0
    if world > 1:
        tel = torch.tensor([el], device=dev)
        dist.all_reduce(tel, op=dist.ReduceOp.MAX)
        el = float(tel.item())
    n_people = sum(len(r["scores"]) for r in res)
    sampler.stop()
    clocks = sampler.summary(t_wall0, t_wall1)
    if rank == 0:
        emit({
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16 convolutions (fp32 DCNv2 / GroupNorm / predictor outputs), f32 decode" if dtype else "tf32 convolutions, f32 decode",
            "data": "synthetic",
            "config": {"workload": w["name"], "images_per_gpu": B, "image": f"{H}x{W}", "J": J, "K": K, "levels": 4,
                       "network": "MSPN2 x2 [3,4,6,3] + FPN(start_level=1, 4 outs) + DAS towers, random init, BatchNorm folded, channels-last",
                       "test_cfg": test_cfg, "l2": "two 327 MB image batches alternated; activations far exceed the 126 MB L2",
                       "parallelism": f"dp{world} batch-sharded, no collective" if world > 1 else "single GPU"},
            "clocks": clocks, "gpu_launches": int(launches) * world,
            "stages_ms": {"network_cudnn": model_ms, "decode_cuda": decode_ms},
            "roofline": None, "roofline_note": "the step is dominated by library (cuDNN) convolutions; the decode kernels' roofline is the default workload's",
            "e2e": {"value": world * B * n_e2e / el, "unit": UNIT, "h2d_bytes_per_step": B * 3 * H * W * 4,
                    "d2h_bytes_per_step": plan.d2h_bytes, "steps": n_e2e, "ms_per_step": el / n_e2e * 1e3,
                    "api": "DASNet(img) -> DASHeadB200.get_poses(*outs, img_metas): pinned host images in, result dicts out",
                    "people_last_step": n_people},
            "cpu_baseline": None,
        })
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None,
                    help="timed steps per repetition (default: 1000 decode steps; 20 for --impl reference; 10 for e2e_model); the timed "
                         "region repeats them until --min-seconds have passed")
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--input-sets", type=int, default=6)
    ap.add_argument("--streams", type=int, default=6,
                    help="independent batches in flight (<= input sets); round 2: 3/6: 1.10 M, 4/4: 1.15 M, 6/6: 1.19 M, 8/8: 1.18 M images/s")
    ap.add_argument("--profile-region", action="store_true", help="cudaProfilerStart/Stop around the timed region (ncu --profile-from-start off)")
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--e2e-depth", type=int, default=4,
                    help="host-entry calls in flight for the e2e figure (plans / streams driven in turn through das_plan_run_host_async); 1 = synchronous calls")
    ap.add_argument("--cpu-budget", type=float, default=12.0)
    ap.add_argument("--min-seconds", type=float, default=0.3,
                    help="the timed region (steps x repetitions) lasts at least this long; the median repetition is reported")
    ap.add_argument("--collect", default="nccl", choices=["p2p", "nccl", "none"],
                    help="N > 1 result collection: nccl (default) = one NCCL all-gather of the packed pose lists per round of input "
                         "sets on its own stream, plans writing straight into double-buffered staging (measured best: 58.6 us per "
                         "step at N = 8 against 55.6 at N = 1); p2p = the ranks' own publish kernel (NVLink stores into IPC-mapped "
                         "peer buffers: 63.9-64.8 us at N = 8); none = no collection (diagnosis only)")
    ap.add_argument("--no-extra", action="store_true", help="skip the short sub-runs of the other BASELINE configs")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--workload", default="panoptic", choices=sorted(WORKLOADS),
                    help="panoptic = BASELINE config #2 (the metric; default); single = #1 (B=1 latency); mupots = #3; crowded = #4; e2e_model = #5 (network + decode)")
    args = ap.parse_args()
    if args.gpus > 1 and "RANK" not in os.environ and args.impl != "reference":
        # `python bench.py --gpus N` without a launcher: start one rank per GPU ourselves (what the driver's torchrun does)
        import socket
        with socket.socket() as sk:
            sk.bind(("127.0.0.1", 0))
            port = sk.getsockname()[1]
        os.execv(sys.executable, [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
                                  "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.abspath(__file__)] + sys.argv[1:])
    if args.steps is None and args.workload != "e2e_model":
        args.steps = 20 if args.impl == "reference" else (1000 if args.workload in ("panoptic", "panoptic_spread", "single") else (300 if args.workload == "crowded" else 40))
    guard_stdout()
    WORKLOAD.clear()
    WORKLOAD.update(WORKLOADS[args.workload])
    TEST_CFG.update(nms_pre=WORKLOAD["K"], nms_post=WORKLOAD["K"])
    if args.workload == "e2e_model":
        if args.impl == "reference":      # mmcv / mmdet are absent: the reference network cannot run here
            emit({"impl": "reference", "unavailable": "the reference network needs mmcv-full/mmdet (not installed); only its decode is restated"})
            return
        if args.steps is None:
            args.steps = 10
        run_model(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
