"""Probe of the host-buffer entry (das_plan_run_host, zero-copy mode): per-call wall time, and -- under
`ncu --metrics pcie__read_bytes.sum,...` -- what each kernel pulls over PCIe.  GPU box only."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from das_b200 import synth
from das_b200.head import DecodePlan

B, H, W, K = 64, 128, 208, 10
head = synth.PANOPTIC
dev = torch.device("cuda", 0)
layers = synth.make_layers(head, seed=1235, device=dev)
metas = synth.make_metas(B, H, W, stride=8, seed=1236)
levels = synth.make_levels(head, B, H, W, seed=1234, device=dev, peaks=16)
plan = DecodePlan(num_joints=head.num_joints, root_idx=head.root_idx, depth_factor=head.depth_factor, z_norm=head.z_norm,
                  strides=head.strides, level_sizes=[(H, W)], batch=B, test_cfg=dict(nms_pre=K, nms_post=K, nms_thr=0.9, score_thr=0.0),
                  num_heads=4, feat_channels=256, num_layers=1, refine=True, device=dev)
plan.set_weights(layers)
lv0 = levels[0]
host_levels = [dict(cls=lv0["cls"].cpu().pin_memory(), ctr=lv0["ctr"].cpu().pin_memory(), pose=lv0["pose_raw"].cpu().pin_memory(),
                    feats=[f.permute(0, 2, 3, 1).cpu().pin_memory().permute(0, 3, 1, 2) for f in lv0["feats"]], scales=lv0["scales"])]
host_out = plan.alloc_host_out(pinned=True)
plan.set_host_mode(True)
for _ in range(3):
    plan.run_host(host_levels, metas, host_out)
n = int(os.environ.get("N", 10))
if os.environ.get("PROFILE"):
    torch.cuda.profiler.start()
t0 = time.perf_counter()
for _ in range(n):
    plan.run_host(host_levels, metas, host_out)
el = time.perf_counter() - t0
if os.environ.get("PROFILE"):
    torch.cuda.profiler.stop()
print("zero-copy run_host: %.3f ms per call" % (el / n * 1e3))
