"""Kernel-time breakdown of the config #5 network forward (torch.profiler; GPU box only)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from das_b200.model import DASNet

B = int(os.environ.get("B", 16))
dtype = None if os.environ.get("DAS_MODEL_DTYPE", "bf16") == "tf32" else torch.bfloat16
net = DASNet().cuda().prepare_inference(dtype)
img = torch.randn(B, 3, 1024, 1664, device="cuda")
with torch.no_grad():
    for _ in range(2):
        net(img)
    torch.cuda.synchronize()
    # per-section device times
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    with torch.autocast("cuda", dtype=dtype or torch.bfloat16, enabled=dtype is not None):
        x = img.contiguous(memory_format=torch.channels_last)
        ev[0].record(); f = net.backbone(x); ev[1].record(); p = net.neck(f); ev[2].record()
        o = [net.towers(t) for t in p]; ev[3].record()
    torch.cuda.synchronize()
    print("backbone %.1f ms  neck %.1f ms  towers %.1f ms" % (ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), ev[2].elapsed_time(ev[3])))
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        net(img)
        torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=70))
