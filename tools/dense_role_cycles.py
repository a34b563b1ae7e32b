"""Per-role cycle counters of dense_project_tc_kernel (BASELINE config #3 shapes: B=64, J=17, 128x208): GPU box only.
python tools/dense_role_cycles.py"""
import sys, os, ctypes as C
_r = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, _r); sys.path.insert(0, os.path.join(_r, 'tests'))
import torch, numpy as np
from das_b200 import synth, _lib
import util
lib = _lib.load()
cfg = synth.MUPOTS17; tc = dict(nms_pre=20, nms_post=20, nms_thr=0.9, score_thr=0.0)
dev = torch.device('cuda')
B = int(os.environ.get('B', 64))
levels = synth.make_levels(cfg, B, 128, 208, seed=1234, device=dev); layers = synth.make_layers(cfg, seed=1235, device=dev); metas = synth.make_metas(B, 128, 208)
case = dict(cfg=cfg, levels=levels, layers=layers, metas=metas, batch=B)
plan = util.make_plan(case, tc)
lv = levels[0]; plan.bind([dict(cls=lv['cls'], ctr=lv['ctr'], pose=lv['pose_raw'], feats=lv['feats'], scales=lv['scales'])]); plan.set_metas(metas)
for _ in range(2): plan.run(use_graph=False)
torch.cuda.synchronize()
dbg = torch.zeros(148, 8, dtype=torch.int64, device=dev)
lib.das_dense_set_debug_buffer(C.c_void_p(dbg.data_ptr()))
plan.run(use_graph=False); torch.cuda.synchronize()
lib.das_dense_set_debug_buffer(None)
d = dbg.cpu().numpy().astype(np.float64) / 2.0          # two dense layers per decode accumulate into the same counters: per layer
names = ['mma: wait acc_free', 'mma: wait a_full', 'mma: issue+execute', 'producer0: wait TMA', 'producer0: wait TMEM slot', 'producer0: work', 'epilogue0: wait acc_full', 'cta total']
act = d[d[:, 2] > 0]
for i, n in enumerate(names): print(f'{n:30s} min {act[:, i].min():12.0f}  mean {act[:, i].mean():12.0f}  max {act[:, i].max():12.0f}')
# the Q CTAs of a tile sequence m = blockIdx // Q move in lock-step; sequences are independent: spread of their totals
Q = (cfg.num_joints + 2) // 3
tot = act[: (len(act) // Q) * Q, 7].reshape(-1, Q).mean(axis=1)
print('cta total per tile sequence (kcycles):', ' '.join(f'{t / 1e3:.0f}' for t in tot))
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
ev[0].record(); plan.run(use_graph=False); ev[1].record(); torch.cuda.synchronize()
print('whole decode, eager: %.3f ms' % ev[0].elapsed_time(ev[1]))
