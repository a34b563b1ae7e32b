import sys, os, ctypes as C; import os as _os; _r=_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))); sys.path.insert(0,_r); sys.path.insert(0,_os.path.join(_r,'tests'))
import torch, numpy as np
from das_b200 import synth, _lib
import util
lib=_lib.load()
P=synth.PANOPTIC; tc=dict(nms_pre=10,nms_post=10,nms_thr=0.9,score_thr=0.0)
dev=torch.device('cuda')
levels=synth.make_levels(P,64,128,208,seed=1234,device=dev); layers=synth.make_layers(P,seed=1235,device=dev); metas=synth.make_metas(64,128,208)
case=dict(cfg=P,levels=levels,layers=layers,metas=metas,batch=64)
mode=int(os.environ.get('DAS_REFINE_MODE','1'))
plan=util.make_plan(case,tc,refine_mode=mode)
lv=levels[0]; plan.bind([dict(cls=lv['cls'],ctr=lv['ctr'],pose=lv['pose_raw'],feats=lv['feats'],scales=lv['scales'])]); plan.set_metas(metas)
dbg=torch.zeros(148,16,dtype=torch.int64,device=dev)
lib.das_tc_set_debug_buffer(C.c_void_p(dbg.data_ptr()))
for _ in range(3): plan.run(use_graph=False)
torch.cuda.synchronize(); dbg.zero_(); plan.run(use_graph=False); torch.cuda.synchronize()
d=dbg.cpu().numpy().astype(np.float64)
names=['mma:wait_afull','mma:issue','mma:wait_accfree','pg0:cpwait+bar','pg0:lds+bar','pg0:gather_issue','pg0:wait_aempty','pg0:tmem_st+arrive','pg0:n_kb','epi0:wait_accfull','epi0:rest','-','-','cta:total','mma:bpanel']
for i,n in enumerate(names): print(f'{n:22s} mean {d[:,i].mean():10.0f}  max {d[:,i].max():10.0f}')
