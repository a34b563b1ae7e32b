import ctypes as C, torch, sys
sys.path.insert(0,'.')
from das_b200 import _lib
lib=_lib.load()
out=torch.zeros(2,dtype=torch.int64,device='cuda')
for N in (16,32,64,128,256):
    for it in (12,96,960):
        _lib.check(lib.das_tc_mma_bench(N,it,C.c_void_p(out.data_ptr()),None)); torch.cuda.synchronize()
        a,b=out.tolist(); print(f'N={N:3d} iters={it:4d} issue {a/it:7.1f} cyc/mma   issue+complete {b/it:7.1f} cyc/mma  total {b}')
