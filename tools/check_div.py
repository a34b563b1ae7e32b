"""Markstein division a/size = fma(fma(-q0,size,a), rsize, q0), q0 = a*rsize, rsize = RN(1/size), against IEEE float32 division
(refine_common.cuh: div_by).  Prints the number of mismatches (0)."""
import numpy as np
rng=np.random.default_rng(0)
f32=np.float32
def fma32(a,b,c):
    # a,b,c float32 arrays -> float32 of exact a*b+c (float64 product is exact; sum rounded to f64 then f32: double rounding is ~2^-29 rare)
    return (a.astype(np.float64)*b.astype(np.float64)+c.astype(np.float64)).astype(f32)
bad_total=0
sizes=list(range(1,600))+[640,768,832,1024,1333,1664,2048,4096]
for size in sizes:
    s=f32(size); rs=(f32(1.0)/s).astype(f32) if isinstance(s,np.ndarray) else f32(f32(1.0)/s)
    a=np.concatenate([rng.uniform(-3*size-64, 4*size+64, 200000), rng.normal(0,1,20000), np.arange(-64,size+64)+0.5, rng.uniform(-1e6,1e6,20000)]).astype(f32)
    q_true=(a/s).astype(f32)   # numpy float32 division is correctly rounded
    q0=(a*rs).astype(f32)
    r=fma32(-q0, np.full_like(a,s), a)
    q=fma32(r, np.full_like(a,rs), q0)
    bad=np.nonzero(q!=q_true)[0]
    bad_total+=len(bad)
    if len(bad): print(size, len(bad), a[bad[:3]], q[bad[:3]], q_true[bad[:3]])
print("total mismatches", bad_total)
