"""GPU: the tcgen05 / TMEM building blocks (tc_common.cuh) against torch on the same device."""
import ctypes as C

import pytest
import torch

from das_b200 import _lib

pytestmark = pytest.mark.gpu


def tf32_trunc(x):
    return (x.view(torch.int32) & -8192).view(torch.float32)


@pytest.mark.parametrize("N,K,split", [(16, 32, 0), (16, 256, 0), (32, 256, 0), (16, 256, 1), (32, 256, 1), (32, 128, 1),
                                       (16, 64, 2), (32, 256, 2), (16, 256, 3), (32, 256, 3)])   # split & 2: A operand through TMEM
def test_tcgen05_tile_gemm(N, K, split):
    lib = _lib.load()
    g = torch.Generator(device="cuda").manual_seed(5)
    A = torch.randn(128, K, device="cuda", generator=g)
    B = torch.randn(N, K, device="cuda", generator=g) * 0.02
    D = torch.full((128, N), float("nan"), device="cuda")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(lib.das_tc_selftest(C.c_void_p(A.data_ptr()), C.c_void_p(B.data_ptr()), C.c_void_p(D.data_ptr()), N, K, split, st))
    torch.cuda.synchronize()
    exact = A.double() @ B.double().T
    scale = (A.double().abs() @ B.double().abs().T)
    if split & 1:
        # 3xTF32: only the lo*lo term (2^-22 relative) and fp32 accumulation are missing
        assert float(((D.double() - exact).abs() / scale).max()) < 2e-6
    else:
        ref = tf32_trunc(A).double() @ tf32_trunc(B).double().T
        assert float(((D.double() - ref).abs() / scale).max()) < 2e-6       # the tensor core truncates to tf32
        assert float(((D.double() - exact).abs() / scale).max()) < 2e-3
