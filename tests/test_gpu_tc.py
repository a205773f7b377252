"""tcgen05 building blocks: split-bf16 GEMM accumulated in TMEM vs a float64 matmul."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("splits,tol", [(1, 2e-2), (3, 1e-4)])
def test_tc_gemm_selftest(splits, tol):
    import msacl_b200
    from msacl_b200 import _lib
    lib = msacl_b200.load_library()
    g = torch.Generator().manual_seed(0)
    A = torch.randn(128, 256, generator=g).cuda()
    W = (torch.randn(256, 256, generator=g) / 16).cuda()
    D = torch.zeros(128, 256, device="cuda")
    _lib.check(lib.msacl_selftest_tc_gemm(A.data_ptr(), W.data_ptr(), D.data_ptr(), splits, _lib.current_stream()))
    torch.cuda.synchronize()
    want = (A.double() @ W.double().t()).cpu().numpy()
    got = D.cpu().numpy()
    err = np.abs(got - want).max() / np.abs(want).max()
    assert err < tol, err
