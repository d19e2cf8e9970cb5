"""-m gpu tests of the tcgen05 (tensor-core) head path: the 3xTF32 GEMM core against an fp64 reference."""
import numpy as np
import pytest
import torch

from contrastive_lift_b200 import lib as L

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("bias", [False, True], ids=["nobias", "bias"])
@pytest.mark.parametrize("k,n", [(256, 256), (3, 256), (8, 32), (150, 128), (144, 27), (256, 21), (128, 3), (16, 16), (17, 33)])
def test_tc_gemm_matches_fp64(k, n, bias):
    lib = L.load()
    g = torch.Generator().manual_seed(k * 1000 + n)
    a = (torch.randn(128, k, generator=g) * 2.0).cuda()
    a[::7] = a[::7].abs()                                   # post-ReLU-like rows
    w = (torch.rand(n, k, generator=g) * 2 - 1).mul_(1.0 / np.sqrt(k)).cuda()
    b = (torch.randn(n, generator=g) * 0.5).cuda() if bias else None
    nf = lib.clift_tc_weight_floats(n, k, int(bias))
    assert nf > 0
    wtc = torch.zeros((nf,), device="cuda")
    L.check(lib.clift_pack_linear_tc(L.ptr(w), L.ptr(b), L.ptr(wtc), n, k, L.stream_ptr(w.device)))
    n_pad = (n + 31) // 32 * 32
    out = torch.full((128, n_pad), float("nan"), device="cuda")
    L.check(lib.clift_debug_tc_gemm(L.ptr(a), L.ptr(wtc), L.ptr(out), k, n, int(bias), L.stream_ptr(w.device)))
    torch.cuda.synchronize()
    ref = a.double() @ w.double().T + (b.double() if bias else 0.0)
    got = out[:, :n].double()
    scale = float(ref.abs().max())
    err = float((got - ref).abs().max()) / scale
    fp32 = float(((a @ w.T + (b if bias else 0.0)).double() - ref).abs().max()) / scale
    print(f"k={k} n={n}: 3xTF32 err {err:.2e}  (fp32 FFMA reference err {fp32:.2e})")
    assert err < 5e-6, err
    assert torch.all(out[:, n:] == 0)                       # padded columns come out as exact zeros
