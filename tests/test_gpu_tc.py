"""-m gpu tests of the tcgen05 (tensor-core) head paths: the 3xTF32 and the fp16-split GEMM cores against an fp64 reference."""
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


@pytest.mark.parametrize("bias", [False, True], ids=["nobias", "bias"])
@pytest.mark.parametrize("scale", [1.0, 1e-3, 300.0], ids=["unit", "small", "large"])
@pytest.mark.parametrize("k,n", [(256, 256), (3, 256), (8, 32), (150, 128), (144, 27), (256, 21), (128, 3), (16, 16), (17, 33)])
def test_tc16_gemm_matches_fp64(k, n, scale, bias):
    """fp16-split core: operands are scaled by the pack-time bound chain, so accuracy must not depend on magnitude."""
    lib = L.load()
    g = torch.Generator().manual_seed(k * 1000 + n)
    a = (torch.randn(128, k, generator=g) * 2.0 * scale).cuda()
    a[::7] = a[::7].abs()
    w = (torch.rand(n, k, generator=g) * 2 - 1).mul_(1.0 / np.sqrt(k)).cuda()
    b = (torch.randn(n, generator=g) * 0.5 * scale).cuda() if bias else None
    nb = lib.clift_tc16_weight_bytes(n, k, int(bias))
    assert nb > 0 and nb % 16 == 0
    wtc = torch.zeros((nb // 4,), device="cuda")
    bound = a.abs().max().reshape(1).contiguous()
    st = L.stream_ptr(w.device)
    L.check(lib.clift_pack_linear_tc16(L.ptr(w), L.ptr(b), L.ptr(wtc), n, k, L.ptr(bound), 0.0, st))
    n_pad = (n + 31) // 32 * 32
    out = torch.full((128, n_pad), float("nan"), device="cuda")
    L.check(lib.clift_debug_tc16_gemm(L.ptr(a), L.ptr(wtc), L.ptr(out), k, n, int(bias), st))
    torch.cuda.synchronize()
    hdr = wtc[:8].cpu()
    ref = a.double() @ w.double().T + (b.double() if bias else 0.0)
    assert float(ref.abs().max()) <= float(hdr[3]) * (1 + 1e-6), "the header's output bound must hold"
    assert float(hdr[0]) * float(bound) <= 2.0 ** 14 and float(hdr[2]) * float(hdr[0]) * float(hdr[1]) == 1.0
    got = out[:, :n].double()
    sc = float(ref.abs().max())
    err = float((got - ref).abs().max()) / sc
    fp32 = float(((a @ w.T + (b if bias else 0.0)).double() - ref).abs().max()) / sc
    print(f"k={k} n={n} scale={scale}: fp16-split err {err:.2e}  (fp32 FFMA reference err {fp32:.2e})")
    assert err < 5e-6, err
    assert torch.all(out[:, n:] == 0)


def test_tc16_bound_chain_covers_a_stack():
    """Chained headers: every layer's recorded bound dominates the activations an fp64 forward produces."""
    lib = L.load()
    g = torch.Generator().manual_seed(5)
    dims = [3, 256, 256, 256, 21]
    x = (torch.rand(512, 3, generator=g) * 2 - 1).double()
    st = L.stream_ptr(torch.device("cuda"))
    prev = None
    for i in range(len(dims) - 1):
        w = ((torch.rand(dims[i + 1], dims[i], generator=g) * 2 - 1) * (3.0 / np.sqrt(dims[i]))).cuda()
        b = ((torch.rand(dims[i + 1], generator=g) * 2 - 1) * 0.7).cuda()
        buf = torch.zeros((lib.clift_tc16_weight_bytes(dims[i + 1], dims[i], 1) // 4,), device="cuda")
        in_ptr = None if prev is None else prev.data_ptr() + 12
        L.check(lib.clift_pack_linear_tc16(L.ptr(w), L.ptr(b), L.ptr(buf), dims[i + 1], dims[i], in_ptr, 1.0 if prev is None else 0.0, st))
        torch.cuda.synchronize()
        x = x @ w.double().cpu().T + b.double().cpu()
        hdr = buf[:8].cpu()
        assert float(x.abs().max()) <= float(hdr[3]), (i, float(x.abs().max()), float(hdr[3]))
        if prev is not None:
            assert float(hdr[4]) == float(prev[3].cpu())
        x = x.clamp_min(0)
        prev = buf


def test_wgrad_tensor_core_matches_fma_kernel(monkeypatch):
    """wgrad_tc_kernel (tcgen05 3xTF32 over the training stashes) against the FP32-FMA wgrad kernel on the same forward /
    backward: every Linear weight gradient within 2e-5 scale-relative (both are fp32-faithful; they differ by summation order)."""
    import golden_util as gu
    import gpu_util as gpu
    for name in ("render_a", "render_d"):
        fx = gu.load(name)
        params, cfg, rays = gu.render_inputs(fx)
        sem_grid, ins_grid = gu.grid_comps(fx)
        grads = {}
        for mode in ("1", "0"):
            monkeypatch.setenv("CLIFT_WGRAD_FMA", mode)
            model, rend = gpu.build(params, cfg.grid_dim, int(fx["n_cls"]), int(fx["n_ins"]), bool(fx["slow_fast"]),
                                    bool(fx["softmax"]), cfg.aabb, float(fx["step_ratio"]), sem_grid=sem_grid, ins_grid=ins_grid)
            torch.manual_seed(7)
            out = rend(model, rays.cuda(), 1.0, False, True)
            gu.train_loss(out, fx, "trn").backward()
            grads[mode] = {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None}
        for k, g in grads["1"].items():
            if k.endswith(".weight"):
                err = gpu.rel_err(grads["0"][k], g.cpu())
                assert err < 2e-5, (name, k, err)


def test_training_forward_tensor_core_stash_matches_fma_forward(monkeypatch):
    """Training forwards run on the tcgen05 fp16-split kernel and record the backward's stash from there; with
    CLIFT_TRAIN_FWD_FMA=1 the FP32-FMA kernel does both.  Same outputs (1e-5) and same parameter gradients
    (gpu_util.grad_close: 2e-3 relative L2, single elements may carry a ReLU-mask flip)."""
    import golden_util as gu
    import gpu_util as gpu
    for name in ("render_a", "render_b", "render_c"):
        fx = gu.load(name)
        params, cfg, rays = gu.render_inputs(fx)
        res = {}
        for mode in ("1", "0"):
            monkeypatch.setenv("CLIFT_TRAIN_FWD_FMA", mode)
            model, rend = gpu.build(params, cfg.grid_dim, int(fx["n_cls"]), int(fx["n_ins"]), bool(fx["slow_fast"]),
                                    bool(fx["softmax"]), cfg.aabb, float(fx["step_ratio"]))
            torch.manual_seed(7)
            out = rend(model, rays.cuda(), 1.0, False, True)
            gu.train_loss(out, fx, "trn").backward()
            res[mode] = ([o.detach().clone() for o in out[:3]],
                         {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None})
        for a, b in zip(res["0"][0], res["1"][0]):
            assert gpu.rel_err(a, b.cpu()) < 1e-5, name
        assert set(res["0"][1]) == set(res["1"][1])
        for k, g in res["1"][1].items():
            assert gpu.grad_close(res["0"][1][k], g.cpu(), "tc16"), (name, k)
