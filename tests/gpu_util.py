"""Helpers for the -m gpu parity tests: build the product's model/renderer from a fixture's parameter dict."""
import torch

import contrastive_lift_b200 as cl


def build(params, grid, n_cls, n_ins, slow_fast, softmax, aabb, step_ratio, device="cuda", sem_grid=None, ins_grid=None):
    model = cl.TensorVMSplit(list(grid), num_semantics_comps=(sem_grid or 32,) * 3, num_instance_comps=(ins_grid or 32,) * 3,
                             num_semantic_classes=n_cls,
                             dim_feature_instance=2 * n_ins if slow_fast else n_ins,
                             output_mlp_semantics=torch.nn.Softmax(dim=-1) if softmax else torch.nn.Identity(),
                             use_semantic_mlp=not sem_grid, use_instance_mlp=not ins_grid, slow_fast_mode=slow_fast)
    missing = model.load_state_dict(params, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    rend = cl.TensoRFRenderer(aabb.clone(), list(grid), semantic_weight_mode="softmax" if softmax else "none",
                              stop_semantic_grad=True, step_ratio=0.5)
    rend.update_step_ratio(step_ratio)
    return model.to(device), rend.to(device)


def rel_err(got, ref):
    """max |got-ref| / max(|ref|) - the scale-relative error used for 'within 1e-4 relative'."""
    ref = ref.to(got.device)
    denom = ref.abs().max().clamp_min(1e-12)
    return float((got - ref).abs().max() / denom)


def flip_risk(params, cfg, rays, jitter, margin=2e-3):
    """Per-ray count of samples whose reference weight sits within `margin` (relative) of raymarch_weight_thres.
    Such a sample may land on the other side of `weight > thres` (renderer:103) under fp32 reassociation of the
    density sum; the head outputs of its ray then differ by up to thres * |head output| per flipped sample.  Tests
    compare those rays with that explicit allowance instead of the plain 1e-4 bound."""
    from oracle import clift_oracle as orc
    with torch.no_grad():
        w = orc._march(params, cfg, rays, jitter)[6]
    return ((w - cfg.weight_thres).abs() < margin * cfg.weight_thres).sum(-1)


def rel_err_rows(got, ref, rows):
    """rel_err restricted to the rays selected by the bool mask `rows` (scale = max |ref| over ALL rays)."""
    ref = ref.to(got.device)
    denom = ref.abs().max().clamp_min(1e-12)
    rows = rows.to(got.device)
    if not bool(rows.any()):
        return 0.0
    return float((got[rows] - ref[rows]).abs().max() / denom)


def l2_rel(got, ref):
    ref = ref.to(got.device)
    return float((got - ref).norm() / ref.norm().clamp_min(1e-20))


def grad_close(got, ref, fwd):
    """Parameter-gradient check.  fwd = "fma": training forward on the FP32-FMA kernel - 2e-3 on the worst element
    (scale-relative).  fwd = "tc16": training forward on the tcgen05 fp16-split kernel - its hidden pre-activations differ
    from the reference's by ~1e-6 relative, so a unit whose pre-activation sits within that distance of zero can take the
    other side of the ReLU: one (unit, record) term of the gradient flips, which moves single elements of a 164-ray test
    gradient by up to a percent without being an error of the arithmetic.  There the tensor as a whole must agree to 2e-3
    (relative L2) and no element may be off by more than 3e-2."""
    if fwd == "fma":
        return rel_err(got, ref) < 2e-3
    return l2_rel(got, ref) < 2e-3 and rel_err(got, ref) < 3e-2
