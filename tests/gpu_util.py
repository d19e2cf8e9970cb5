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
