#!/usr/bin/env python
"""Design study for the next kernel (tcgen05 data-gradient chain), CPU only.
TEST / DESIGN INFRASTRUCTURE ONLY (lives under oracle/ because it runs the oracle; nothing in the product imports it).


Question: is the forward's operand format - power-of-two scale, fp16 (hi, lo) split, three products accumulated in fp32 -
accurate enough for the BACKWARD data-gradient GEMMs dA = dZ . W, whose operand dZ has a runtime, wide dynamic range, when the
scale comes from a cheap a-priori BOUND chain (absmax of the incoming ray gradients x column-norm bounds of the weights)
instead of the true per-layer maximum?

Method: one training main pass of the oracle (BASELINE config 3 shape, reduced ray count) with hooks on every hidden
pre-activation of the semantic / instance / rgb stacks, so the true dZ_l and the true dA_{l-1} are known.  For every layer the
data-gradient GEMM is re-evaluated with
    s_z = 2^floor(log2(2^14 / bound_l))          bound_l from the chain, >= true max |dZ_l|
    Z_hi = fp16(s_z dZ), Z_lo = fp16(s_z dZ - Z_hi);  W_hi, W_lo likewise at the forward's weight scale
    dA ~= (Z_hi W_hi + Z_hi W_lo + Z_lo W_hi) / (s_z s_w)      fp32 accumulation
and compared with the fp64 product.  Reported per layer: how loose the bound is (bound / true max), the scale-relative max
error and the relative L2 error, next to the same figures for a single fp16 product (Z_hi W_hi) and for the exact per-layer
maximum as scale.  Writes profiles/r01_dgrad_fp16_split_study.md.
"""
import math
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from contrastive_lift_b200 import synthetic as syn          # noqa: E402
from oracle import clift_oracle as orc                      # noqa: E402

GRID, N_CLS, N_INS = (64, 64, 64), 21, 3


def split16(x, scale):
    y = (x.double() * scale).float()
    hi = y.half()
    lo = (y - hi.float()).half()
    return hi.float(), lo.float()


def pow2_scale(bound, target_log2=14):
    return 2.0 ** math.floor(target_log2 - math.log2(max(bound, 1e-38)))


def three_product(z, w, s_z, s_w):
    """z [n, out], w [out, in] -> z @ w through the fp16 split; fp32 accumulation like the tensor core's."""
    zh, zl = split16(z, s_z)
    wh, wl = split16(w, s_w)
    acc = zh @ wh + zh @ wl + zl @ wh
    one = zh @ wh
    return acc.double() / (s_z * s_w), one.double() / (s_z * s_w)


def main():
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count() or 1)
    params = syn.make_field_params(0, GRID, N_CLS, N_INS)
    aabb = syn.default_aabb()
    cfg = orc.RenderConfig(aabb=aabb, grid_dim=GRID, step_ratio=0.5).refresh()
    k, c2w = syn.camera(48, 48)
    rays = orc.make_rays(48, 48, k, c2w)
    p = {kk: v.clone().requires_grad_(True) for kk, v in params.items()}

    records = []            # (stack name, layer index, weight, pre-activation tensor)
    real_run_mlp = orc.run_mlp

    def run_mlp_recording(x, layers):
        name = {4: "instance (3 hidden)", 5: "semantic (4 hidden)", 3: "rgb (2 hidden)"}.get(len(layers), f"{len(layers)} layers")
        for i, (w, b) in enumerate(layers):
            x = F.linear(x, w, b)
            x.retain_grad()
            records.append((name, i, w, x))
            if i + 1 < len(layers):
                x = torch.relu(x)          # out of place: the recorded pre-activation must survive
        return x

    orc.run_mlp = run_mlp_recording
    try:
        jitter = torch.rand((rays.shape[0], 1))
        out = orc.render_forward(p, cfg, rays, jitter, True)
    finally:
        orc.run_mlp = real_run_mlp
    # BASELINE config 3's main-pass loss shape (trainer:148-199): MSE + dist-reg + CE on the semantic log-probabilities,
    # plus a term on the embeddings so the instance stacks receive a gradient too (their own pass has the same structure)
    tgt = torch.full_like(out[0], 0.5)
    labels = torch.randint(0, N_CLS, (rays.shape[0],))
    loss = ((out[0] - tgt) ** 2).mean() + 0.01 * out[5] + F.nll_loss(out[1], labels) + 0.1 * out[2].pow(2).mean()
    loss.backward()

    # incoming ray gradients: their absmax starts the bound chain (one small reduction kernel in the plan)
    lines = ["# Round 1 - design study: fp16-split operands for the data-gradient GEMMs (CPU emulation, `oracle/dgrad_fp16_split_study.py`)\n",
             f"One oracle training pass, {rays.shape[0]} rays, S = {cfg.n_samples}, G = 64^3, C = {N_CLS}: true dZ of every Linear captured by "
             "autograd hooks; `dA = dZ . W` re-evaluated with the forward kernel's operand format and a scale from an a-priori bound "
             "chain. Errors are against the fp64 product, `max` = max |err| / max |dA| (the scale-relative measure of the parity tests), "
             "`L2` = relative L2.\n",
             "| stack | layer (out -> in) | records | max abs dZ | bound / max | 3-product max | 3-product L2 | exact-max scale: max | single fp16 product: max |",
             "|---|---|---|---|---|---|---|---|---|"]
    worst = 0.0
    # walk every recorded stack instance from its last layer to its first, carrying the bound
    stacks, cur = [], []
    for rec in records:
        if rec[1] == 0 and cur:
            stacks.append(cur)
            cur = []
        cur.append(rec)
    if cur:
        stacks.append(cur)
    for stack in stacks:
        name = stack[0][0]
        last = stack[-1]
        g_last = last[3].grad
        if g_last is None:
            continue
        bound = float(g_last.abs().max())          # the output-layer dZ comes straight from the ray gradients: its absmax is measured
        for (nm, i, w, x) in reversed(stack):
            dz = x.grad.detach()
            wd = w.detach()
            true_max = float(dz.abs().max())
            if i == 0:
                break                               # the first layer's data gradient (w.r.t. xyz / the rgb input) is handled by the PE / gather backward
            exact = dz.double() @ wd.double()
            s_w = pow2_scale(float(wd.abs().max()))
            s_chain = pow2_scale(max(bound, 1e-30))
            s_exact = pow2_scale(true_max)
            got, one = three_product(dz, wd, s_chain, s_w)
            got_exact, _ = three_product(dz, wd, s_exact, s_w)
            ref_max = float(exact.abs().max())
            e3 = float((got - exact).abs().max()) / ref_max
            l2 = float((got - exact).norm() / exact.norm())
            e3x = float((got_exact - exact).abs().max()) / ref_max
            e1 = float((one - exact).abs().max()) / ref_max
            worst = max(worst, e3)
            lines.append(f"| {name} | {i} ({wd.shape[0]} -> {wd.shape[1]}) | {dz.shape[0]} | {true_max:.3e} | {bound / true_max:.1f} | "
                         f"{e3:.2e} | {l2:.2e} | {e3x:.2e} | {e1:.2e} |")
            # chain: |dZ_{l-1}| <= |dA_{l-1}| <= max_j sum_o |dZ_l[o]| |W[o][j]| <= bound_l * max column L1 norm of W_l
            bound = bound * float(wd.abs().sum(0).max())
    lines.append("")
    lines.append(f"Worst 3-product error with the chained bound: {worst:.2e} (gradient parity tolerance: 2e-3). The chained bound is up to "
                 "three orders of magnitude loose after four layers and costs nothing: fp16 keeps 11 significant bits of `hi` down to "
                 "2^-14 and the `lo` remainders degrade gracefully through the subnormals, so the 3-product result sits at the fp32 "
                 "accumulation floor (~2^-21) either way - no per-layer max reduction is needed between the GEMMs of the chain. A single "
                 "fp16 product is 3-5e-4 per GEMM: inside the tolerance for one layer, but the error compounds over a 4-layer chain and "
                 "into the weight gradients, so the plan keeps three products (dropping only one cross term does not help: either "
                 "remaining 2^-12 operand error gives the same 3e-4).")
    out_path = os.path.join(ROOT, "profiles", "r01_dgrad_fp16_split_study.md")
    open(out_path, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
