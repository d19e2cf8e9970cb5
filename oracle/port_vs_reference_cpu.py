#!/usr/bin/env python
"""Build-container check (needs /root/reference): is the oracle port a FAIR stand-in for the reference on the CPU?
TEST INFRASTRUCTURE ONLY (lives under oracle/ because it runs the oracle and the imported reference; nothing in the product imports it).


Times the unmodified reference renderer (imported through oracle/refload.py) and the oracle port
(oracle/clift_oracle.py, what `bench.py --impl reference` and `cpu_baseline` run on the GPU box) on the same bench
workload: chunks of 2048 rays strided from the synthetic 800x800 frame, S=512, G=128^3, all heads, no_grad,
render_panopli.py:114-119's chunk loop.  Prints one JSON line; committed as profiles/r01_port_vs_reference_cpu.json.
"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from contrastive_lift_b200 import synthetic as syn          # noqa: E402
from oracle import clift_oracle as orc                      # noqa: E402
from oracle import refload                                   # noqa: E402

GRID, N_CLS, N_INS, S, FRAME, CHUNK = (128, 128, 128), 21, 3, 512, 800, 2048


def main():
    n_chunks = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    torch.set_num_threads(os.cpu_count() or 1)
    params = syn.make_field_params(0, GRID, N_CLS, N_INS)
    aabb = syn.default_aabb()
    ratio = orc.ratio_for_samples(aabb, GRID, S)
    cfg = orc.RenderConfig(aabb=aabb, grid_dim=GRID, step_ratio=ratio).refresh()
    k, c2w = syn.camera(FRAME, FRAME)
    rays = orc.make_rays(FRAME, FRAME, k, c2w)
    stride = rays.shape[0] // (n_chunks * CHUNK)
    sub = rays[::stride][:n_chunks * CHUNK].contiguous()

    model = refload.build_model(params, GRID, N_CLS, N_INS, slow_fast=True, semantic_softmax=True)
    rend = refload.build_renderer(aabb, GRID, semantic_softmax=True)
    rend.update_step_ratio(ratio)
    assert rend.n_samples == S == cfg.n_samples

    def run_reference():
        outs = []
        with torch.no_grad():
            for i in range(0, sub.shape[0], CHUNK):
                outs.append(rend(model, sub[i:i + CHUNK], 1.0, False, False)[:4])
        return [torch.cat([o[j] for o in outs]) for j in range(4)]

    def run_port():
        return orc.render_chunked(params, cfg, sub, chunk=CHUNK)

    res = {}
    outs = {}
    for name, fn in (("reference", run_reference), ("port", run_port), ("reference_again", run_reference), ("port_again", run_port)):
        fn_out = None
        if name in ("reference", "port"):
            with torch.no_grad():
                (rend(model, sub[:256], 1.0, False, False) if name == "reference" else orc.render_chunked(params, cfg, sub[:256], chunk=CHUNK))
        t0 = time.perf_counter()
        fn_out = fn()
        dt = time.perf_counter() - t0
        res[name] = dt
        outs[name] = fn_out
    same = all(torch.equal(a, b) for a, b in zip(outs["reference"][:4], outs["port"][:4]))
    t_ref = min(res["reference"], res["reference_again"])
    t_port = min(res["port"], res["port_again"])
    print(json.dumps({"workload": f"{sub.shape[0]} rays strided from the synthetic {FRAME}x{FRAME} frame, S={S}, G=128^3, all heads, chunk={CHUNK}, no_grad",
                      "cores": os.cpu_count(), "reference_s": t_ref, "port_s": t_port,
                      "reference_Mrays_s": sub.shape[0] / t_ref / 1e6, "port_Mrays_s": sub.shape[0] / t_port / 1e6,
                      "port_over_reference_time": t_port / t_ref, "outputs_bit_equal": bool(same)}))


if __name__ == "__main__":
    main()
