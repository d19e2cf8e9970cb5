"""Grid-mode heads (allgrid.yaml family: semantic and instance heads read their own 32-component VM factor set through a basis
Linear, tensoRF.py:72-85) on the fp16-split tensor-core kernel against the FP32-FMA kernel: head-stage ms of one 800x800 frame
at S = 512, and the largest difference between the two renders.  Prints one JSON line."""
import ctypes as C
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import contrastive_lift_b200 as cl
from contrastive_lift_b200 import lib as L, synthetic as syn
import bench


def main():
    dev = torch.device("cuda", 0)
    out = {}
    for label, sem_g, ins_g in (("allgrid (semantic + instance grid heads)", 32, 32), ("instGRIDsemMLP (instance grid head)", None, 32)):
        params = syn.make_field_params(0, bench.GRID, bench.N_CLS, bench.N_INS, sem_grid_comps=sem_g, ins_grid_comps=ins_g)
        aabb = syn.default_aabb()
        model = cl.TensorVMSplit(list(bench.GRID), num_semantics_comps=(sem_g or 32,) * 3, num_instance_comps=(ins_g or 32,) * 3,
                                 num_semantic_classes=bench.N_CLS, dim_feature_instance=2 * bench.N_INS,
                                 use_semantic_mlp=not sem_g, use_instance_mlp=not ins_g, slow_fast_mode=True)
        model.load_state_dict(params)
        rend = cl.TensoRFRenderer(aabb, list(bench.GRID), semantic_weight_mode="softmax")
        rend.update_step_ratio(syn.ratio_for_samples(aabb, bench.GRID, 512))
        model, rend = model.to(dev), rend.to(dev)
        rend.max_active_per_ray = bench.MAX_ACTIVE_PER_RAY
        rend.check_overflow = False
        k, c2w = syn.camera(800, 800)
        rays = cl.get_rays_checked(800, 800, k.numpy(), c2w.numpy(), device=dev)
        lib = L.load()
        L.check(lib.clift_profile_enable(1))
        res, maps = {}, {}
        for name, path in (("fma", L.HEADS_FMA), ("tensor16", L.HEADS_TENSOR16)):
            rend.head_path = path
            ts = []
            with torch.no_grad():
                for i in range(4):
                    o = rend._run(model, rays, None, False, L.HEAD_ALL, False)
                    buf = (C.c_float * 4)()
                    L.check(lib.clift_profile_stage_ms(buf))
                    if i > 0:
                        ts.append(float(buf[2]))
            assert lib.clift_debug_last_head_path() & 15 == path
            res[name + "_heads_ms"] = sum(ts) / len(ts)
            maps[name] = [t.clone() for t in o[:3]]
        rel = lambda a, b: float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))
        res["max_rel_diff_rgb_semprob_ins"] = [rel(maps["tensor16"][0], maps["fma"][0]),
                                               rel(maps["tensor16"][1].exp(), maps["fma"][1].exp()),
                                               rel(maps["tensor16"][2], maps["fma"][2])]
        res["speedup"] = res["fma_heads_ms"] / res["tensor16_heads_ms"]
        out[label] = res
    print(json.dumps({"workload": "800x800 frame, S=512, G=128^3, C=21, d=3+3, grid-mode heads with 32 components per mode", "results": out}))


main()
