"""The bench frame (800x800, S = 512, all heads) rendered through the reference's own chunk loop (render_panopli.py:108-121:
`for i in range(0, rays.shape[0], chunk)`, chunk = 2048 in the shipped configs) instead of one call: what an unmodified
inference script gets from the drop-in renderer.  Prints ms per frame for several chunk sizes (wall clock around the loop +
final synchronize, best of 3)."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import contrastive_lift_b200 as cl
from contrastive_lift_b200 import synthetic as syn
import bench


def main():
    dev = torch.device("cuda", 0)
    params = syn.make_field_params(0, bench.GRID, bench.N_CLS, bench.N_INS)
    aabb = syn.default_aabb()
    model = cl.TensorVMSplit(list(bench.GRID), num_semantic_classes=bench.N_CLS, dim_feature_instance=2 * bench.N_INS,
                             use_semantic_mlp=True, use_instance_mlp=True, slow_fast_mode=True)
    model.load_state_dict(params)
    rend = cl.TensoRFRenderer(aabb, list(bench.GRID), semantic_weight_mode="softmax")
    rend.update_step_ratio(syn.ratio_for_samples(aabb, bench.GRID, 512))
    model, rend = model.to(dev), rend.to(dev)
    k, c2w = syn.camera(800, 800)
    rays = cl.get_rays_checked(800, 800, k.numpy(), c2w.numpy(), device=dev)
    out = {}
    for chunk in (2048, 8192, 65536, rays.shape[0]):
        best = None
        for rep in range(4):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            with torch.no_grad():
                parts = [rend(model, rays[i:i + chunk], 1.0, False, False) for i in range(0, rays.shape[0], chunk)]
                rgb = torch.cat([p[0] for p in parts])
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) * 1e3
            if rep > 0:
                best = dt if best is None else min(best, dt)
        out[str(chunk)] = {"ms_per_frame": best, "calls": (rays.shape[0] + chunk - 1) // chunk,
                           "Mrays_per_s": rays.shape[0] / best / 1e3}
    print(json.dumps({"workload": "800x800 x 512, all heads, renderer called per chunk of rays, default settings (calls whose "
                                  "worst case fits 1 GiB of records get room for every sample: no capacity check, no host sync)", "by_chunk": out}))


main()
