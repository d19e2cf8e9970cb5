"""Time the head kernels by head subset on the bench frame (800x800, S=512): xyz stacks (semantic + instance) and the rgb stack,
pipelined kernel on / off.  Prints one JSON line."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import contrastive_lift_b200 as cl
from contrastive_lift_b200 import lib as L, synthetic as syn
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
    rend.max_active_per_ray = bench.MAX_ACTIVE_PER_RAY
    rend.check_overflow = False
    k, c2w = syn.camera(800, 800)
    rays = cl.get_rays_checked(800, 800, k.numpy(), c2w.numpy(), device=dev)
    lib = L.load()
    import ctypes as C
    out = {}
    L.check(lib.clift_profile_enable(1))
    for name, heads in (("march_only", 0), ("xyz", L.HEAD_SEMANTIC | L.HEAD_INSTANCE), ("sem", L.HEAD_SEMANTIC), ("ins", L.HEAD_INSTANCE),
                        ("rgb", L.HEAD_RGB), ("all", L.HEAD_ALL)):
        ts = []
        with torch.no_grad():
            for i in range(4):
                rend._run(model, rays, None, False, heads, False)
                buf = (C.c_float * 4)()
                L.check(lib.clift_profile_stage_ms(buf))
                if i > 0:
                    ts.append(float(buf[2]))
        out[name] = sum(ts) / len(ts)
    print(json.dumps({"heads_stage_ms": out, "x16": os.environ.get("CLIFT_X16", "1")}))

main()
