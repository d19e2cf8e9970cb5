"""clock64 timeline of CTA 0's first tiles in the pipelined xyz-stack kernel (heads_x16.cu): per accumulator unit, when the
issuer started it, saw its first weight stage, finished issuing; when the row threads saw it complete, had it in registers,
finished its epilogue."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import contrastive_lift_b200 as cl
from contrastive_lift_b200 import lib as L, synthetic as syn

grid = (128, 128, 128)
params = syn.make_field_params(0, grid, 21, 3)
aabb = syn.default_aabb()
model = cl.TensorVMSplit(list(grid), num_semantic_classes=21, dim_feature_instance=6, use_semantic_mlp=True,
                         use_instance_mlp=True, slow_fast_mode=True)
model.load_state_dict(params)
rend = cl.TensoRFRenderer(aabb, list(grid), semantic_weight_mode="softmax")
rend.update_step_ratio(syn.ratio_for_samples(aabb, grid, 512))
model, rend = model.cuda(), rend.cuda()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 800
k, c2w = syn.camera(n, n)
rays = cl.get_rays_checked(n, n, k.numpy(), c2w.numpy())
lib = L.load()
heads = L.HEAD_SEMANTIC | L.HEAD_INSTANCE
with torch.no_grad():
    rend._run(model, rays, None, False, heads, False)
    buf = torch.zeros((4, 32, 8), dtype=torch.int64, device="cuda")
    lib.clift_debug_tc_trace(L.ptr(buf))
    rend._run(model, rays, None, False, heads, False)
    torch.cuda.synchronize()
    lib.clift_debug_tc_trace(None)
t = buf.cpu()
names = ["S0a", "S0b", "S1a", "S1b", "S2a", "S2b", "S3a", "S3b", "S4", "F0a", "F0b", "F1a", "F1b", "F2a", "F2b", "F3", "W0a", "W0b", "W1a", "W1b", "W2a", "W2b", "W3"]
for tile in (1, 2):
    t0 = int(t[tile, 0, 0])
    print(f"tile {tile}: cycles relative to the issuer starting the tile's first unit")
    print(f"{'unit':5s} {'start':>8s} {'w_seen':>8s} {'issued':>8s} {'D_seen':>8s} {'D_regs':>8s} {'epi_end':>8s} | {'issue':>6s} {'mma_tail':>8s} {'ld':>5s} {'epi':>6s} {'k8wait':>7s}")
    for u, nme in enumerate(names):
        v = [int(t[tile, u, i]) - t0 if int(t[tile, u, i]) else 0 for i in range(8)]
        print(f"{nme:5s} {v[0]:8d} {v[1]:8d} {v[2]:8d} {v[3]:8d} {v[4]:8d} {v[5]:8d} | {v[2]-v[0]:6d} {v[3]-v[2]:8d} {v[4]-v[3]:5d} {v[5]-v[4] if v[5] else 0:6d} {v[7]-v[6] if v[6] else 0:7d}")
    print(f"next tile starts at {int(t[tile + 1, 0, 0]) - t0}")
