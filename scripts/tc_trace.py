"""Print the clock64 timeline of CTA 0's first tiles in the tensor-core head kernel (debug aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import contrastive_lift_b200 as cl
from contrastive_lift_b200 import lib as L, synthetic as syn

grid = (128, 128, 128)
params = syn.make_field_params(0, grid, 21, 3)
aabb = syn.default_aabb()
ratio = syn.ratio_for_samples(aabb, grid, 512)
model = cl.TensorVMSplit(list(grid), num_semantic_classes=21, dim_feature_instance=6, use_semantic_mlp=True,
                         use_instance_mlp=True, slow_fast_mode=True)
model.load_state_dict(params)
rend = cl.TensoRFRenderer(aabb, list(grid), semantic_weight_mode="softmax")
rend.update_step_ratio(ratio)
model, rend = model.cuda(), rend.cuda()
k, c2w = syn.camera(200, 200)
rays = cl.get_rays_checked(200, 200, k.numpy(), c2w.numpy())
lib = L.load()
with torch.no_grad():
    rend(model, rays, 1.0, False, False)
    buf = torch.zeros((4, 24, 10), dtype=torch.int64, device="cuda")
    lib.clift_debug_tc_trace(L.ptr(buf))
    rend(model, rays, 1.0, False, False)
    torch.cuda.synchronize()
    lib.clift_debug_tc_trace(None)
t = buf.cpu()
names = ["sem0", "sem1", "sem2", "sem3", "sem4", "insF0", "insF1", "insF2", "insF3", "insS0", "insS1", "insS2", "insS3", "basis", "rgb0", "rgb1", "rgb2"]
for tile in (1, 2):
    t0 = int(t[tile, 0, 1])
    print(f"tile {tile}: (cycles relative to first operand build done)")
    print(f"{'gemm':6s} {'A_written':>10s} {'published':>10s} {'mma_saw_A':>10s} {'mma_issued':>10s} {'D_seen':>10s} | {'build/epi':>9s} {'publish':>8s} {'handoff':>8s} {'issue':>7s} {'mma+wake':>9s} {'w_wait':>7s}")
    prev_d = None
    for g, n in enumerate(names):
        a_w, pub, saw, iss, d = (int(t[tile, g, i]) - t0 for i in (1, 1, 3, 4, 0))   # streamed operands: written == published
        epi = a_w - prev_d if prev_d is not None else 0
        w_wait = int(t[tile, g, 5]) - t0 - saw
        print(f"{n:6s} {a_w:10d} {pub:10d} {saw:10d} {iss:10d} {d:10d} | {epi:9d} {pub - a_w:8d} {saw - pub:8d} {iss - saw:7d} {d - iss:9d} {w_wait:7d}")
        prev_d = d
        if int(t[tile, g, 6]) or int(t[tile, g, 7]):
            print(f"       marks: {int(t[tile, g, 6]) - t0:10d} {int(t[tile, g, 7]) - t0:10d}")
    nxt = int(t[tile + 1, 0, 1]) - t0
    print(f"next tile's first operand written at {nxt}")
