"""Per-CTA timeline of wgrad_tc_kernel for one 2048-ray training render (debug aid): layer, ring stages, cycles."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import contrastive_lift_b200 as cl
from contrastive_lift_b200 import lib as L, synthetic as syn

grid = (128, 128, 128)
params = syn.make_field_params(0, grid, 21, 3)
model = cl.TensorVMSplit(list(grid), num_semantic_classes=21, dim_feature_instance=6, use_semantic_mlp=True,
                         use_instance_mlp=True, slow_fast_mode=True)
model.load_state_dict(params)
rend = cl.TensoRFRenderer(syn.default_aabb(), list(grid), semantic_weight_mode="softmax")
model, rend = model.cuda(), rend.cuda()
k, c2w = syn.camera(400, 400)
frame = cl.get_rays_checked(400, 400, k.numpy(), c2w.numpy())
rays = frame[torch.randperm(frame.shape[0], generator=torch.Generator().manual_seed(3))[:2048].cuda()].contiguous()
lib = L.load()
buf = torch.zeros((240, 4), dtype=torch.int64, device="cuda")
for it in range(2):
    if it == 1:
        lib.clift_debug_tc_trace(L.ptr(buf))
    out = rend(model, rays, 1.0, False, True)
    ((out[0] - 0.5) ** 2).mean().add(out[1].mean()).backward()
    torch.cuda.synchronize()
lib.clift_debug_tc_trace(None)
t = buf.cpu()
n_act = rend.last_stats("cuda:0")[0]
print(f"active records {n_act} = {(n_act + 127) // 128} tiles")
print("cta layer stages cyc_to_last_mma cyc_total cyc/stage")
for i in range(240):
    if int(t[i, 3]) == 0:
        continue
    st = max(1, int(t[i, 1]))
    print(i, int(t[i, 0]), int(t[i, 1]), int(t[i, 2]), int(t[i, 3]), int(t[i, 2]) // st)
