"""GPU-box probe: which warm-up puts the reference CPU leg into its fast regime (0.22 s vs 0.41 s per 2048-ray chunk)?"""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
render, rays, kind = bench.cpu_renderer(800, 512)
render(rays[::312][:256])
def t(sub):
    t0 = time.perf_counter(); render(sub); return round(time.perf_counter() - t0, 3)
mode = sys.argv[1]
before = [t(rays[i::312][:2048].contiguous()) for i in range(2)]
if mode == "bigalloc":
    xs = [torch.zeros(64 << 20) for _ in range(8)]; del xs
elif mode == "grad":
    w = torch.randn(512, 512, requires_grad=True)
    (torch.randn(100000, 512) @ w).relu().sum().backward()
elif mode == "gradsmall":
    w = torch.randn(8, 8, requires_grad=True)
    (torch.randn(10, 8) @ w).sum().backward()
elif mode == "threads":
    torch.set_num_threads(8); torch.set_num_threads(16)
elif mode == "oracle_fwd_grad":
    from contrastive_lift_b200 import synthetic as syn
    from oracle import clift_oracle as orc
    params = syn.make_field_params(0, bench.GRID, 21, 3)
    cfg = orc.RenderConfig(aabb=syn.default_aabb(), grid_dim=bench.GRID).refresh()
    p = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    out = orc.render_forward(p, cfg, rays[::312][:2048].contiguous(), torch.rand(2048, 1), False)
    if "bwd" in sys.argv:
        out[0].sum().backward()
print(mode, sys.argv[2:], "before", before, "after", [t(rays[i::312][:2048].contiguous()) for i in range(3)], flush=True)
