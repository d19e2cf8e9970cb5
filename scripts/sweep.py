"""BASELINE config 5: render-throughput sweep (frame 256^2..1600^2 x 128..1024 samples/ray) on one GPU.

Prints a markdown table: Mrays/s (device-resident, CUDA events, 3 timed frames after 2 warm-ups), the stage split and the
algorithmic roofline figures (march gather GB/s, head TFLOP/s).  Run per GPU count with bench.py for the scaling numbers.
"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import contrastive_lift_b200 as cl  # noqa: E402
from contrastive_lift_b200 import lib as L, synthetic as syn  # noqa: E402
import bench  # noqa: E402

GRID = (128, 128, 128)


def main():
    dev = torch.device("cuda", 0)
    lib = L.load()
    params = syn.make_field_params(0, GRID, 21, 3)
    aabb = syn.default_aabb()
    model = cl.TensorVMSplit(list(GRID), num_semantic_classes=21, dim_feature_instance=6, use_semantic_mlp=True,
                             use_instance_mlp=True, slow_fast_mode=True)
    model.load_state_dict(params)
    model = model.to(dev)
    flops_per = bench.head_flops_per_sample(params)
    print("| frame | samples/ray | Mrays/s | ms/frame | march ms | heads ms | active/ray | in-box/ray | march GB/s (algorithmic) | heads TFLOP/s (algorithmic) |")
    print("|---|---|---|---|---|---|---|---|---|---|")
    for frame in (256, 400, 800, 1600):
        for S in (128, 256, 512, 1024):
            rend = cl.TensoRFRenderer(aabb, list(GRID), semantic_weight_mode="softmax").to(dev)
            rend.update_step_ratio(syn.ratio_for_samples(aabb, GRID, S))
            assert rend.n_samples == S
            rend.max_active_per_ray = 192
            k, c2w = syn.camera(frame, frame)
            rays = cl.get_rays_checked(frame, frame, k.numpy(), c2w.numpy(), device=dev)
            with torch.no_grad():
                for _ in range(2):
                    rend(model, rays, 1.0, False, False)
                rend.check_overflow = False
                torch.cuda.synchronize()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(3):
                    rend(model, rays, 1.0, False, False)
                b.record()
                torch.cuda.synchronize()
                ms = a.elapsed_time(b) / 3
                stage = [0.0] * 4
                n_act = n_in = 0
                if rays.shape[0] * S < 2 ** 31:          # stage events describe one C call
                    L.check(lib.clift_profile_enable(1))
                    rend(model, rays, 1.0, False, False)
                    buf = (C.c_float * 4)()
                    L.check(lib.clift_profile_stage_ms(buf))
                    L.check(lib.clift_profile_enable(0))
                    stage = [float(v) for v in buf]
                    n_act, n_in, _, _ = rend.last_stats(dev)
            n = rays.shape[0]
            gbs = (32.0 * n + 1152.0 * n_in + 4.0 * n * S) / (stage[0] * 1e-3) / 1e9 if stage[0] else float("nan")
            tfl = flops_per * n_act / (stage[2] * 1e-3) / 1e12 if stage[2] else float("nan")
            print(f"| {frame}x{frame} | {S} | {n / ms / 1e3:.3f} | {ms:.1f} | {stage[0]:.2f} | {stage[2]:.1f} | {n_act / n:.1f} | {n_in / n:.1f} | {gbs:.0f} | {tfl:.1f} |", flush=True)
            del rays
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
