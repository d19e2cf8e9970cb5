"""CPU-side checks of bench.py's contract: the reference arm (the reference algorithm on the host cores) prints one JSON
line with the agreed keys; the GPU arm refuses to run without a CUDA device instead of falling back."""
import json
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--frame", "64", "--samples", "64"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "Mrays/s" and d["higher_is_better"] is True and d["value"] > 0
    from oracle import refload
    assert d["cpu_baseline"]["kind"] == ("reference" if refload.available() else "port")
    assert d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["config"]["rays_per_step_on_host"] == 2048 and d["config"]["rays_per_step_per_gpu"] == 0
    assert d["e2e"] == {"value": d["value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0 and d["vs_baseline"] is None and "workload" in d["config"]


def test_reference_arm_non_zero_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_gpu_arm_has_no_cpu_fallback():
    if torch.cuda.is_available():
        return      # on a GPU box the arm runs; the -m gpu suite and the driver cover it
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1"], capture_output=True, text=True,
                         timeout=300, cwd=ROOT)
    assert out.returncode != 0 and "no CUDA device" in (out.stderr + out.stdout)


def test_roofline_flop_basis_matches_the_survey_figure():
    """roofline.achieved = algorithmic FLOPs / kernel time: the per-active-sample figure is SURVEY 8(d)'s
    2 x (3888 + 35968 + 202752 + 265216) MACs = 1.016 MFLOP at C = 21, d = 3 (basis + rgb + semantic + fast/slow instance)."""
    import bench
    from contrastive_lift_b200 import synthetic as syn
    params = syn.make_field_params(0, (8, 8, 8), bench.N_CLS, bench.N_INS)
    assert bench.head_flops_per_sample(params) == 2 * (3888 + 35968 + 202752 + 265216)


def test_cpu_baseline_leg_keeps_what_the_psnr_match_needs():
    import bench
    rate, n, dt, kind, rays, rgb, sem, ins, depth = bench.cpu_reference_rate(16, 32, 5.0, 256, keep=True)
    assert rate > 0 and n == 256 == rays.shape[0] and rgb.shape == (256, 3) and depth.shape == (256,)
    assert sem.shape == (256, bench.N_CLS) and ins.shape == (256, 2 * bench.N_INS) and kind in ("reference", "port")

    class SameMaps:                                   # stands in for the CUDA renderer: returns the reference's own maps
        def __call__(self, model, r, *a):
            return rgb + 1e-6, sem, ins, depth
    out = bench.psnr_match(SameMaps(), None, rays, (rgb, sem, ins, depth), "cpu")
    assert 110.0 < out["psnr_db"] <= 200.0 and out["rays"] == 256 and out["depth_max_rel_err"] == 0.0
    assert out["semantic_prob_max_abs_err"] == 0.0 and out["instance_max_rel_err"] == 0.0 and out["ok"]
    same = type("Z", (), {"__call__": lambda s, m, r, *a: (rgb, sem, ins, depth)})()
    assert bench.psnr_match(same, None, rays, (rgb, sem, ins, depth), "cpu")["psnr_db"] == 200.0   # capped, never inf (JSON)


def test_reference_and_port_cpu_legs_agree_bit_for_bit(monkeypatch):
    """The staged / live reference modules and the oracle port render the bench scene's rays identically (when the reference
    is reachable), so either is a valid cpu_baseline and psnr_match checker."""
    import bench
    from oracle import refload
    if not refload.available():
        return
    render_ref, rays_ref, kind = bench.cpu_renderer(24, 48)
    assert kind == "reference"
    monkeypatch.setenv("CLIFT_BENCH_CPU_PORT", "1")
    render_port, rays_port, kind2 = bench.cpu_renderer(24, 48)
    assert kind2 == "port" and torch.equal(rays_ref, rays_port)
    for a, b in zip(render_ref(rays_ref[:300]), render_port(rays_port[:300])):
        assert torch.equal(a, b)


def test_oracle_is_only_reachable_from_the_allowed_places():
    """oracle/ is test infrastructure: the product package never imports it; outside tests/ and oracle/ itself only
    __graft_entry__.smoke()/build(), bench.py's CPU legs (cpu_baseline, --impl reference) and the CPU leg of the
    training-step bench (its cpu_baseline analogue) do."""
    import re
    pat = re.compile(r"^\s*(from\s+oracle\b|import\s+oracle\b)", re.M)
    users = {}
    for base, _, files in os.walk(ROOT):
        rel = os.path.relpath(base, ROOT)
        if rel.split(os.sep)[0] in ("tests", "oracle", ".git", "gpurun_out", "baseline"):
            continue
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(base, f)).read()
                if pat.search(src):
                    users[os.path.normpath(os.path.join(rel, f))] = src
    assert sorted(users) == ["__graft_entry__.py", "bench.py", os.path.join("scripts", "train_step_bench.py")], sorted(users)
    assert not any(k.startswith("contrastive_lift_b200") for k in users)
    # bench.py: the GPU arm (run_ours) reaches the oracle only through cpu_reference_rate (the cpu_baseline leg)
    bench_src = users["bench.py"]
    body = bench_src[bench_src.index("def run_ours"):bench_src.index("def main")]
    assert "oracle" not in body.replace("cpu_reference_rate", "")
