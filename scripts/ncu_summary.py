"""Summarise an .ncu-rep (one kernel, --set full) into a small markdown table for profiles/.

    python scripts/ncu_summary.py gpurun_out/prof_heads_tc.ncu-rep profiles/r01_ncu_heads_tc.md "title"
Also writes <out>.json with the numbers bench.py uses for roofline.traffic.
"""
import csv
import io
import json
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "kernel duration"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem / block"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput (% of peak)"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput (% of peak)"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit rate"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_active", "L1/TEX throughput (% of peak)"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput (% of peak)"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active (%)"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe active (%)"),
    ("sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active", "TMEM pipe inst (%)"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy (%)"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy (%)"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall: long scoreboard"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall: short scoreboard"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall: barrier"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall: math pipe throttle"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall: wait"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall: not selected"),
    ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall: branch resolving"),
]


def main():
    rep, out, title = sys.argv[1], sys.argv[2], sys.argv[3]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    idx = {h: i for i, h in enumerate(hdr)}
    name = vals[idx["Kernel Name"]] if "Kernel Name" in idx else "?"
    lines = [f"# {title}", "", f"Kernel: `{name}`  (one launch, `ncu --set full --clock-control none`; times under ncu are "
             "replayed/cold-cache, use bench.py's CUDA-event numbers for durations)", "", "| metric | value | unit |", "|---|---|---|"]
    js = {"kernel": name}
    for key, label in WANT:
        if key in idx:
            lines.append(f"| {label} (`{key}`) | {vals[idx[key]]} | {units[idx[key]]} |")
            try:
                js[key] = float(vals[idx[key]].replace(",", ""))
                js[key + "__unit"] = units[idx[key]]
            except ValueError:
                pass
    open(out, "w").write("\n".join(lines) + "\n")
    json.dump(js, open(out.replace(".md", ".json"), "w"), indent=1)
    print("\n".join(lines))


if __name__ == "__main__":
    main()
