"""Summarise an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file X.csv CMD`)
into a per-kernel markdown table (launch count, total time, share) for profiles/.

    python scripts/launch_summary.py gpurun_out/launches.csv profiles/r01_launches_x.md "title" ["note line"] [--minus other.csv]

``--minus other.csv``: subtract the launch list of a shorter run of the same command (setup launches cancel, what is left is
the steady-state steps in between).
"""
import csv
import sys
from collections import OrderedDict


def aggregate(src):
    lines = [l for l in open(src, errors="replace") if not l.startswith("==")]
    rows = list(csv.reader(lines))
    hdr = rows[0]
    idx = {h: i for i, h in enumerate(hdr)}
    k_name, k_val, k_unit = idx["Kernel Name"], idx["Metric Value"], idx["Metric Unit"]
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6}
    agg = OrderedDict()
    for r in rows[1:]:
        if len(r) <= k_val or "gpu__time_duration" not in ",".join(r):
            continue
        us = float(r[k_val].replace(",", "")) * scale.get(r[k_unit], 1.0)
        n, t = agg.get(r[k_name], (0, 0.0))
        agg[r[k_name]] = (n + 1, t + us)
    return agg


def main():
    argv = list(sys.argv)
    minus = None
    if "--minus" in argv:
        i = argv.index("--minus")
        minus = argv[i + 1]
        del argv[i:i + 2]
    src, out, title = argv[1], argv[2], argv[3]
    note = argv[4] if len(argv) > 4 else ""
    agg = aggregate(src)
    if minus:
        for name, (n, t) in aggregate(minus).items():
            n0, t0 = agg.get(name, (0, 0.0))
            agg[name] = (n0 - n, t0 - t)
        agg = OrderedDict((k, v) for k, v in agg.items() if v[0] > 0)
    total = sum(t for _, t in agg.values())
    md = [f"# {title}", ""]
    if note:
        md += [note, ""]
    md += ["Per-launch times under ncu are cold-cache and serialised; the SHARE column is the comparable figure.", "",
           "| kernel | launches | total us | share |", "|---|---|---|---|"]
    for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        md.append(f"| `{name[:110]}` | {n} | {t:.1f} | {100.0 * t / total:.2f} % |")
    md.append(f"| total | {sum(n for n, _ in agg.values())} | {total:.1f} | 100 % |")
    open(out, "w").write("\n".join(md) + "\n")
    print("\n".join(md))


if __name__ == "__main__":
    main()
