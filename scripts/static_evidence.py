#!/usr/bin/env python
"""Static evidence from the built library (no GPU needed): per-kernel registers / stack / static shared memory from
`cuobjdump --dump-resource-usage`, and the count of Blackwell-specific SASS mnemonics per kernel from `cuobjdump -sass`
(B200_PROFILING.md: tcgen05.mma -> UTC*MMA, tcgen05.ld/st -> LDTM/STTM, bulk TMA -> UBLKCP/UTMALDG, cp.async -> LDGSTS,
red.global -> REDG).  Writes profiles/r02_static_sass.md (CLIFT_STATIC_OUT names another file).

    python scripts/static_evidence.py
"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "contrastive_lift_b200", "libclift_b200.so")
OUT = os.path.join(ROOT, "profiles", os.environ.get("CLIFT_STATIC_OUT", "r02_static_sass.md"))
MNEMONICS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UTMALDG", "SYNCS", "LDGSTS", "REDG", "HMMA"]


def demangle(names):
    out = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.strip().split("\n")
    clean = []
    for d in out:
        d = d.replace("clift::(anonymous namespace)::", "").replace("clift::", "").replace("void ", "")
        clean.append(re.sub(r"\(.*", "", d))
    return clean


def main():
    res = subprocess.run(["cuobjdump", "--dump-resource-usage", LIB], capture_output=True, text=True).stdout
    items = re.findall(r"Function (\S+):\s*\n\s*REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", res)
    usage = {n: (int(r), int(st), int(sh), int(lo)) for n, r, st, sh, lo in items}
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    counts = collections.defaultdict(collections.Counter)
    pat = re.compile(r"\b(" + "|".join(MNEMONICS) + r")\b")
    fn = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = m.group(1)
            continue
        if fn:
            for t in pat.findall(line):
                counts[fn][t] += 1
    names = sorted(usage)
    nice = dict(zip(names, demangle(names)))
    rows = sorted((nice[n], n) for n in names)
    with open(OUT, "w") as f:
        f.write("# Round 2 - static evidence from libclift_b200.so (sm_100a cubin, `scripts/static_evidence.py`, no GPU involved)\n\n")
        f.write("Registers / stack bytes / static shared bytes per kernel (`cuobjdump --dump-resource-usage`; dynamic shared memory is set "
                "at launch) and counts of the SASS mnemonics that identify the Blackwell paths (`cuobjdump -sass`): `UTCHMMA` = tcgen05.mma, "
                "`UTCBAR` = tcgen05.commit, `LDTM`/`STTM` = tcgen05.ld/st (tensor memory), `UBLKCP` = bulk TMA copy (cp.async.bulk), "
                "`SYNCS` = mbarrier ops, `LDGSTS` = cp.async, `REDG` = red.global. `HMMA` (legacy mma.sync) appears nowhere.\n\n")
        f.write("| kernel | regs | stack B | static smem B | " + " | ".join(MNEMONICS) + " |\n")
        f.write("|---|---|---|---|" + "---|" * len(MNEMONICS) + "\n")
        for label, n in rows:
            r, st, sh, _ = usage[n]
            c = counts.get(n, {})
            f.write(f"| `{label}` | {r} | {st} | {sh} | " + " | ".join(str(c.get(k, 0) or "") for k in MNEMONICS) + " |\n")
        spills = [(nice[n], usage[n][1]) for n in names if usage[n][1] > 0]
        f.write(f"\n{len(names)} kernels; kernels with a non-zero stack frame (local arrays or spills): "
                + (", ".join(f"`{a}` ({b} B)" for a, b in sorted(spills)) if spills else "none") + ".\n")
        f.write("\nWhere the stack frames come from (round 1's `nvdisasm -g` line mapping of every `LDL`/`STL`, "
                "`profiles/r01_static_sass.md`; the code around them is unchanged): in `heads_tc16_forward_kernel` and "
                "`heads_tc_forward_kernel` they sit on the `sinf`/`cosf` calls of the positional-encoding branch `pe_sem / pe_ins > 0` "
                "(libdevice's slow-path argument reduction keeps its table walk in local memory); the shipped configurations "
                "(`pe_sem = pe_ins = 0`) take the branch above it and never execute them.  `march_kernel` keeps the ray origin / "
                "direction (6 words, written once per ray) in local memory and re-reads them in `sample_point` (L1 hits); "
                "`march_backward_kernel<3>` spills a few loop-invariant scalars set up before its ray loop.  The FP32-FMA head "
                "kernels' frames are the same libdevice slow path.  `heads_backward_kernel` has no frame any more: its 96 bytes were the "
                "data-gradient engine's state (`DgEngine`'s fill counters were indexed by the operand-buffer parity, which put the "
                "whole struct in local memory - found with `nvdisasm -g`, now two scalars).  The tensor-core kernels touch no local "
                "memory in their steady-state MMA loops.\n")
    print(OUT, len(names), "kernels")


if __name__ == "__main__":
    main()
