#!/usr/bin/env python
"""Stage the reference's own hot-path modules under the git-ignored oracle/_ref/ so they travel to the GPU box.

    python oracle/vendor_reference.py             (also run by __graft_entry__.build() when /root/reference is present)

/root/reference does not exist on the GPU box, but the files of the path are plain Python: this script imports the
UNMODIFIED reference through oracle/refload.py (stand-ins for its non-arithmetic imports), records which files under
/root/reference that import actually executed, and copies exactly those - byte for byte, same relative paths - into
oracle/_ref/.  oracle/_ref/ is listed in .gitignore (no reference source enters the history) and not in .gpurunignore
(it ships with the snapshot like the built .so).  `bench.py --impl reference`, the `cpu_baseline` leg and the on-box
parity checks then run the reference itself (`cpu_baseline.kind == "reference"`); without oracle/_ref they fall back to
the reference-pinned port in oracle/clift_oracle.py (`kind == "port"`).  TEST / MEASUREMENT INFRASTRUCTURE ONLY.
"""
import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference"
DST = os.path.join(ROOT, "oracle", "_ref")


def main() -> int:
    if not os.path.isdir(os.path.join(SRC, "model", "renderer")):
        print(f"vendor_reference: {SRC} not present - nothing staged")
        return 0
    sys.path.insert(0, ROOT)
    os.environ["CLIFT_REFERENCE_ROOT"] = SRC
    from oracle import refload
    refload.load()
    files = sorted({os.path.realpath(m.__file__) for m in list(sys.modules.values())
                    if getattr(m, "__file__", None) and os.path.realpath(m.__file__).startswith(SRC + os.sep)})
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    manifest = {}
    for f in files:
        rel = os.path.relpath(f, SRC)
        out = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(out), exist_ok=True)
        shutil.copyfile(f, out)
        manifest[rel] = hashlib.sha256(open(f, "rb").read()).hexdigest()
    json.dump({"source": SRC, "files": manifest}, open(os.path.join(DST, "MANIFEST.json"), "w"), indent=1)
    print(f"vendor_reference: staged {len(files)} files under oracle/_ref/: " + ", ".join(manifest))
    return 0


if __name__ == "__main__":
    sys.exit(main())
