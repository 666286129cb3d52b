#!/usr/bin/env python
"""Aggregates warp-stall samples of an .ncu-rep per CUDA source line (needs -lineinfo + --import-source on).

    python tools/ncu_lines.py gpurun_out/prof_x.ncu-rep [top_n]
"""
import collections
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    rows = csv.reader(io.StringIO(out))
    cur_file, hdr = None, None
    per_line = collections.defaultdict(lambda: collections.Counter())
    src = {}
    total = 0
    kern = None
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            if kern is None:
                kern = r[1]
            elif r[1] != kern and len(per_line) and "--all" not in sys.argv:
                pass
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or len(r) < len(hdr):
            continue
        d = dict(zip(hdr[:2] + ["Address", "Sass"] + hdr[4:], r))
        if d["Line No"] == "":
            continue  # SASS rows: already aggregated in their source-line row
        try:
            n = int(d["# Samples"] or 0)
        except ValueError:
            continue
        key = (cur_file, int(d["Line No"]))
        src[key] = d["Source"].strip()
        per_line[key]["samples"] += n
        total += n
        for k, v in d.items():
            if k.startswith("stall_") and "Not Issued" not in k and v not in ("", "0"):
                per_line[key][k] += int(v)
        per_line[key]["inst"] += int(d["Instructions Executed"] or 0)
    print(kern, "total samples", total)
    for key, c in sorted(per_line.items(), key=lambda kv: -kv[1]["samples"])[:top]:
        stalls = ", ".join(f"{k[6:]}:{v}" for k, v in c.most_common(6) if k.startswith("stall_"))
        print(f"{100 * c['samples'] / max(total, 1):5.1f}% {key[0]}:{key[1]:<4d} inst={c['inst']:<9d} {src[key][:90]}\n        [{stalls}]")


if __name__ == "__main__":
    main()
