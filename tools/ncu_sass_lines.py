#!/usr/bin/env python
"""Per-source-line stall samples from a SASS-only `ncu --page source --csv` export, using the line table of the local
.so (same binary as on the GPU box):   python tools/ncu_sass_lines.py gpurun_out/x_source.csv <mangled-name-substr> [top]"""
import collections, csv, re, subprocess, sys, os, tempfile

def line_table(substr):
    d = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(os.path.dirname(__file__), "..", "etude_b200", "libetude_b200.so")], cwd=d, capture_output=True)
    cub = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
    out = subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(d, cub)], capture_output=True, text=True).stdout
    table, cur, infn, frames = {}, None, False, []
    for ln in out.splitlines():
        m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
        if m:
            infn = substr in m.group(1)
            continue
        if not infn:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            frames.append((os.path.basename(m.group(1)), int(m.group(2))))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
        if m:
            # innermost frame that lies in a kernel file (not the shared helpers / CUDA headers)
            own = [f for f in frames if f[0] not in ("common.cuh",) and not f[0].endswith(".hpp") and not f[0].endswith(".h")]
            cur = own[0] if own else (frames[0] if frames else cur)
            table[int(m.group(1), 16)] = cur
            frames = []
    return table

def main():
    path, substr = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    table = line_table(substr)
    rows = list(csv.reader(open(path)))
    hdr = rows[1]
    base = None
    agg = collections.defaultdict(collections.Counter)
    total = 0
    for r in rows[2:]:
        if len(r) < len(hdr) or not r[0].startswith("0x"):
            continue
        a = int(r[0], 16)
        base = a if base is None else base
        key = table.get(a - base, ("?", 0))
        d = dict(zip(hdr, r))
        n = int(d["# Samples"] or 0)
        total += n
        agg[key]["samples"] += n
        agg[key]["inst"] += int(d["Instructions Executed"] or 0)
        for k, v in d.items():
            if k.startswith("stall_") and "Not Issued" not in k and v not in ("", "0"):
                agg[key][k] += int(v)
    srcs = {}
    def src(key):
        f = os.path.join(os.path.dirname(__file__), "..", "etude_b200", "csrc", key[0])
        if key[0] not in srcs:
            srcs[key[0]] = open(f).read().splitlines() if os.path.exists(f) else []
        L = srcs[key[0]]
        return L[key[1] - 1].strip() if 0 < key[1] <= len(L) else ""
    print("total samples", total)
    for key, c in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:top]:
        stalls = ", ".join(f"{k[6:]}:{v}" for k, v in c.most_common(6) if k.startswith("stall_"))
        print(f"{100 * c['samples'] / max(total, 1):5.1f}% {key[0]}:{key[1]:<4d} inst={c['inst']:<9d} {src(key)[:95]}\n        [{stalls}]")

if __name__ == "__main__":
    main()
