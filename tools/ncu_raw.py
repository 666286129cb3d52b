#!/usr/bin/env python
"""Prints selected metrics from an `ncu --page raw --csv` export (one column per launch)."""
import csv, sys, re
pat = re.compile(sys.argv[2] if len(sys.argv) > 2 else r"gpu__time_duration.sum|dram__bytes_(read|write).sum$|sm__pipe_tensor.*cycles_active.avg.pct|sm__inst_executed_pipe_(xu|alu|fma|fmaheavy|uniform|lsu|tmem).*pct|smsp__issue_active.avg.pct|sm__warps_active.avg.pct|launch__registers_per_thread|lts__t_bytes.sum$|l1tex__m_xbar2l1tex_read_bytes.sum$|sm__cycles_elapsed.max|smsp__cycles_active.avg|lts__t_sectors_srcunit_tex_op_read.sum$|gpu__dram_throughput|sm__throughput.avg.pct|smsp__inst_executed.sum$")
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
for i, h in enumerate(hdr):
    if pat.search(h):
        print(f"{h:75s} {units[i]:12s} " + "  ".join(r[i] for r in rows[2:]))
