#!/usr/bin/env python
"""Warp-stall samples and executed instructions of a kernel aggregated over SASS line ranges.
    python tools/ncu_regions.py rep kernel_regex [bucket]"""
import csv, io, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
bucket = int(sys.argv[3]) if len(sys.argv) > 3 else 250
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{pat}"],
                     capture_output=True, text=True).stdout
blk = '"Kernel Name"' + out.split('"Kernel Name"')[1]
rows = list(csv.reader(io.StringIO("\n".join(blk.splitlines()[1:]))))
hdr, rows = rows[0], rows[1:]
ci = {h: i for i, h in enumerate(hdr)}
def num(r, h):
    try: return float(r[ci[h]])
    except Exception: return 0.0
for b0 in range(0, len(rows), bucket):
    rs = rows[b0:b0 + bucket]
    s = sum(num(r, "# Samples") for r in rs); i = sum(num(r, "Instructions Executed") for r in rs)
    ops = {}
    for r in rs:
        src = r[ci["Source"]].strip().split()
        if not src: continue
        op = src[1] if src[0].startswith("@") and len(src) > 1 else src[0]
        ops[op] = ops.get(op, 0) + num(r, "Instructions Executed")
    top = ", ".join(f"{k}:{v/1e3:.0f}k" for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:5])
    print(f"{b0:5d}-{b0+len(rs):5d} samples={s:6.0f} instr={i/1e6:7.2f}M  {top}")
