#!/usr/bin/env python
"""Top SASS instructions of a kernel by warp-stall samples / executed count, from an .ncu-rep source page.

    python tools/ncu_hot.py gpurun_out/x.ncu-rep encode_topk [N]
"""
import csv, io, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{pat}"],
                     capture_output=True, text=True).stdout
blocks = out.split('"Kernel Name"')
blk = '"Kernel Name"' + blocks[1]
lines = blk.splitlines()
rows = list(csv.reader(io.StringIO("\n".join(lines[1:]))))
hdr, rows = rows[0], rows[1:]
ci = {h: i for i, h in enumerate(hdr)}
def num(r, h):
    try: return float(r[ci[h]])
    except Exception: return 0.0
tot_s = sum(num(r, "# Samples") for r in rows)
tot_i = sum(num(r, "Instructions Executed") for r in rows)
print(f"kernel block 1 of {len(blocks)-1}: {len(rows)} SASS lines, samples={tot_s:.0f}, warp-instructions={tot_i:.0f}")
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {h: sum(num(r, h) for r in rows) for h in stalls}
print("stall totals:", {k: int(v) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
print("--- top by samples ---")
for idx in sorted(range(len(rows)), key=lambda i: -num(rows[i], "# Samples"))[:n]:
    r = rows[idx]
    top = max(stalls, key=lambda h: num(r, h))
    print(f"{idx:5d} {num(r,'# Samples'):8.0f} {num(r,'Instructions Executed'):10.0f}  {r[ci['Source']].strip()[:70]:70s} {top}")
# opcode histogram by executed instructions
ops = {}
for r in rows:
    op = r[ci["Source"]].strip().split()[0] if r[ci["Source"]].strip() else "?"
    if op.startswith("@"):
        op = r[ci["Source"]].strip().split()[1]
    ops[op] = ops.get(op, 0) + num(r, "Instructions Executed")
print("--- executed warp-instructions by opcode ---")
for op, c in sorted(ops.items(), key=lambda kv: -kv[1])[:25]:
    print(f"{op:24s} {c:12.0f} {100*c/tot_i:5.1f}%")
