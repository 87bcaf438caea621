#!/usr/bin/env python
"""Per-kernel SASS digest of libwsae_sm100.so: instruction count, registers, and how many of the
Blackwell mnemonics each kernel carries (tcgen05.mma = UTCHMMA / UTCQMMA, tcgen05.ld = LDTM,
tcgen05.commit = UTCBAR, TMA = UTMALDG / UTMAPF / UBLKCP, bf16 FMA = HFMA2.BF16 / FHFMA, mma.sync = HMMA,
cp.async = LDGSTS, reductions = RED / REDG / ATOMG).  No GPU needed.

    python tools/sass_digest.py [lib.so] > profiles/<round>_sass_digest.txt
"""
import re
import subprocess
import sys
from collections import Counter
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
lib = sys.argv[1] if len(sys.argv) > 1 else str(ROOT / "whisper_sae_b200" / "lib" / "libwsae_sm100.so")
WATCH = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTCBAR", "UTCCP", "UTMALDG", "UTMASTG", "UTMAPF", "UBLKCP",
         "SYNCS", "HMMA", "FHFMA", "HFMA2", "LDGSTS", "LDSM", "REDG", "RED", "ATOMG", "ATOMS", "LDG", "STG",
         "LDS", "STS", "SHFL", "BAR", "LDL", "STL"]
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout
regs = dict(re.findall(r"Function (\S+):\s*\n\s*REG:(\d+)", res))
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
print(f"# {Path(lib).name}: per-kernel SASS digest (cuobjdump -sass, CUDA {subprocess.run(['nvcc', '--version'], capture_output=True, text=True).stdout.split('release ')[-1].split(',')[0]})")
print("# columns: kernel | instructions | registers | mnemonic counts (multicast / 2-CTA forms are listed with their suffix)")
for blk in re.split(r"\n\s*Function : ", sass)[1:]:
    name = blk.splitlines()[0].strip()
    ops = re.findall(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]+)", blk)
    cnt = Counter()
    for op in ops:
        base = op.split(".")[0]
        if base in WATCH:
            key = base
            if base in ("UTMALDG", "UTCHMMA", "UTCBAR", "HFMA2", "HMMA", "UBLKCP", "REDG", "RED", "ATOMG", "LDG"):
                suffix = [p for p in op.split(".")[1:] if p in ("MULTICAST", "2CTA", "BF16", "BF16_V2", "2D", "ADD",
                                                                "F32", "128", "64", "16816", "GATHER4")]
                key = ".".join([base] + suffix)
            cnt[key] += 1
    pretty = demangle(name)
    pretty = re.sub(r"\(.*", "", pretty).replace("void wsae::", "").replace("(anonymous namespace)::", "")
    marks = "  ".join(f"{k}={v}" for k, v in sorted(cnt.items()))
    print(f"{pretty} | {len(ops)} | {regs.get(name, '?')} | {marks}")
