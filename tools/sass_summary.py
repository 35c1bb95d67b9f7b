"""Per-kernel SASS summary of libmmf_b200.so (cuobjdump -sass): instruction totals and the mnemonics that show which
hardware paths a kernel uses -- UTCHMMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UTMALDG (TMA), FFMA2 / FADD2 /
FMUL2 (packed FP32), DFMA (FP64), HMMA (mma.sync).  Runs without a GPU.  Usage: python tools/sass_summary.py > profiles/rN_sass_summary.txt"""
import collections
import re
import subprocess
import sys
from pathlib import Path

so = Path(__file__).resolve().parents[1] / "modulation_mfcc_b200" / "libmmf_b200.so"
out = subprocess.run(["cuobjdump", "-sass", str(so)], capture_output=True, text=True, check=True).stdout
demangle = lambda n: subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip() or n
KEYS = ["UTCHMMA", "UTCCP", "LDTM", "STTM", "UTMALDG", "SYNCS", "FFMA2", "FADD2", "FMUL2", "FFMA", "DFMA", "HMMA", "MUFU", "LDS", "STS", "LDG", "STG", "SHFL", "BAR"]
cur, counts = None, {}
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        op = m.group(1)
        counts[cur]["total"] += 1
        for k in KEYS:
            if op == k or (k in ("LDS", "STS", "LDG", "STG", "BAR", "SHFL", "MUFU", "SYNCS", "HMMA", "UTMALDG", "UTCHMMA") and op.startswith(k)):
                counts[cur][k] += 1
print(f"# SASS summary of {so.name} (cuobjdump -sass, sm_100a): static instruction counts per kernel")
print(f"# {'kernel':78s} total  " + " ".join(f"{k:>7s}" for k in KEYS))
for fn, c in sorted(counts.items(), key=lambda kv: -kv[1]["total"]):
    name = demangle(fn)
    name = (name.split(">(")[0] + ">" if ">(" in name else name.split("(")[0]).replace("(int)", "").replace("(bool)", "").replace("mmf::", "").replace("void ", "")[:78]
    print(f"{name:80s} {c['total']:5d}  " + " ".join(f"{c[k]:7d}" for k in KEYS))
