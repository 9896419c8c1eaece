"""SASS evidence for profiles/: per kernel of libsdp_b200.so the instruction counts that matter -
UBLKCP (TMA bulk copies), SYNCS (mbarrier), LDS / LDG, DADD / DMUL / DFMA (the sweep kernels must
contain NO DFMA: the parity contract forbids contraction), MEMBAR.
    python scripts/sass_extract.py > profiles/r2_sass_extract.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "stodynprog_b200", "_lib", "libsdp_b200.so")
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
arch = re.search(r"arch = (sm_\w+)", out)
print("# cuobjdump -sass stodynprog_b200/_lib/libsdp_b200.so ; arch =", arch.group(1) if arch else "?")
print("# columns: UBLKCP SYNCS LDS LDG DADD DMUL DFMA MEMBAR total-instructions  kernel")
WANT = ["UBLKCP", "SYNCS", "LDS", "LDG", "DADD", "DMUL", "DFMA", "MEMBAR"]
cur, counts = None, None
rows = []
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        if cur:
            rows.append((cur, counts))
        cur, counts = m.group(1), collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        counts[m.group(1)] += 1
        counts["total"] += 1
if cur:
    rows.append((cur, counts))
demangle = subprocess.run(["c++filt"], input="\n".join(r[0] for r in rows), capture_output=True, text=True).stdout.splitlines()
bad = 0
for (name, c), dm in sorted(zip(rows, demangle), key=lambda x: x[1]):
    short = re.sub(r"\(.*", "", dm).replace("void ", "")
    print("%6d %5d %5d %5d %5d %5d %5d %6d %8d  %s" % tuple([c[k] for k in WANT] + [c["total"], short]))
    if short.startswith("k_sweep") and c["DFMA"]:
        bad += 1
print("# sweep kernels containing DFMA: %d" % bad)
sys.exit(1 if bad else 0)
