"""SASS mnemonic summary of the shipped library, per kernel: which functions carry the Blackwell tensor / copy-engine
instructions (B200_PROFILING.md: tcgen05.mma = UTC*MMA, tcgen05.ld/st = LDTM/STTM, TMA = UTMALDG/UBLKCP, mma.sync = HMMA).

    python scripts/sass_summary.py > profiles/r2_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "umbrella_b200", "libumbrella_b200.so")
KEYS = ("UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTMAPF", "HMMA", "LDSM", "LDGSTS",
        "SYNCS", "ELECT", "ACQBULK", "REDUX")
out = subprocess.run(["cuobjdump", "-sass", lib], stdout=subprocess.PIPE, text=True, check=True).stdout
per, name = collections.OrderedDict(), None
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], stdout=subprocess.PIPE, text=True).stdout.strip()
        name = re.sub(r"\(.*", "", name)
        per[name] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and name:
        per[name]["_total"] += 1
        if m.group(1) in KEYS:
            per[name][m.group(1)] += 1
print(f"{os.path.relpath(lib, ROOT)}: {len(per)} kernels (cuobjdump -sass, sm_100a)")
tot = collections.Counter()
for k, c in per.items():
    tot.update(c)
    marks = "  ".join(f"{m} x{c[m]}" for m in KEYS if c[m])
    print(f"{c['_total']:7d} instr  {k[:90]:90s} {marks}")
print("whole library: " + "  ".join(f"{m} x{tot[m]}" for m in KEYS if tot[m]))
