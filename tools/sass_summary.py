#!/usr/bin/env python3
"""Per-kernel counts of the Blackwell-specific SASS mnemonics in the built library (B200_PROFILING.md: UTCHMMA = tcgen05.mma,
UTMALDG = TMA tensor loads, LDTM / STTM = tcgen05.ld / st, UTCBAR = tcgen05.commit, USETMAXREG = setmaxnreg, SYNCS = mbarrier).
   python tools/sass_summary.py > profiles/r2_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(REPO, "temporal_latticenet_b200", "csrc", "libltn_b200.so")
MNEMONICS = ("UTCHMMA", "UTMALDG", "LDTM", "STTM", "UTCBAR", "USETMAXREG", "SYNCS", "UTCATOMSWS", "HMMA", "FFMA", "ATOM", "RED", "STG", "LDG")

out = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
kernel, counts = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kernel = re.sub(r"\(anonymous namespace\)::", "", name).split("(")[0]
        counts[kernel] = collections.Counter()
        continue
    if kernel is None:
        continue
    m = re.search(r"/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m:
        op = m.group(1)
        counts[kernel]["instructions"] += 1
        for mn in MNEMONICS:
            if op.startswith(mn):
                counts[kernel][mn] += 1
print("SASS summary of %s (cuobjdump -sass, sm_100a)" % os.path.relpath(SO, REPO))
print("%-44s %7s " % ("kernel", "instr") + " ".join("%9s" % m for m in MNEMONICS))
for k, c in counts.items():
    print("%-44s %7d " % (k[:44], c["instructions"]) + " ".join("%9d" % c[m] for m in MNEMONICS))
tot = collections.Counter()
for c in counts.values():
    tot.update(c)
print("%-44s %7d " % ("TOTAL", tot["instructions"]) + " ".join("%9d" % tot[m] for m in MNEMONICS))
