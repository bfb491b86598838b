#!/usr/bin/env python
"""Per-kernel SASS opcode counts of the in-tree library (no GPU needed):
python tools/sass_summary.py visual_underwater_slam_b200/libvus.so > profiles/r2_sass_summary.txt"""
import re
import subprocess
import sys
from collections import OrderedDict

lib = sys.argv[1] if len(sys.argv) > 1 else "visual_underwater_slam_b200/libvus.so"
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
names = {}
ops = ["DMMA", "DFMA", "UBLKCP", "SYNCS", "BAR", "LDG", "STG", "LDS", "STS", "SHFL", "MUFU", "RED", "ATOM", "UTMALDG", "UTCHMMA", "LDTM"]
cur = None
tab = OrderedDict()
arch = set()
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        tab[cur] = dict.fromkeys(ops, 0)
        tab[cur]["total"] = 0
        continue
    m = re.match(r"\s*arch = (\S+)", line)
    if m:
        arch.add(m.group(1))
    if cur is None:
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if not m:
        continue
    op = m.group(1).split(".")[0]
    if op.startswith("ATOM"):          # ATOM / ATOMG / ATOMS
        op = "ATOM"
    elif op in ("REDG", "REDS", "REDUX"):
        op = "RED" if op != "REDUX" else op
    tab[cur]["total"] += 1
    if op in tab[cur]:
        tab[cur][op] += 1
dem = subprocess.run(["c++filt"], input="\n".join(tab.keys()), capture_output=True, text=True).stdout.splitlines()
print("# SASS opcode counts per kernel, %s (arch %s); cuobjdump -sass, static instruction counts" % (lib, ",".join(sorted(arch))))
print("# tcgen05 / TMEM / tensor-map opcodes (UTCHMMA, LDTM, UTMALDG) are expected to be 0: the arithmetic is FP64 (no tcgen05 kind), blocks move by 1-D bulk copies (UBLKCP)")
hdr = ["total"] + ops
print("%-86s %s" % ("kernel", " ".join("%7s" % h for h in hdr)))
tot = dict.fromkeys(hdr, 0)
for (mangled, c), d in zip(tab.items(), dem):
    d = d.replace("vus::", "").replace("void rt::", "").replace("void ", "")
    d = re.sub(r"cub::CUB_\d+_SM_\d+::", "cub::", d)
    print("%-86s %s" % (d[:86], " ".join("%7d" % c[h] for h in hdr)))
    for h in hdr:
        tot[h] += c[h]
print("%-86s %s" % ("ALL KERNELS", " ".join("%7d" % tot[h] for h in hdr)))
