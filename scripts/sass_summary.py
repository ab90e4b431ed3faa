"""Per-kernel census of the Blackwell-native SASS in libpcv_b200.so (cuobjdump -sass): tcgen05 MMA (UTC*MMA), TMA loads /
stores (UTMALDG / UTMASTG), TMEM loads (LDTM), tcgen05 commit barriers (UTCBAR), TMEM allocation (UTCATOMSWS), packed
fp32x2 math (FFMA2 / FADD2), cluster barriers.  Usage: python scripts/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "pytorchcv_b200", "libpcv_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
MNEM = ["UTCHMMA", "UTCQMMA", "UTCMMA", "UTMALDG", "UTMASTG", "UTMAPF", "LDTM", "UTCBAR", "UTCATOMSWS", "SYNCS", "UCGABAR",
        "FFMA2", "FADD2", "F2FP", "HMNMX2", "MUFU.TANH", "LDS", "STS", "LDG", "STG", "ATOM", "RED"]
kern, name = collections.OrderedDict(), None
arch = None
for ln in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", ln)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(.*", "", name)
        kern[name] = collections.Counter()
        continue
    m = re.match(r"\s*arch = (\S+)", ln)
    if m:
        arch = m.group(1)
    if name is None:
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
    if m:
        op = m.group(1)
        kern[name]["_instr"] += 1
        for k in MNEM:
            if op.startswith(k):
                kern[name][k] += 1
print(f"# SASS census of {os.path.relpath(lib, ROOT)} (arch {arch}; cuobjdump -sass; {len(kern)} kernels)")
print("# columns: instructions | " + " ".join(MNEM))
tot = collections.Counter()
for n, c in kern.items():
    tot.update(c)
    hot = " ".join(f"{k}={c[k]}" for k in MNEM if c[k])
    print(f"{n[:120]:120s} {c['_instr']:6d} | {hot}")
print("# totals: " + " ".join(f"{k}={tot[k]}" for k in MNEM if tot[k]))
