#!/usr/bin/env python
"""Opcode histogram of the built library's SASS (cuobjdump -sass), per kernel family: the evidence that the hot
path is tcgen05 (UTC*MMA) + TMEM (LDTM/STTM) + TMA (UTMALDG) code and carries no legacy HMMA.  CPU-only.

    python tools/sass_opcodes.py > profiles/r02_sass_opcodes.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "wav2vec_contr_loss_b200", "lib", "libsupcon_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
per_kernel, cur = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = per_kernel.setdefault(m.group(1), collections.Counter())
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and cur is not None:
        cur[m.group(1)] += 1
KEY = ("UTCHMMA", "UTCBAR", "UTCATOMSWS", "LDTM", "STTM", "UTMALDG", "UTMAPF", "SYNCS", "MUFU", "FFMA", "FFMA2", "FADD2", "FMUL2", "HMMA", "LDGSTS")
print(f"# cuobjdump -sass {os.path.relpath(lib, ROOT)} -- instruction counts per kernel (static SASS)")
print(f"# {'kernel':70s} {'total':>7s} " + " ".join(f"{k:>8s}" for k in KEY))
tot = collections.Counter()
for name, c in per_kernel.items():
    short = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    short = re.sub(r"\(anonymous namespace\)::", "", short).split("(")[0].replace("void ", "").replace("supcon::", "")
    print(f"{short[:72]:72s} {sum(c.values()):7d} " + " ".join(f"{c.get(k, 0):8d}" for k in KEY))
    tot.update(c)
print(f"{'ALL KERNELS':72s} {sum(tot.values()):7d} " + " ".join(f"{tot.get(k, 0):8d}" for k in KEY))
print("\n# full opcode histogram over all kernels")
for op, n_ in tot.most_common():
    print(f"{op:14s} {n_}")
