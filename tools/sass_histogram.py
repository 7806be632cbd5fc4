#!/usr/bin/env python
"""Opcode histogram of the shipped kernels (cuobjdump -sass of libfractencode_b200.so): the tensor / TMEM / bulk-copy
mnemonics that prove the tcgen05 path, per kernel.  Usage: python tools/sass_histogram.py > profiles/sass_r2.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "fractencode_b200", "libfractencode_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
KEY = ("UTCHMMA", "UTCIMMA", "UTCQMMA", "LDTM", "STTM", "UTCBAR", "UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "USETMAXREG", "FMNMX3", "FMNMX", "VIMNMX3", "VIMNMX",
       "DFMA", "DMUL", "DADD", "IDP", "HFMA2", "PRMT", "ATOMG", "REDG", "LDG", "STG", "LDS", "STS", "BAR", "MUFU")
cur, hist = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", name.replace("(anonymous namespace)::", "").replace("void ", ""))
        hist[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        hist[cur][m.group(1)] += 1
print("# SASS opcode histogram per kernel (sm_100a), %s" % os.path.basename(so))
print("# columns: total instructions, then the mnemonics that matter for the roofline argument (prefix match)\n")
for k, c in hist.items():
    tot = sum(c.values())
    parts = []
    for key in KEY:
        n = sum(v for op, v in c.items() if op.startswith(key) and not any(op.startswith(k2) and len(k2) > len(key) for k2 in KEY if k2 != key and k2.startswith(key)))
        if n:
            parts.append("%s %d" % (key, n))
    print("%-64s %6d | %s" % (k[:64], tot, ", ".join(parts)))
