#!/usr/bin/env python
"""Opcode histogram per kernel of the built library (cuobjdump -sass): the static evidence behind the claims in
profiles/README.md — packed FFMA2 in the matrix products, no contracted FFMA where the reference has a*b+c (the only
FFMA are the correction steps of IEEE division / square root and the exact-product spellings of gr_math.cuh), MUFU.RCP /
FCHK per IEEE divide, 128-bit loads and stores, shared-memory 64-bit atomics.

    python scripts/sass_opcounts.py [lib.so] > profiles/rNN_sass_opcounts.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gorender_b200", "lib", "libgorender_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
fn, counts = None, collections.OrderedDict()
for ln in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", ln)
    if m:
        fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
        counts[fn] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", ln)
    if m and fn:
        counts[fn][m.group(1)] += 1
print(f"# cuobjdump -sass {os.path.relpath(lib, ROOT)}  (sm_100a; static instruction counts per kernel)")
KEY = ["FFMA2", "FFMA", "FMUL", "FADD", "MUFU.RCP", "MUFU.RSQ", "FCHK", "CALL.REL.NOINC", "IMAD", "IADD3", "LOP3.LUT", "ISETP",
       "LDG.E.128.CONSTANT", "LDG.E.128", "STG.E.128", "ATOMS", "ATOMG", "RED", "MATCH.ANY", "REDUX", "VOTE", "SHFL", "BAR.SYNC", "LDS", "STS", "LDL", "STL"]
for fn, c in counts.items():
    total = sum(c.values())
    if total < 20:
        continue
    print(f"\n## {fn}\n   {total} instructions")

    def grp(prefix):
        return sum(v for k, v in c.items() if k == prefix or k.startswith(prefix + "."))

    print("   " + "  ".join(f"{k}={grp(k)}" for k in KEY if grp(k)))
    top = ", ".join(f"{k} {v}" for k, v in c.most_common(14))
    print(f"   top: {top}")
