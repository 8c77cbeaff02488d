#!/usr/bin/env python
"""profiles/sass/*.sass.txt: opcode histogram of one kernel plus its tensor-core / tensor-memory / TMA / barrier / mixed-precision
instructions as emitted (cuobjdump -sass), in program order.   usage: sass_excerpt.py <object.o> <mangled-name> <title> > out"""
import collections, re, subprocess, sys
obj, fn, title = sys.argv[1:4]
sass = subprocess.run(["cuobjdump", "-sass", "-fun", fn, obj], capture_output=True, text=True).stdout.splitlines()
ops, keep = collections.Counter(), []
pat = re.compile(r"UTC|LDTM|STTM|UBLKCP|UTMA|SYNCS|FHFMA|FHADD|LDGMC|MULTIMEM|USETMAXREG|R2UR.*TMEM|TCGEN|ELECT|NANOSLEEP")
seen = collections.Counter()
for ln in sass:
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", ln)
    if not m:
        if "Function :" in ln:
            keep.append(ln)
        continue
    ins = re.sub(r"^@!?U?P\w+\s+", "", m.group(2)).split()[0]
    ops[ins.split(".")[0]] += 1
    if pat.search(m.group(2)):
        key = ins
        seen[key] += 1
        if seen[key] <= 6:                      # the first few of each kind; the histogram has the totals
            keep.append(ln.rstrip())
print(title)
print("opcode histogram (top 40): " + ", ".join(f"{k} {v}" for k, v in ops.most_common(40)))
print("tensor-core / tensor-memory / TMA / barrier / mixed-precision instructions as emitted (first 6 of each kind, program order):")
print("\n".join(keep))
