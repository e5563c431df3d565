#!/usr/bin/env python
"""Opcode histogram of the kernels of a cubin/.so whose name matches a substring.
    python tools/sass_mix.py hexl_b200/lib/libhexl_b200.so ntt_row_fwdILi1ELi12"""
import collections
import re
import subprocess
import sys

path, pat = sys.argv[1], sys.argv[2]
txt = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
for part in re.split(r"\n\s*Function : ", txt)[1:]:
    name = part.split("\n")[0]
    if pat not in name:
        continue
    c = collections.Counter()
    for line in part.split("\n"):
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
        if m:
            c[m.group(2)] += 1
    fam = collections.Counter()
    for k, v in c.items():
        base = k.split(".")[0]
        if k.startswith("IMAD.WIDE"):
            fam["fma-heavy wide"] += v
        elif base in ("IMAD",) or k.startswith("IMAD."):
            fam["fma-heavy narrow (IMAD*)"] += v
        elif base in ("DFMA", "DADD", "DMUL"):
            fam["fp64"] += v
        elif base in ("IADD3", "LOP3", "SHF", "PRMT", "MOV", "SEL", "ISETP", "VIADD", "LEA", "IABS", "VIMNMX"):
            fam["alu"] += v
        elif base in ("LDG", "STG", "LDS", "STS", "LDL", "STL", "LDC", "LDCU"):
            fam["mem"] += v
        else:
            fam["other"] += v
    print(name[:110])
    print("   total", sum(c.values()), dict(fam))
    print("   ", ", ".join(f"{k} {v}" for k, v in c.most_common(18)))
