#!/usr/bin/env python
"""Opcode histogram per kernel of the shipped library (cuobjdump -sass), the evidence for which hardware paths
each kernel uses: LDG.E.128 / REDG.E.ADD.F32x4 (vector row traffic), UBLKCP / UBLKRED (TMA bulk copy / bulk
reduction), UTMALDG (TMA tensor loads), UTCHMMA / LDTM (tcgen05 MMA / TMEM loads), SYNCS (mbarrier), MUFU.

    python scripts/sass_histogram.py > profiles/r02/sass_histogram.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "mkb_b200", "lib", "libkge_b200.so")
KEY = ("LDG.E.128", "LDG.E.64", "LDG.E", "STG", "REDG", "RED.", "ATOMG", "LDS", "STS", "UBLKCP", "UBLKRED", "UTMALDG",
       "UTMAREDG", "UTCHMMA", "UTCBAR", "LDTM", "SYNCS", "MUFU", "BAR.", "SHFL", "FFMA", "FADD", "FMUL")


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
    kernels, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            kernels[cur][m.group(1)] += 1
    want = sys.argv[1:] or None
    print(f"# SASS opcode histogram of {os.path.relpath(LIB, ROOT)} (cuobjdump -sass; sm_100a) — instructions per kernel\n")
    for name, ops in kernels.items():
        dn = demangle(name).replace("kge::", "")
        dn = re.sub(r"\(.*", "", dn)
        if want and not any(w in dn for w in want):
            continue
        total = sum(ops.values())
        picked = []
        for k in KEY:
            n = sum(v for o, v in ops.items() if o.startswith(k))
            if n:
                picked.append(f"{k.rstrip('.')}={n}")
        print(f"{dn}  [{total} instr]\n    " + " ".join(picked))


if __name__ == "__main__":
    main()
