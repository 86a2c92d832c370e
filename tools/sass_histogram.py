#!/usr/bin/env python
"""SASS opcode histogram per kernel of libpde_b200.so (cuobjdump -sass): static instruction counts, with the
Blackwell-specific opcodes called out (UBLKCP = TMA bulk copy, SYNCS = mbarrier, LDTM / STTM = tensor-memory
load / store, LDGSTS = cp.async, FFMA2 / FMUL2 / FADD2 = packed fp32).

    python tools/sass_histogram.py [kernel substring ...] > profiles/r02_sass_histogram.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "cnn-with-pde_b200", "libpde_b200.so")
MARK = ("UBLKCP", "UBLKPF", "SYNCS", "LDTM", "STTM", "LDGSTS", "FFMA2", "FMUL2", "FADD2", "UTCBAR", "UTCMMA", "HMMA", "UTMALDG")


def main():
    want = sys.argv[1:]
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = kernels.setdefault(re.sub(r"\(.*", "", name), collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and cur is not None:
            cur[m.group(1)] += 1
    total_mark = collections.Counter()
    print(f"# {os.path.relpath(LIB, ROOT)}: static SASS opcode counts per kernel (sm_100a); * = Blackwell / async / packed-fp32 opcodes")
    for name, c in kernels.items():
        if want and not any(w in name for w in want):
            continue
        n = sum(c.values())
        marks = ", ".join(f"{k} {c[k]}" for k in MARK if c[k])
        top = ", ".join(f"{k} {v}" for k, v in c.most_common(14))
        print(f"\n== {name}\n   {n} instructions;  * {marks or '-'}\n   {top}")
        for k in MARK:
            total_mark[k] += c[k]
    print("\n# library totals: " + ", ".join(f"{k} {v}" for k, v in total_mark.items() if v))


if __name__ == "__main__":
    main()
