#!/usr/bin/env python
"""Registers / spills / stack of every kernel in a .cu file (nvcc -Xptxas -v, sm_100a).

    python tools/ptxas_report.py cnn-with-pde_b200/csrc/adi_split.cu [-I dir ...] [filter]
"""
import re
import subprocess
import sys


def report(src, includes, extra=()):
    cmd = ["nvcc", "-ccbin", "/usr/bin/g++", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
           "--ftz=false", "--prec-div=true", "--prec-sqrt=true", "-Xptxas", "-v", "-c", src, "-o", "/dev/null"]
    for i in includes:
        cmd += ["-I", i]
    cmd += list(extra)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode:
        sys.stderr.write(r.stderr)
        raise SystemExit(1)
    out, cur = [], None
    for line in r.stderr.splitlines():
        m = re.search(r"Compiling entry function '(\S+)'", line)
        if m:
            cur = {"name": subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()}
            out.append(cur)
            continue
        if cur is None:
            continue
        m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
        if m:
            cur["stack"], cur["sst"], cur["sld"] = map(int, m.groups())
        m = re.search(r"Used (\d+) registers", line)
        if m:
            cur["regs"] = int(m.group(1))
            m2 = re.search(r"(\d+) bytes smem", line)
            cur["smem"] = int(m2.group(1)) if m2 else 0
    return out


if __name__ == "__main__":
    args = sys.argv[1:]
    src = args.pop(0)
    inc, filt, extra = [], None, []
    while args:
        a = args.pop(0)
        if a == "-I":
            inc.append(args.pop(0))
        elif a.startswith("-D"):
            extra.append(a)
        else:
            filt = a
    for k in report(src, inc, extra):
        if filt and filt not in k["name"]:
            continue
        name = re.sub(r"\(.*", "", k["name"])
        print(f"{k.get('regs', 0):4d} regs  stack {k.get('stack', 0):4d}  spill st/ld {k.get('sst', 0):4d}/{k.get('sld', 0):4d}  smem {k.get('smem', 0):6d}  {name}")
