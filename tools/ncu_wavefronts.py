#!/usr/bin/env python
"""Shared-memory wavefronts per SASS instruction class of a kernel in an .ncu-rep (source page):
where the L1 / shared-memory pipe's work comes from, and which instructions have bank conflicts.

    python tools/ncu_wavefronts.py capture.ncu-rep <kernel substring> [--top 25]
"""
import collections
import csv
import io
import subprocess
import sys


def main():
    rep, want = sys.argv[1], sys.argv[2]
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 25
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    blocks, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "hdr": None, "rows": []}
            blocks.append(cur)
        elif cur is not None and cur["hdr"] is None:
            cur["hdr"] = r
        elif cur is not None and len(r) == len(cur["hdr"]):
            cur["rows"].append(r)
    for b in blocks:
        if want not in b["name"]:
            continue
        ix = {n: i for i, n in enumerate(b["hdr"])}
        f = lambda r, k: float(r[ix[k]] or 0) if k in ix else 0.0   # noqa: E731
        tot_w = sum(f(r, "L1 Wavefronts Shared") for r in b["rows"])
        tot_x = sum(f(r, "L1 Wavefronts Shared Excessive") for r in b["rows"])
        tot_i = sum(f(r, "Instructions Executed") for r in b["rows"])
        print(f"== {b['name'][:90]}\n   instructions {tot_i:.3e}  shared wavefronts {tot_w:.3e} (excessive {tot_x:.3e} = {100 * tot_x / max(tot_w, 1):.1f} %)")
        byop = collections.defaultdict(lambda: [0.0, 0.0, 0.0])
        for r in b["rows"]:
            toks = r[ix["Source"]].split()
            op = toks[1] if toks and toks[0].startswith("@") and len(toks) > 1 else (toks[0] if toks else "?")
            a = byop[op]
            a[0] += f(r, "Instructions Executed"); a[1] += f(r, "L1 Wavefronts Shared"); a[2] += f(r, "L1 Wavefronts Shared Excessive")
        print("   by opcode (instr, wavefronts, excessive):")
        for op, a in sorted(byop.items(), key=lambda kv: -kv[1][1])[:12]:
            if a[1] > 0:
                print(f"     {op:28s} {a[0]:.3e} {a[1]:.3e} {a[2]:.3e}   {a[1] / max(a[0], 1):.2f} wf/instr")
        print("   instruction mix:")
        for op, a in sorted(byop.items(), key=lambda kv: -kv[1][0])[:22]:
            print(f"     {op:28s} {a[0]:.3e}  {100 * a[0] / tot_i:.1f} %")
        print("   top instructions by excessive wavefronts:")
        for r in sorted(b["rows"], key=lambda r: -f(r, "L1 Wavefronts Shared Excessive"))[:top]:
            if f(r, "L1 Wavefronts Shared Excessive") > 0:
                print(f"     {f(r, 'L1 Wavefronts Shared Excessive'):.3e} of {f(r, 'L1 Wavefronts Shared'):.3e}  x{f(r, 'Instructions Executed'):.2e}  {r[ix['Source']][:70]}")


if __name__ == "__main__":
    main()
