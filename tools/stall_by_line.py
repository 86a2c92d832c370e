"""Join ncu's per-SASS-instruction stall samples with nvdisasm line info -> stalls per CUDA line.

    python tools/stall_by_line.py <report.ncu-rep> <kernel mangled-name substring> [top]
Needs the same libpde_b200.so that was profiled (uses cuobjdump / nvdisasm on it).
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    rep, kname = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 and sys.argv[3].isdigit() else 30
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "cnn-with-pde_b200", "libpde_b200.so")], cwd=tmp,
                   capture_output=True)
    dis = ""
    for cubin in sorted(os.listdir(tmp)):      # one cubin per .cu file: take the one that holds the kernel
        if cubin.endswith(".cubin"):
            d = subprocess.run(["nvdisasm", "--print-line-info", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
            if kname in d:
                dis = d
                break
    # walk the disassembly: track current function and current line
    lines_of = []          # per instruction (in order) of the selected function: line number
    cur_fn, cur_line, in_fn = None, None, False
    for ln in dis.splitlines():
        m = re.match(r"\s*\.text\.(\S+):", ln)
        if m:
            cur_fn = m.group(1)
            in_fn = kname in cur_fn
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur_line = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        if in_fn and re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", ln):
            lines_of.append(cur_line)
    short = next((k for k in ("sbwd_multi_kernel", "sfwd_multi_kernel", "sbwd_kernel", "sfwd_kernel", "emo_bwd_tiled", "emo_fwd_tiled",
                               "tiny_bwd_kernel", "tiny_fwd_kernel", "bwd_kernel", "fwd_kernel") if k in kname), kname)
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + short],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h = rows[1]
    ix = {n: i for i, n in enumerate(h)}
    body, seen = [], set()
    for r in rows[2:]:
        if len(r) == len(h) and r[ix["Address"]] not in seen and r[ix["Address"]] != "Address":
            seen.add(r[ix["Address"]])
            body.append(r)
    if len(body) != len(lines_of):
        print(f"warning: {len(body)} profiled instructions vs {len(lines_of)} disassembled", file=sys.stderr)
    agg = collections.defaultdict(lambda: collections.Counter())
    tot = 0.0
    for r, ln in zip(body, lines_of):
        n = float(r[ix["# Samples"]] or 0)
        tot += n
        a = agg[ln]
        a["samples"] += n
        a["inst"] += float(r[ix["Instructions Executed"]] or 0)
        for k in ("stall_long_sb", "stall_short_sb", "stall_wait", "stall_mio", "stall_lg", "stall_barrier",
                  "stall_no_inst", "stall_selected", "stall_math", "stall_dispatch", "stall_branch_resolving"):
            if k in ix:
                a[k] += float(r[ix[k]] or 0)
    src = {}
    for (f, _l) in agg:
        if f and f not in src:
            for cand in (os.path.join(ROOT, "cnn-with-pde_b200", "csrc", f),):
                if os.path.exists(cand):
                    src[f] = open(cand).read().splitlines()
    tinst = sum(a["inst"] for a in agg.values()) or 1.0
    print(f"total samples {tot:.0f}")
    by = "inst" if "--by-inst" in sys.argv else "samples"
    for ln, a in sorted(agg.items(), key=lambda kv: -kv[1][by])[:top]:
        f, l = ln if ln else ("?", 0)
        text = src.get(f, [""] * (l + 1))[l - 1].strip()[:70] if l else ""
        print(f"{100 * a['samples'] / tot:5.1f}% smp {100 * a['inst'] / tinst:5.1f}% inst | lsb {a['stall_long_sb']:6.0f} ssb {a['stall_short_sb']:5.0f} "
              f"wait {a['stall_wait']:5.0f} mio {a['stall_mio']:4.0f} lg {a['stall_lg']:4.0f} noi {a['stall_no_inst']:5.0f} | {f}:{l} {text}")


if __name__ == "__main__":
    main()
