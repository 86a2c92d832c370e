#!/usr/bin/env python
"""Fold the per-launch counters bench.py quotes (DRAM bytes, executed warp instructions) of one
layer's forward / backward kernels from an `ncu --set full` capture into profiles/kernel_counters.json.

    python tools/ncu_counters.py <layer> <capture.ncu-rep> [--sha <csrc sha the capture ran on>]
"""
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    layer, rep = sys.argv[1], sys.argv[2]
    sha = sys.argv[sys.argv.index("--sha") + 1] if "--sha" in sys.argv else None
    if sha is None:
        import bench
        sha = bench._csrc_sha()
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[0]
    res = {}
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        name = d.get("Kernel Name", "")
        kind = "bwd" if re.search(r"bwd", name) else ("fwd" if re.search(r"fwd", name) else None)
        if kind is None:
            continue
        num = lambda k: float(d[k].replace(",", "")) if d.get(k) else 0.0   # noqa: E731
        unit = lambda k: rows[1][hdr.index(k)] if k in hdr else ""          # noqa: E731
        scale = lambda k: {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(unit(k), 1.0)   # noqa: E731
        dram = num("dram__bytes_read.sum") * scale("dram__bytes_read.sum") + num("dram__bytes_write.sum") * scale("dram__bytes_write.sum")
        res[kind + "_kernel"] = re.sub(r"\(.*", "", name)
        res[kind + "_dram_bytes"] = int(dram)
        res[kind + "_inst_executed"] = int(num("smsp__inst_executed.sum"))
        res[kind + "_ncu_ms"] = round(num("gpu__time_duration.sum") * {"ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}.get(unit("gpu__time_duration.sum"), 1.0), 4)
    path = os.path.join(ROOT, "profiles", "kernel_counters.json")
    allc = json.load(open(path)) if os.path.exists(path) else {}
    allc[layer] = res
    allc["_csrc_sha"] = sha
    allc["_note"] = ("per launch at the layer's bench batch, from `ncu --set full --clock-control none` (profiles/r02_ncu_*.txt); "
                     "dram bytes = dram__bytes_read.sum + dram__bytes_write.sum, inst = smsp__inst_executed.sum (warp instructions)")
    json.dump(allc, open(path, "w"), indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
