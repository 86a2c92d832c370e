"""Summarise an .ncu-rep: key launch/throughput/stall metrics per kernel + hottest SASS lines.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--top 15]
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_active.avg.per_cycle_active",
    "smsp__warps_eligible.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum",
]


def ncu_csv(rep, page, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv", *extra], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 12
    rows = ncu_csv(rep, "raw")
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print("==", d.get("Kernel Name"))
        for k in KEYS:
            if k in d:
                print(f"  {k:72s} {d[k]:>18s} {units[hdr.index(k)]}")
        st = []
        for k in hdr:
            if k.startswith("smsp__pcsamp_warps_issue_stalled_") and not k.endswith("_not_issued"):
                try:
                    st.append((float(d[k]), k[len("smsp__pcsamp_warps_issue_stalled_"):]))
                except ValueError:
                    pass
        tot = sum(v for v, _ in st) or 1.0
        print("  stalls: " + ", ".join(f"{n} {100 * v / tot:.0f}%" for v, n in sorted(st, reverse=True)[:8]))
    src = ncu_csv(rep, "source")
    if len(src) > 2:
        h = src[1]
        ix = {n: i for i, n in enumerate(h)}
        body = [r for r in src[2:] if len(r) == len(h)]

        def f(r, k):
            try:
                return float(r[ix[k]])
            except (ValueError, KeyError):
                return 0.0
        total = sum(f(r, "# Samples") for r in body) or 1.0
        byop = collections.Counter()
        for r in body:
            toks = r[ix["Source"]].split()
            op = toks[1] if toks and toks[0].startswith("@") and len(toks) > 1 else (toks[0] if toks else "?")
            byop[op.split(".")[0]] += f(r, "# Samples")
        print("  samples by opcode: " + ", ".join(f"{o} {100 * v / total:.0f}%" for o, v in byop.most_common(10)))
        execd = collections.Counter()
        for r in body:
            toks = r[ix["Source"]].split()
            op = toks[1] if toks and toks[0].startswith("@") and len(toks) > 1 else (toks[0] if toks else "?")
            execd[op.split(".")[0]] += f(r, "Instructions Executed")
        te = sum(execd.values()) or 1.0
        print("  instructions by opcode: " + ", ".join(f"{o} {100 * v / te:.0f}%" for o, v in execd.most_common(12)))
        for r in sorted(body, key=lambda r: -f(r, "# Samples"))[:top]:
            print(f'   {100 * f(r, "# Samples") / total:5.2f}%  lsb={f(r, "stall_long_sb"):6.0f} ssb={f(r, "stall_short_sb"):6.0f} '
                  f'wait={f(r, "stall_wait"):6.0f} mio={f(r, "stall_mio"):5.0f} lg={f(r, "stall_lg"):5.0f}  {r[ix["Source"]][:80]}')


if __name__ == "__main__":
    main()
