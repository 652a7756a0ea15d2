#!/usr/bin/env python
"""One text block per kernel launch of an ncu report (`ncu --set full ... -o x` -> x.ncu-rep), the form used in
profiles/*_trace_kernel_summary.md. Runs on the CPU (`ncu -i`); nothing here touches a GPU.

  python tools/ncu_summary.py gpurun_out/r2f_c2.ncu-rep [warp-pops of the launch, for the per-pop figure]
"""
import collections
import csv
import io
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct",
    "lts__t_sector_hit_rate.pct", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__cycles_elapsed.avg",
    "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
]
STALL = "smsp__average_warps_issue_stalled_"


def page(rep, name, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv", *extra], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    pops = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
    rows = page(rep, "raw")
    names, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(names, r))
        u = dict(zip(names, units))
        print(d["Kernel Name"])
        for m in METRICS:
            if m in d:
                print("  %-70s %s %s" % (m, d[m], u[m]))
        stalls = sorted(((float(v), k[len(STALL):-len("_per_issue_active.ratio")]) for k, v in d.items()
                         if k.startswith(STALL) and k.endswith("_per_issue_active.ratio") and v not in ("", "n/a")), reverse=True)
        for v, k in stalls[:8]:
            print("  stall %-40s %.3f" % (k, v))
        if pops:
            print("  warp instructions per warp-pop: %.0f" % (float(d["smsp__inst_executed.sum"]) / pops))
    # opcode mix from the SASS page of the (first) kernel
    src = page(rep, "source")
    hdr = next((i for i, r in enumerate(src) if "Instructions Executed" in r), None)
    if hdr is not None:
        iS, iI = src[hdr].index("Source"), src[hdr].index("Instructions Executed")
        mix = collections.Counter()
        for r in src[hdr + 1:]:
            if len(r) > iI and r[iI].isdigit():
                op = r[iS].split()
                if op and op[0].startswith("@"):
                    op = op[1:]
                if op:
                    mix[op[0].split(".")[0]] += int(r[iI])
        tot = sum(mix.values())
        print("  opcode mix: " + ", ".join("%s %.1f%%" % (k, 100.0 * v / tot) for k, v in mix.most_common(16)))
        print("  total warp instr %d" % tot)


if __name__ == "__main__":
    main()
