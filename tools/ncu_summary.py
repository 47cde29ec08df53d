#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed) into a text file for profiles/:
headline metrics, stall reasons, and instructions per routine from the source page."""
import collections
import csv
import re
import subprocess
import sys


def main(rep, src, niter, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
            "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
            "launch__block_size", "launch__grid_size", "launch__shared_mem_per_block_dynamic",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct",
            "lts__t_sector_hit_rate.pct", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
    lines = [f"# ncu summary of {rep}", ""]
    for i, h in enumerate(hdr):
        if h in want or ("issue_stalled" in h and "per_issue_active" in h and float(vals[i] or 0) > 0.05):
            lines.append(f"{h:88s} {units[i]:16s} {vals[i]}")
    if src:
        sp = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                            capture_output=True, text=True).stdout
        rows = list(csv.reader(sp.splitlines()))
        hi = [i for i, r in enumerate(rows) if r and r[0] == "Line No"][0]
        h = rows[hi]
        iI, iS, iT = h.index("Instructions Executed"), h.index("# Samples"), h.index("Thread Instructions Executed")
        text = open(src).read().split("\n")
        marks = []
        for i, l in enumerate(text):
            m = re.search(r"__device__ .*? (\w+)\(", l)
            if m:
                marks.append((i + 1, m.group(1)))
            if "__global__" in l:
                marks.append((i + 1, "kernel"))

        def region(ln):
            name = "?"
            for a, n in marks:
                if a <= ln:
                    name = n
            return name

        per = collections.OrderedDict()
        tot = tots = 0
        for r in rows[hi + 1:]:
            try:
                ln, ie, sm, th = int(r[0]), int(r[iI]), int(r[iS]), int(r[iT])
            except (ValueError, IndexError):
                continue
            d = per.setdefault(region(ln), [0, 0, 0])
            d[0] += ie
            d[1] += sm
            d[2] += th
            tot += ie
            tots += sm
        lines += ["", f"# warp-instructions by routine ({niter} Newton iterations in this launch)"]
        for k, d in sorted(per.items(), key=lambda kv: -kv[1][0]):
            if d[0] == 0:
                continue
            lines.append(f"{k:20s} inst {100 * d[0] / tot:5.1f}%  samples {100 * d[1] / max(1, tots):5.1f}%  "
                         f"threads/inst {d[2] / d[0]:5.1f}  warp-inst/cell-iteration {d[0] / niter:8.1f}")
        lines.append(f"total warp-inst/cell-iteration {tot / niter:.1f}")
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:12]))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 and sys.argv[2] != "-" else None,
         int(sys.argv[3]) if len(sys.argv) > 3 else 1, sys.argv[4])
