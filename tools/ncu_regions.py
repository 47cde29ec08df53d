#!/usr/bin/env python
"""Stall samples and executed instructions of an .ncu-rep per source routine of the cubin it
profiled (ncu --page source gives per-SASS-address samples; nvdisasm -g maps addresses to
source lines; the enclosing function is found like tools/sass_regions.py does).
usage: ncu_regions.py rep cubin [warp_iterations]"""
import collections
import csv
import re
import subprocess
import sys

sys.path.insert(0, __file__.rsplit("/", 1)[0])
from sass_regions import FN  # noqa: E402


def addr_regions(cubin):
    sass = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
    cache, reg, cur = {}, {}, None

    def region(f, l):
        if f not in cache:
            try:
                cache[f] = open(f).read().split("\n")
            except OSError:
                cache[f] = None
        text = cache[f]
        if text is None or not (f.endswith(".cu") or f.endswith(".cuh")):
            return f.split("/")[-1]
        for i in range(min(l, len(text)) - 1, -1, -1):
            if "__global__" in text[i]:
                return "kernel"
            m = FN.search(text[i])
            if m:
                return m.group(1)
        return "?"

    for line in sass.split("\n"):
        m = re.search(r'//## File "(.*?)", line (\d+)', line)
        if m:
            cur = region(m.group(1), int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+", line)
        if m and cur:
            reg[int(m.group(1), 16)] = cur
    return reg


def main(rep, cubin, witer=None):
    reg = addr_regions(cubin)
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h = rows[1]
    iA, iN, iI = h.index("Address"), h.index("# Samples"), h.index("Instructions Executed")
    stalls = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
    idx = {c: h.index(c) for c in stalls}
    base = int(rows[2][iA], 16)
    agg = collections.defaultdict(collections.Counter)
    for r in rows[2:]:
        if len(r) <= iI:
            continue
        k = reg.get(int(r[iA], 16) - base, "?")
        agg[k]["samples"] += int(r[iN] or 0)
        agg[k]["inst"] += int(r[iI] or 0)
        for c in stalls:
            agg[k][c] += int(r[idx[c]] or 0)
    tot = sum(v["samples"] for v in agg.values())
    toti = sum(v["inst"] for v in agg.values())
    print(f"total samples {tot}, warp instructions executed {toti}" + (f", per warp-iteration {toti / witer:.0f}" if witer else ""))
    show = ["stall_long_sb", "stall_wait", "stall_selected", "stall_not_selected", "stall_math", "stall_short_sb",
            "stall_no_inst", "stall_barrier", "stall_lg", "stall_mio", "stall_dispatch", "stall_branch_resolving"]
    print(f"{'routine':24s} {'samples':>8s} {'%':>6s} {'instr':>11s} " + " ".join(f"{c[6:][:8]:>8s}" for c in show))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1]["samples"]):
        if v["samples"] < tot * 0.002:
            continue
        print(f"{k:24s} {v['samples']:8d} {100.0 * v['samples'] / tot:6.1f} {v['inst']:11d} " +
              " ".join(f"{v[c]:8d}" for c in show))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], float(sys.argv[3]) if len(sys.argv) > 3 else None)
