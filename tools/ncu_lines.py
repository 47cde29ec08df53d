#!/usr/bin/env python
"""per-source-line executed instructions and stall samples of an .ncu-rep
usage: ncu_lines.py rep file-suffix [top]"""
import csv, subprocess, sys, collections
rep, suffix = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
# sections: a file header row precedes each 'Line No' header
tot = 0
per = collections.OrderedDict()
cur_file = None
i = 0
while i < len(rows):
    r = rows[i]
    if r and r[0] == "Line No":
        h = r
        iI, iS = h.index("Instructions Executed"), h.index("# Samples")
        cols = {k: h.index(k) for k in ("stall_barrier", "stall_short_sb", "stall_long_sb", "stall_no_inst", "stall_wait")}
        fname = ""
        for b in range(i - 1, max(-1, i - 4), -1):
            if rows[b] and rows[b][0] == "File Path":
                fname = rows[b][1]
        j = i + 1
        while j < len(rows) and rows[j] and rows[j][0] != "Line No" and len(rows[j]) > iS:
            rr = rows[j]
            if rr[0] == "File Path":
                break
            try:
                ln = int(rr[0])
                ie = int(rr[iI] or 0)
                sm = int(rr[iS] or 0)
            except ValueError:
                j += 1
                continue
            if fname.endswith(suffix) or suffix == "*":
                d = per.setdefault((fname.split("/")[-1], ln), [rr[1], 0, 0, collections.Counter()])
                d[1] += ie
                d[2] += sm
                for k, c in cols.items():
                    try:
                        d[3][k] += int(rr[c] or 0)
                    except ValueError:
                        pass
            tot += ie
            j += 1
        i = j
    else:
        i += 1
print("total instructions", tot)
for (f, ln), (src, ie, sm, st) in sorted(per.items(), key=lambda kv: -kv[1][2])[:top]:
    print(f"{f}:{ln:5d} inst {ie:12d} samples {sm:7d} {dict(st)}  | {src.strip()[:90]}")
