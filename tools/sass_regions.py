#!/usr/bin/env python
"""Static SASS instruction count per source routine of a specialised cubin
(nvdisasm -g line info -> enclosing function), plus the sum of the stall fields
(bits 105..108 of each instruction = cycles before the next issue) per routine: a
lower bound of the issue time of one warp running that code alone."""
import collections
import re
import subprocess
import sys

FN = re.compile(r"^\s*(?:S2_FN|S2_CE|static inline|static __device__|__device__|__host__|extern \"C\" __global__).*?(\w+)\s*\(")


def stalls(cubin):
    """address -> stall count from the raw encodings (cuobjdump -sass prints them)"""
    out = subprocess.run(["cuobjdump", "-sass", cubin], capture_output=True, text=True).stdout.split("\n")
    st = {}
    addr = None
    for ln in out:
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+.*/\* 0x([0-9a-f]{16}) \*/", ln)
        if m:
            addr = int(m.group(1), 16)
            continue
        m = re.match(r"\s+/\* 0x([0-9a-f]{16}) \*/", ln)
        if m and addr is not None:
            hi = int(m.group(1), 16)
            st[addr] = (hi >> 41) & 0xF
            addr = None
    return st


def main(cubin):
    sass = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
    st = stalls(cubin)
    cache = {}

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

    cur = None
    cnt, ops = collections.Counter(), collections.defaultdict(collections.Counter)
    cyc = collections.Counter()
    for line in sass.split("\n"):
        m = re.search(r'//## File "(.*?)", line (\d+)', line)
        if m:
            cur = region(m.group(1), int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(@!?U?P\d\s+)?([A-Z0-9_]+)", line)
        if m and cur:
            cnt[cur] += 1
            ops[cur][m.group(3)] += 1
            cyc[cur] += st.get(int(m.group(1), 16), 1)
    print("total", sum(cnt.values()), "instructions, stall-field sum", sum(cyc.values()))
    for k, v in cnt.most_common():
        print(f"{k:24s} {v:7d} {cyc[k]:7d}  " + " ".join(f"{o}:{n}" for o, n in ops[k].most_common(10)))


if __name__ == "__main__":
    main(sys.argv[1])
