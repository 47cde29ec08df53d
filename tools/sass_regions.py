#!/usr/bin/env python
"""Static SASS instruction count per source routine of a specialised cubin
(nvdisasm -g line info -> enclosing __device__ function)."""
import collections
import re
import subprocess
import sys


def main(cubin):
    sass = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
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
            m = re.search(r"__device__ .*?(\w+)\(", text[i])
            if m:
                return m.group(1)
            if "__global__" in text[i]:
                return "kernel"
        return "?"

    cur = None
    cnt, ops = collections.Counter(), collections.defaultdict(collections.Counter)
    for line in sass.split("\n"):
        m = re.search(r'//## File "(.*?)", line (\d+)', line)
        if m:
            cur = region(m.group(1), int(m.group(2)))
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(@!?U?P\d\s+)?([A-Z0-9_]+)", line)
        if m and cur:
            cnt[cur] += 1
            ops[cur][m.group(2)] += 1
    print("total", sum(cnt.values()))
    for k, v in cnt.most_common():
        print(f"{k:24s} {v:7d}  " + " ".join(f"{o}:{n}" for o, n in ops[k].most_common(9)))


if __name__ == "__main__":
    main(sys.argv[1])
