#!/usr/bin/env python
"""Build libpfrx_b200.so in-tree with nvcc for sm_100a (no GPU needed).

One translation unit per padded system size N (pfrx_kern.cu with -DPFRX_N),
compiled in parallel, plus the host API (pfrx_api.cu).  The .so is written next
to the Python package so that it travels with the repo snapshot.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
OBJ = os.path.join(HERE, "_obj")
OUT = os.path.join(HERE, "..", "libpfrx_b200.so")

# padded size N -> lane counts instantiated (keep in sync with pfrx_api.cu)
VARIANTS = {
    3: [1, 4],
    4: [1, 4],
    8: [4],
    13: [4, 8, 16],
    15: [4, 8, 16],
    16: [16],
    32: [32],
}

NVCC = os.environ.get("NVCC", "nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def _newer(src_list, out):
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.getmtime(s) > t for s in src_list)


def _run(cmd, log):
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    with open(log, "w") as f:
        f.write(" ".join(cmd) + "\n" + p.stdout)
    if p.returncode != 0:
        sys.stderr.write(p.stdout)
        raise RuntimeError("nvcc failed: " + " ".join(cmd))
    return p.stdout


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    hdrs = [os.path.join(HERE, "pfrx_device.cuh"), os.path.join(HERE, "pfrx_types.cuh"),
            os.path.join(HERE, "pfrx_sandbox.cuh"), os.path.join(HERE, "..", "..", "include", "pfrx.h"),
            os.path.abspath(__file__)]
    jobs = []
    objs = []
    for n, lanes in VARIANTS.items():
        o = os.path.join(OBJ, f"kern_{n}.o")
        objs.append(o)
        src = os.path.join(HERE, "pfrx_kern.cu")
        if force or _newer(hdrs + [src, os.path.join(HERE, "pfrx_tpc.cuh")], o):
            defs = [f"-DPFRX_N={n}"] + [f"-DPFRX_L{i}={l}" for i, l in enumerate(lanes)]
            jobs.append((FLAGS_CMD(defs, src, o), os.path.join(OBJ, f"kern_{n}.log")))
    o = os.path.join(OBJ, "api.o")
    objs.append(o)
    src = os.path.join(HERE, "pfrx_api.cu")
    if force or _newer(hdrs + [src], o):
        jobs.append((FLAGS_CMD([], src, o), os.path.join(OBJ, "api.log")))
    if jobs:
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
            outs = list(ex.map(lambda j: _run(*j), jobs))
        if verbose:
            for t in outs:
                print(t)
    if jobs or not os.path.exists(OUT):
        _run([NVCC, "-shared", "-o", OUT] + objs + ["-ldl", "-lcudart"], os.path.join(OBJ, "link.log"))
    return OUT


def FLAGS_CMD(defs, src, o):
    return [NVCC] + FLAGS + defs + ["-c", src, "-o", o]


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
