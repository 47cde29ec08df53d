#!/usr/bin/env python
"""bench.py -- RStep cell-solves per second of the operator-split chemistry step.

    python bench.py --gpus N --steps K --warmup W            (our CUDA path)
    python bench.py --impl reference --gpus N --steps K ...  (CPU reference arm)

One "step" = one pass of the hot path (RStep over tran_dt) over one batch of
synthetic post-transport cell states.  The default workload is C3 of
BASELINE.json (Hanford U(VI), 256x256x64 = 4 194 304 cells, equilibrium surface
complexation + 2 kinetic minerals): the largest single-GPU configuration.  The
metric's C2 (10k-cell calcite column) cannot load a B200 (3 MB of state) and is
a parity-test case; `--workload c2` benches it anyway.

Multi-GPU: one rank per GPU (torchrun), cells sharded by contiguous ownership
ranges with no data-path collective; the only communication is the NCCL
allreduce of the step flags, enqueued by pfrx_rstep_async on the kernel stream
behind the kernel and therefore inside the timed region.  Scaling is WEAK: every
rank steps a full-size C3 shard.  At N >= 2 the line also carries `c5_baseline`:
BASELINE.json's 512 x 512 x 256 multi-mineral grid split over the N ranks
(STRONG scaling of a fixed 67 108 864-cell grid).

`e2e` is the call that replaces the cell loop of PMCSubsurfaceOSRTStepDT
(pmc_subsurface_osrt.F90:303-378): pfrx_os_step_host with HOST block vectors,
the chemistry state resident in device memory from step to step as rt_auxvars
are in the reference.  `e2e_full_state` moves the whole state (pfrx_rstep_host).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

DEFAULT_CELLS = {"c2": 10000, "c3": 256 * 256 * 64, "c3mr": 1 << 20, "c4": 2 * 1024 * 1024, "c4s": 2 * 1024 * 1024,
                 "c4se": 2 * 1024 * 1024, "c4fe": 2 * 1024 * 1024, "c5": 256 * 256 * 64, "c6": 1 << 20,
                 "c7": 1 << 20, "c8": 1 << 20}
DEFAULT_DT = {"c2": 3600.0, "c3": 3600.0, "c3mr": 3600.0, "c4": 1800.0, "c4s": 1800.0, "c4se": 1800.0, "c4fe": 1800.0,
              "c5": 86400.0, "c6": 86400.0, "c7": 86400.0, "c8": 86400.0}


def _kernel_name(info) -> str:
    if info["lanes"] < 0:
        return f"pfrx_spec_kernel[N={info['N']}]"
    if info["lanes"] == 0:
        return f"pfrx_rstep_tpc_kernel<{info['N']}>"
    return f"pfrx_rstep_kernel<{info['N']},{info['lanes']}>"


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=5)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="b200", choices=["b200", "reference"])
    p.add_argument("--workload", default="c3", choices=sorted(DEFAULT_CELLS))
    p.add_argument("--cells", type=int, default=None, help="cells per GPU")
    p.add_argument("--dt", type=float, default=None)
    p.add_argument("--kernel", default="auto", choices=["auto", "generic", "spec"],
                   help="auto: the network-specialised kernel when its cubin was built, else the generic one")
    p.add_argument("--no-e2e", action="store_true")
    p.add_argument("--no-cpu", action="store_true")
    p.add_argument("--cpu-seconds", type=float, default=12.0)
    return p.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return float(d.get("hbm_gbs", 6650.0)), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region"""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                 "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def bind_numa(local_rank: int) -> str:
    """pin this rank (and the pinned host buffers it allocates afterwards) to the CPUs next to its GPU"""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        n = os.cpu_count() or 1
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, (n + 63) // 64)
        cpus = [i for i in range(n) if (int(mask[i // 64]) >> (i % 64)) & 1]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return f"{len(cpus)} cpus near gpu {local_rank}"
    except Exception as e:  # diagnostics only
        return f"not bound ({type(e).__name__})"
    return "not bound"


class _Sample:
    """what cpu_baseline needs of a workload: cfg, tran_dt, name and a host state"""

    def __init__(self, wl, n):
        from pflotran_elm_interface_b200 import abi

        self.cfg, self.tran_dt, self.name = wl.cfg, wl.tran_dt, wl.name
        n = int(min(n, wl.state.ncell))
        self.state = abi.HostState(wl.cfg, n)
        for k, v in wl.state.a.items():
            self.state.a[k][...] = v[:, :n]


def cpu_baseline(wl, seconds, threads):
    """the oracle (a CPU port of the reference path: `kind` = port) on a bounded
    sample of the same workload, all host threads"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as orc
    from pflotran_elm_interface_b200 import abi

    def sample(n):
        st = abi.HostState(wl.cfg, n)
        for k, v in wl.state.a.items():
            st.a[k][...] = v[:, :n]
        return st

    probe_n = min(wl.state.ncell, 2048)
    st = sample(probe_n)
    t0 = time.perf_counter()
    orc.rstep(wl.cfg, st, wl.tran_dt, threads)
    rate = probe_n / max(time.perf_counter() - t0, 1e-6)
    n = int(min(wl.state.ncell, max(probe_n, rate * seconds)))
    st = sample(n)
    t0 = time.perf_counter()
    res = orc.rstep(wl.cfg, st, wl.tran_dt, threads)
    dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": "cell-solves/s", "cores": threads, "kind": "port",
            "sample": f"first {n} cells of {wl.name} (same seeded states), {dt:.1f} s wall, "
                      f"{res.sum_newton_iterations / max(1, res.ncell_active):.2f} Newton its/cell",
            "per_core": n / dt / threads, "seconds": dt, "cells_stepped": n}


def parity_sample(wl_cfg, pristine, work, tran_dt, nsample=8192, seed=20261018):
    """the launch that was just timed against the oracle on a random sample of its cells: inputs
    from the pristine copy, outputs from the stepped state"""
    import torch

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as orc
    from pflotran_elm_interface_b200 import abi

    n = pristine.ncell
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    idx = torch.randperm(n, generator=g)[: min(nsample, n)].sort().values.to(pristine.device)
    ref = abi.HostState(wl_cfg, idx.numel())
    got = abi.HostState(wl_cfg, idx.numel())
    for k in ref.a:
        ref.a[k][...] = pristine.t[k][:, idx].cpu().numpy()
        got.a[k][...] = work.t[k][:, idx].cpu().numpy()
    orc.rstep(wl_cfg, ref, tran_dt, os.cpu_count() or 1)
    counts_equal = all(bool(np.array_equal(ref.a[f], got.a[f])) for f in abi.STATE_RESULT_FIELDS)
    worst, where = 0.0, None
    for f in ("total", "pri_molal", "immobile", "mnrl_volfrac", "total_sorb_eq", "sec_molal"):
        a, b = ref.a[f], got.a[f]
        if a.size == 0:
            continue
        scale = np.maximum(np.abs(a), np.abs(b))
        err = np.where(scale < 1e-30, 0.0, np.abs(a - b) / np.maximum(scale, 1e-300))
        if err.max() > worst:
            worst, where = float(err.max()), f
    return {"cells": int(idx.numel()), "max_rel_err": worst, "field": where, "counts_equal": bool(counts_equal),
            "note": "random cells of the last timed launch vs the CPU oracle on the same inputs (tolerance 1e-10)"}


def c5_baseline(step_c3, work_c3, pristine_c3, rank, world, local_rank, dev, total_cells=512 * 512 * 256, small=1 << 20,
                steps=2, warmup=1):
    import torch
    import torch.distributed as dist
    from pflotran_elm_interface_b200 import rstep, workloads

    # free the C3 shard first
    step_c3.close()
    for d in (work_c3, pristine_c3):
        d.t.clear()
    torch.cuda.empty_cache()
    n_rank = total_cells // world
    dt = DEFAULT_DT["c5"]
    wl = workloads.by_name("c5", ncell=small, tran_dt=dt)
    step = rstep.ChemistryStep(wl.cfg, local_rank)
    step.init_comm()
    step.specialize(required=True)
    tile = rstep.DeviceState.from_host(wl.state, dev)
    shard = rstep.DeviceState(wl.cfg, n_rank, dev)

    def retile():
        for k, t in shard.t.items():
            src = tile.t[k]
            for c0 in range(0, n_rank, small):
                c1 = min(n_rank, c0 + small)
                t[:, c0:c1].copy_(src[:, : c1 - c0])
        torch.cuda.synchronize(dev)

    step.bind(shard)
    ks = torch.cuda.ExternalStream(step.stream_ptr, device=dev)
    ms = []
    res = None
    for i in range(warmup + steps):
        retile()
        dist.barrier()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ks)
        step.rstep_async(dt)
        e1.record(ks)
        res = step.allreduce(step.rstep_finish())
        if i >= warmup:
            ms.append(e0.elapsed_time(e1))
    t = torch.tensor([float(np.sum(ms)) * 1e-3], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    t_max = float(t.item())
    out = {"workload": "c5", "grid": "512x512x256", "cells": int(res.ncell_active), "cells_per_gpu": n_rank,
           "scaling": "strong", "steps": steps, "warmup": warmup, "ms_per_step": 1000.0 * t_max / steps,
           "value": int(res.ncell_active) * steps / t_max, "unit": "cell-solves/s", "tran_dt_s": dt,
           "kernel_variant": step.variant, "newton_its_per_cell": res.sum_newton_iterations / max(1, res.ncell_active),
           "inputs": f"shard tiled on the device from {small} seeded cells, re-tiled between steps",
           "timing": "CUDA events on the kernel stream around kernel + NCCL reduction, max over ranks"}
    step.close()
    return out


_REAL_STDOUT = None


def emit(obj) -> None:
    """the ONE JSON line of the contract, on the process's real stdout"""
    line = (json.dumps(obj) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(line.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, line)


def main():
    global _REAL_STDOUT
    a = parse()
    # libraries write to fd 1 (NCCL prints its version banner there at communicator set-up): point
    # fd 1 at stderr for the run and keep the real stdout for the JSON line
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    ncell = a.cells or DEFAULT_CELLS[a.workload]
    dt = a.dt or DEFAULT_DT[a.workload]
    from pflotran_elm_interface_b200 import workloads

    cfg_json = {"workload": a.workload, "cells_per_gpu": ncell, "tran_dt_s": dt,
                "inputs": "larger than L2; state restored from a pristine HBM copy between steps",
                "parallelism": f"cells sharded over {world} rank(s), no halo"}

    # ------------------------------------------------------------------ CPU arm
    if a.impl == "reference":
        if rank != 0:
            return 0
        threads = os.cpu_count() or 1
        wl = workloads.by_name(a.workload, ncell=min(ncell, 1 << 18), tran_dt=dt)
        wl.name = workloads.by_name(a.workload, ncell=8, tran_dt=dt).name
        cfg_json["cells_generated"] = int(wl.state.ncell)
        cfg_json["note"] = ("the CPU arm steps a bounded sample per step (cells_stepped_per_step) of states drawn "
                            "like the GPU arm's; throughput is per cell, the grid size is nominal")
        per_step = max(2.0, min(a.cpu_seconds, 120.0 / max(1, a.steps + a.warmup)))
        vals = []
        for i in range(a.warmup + a.steps):
            r = cpu_baseline(wl, per_step, threads)
            if i >= a.warmup:
                vals.append(r)
        v = float(np.mean([r["value"] for r in vals]))
        ms = 1000.0 * float(np.mean([r["seconds"] for r in vals]))
        out = {"impl": "reference", "metric": "rstep_cell_solves_per_sec", "value": v, "unit": "cell-solves/s",
               "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True,
               "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg_json,
               "cpu_baseline": {"value": v, "unit": "cell-solves/s", "cores": threads, "kind": "port",
                                "sample": vals[-1]["sample"], "cells_stepped": vals[-1]["cells_stepped"]},
               "cells_stepped_per_step": vals[-1]["cells_stepped"],
               "e2e": {"value": v, "unit": "cell-solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
               "note": "reference Fortran cannot be built in this image (no Fortran compiler/PETSc/MPI); this arm "
                       "times the C oracle port of the same path with back-substitution enabled (SURVEY 0.2)"}
        emit(out)
        return 0

    # ------------------------------------------------------------------ GPU arm
    import torch
    import torch.distributed as dist
    from pflotran_elm_interface_b200 import abi, rstep

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback)"
    if world > 1:
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    numa = bind_numa(local_rank)
    wl = workloads.by_name(a.workload, ncell=ncell, tran_dt=dt)
    cpu_sample = _Sample(wl, 1 << 18) if (rank == 0 and not a.no_cpu) else None
    step = rstep.ChemistryStep(wl.cfg, local_rank)
    step.init_comm()
    if a.kernel != "generic":
        step.specialize(required=(a.kernel == "spec"))
    info = step.kernel_info()
    pristine = rstep.DeviceState.from_host(wl.state, dev)
    work = rstep.DeviceState(wl.cfg, ncell, dev)
    step.bind(work)
    tuned = None
    if a.kernel == "auto" and info["lanes"] < 0 and not os.environ.get("PFRX_SPEC_VARIANT"):
        # untimed set-up: time both skeletons of the specialised kernel on a copy of the
        # first cells of this shard and keep the faster one (rstep.ChemistryStep.autotune)
        tuned = step.autotune(pristine, dt)
        step.bind(work)
        info = step.kernel_info()
    kstream = torch.cuda.ExternalStream(step.stream_ptr, device=dev)

    def restore():
        for k in work.t:
            work.t[k].copy_(pristine.t[k])
        torch.cuda.synchronize(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    res = None
    for _ in range(a.warmup):
        restore()
        res = step.allreduce(step.rstep(dt))
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    launches0 = step.launch_count
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
    kern_ms = []
    wall0 = time.perf_counter()
    for i in range(a.steps):
        restore()
        barrier()
        ev[i][0].record(kstream)
        step.rstep_async(dt)     # kernel + (N > 1) the NCCL reduction of the shard summaries, same stream
        ev[i][1].record(kstream)
        res = step.rstep_finish()
        res = step.allreduce(res)  # N > 1: reads back what the stream already reduced
        kern_ms.append(ev[i][0].elapsed_time(ev[i][1]))
    barrier()
    wall = time.perf_counter() - wall0
    launches = step.launch_count - launches0
    clocks = sampler.stop() if rank == 0 else None
    t_local = float(np.sum(kern_ms)) * 1e-3
    if world > 1:
        t = torch.tensor([t_local], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_max = float(t.item())
    else:
        t_max = t_local
    total_cells = int(res.ncell_active) * a.steps
    value = total_cells / t_max
    ms_per_step = 1000.0 * t_max / a.steps
    parity = None
    if rank == 0 and not a.no_cpu:
        parity = parity_sample(wl.cfg, pristine, work, dt)

    # ---- e2e through the C ABI with host buffers (H2D + kernel + D2H) ----------
    e2e = None
    if not a.no_e2e:
        host = rstep.PinnedHostState(wl.cfg, ncell)
        rows = wl.cfg.field_rows()
        h2d = 8 * sum(rows[f] for f in abi.STATE_DOUBLE_FIELDS if f != "eqsrfcplx_conc") * ncell + 4 * ncell
        d2h = 8 * sum(rows[f] for f in abi.STATE_IO_FIELDS) * ncell + 16 * ncell
        ne = max(2, min(a.steps, 3))
        if world > 1:
            # N ranks on one box: keep ONE host copy of the shard per rank (the pinned
            # one) and refresh it from the pristine device copy between steps
            wl.state.a.clear()

        def refresh():
            for k, v in host.a.items():
                if k in pristine.t:
                    torch.from_numpy(v).copy_(pristine.t[k])
            torch.cuda.synchronize(dev)

        # what the step derives and only the next step reads stays in the device mirror
        resident = [f for f in ("sec_molal", "sec_act_coef", "pri_act_coef") if rows[f]]
        if wl.cfg.c.act_coef_update_frequency != 2:   # frozen coefficients are inputs the caller owns
            resident = [f for f in resident if f == "sec_molal"]
        step.rstep_host_resident(resident)
        refresh()
        step.rstep_host(host, dt)  # warm-up (allocates the device mirror)
        tt = []
        for _ in range(ne):
            refresh()
            barrier()
            t0 = time.perf_counter()
            r2 = step.rstep_host(host, dt)
            r2 = step.allreduce(r2)
            torch.cuda.synchronize(dev)
            tt.append(time.perf_counter() - t0)
        te = float(np.mean(tt))
        h2d, d2h = step.last_transfer_bytes()  # counted by the library from the copies it issued
        if world > 1:
            t = torch.tensor([te], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            te = float(t.item())
        e2e = {"value": int(r2.ncell_active) / te, "unit": "cell-solves/s", "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "ms_per_step": 1000.0 * te, "resident_fields": resident,
               "api": "pfrx_rstep_host (C ABI, pinned host SoA buffers: the whole rt_auxvar state crosses the link; "
                      "derived fields resident on the device via pfrx_rstep_host_resident)"}
        step.rstep_host_resident([])

    # ---- the same step as PMCSubsurfaceOSRT sees it: chemistry state resident in HBM, only the
    # PETSc block vectors (solved totals in, tran_xx in/out) cross the host link every step ------
    e2e_os = None
    if not a.no_e2e:
        naq_, ncomp_ = int(wl.cfg.c.naqcomp), int(wl.cfg.ncomp)
        h_solved = torch.empty((ncell, ncomp_), dtype=torch.float64).pin_memory()
        h_xx = torch.zeros((ncell, ncomp_), dtype=torch.float64).pin_memory()
        h_solved.zero_()
        h_solved[:, :naq_].copy_(pristine.t["total"].t())
        if ncomp_ > naq_:
            h_xx[:, naq_:].copy_(pristine.t["immobile"].t())
        # what the host link gives THIS rank while every rank uses it: 256 MiB up and 256 MiB down at the same
        # time from pinned memory, all ranks between the same barriers (the floor of any host-vector e2e)
        link = None
        try:
            nb = 1 << 28
            hp_u, hp_d = torch.empty(nb, dtype=torch.uint8).pin_memory(), torch.empty(nb, dtype=torch.uint8).pin_memory()
            dv_u, dv_d = torch.empty(nb, dtype=torch.uint8, device=dev), torch.zeros(nb, dtype=torch.uint8, device=dev)
            s_u, s_d = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
            best = None
            for _ in range(3):
                barrier()
                t0 = time.perf_counter()
                with torch.cuda.stream(s_u):
                    dv_u.copy_(hp_u, non_blocking=True)
                with torch.cuda.stream(s_d):
                    hp_d.copy_(dv_d, non_blocking=True)
                torch.cuda.synchronize(dev)
                tl = time.perf_counter() - t0
                best = tl if best is None else min(best, tl)
            if world > 1:
                t = torch.tensor([best], dtype=torch.float64, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                best = float(t.item())
            link = {"GBps_each_direction_per_rank": nb / best / 1e9, "ranks_at_once": world,
                    "note": "256 MiB up and 256 MiB down concurrently from pinned host memory on every rank, slowest rank"}
            del hp_u, hp_d, dv_u, dv_d
        except RuntimeError:
            link = None
        restore()
        for _ in range(4):                          # warm-up: device staging (one chunk, many chunks) and the
            restore()                               # library's two timed chunking trials
            step.os_step_host(h_solved, h_xx, dt)
        tt = []
        for _ in range(max(2, min(a.steps, 3))):
            restore()
            barrier()
            t0 = time.perf_counter()
            r3 = step.os_step_host(h_solved, h_xx, dt)
            r3 = step.allreduce(r3)
            torch.cuda.synchronize(dev)
            tt.append(time.perf_counter() - t0)
        te = float(np.mean(tt))
        h2d3, d2h3 = step.last_transfer_bytes()
        if world > 1:
            t = torch.tensor([te], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            te = float(t.item())
        e2e_os = {"value": int(r3.ncell_active) / te, "unit": "cell-solves/s", "h2d_bytes_per_step": int(h2d3),
                  "d2h_bytes_per_step": int(d2h3), "ms_per_step": 1000.0 * te,
                  "sum_newton_iterations": int(r3.sum_newton_iterations),
                  "host_link": link,
                  "link_floor_ms": (None if not link else
                                    1000.0 * max(int(h2d3), int(d2h3)) / (link["GBps_each_direction_per_rank"] * 1e9)),
                  "api": "pfrx_os_step_host (C ABI): pinned host block vectors solved_total / tran_xx, "
                         "rt_auxvar state bound in device memory between steps (pmc_subsurface_osrt.F90:303-378)"}
        # the same step for a host whose PETSc vectors live in device memory (VECCUDA): pfrx_os_load, pfrx_rstep,
        # pfrx_os_store on device pointers -- the C-ABI path without the host link (not the contract's e2e: no copies)
        try:
            d_solved, d_xx = h_solved.to(dev), h_xx.to(dev)
            td = []
            for k in range(2 + max(2, min(a.steps, 3))):
                restore()
                barrier()
                t0 = time.perf_counter()
                step.os_load(d_solved, d_xx if ncomp_ > naq_ else None)
                r4 = step.rstep(dt)
                step.os_store(d_xx)
                r4 = step.allreduce(r4)
                torch.cuda.synchronize(dev)
                if k >= 2:
                    td.append(time.perf_counter() - t0)
            tdm = float(np.mean(td))
            if world > 1:
                t = torch.tensor([tdm], dtype=torch.float64, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                tdm = float(t.item())
            e2e_os["device_vectors"] = {"ms_per_step": 1000.0 * tdm, "value": int(r4.ncell_active) / tdm,
                                        "api": "pfrx_os_load + pfrx_rstep + pfrx_os_store on device block vectors "
                                               "(wall clock around the three C-ABI calls and the reduction)"}
            del d_solved, d_xx
        except RuntimeError:
            pass
        del h_solved, h_xx

    # ---- BASELINE.json's fifth configuration at its full size: 512 x 512 x 256 cells of the Hanford
    # basis with six kinetic minerals, split over the ranks (STRONG scaling).  The shard is tiled on the
    # device from a 2^20-cell host state and re-tiled between steps: no pristine copy of the shard.
    bytes_per_cell, variant_name = step.bytes_per_cell, step.variant
    c5 = None
    if world >= 2 and a.workload == "c3" and not os.environ.get("PFRX_BENCH_NO_C5"):
        try:
            c5 = c5_baseline(step, work, pristine, rank, world, local_rank, dev)
        except Exception as e:  # the headline line must survive a failure here
            c5 = {"error": f"{type(e).__name__}: {e}"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant (only) kernel --------------------------------
    hbm_gbs, which = peaks()
    fp64_tf, mhz = rstep.fp64_peak_tflops(local_rank)
    f_eval, f_solve = workloads.flops_model(wl.net)
    cells_local = ncell
    its_local = res.sum_newton_iterations / max(1, world)        # per rank (weak scaling: equal shards)
    subs = cells_local                                          # >= one sub-step per cell
    flops_closed = its_local * f_eval + max(0.0, its_local - subs) * f_solve
    # algorithmic flops: COUNTED by the op-counting build of the oracle on a sample of the same cells
    # (add/sub/mul/div/compare = 1, transcendental = 20), scaled by the launch's Newton iterations;
    # the closed form of SURVEY 8(d) is kept beside it
    flops_launch, flops_src = flops_closed, "closed form (workloads.flops_model)"
    if cpu_sample is not None:
        try:
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            import oracle_lib as orc

            ns = min(4096, cpu_sample.state.ncell)
            smp = abi.HostState(wl.cfg, ns)
            for k, v in cpu_sample.state.a.items():
                smp.a[k][...] = v[:, :ns]
            rc, ops = orc.count_ops(wl.cfg, smp, dt, os.cpu_count() or 1)
            ops_per_it = ops / max(1, rc.sum_newton_iterations)
            flops_launch = ops_per_it * its_local
            flops_src = (f"oracle op-counter (oracle/pfrx_oracle_count.cpp) on the first {ns} cells: "
                         f"{ops_per_it:.0f} flop per Newton iteration incl. its share of the solves")
        except Exception as e:  # the closed form stays
            flops_src += f" (op-counter unavailable: {type(e).__name__})"
    bytes_launch = bytes_per_cell * cells_local
    t_launch = t_max / a.steps
    ach_tf = flops_launch / t_launch / 1e12
    ach_gbs = bytes_launch / t_launch / 1e9
    frac_fp64 = ach_tf / fp64_tf if fp64_tf > 0 else None
    frac_hbm = ach_gbs / hbm_gbs
    if frac_fp64 is not None and frac_fp64 >= frac_hbm:
        roof = {"bound": "fp64", "achieved": ach_tf, "peak": fp64_tf, "unit": "TFLOP/s", "frac": frac_fp64,
                "peak_source": "DFMA micro-benchmark in this run (pfrx_diag_fp64_peak)"}
    else:
        roof = {"bound": "hbm", "achieved": ach_gbs, "peak": hbm_gbs, "unit": "GB/s", "frac": frac_hbm,
                "peak_source": f"MEASURED_PEAKS.json ({which})"}
    # DRAM bytes of the dominant kernel from the committed ncu capture, scaled to this launch
    traffic, traffic_src, ncu_counters = None, None, None
    try:
        with open(os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")) as f:
            tr = json.load(f).get(a.workload)
        if tr and tr["kernel"] == _kernel_name(info):
            traffic = (tr["dram_read_bytes"] + tr["dram_write_bytes"]) / tr["cells"] * cells_local
            traffic_src = tr["source"]
            if "fp64_pipe_busy_pct" in tr:
                # what the hardware did in that capture, next to the algorithmic numerator above: the share of
                # cycles the FP64 pipe was busy and the issue slots used (a frac above 1 means the kernel reaches
                # the reference's result with fewer operations than the reference algorithm counts)
                ncu_counters = {"fp64_pipe_busy_pct": tr["fp64_pipe_busy_pct"], "issue_active_pct": tr["issue_active_pct"],
                                "warp_instructions_per_cell_iteration": tr["warp_instructions"] / tr["newton_iterations"],
                                "source": tr["source"]}
    except (OSError, ValueError, KeyError):
        pass
    roof.update({"traffic": traffic, "traffic_source": traffic_src, "frac_fp64": frac_fp64, "frac_hbm": frac_hbm,
                 "algorithmic_flops_per_launch": flops_launch, "algorithmic_flops_source": flops_src,
                 "algorithmic_flops_closed_form": flops_closed, "counted_over_closed_form": flops_launch / flops_closed,
                 "algorithmic_bytes_per_launch": bytes_launch,
                 "flops_per_newton_iteration": f_eval + f_solve, "bytes_per_cell": bytes_per_cell,
                 "newton_its_per_cell": res.sum_newton_iterations / max(1, res.ncell_active),
                 "kernel": _kernel_name(info), "kernel_ms": 1000.0 * t_launch,
                 "fp64_peak_sm_mhz": mhz, "ncu": ncu_counters})

    # ---- the block-vector transposes either side of the cell loop (HBM-bound) ------
    osv = None
    if world == 1 and not a.no_e2e:
        naq, ncomp = int(wl.cfg.c.naqcomp), int(wl.cfg.ncomp)
        vec = torch.zeros((ncell, ncomp), dtype=torch.float64, device=dev)
        vec2 = torch.rand((ncell, ncomp), dtype=torch.float64, device=dev)
        restore()
        algo = {"fixed_accum": 8 * (2 * naq + 3), "load": 16 * ncomp, "store": 16 * ncomp}
        calls = {"fixed_accum": lambda: step.os_fixed_accum(vec), "load": lambda: step.os_load(vec2, vec2),
                 "store": lambda: step.os_store(vec)}
        osv = {}
        for nm in ("fixed_accum", "store", "load"):
            calls[nm]()
            ts = []
            for _ in range(5):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(kstream)
                calls[nm]()
                e1.record(kstream)
                torch.cuda.synchronize(dev)
                ts.append(e0.elapsed_time(e1) * 1e-3)
            t = float(np.median(ts))
            gbs = algo[nm] * ncell / t / 1e9
            osv[nm] = {"ms": 1e3 * t, "algorithmic_bytes_per_cell": algo[nm], "GB/s": gbs, "frac_hbm": gbs / hbm_gbs}
        osv["note"] = ("pfrx_os_fixed_accum / pfrx_os_store / pfrx_os_load: PETSc block vectors <-> SoA "
                       "(pmc_subsurface_osrt.F90:260-274, 303-376), vectors larger than L2")
        del vec, vec2

    cpu = None
    if not a.no_cpu:
        cpu = cpu_baseline(cpu_sample, a.cpu_seconds, os.cpu_count() or 1)

    out = {"metric": "rstep_cell_solves_per_sec", "value": value, "unit": "cell-solves/s", "n_gpus": world,
           "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": dict(cfg_json, name=wl.name, ncomp=wl.cfg.ncomp, neqcplx=int(wl.cfg.c.neqcplx),
                          kernel=info, kernel_variant=variant_name, autotune_s=tuned, note=wl.note,
                          cell_order=(None if not (variant_name or "").startswith(("q", "w", "p"))
                                      or os.environ.get("PFRX_CELL_ORDER") == "0" else
                                      "refill kernel: cells handed out by the previous launch's Newton counts, longest "
                                      "first (pfrx_cell_order); every timed step solves the same restored state, so "
                                      "here the prediction is exact; the sort runs inside the timed region")),
           "gpu_launches": int(launches), "clocks": clocks, "e2e": e2e_os if e2e_os is not None else e2e,
           "e2e_full_state": e2e, "parity_sample": parity, "c5_baseline": c5, "numa": numa,
           "roofline": roof, "cpu_baseline": cpu,
           "os_block_vectors": osv,
           "result": res.as_dict(), "wall_s_timed_region": wall}
    emit(out)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
