"""N>1 host logic on CPU: two gloo ranks own contiguous cell ranges (like PETSc
DMDA ownership, pmc_subsurface_osrt.F90:349-350), step their shards, and
reduce the step flags with the collective that replaces the reference's
MPI_Allreduce(MAX)+MPI_Barrier (:381-383).  The per-shard step is played by the
oracle here (no GPU in this container); the GPU twin is exercised by bench.py
under torchrun."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pflotran_elm_interface_b200 import abi, rstep, workloads as W


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import sys
    sys.path.insert(0, os.path.dirname(__file__))
    import oracle_lib as orc

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    wl = W.by_name("c2", ncell=1001, tran_dt=3600.0)
    wl.cfg.c.maximum_reaction_iterations = 4   # forces cuts in part of the column
    wl.cfg.c.maximum_reaction_cuts = 2
    lo, hi = rstep.shard_range(wl.state.ncell, rank, world)
    shard = abi.HostState(wl.cfg, hi - lo)
    for k, v in wl.state.a.items():
        shard.a[k][...] = v[:, lo:hi]
    local = orc.rstep(wl.cfg, shard, wl.tran_dt, 1)
    red = rstep.reduce_results(local)
    q.put((rank, lo, hi, local.as_dict(), red.as_dict(), shard.a["pri_molal"].copy(), shard.a["ierror"].copy()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single_rank():
    import oracle_lib as orc

    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    outs = sorted([q.get(timeout=120) for _ in range(world)])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    wl = W.by_name("c2", ncell=1001, tran_dt=3600.0)
    wl.cfg.c.maximum_reaction_iterations = 4
    wl.cfg.c.maximum_reaction_cuts = 2
    ref = wl.state.copy()
    full = orc.rstep(wl.cfg, ref, wl.tran_dt, 1).as_dict()
    # ownership ranges tile the grid
    assert outs[0][1] == 0 and outs[0][2] == outs[1][1] and outs[1][2] == 1001
    # every rank sees the same reduced flags, equal to the single-rank run
    for _, _, _, _, red, _, _ in outs:
        for k in ("ncell_active", "sum_newton_iterations", "num_cut_cells", "max_newton_iterations",
                  "max_num_kinetic_state_updates", "rstep_error", "max_sub_steps"):
            assert red[k] == full[k], (k, red[k], full[k])
    assert full["num_cut_cells"] > 0
    # cells are independent: shard results equal the single-rank results bit for bit
    got = np.concatenate([o[5] for o in outs], axis=1)
    assert np.array_equal(got, ref.a["pri_molal"])
    assert np.array_equal(np.concatenate([o[6] for o in outs], axis=1), ref.a["ierror"])


def test_shard_range_covers_everything():
    for n in (0, 1, 7, 1000, 4194304):
        for w in (1, 2, 3, 8):
            r = [rstep.shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(w - 1))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1
