"""Host-side mirror of the reference's operator-split reaction step on top of
the C ABI (``include/pfrx.h`` -> ``libpfrx_b200.so``).

``ChemistryStep.rstep`` is the cell loop of ``PMCSubsurfaceOSRTStepDT``
(src/pflotran/pmc_subsurface_osrt.F90:346-383): every bound cell goes through
``RStep`` (src/pflotran/reaction.F90:3564) over ``tran_dt``; the return value
carries what the coupler accumulates after the loop (:364-388).  There is no
CPU fallback: if the CUDA library is missing or no device is present the calls
raise.

PyTorch is used only for device memory, streams and ``torch.distributed``.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, Optional

import numpy as np

from . import abi

_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libpfrx_b200.so")
_lib = None


class PfrxError(RuntimeError):
    pass


def lib():
    """load libpfrx_b200.so (built in-tree by csrc/build.py); loud on failure"""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise PfrxError(f"{_LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(there is no CPU fallback for the chemistry step)")
    L = C.CDLL(_LIB_PATH)
    cfgp, stp, resp = C.POINTER(abi.PfrxConfig), C.POINTER(abi.PfrxState), C.POINTER(abi.PfrxStepResult)
    hp = C.c_void_p
    L.pfrx_abi_version.restype = C.c_int
    L.pfrx_last_error.restype = C.c_char_p
    L.pfrx_sizeof.argtypes = [C.c_int]
    L.pfrx_sizeof.restype = C.c_int64
    L.pfrx_create.argtypes = [cfgp, C.c_int, C.POINTER(hp)]
    L.pfrx_destroy.argtypes = [hp]
    L.pfrx_destroy.restype = None
    L.pfrx_bind_state.argtypes = [hp, C.c_int64, stp]
    L.pfrx_rstep_async.argtypes = [hp, C.c_double]
    L.pfrx_rstep_finish.argtypes = [hp, resp]
    L.pfrx_rstep.argtypes = [hp, C.c_double, resp]
    L.pfrx_rstep_host.argtypes = [hp, C.c_int64, stp, C.c_double, resp]
    L.pfrx_comm_unique_id.argtypes = [C.c_void_p]
    L.pfrx_comm_init.argtypes = [hp, C.c_int, C.c_int, C.c_void_p]
    L.pfrx_allreduce.argtypes = [hp, resp]
    L.pfrx_stream.argtypes = [hp]
    L.pfrx_stream.restype = C.c_void_p
    L.pfrx_cell_order.argtypes = [hp, C.c_int]
    L.pfrx_launch_count.argtypes = [hp]
    L.pfrx_launch_count.restype = C.c_int64
    L.pfrx_bytes_per_cell.argtypes = [hp]
    L.pfrx_bytes_per_cell.restype = C.c_int64
    L.pfrx_kernel_info.argtypes = [hp, C.POINTER(C.c_int)]
    L.pfrx_reaction.argtypes = [hp, C.c_double, C.c_int, C.c_void_p, C.c_void_p]
    L.pfrx_update_auxvars.argtypes = [hp, C.c_void_p, C.c_int]
    L.pfrx_kinmr_checkpoint_rows.argtypes = [C.POINTER(abi.PfrxConfig), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    L.pfrx_equilibrate_constraint.argtypes = [hp, C.POINTER(abi.PfrxConstraint), C.c_void_p, C.c_void_p, C.c_void_p]
    L.pfrx_os_fixed_accum.argtypes = [hp, C.c_void_p]
    L.pfrx_os_load.argtypes = [hp, C.c_void_p, C.c_void_p]
    L.pfrx_os_store.argtypes = [hp, C.c_void_p]
    L.pfrx_os_step_host.argtypes = [hp, C.c_void_p, C.c_void_p, C.c_double, resp]
    L.pfrx_config_dump.argtypes = [hp, C.c_char_p]
    L.pfrx_config_write.argtypes = [cfgp, C.c_char_p]
    L.pfrx_config_signature_of.argtypes = [cfgp]
    L.pfrx_config_signature_of.restype = C.c_uint64
    L.pfrx_rstep_host_resident.argtypes = [hp, C.c_uint64]
    L.pfrx_rstep_host_fetch.argtypes = [hp, C.c_int64, stp]
    L.pfrx_last_transfer_bytes.argtypes = [hp, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    L.pfrx_load_specialized.argtypes = [hp, C.c_char_p]
    L.pfrx_config_signature.argtypes = [hp]
    L.pfrx_config_signature.restype = C.c_uint64
    L.pfrx_diag_fp64_peak.argtypes = [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    if L.pfrx_abi_version() != abi.PFRX_ABI_VERSION:
        raise PfrxError("libpfrx_b200.so ABI version mismatch")
    if (L.pfrx_sizeof(0) != C.sizeof(abi.PfrxConfig) or L.pfrx_sizeof(1) != C.sizeof(abi.PfrxState)
            or L.pfrx_sizeof(2) != C.sizeof(abi.PfrxStepResult)):
        raise PfrxError("ctypes struct layout does not match include/pfrx.h")
    _lib = L
    return L


def _check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().pfrx_last_error().decode("utf-8", "replace")
        raise PfrxError(f"{what} failed (code {rc}): {msg}")


class DeviceState:
    """Cell-major SoA state resident in HBM: one torch tensor ``[rows, ncell]``
    per field of ``pfrx_state``."""

    def __init__(self, cfg: abi.ReactionConfig, ncell: int, device):
        import torch

        self.cfg = cfg
        self.ncell = int(ncell)
        self.device = torch.device(device)
        rows = cfg.field_rows()
        self.t: Dict[str, "torch.Tensor"] = {}
        for f in abi.STATE_DOUBLE_FIELDS + abi.STATE_ELM_FIELDS:
            self.t[f] = torch.zeros((rows[f], self.ncell), dtype=torch.float64, device=self.device)
        for f in abi.STATE_INT_FIELDS:
            self.t[f] = torch.zeros((rows[f], self.ncell), dtype=torch.int32, device=self.device)

    @classmethod
    def from_host(cls, host: abi.HostState, device) -> "DeviceState":
        import torch

        o = cls(host.cfg, host.ncell, device)
        for k, v in host.a.items():
            o.t[k].copy_(torch.from_numpy(v))
        return o

    def load(self, host: abi.HostState) -> None:
        import torch

        for k, v in host.a.items():
            self.t[k].copy_(torch.from_numpy(v), non_blocking=True)

    def prefix(self, ncell: int) -> "DeviceState":
        """an independent copy of the first ``ncell`` cells"""
        o = DeviceState(self.cfg, min(int(ncell), self.ncell), self.device)
        for k, v in self.t.items():
            o.t[k].copy_(v[:, : o.ncell])
        return o

    def to_host(self) -> abi.HostState:
        h = abi.HostState(self.cfg, self.ncell)
        for k in h.a:
            h.a[k][...] = self.t[k].cpu().numpy()
        return h

    def struct(self) -> abi.PfrxState:
        s = abi.PfrxState()
        s.ld = self.ncell
        for f in abi.STATE_DOUBLE_FIELDS + abi.STATE_ELM_FIELDS:
            t = self.t[f]
            setattr(s, f, C.cast(t.data_ptr() if t.numel() else None, abi.c_double_p))
        for f in abi.STATE_INT_FIELDS:
            t = self.t[f]
            setattr(s, f, C.cast(t.data_ptr() if t.numel() else None, abi.c_int32_p))
        return s


class PinnedHostState(abi.HostState):
    """HostState whose arrays live in page-locked memory (what a coupler would
    register once), so that the H2D/D2H legs of ``rstep_host`` are true DMA."""

    def __init__(self, cfg: abi.ReactionConfig, ncell: int):
        import torch

        super().__init__(cfg, ncell)
        self._pinned = {}
        for k, v in list(self.a.items()):
            t = torch.empty(v.shape, dtype=torch.float64 if v.dtype == np.float64 else torch.int32).pin_memory()
            t.copy_(torch.from_numpy(v))
            self._pinned[k] = t
            self.a[k] = t.numpy()

    def assign(self, host: abi.HostState) -> None:
        for k, v in host.a.items():
            self.a[k][...] = v


class ChemistryStep:
    """One reaction network on one GPU (``pfrx_handle``)."""

    def __init__(self, cfg: abi.ReactionConfig, device: int = 0):
        self.cfg = cfg
        self.device = int(device)
        self._h = C.c_void_p()
        self._state = None
        self.variant = "generic"
        _check(lib().pfrx_create(C.byref(cfg.c), self.device, C.byref(self._h)), "pfrx_create")

    def close(self) -> None:
        if self._h:
            lib().pfrx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- device-resident path ------------------------------------------------- #
    def bind(self, state: DeviceState) -> None:
        st = state.struct()
        _check(lib().pfrx_bind_state(self._h, state.ncell, C.byref(st)), "pfrx_bind_state")
        self._state = state  # keep tensors alive

    def rstep_async(self, tran_dt: float) -> None:
        _check(lib().pfrx_rstep_async(self._h, float(tran_dt)), "pfrx_rstep_async")

    def rstep_finish(self) -> abi.PfrxStepResult:
        res = abi.PfrxStepResult()
        _check(lib().pfrx_rstep_finish(self._h, C.byref(res)), "pfrx_rstep_finish")
        return res

    def rstep(self, tran_dt: float) -> abi.PfrxStepResult:
        res = abi.PfrxStepResult()
        _check(lib().pfrx_rstep(self._h, float(tran_dt), C.byref(res)), "pfrx_rstep")
        return res

    def reaction(self, want_jacobian: bool = True, tran_dt: float = 1.0):
        """batched RReaction (+ RReactionDerivative) on the bound state: returns
        ``res[ncomp, ncell]`` and ``jac[ncomp, ncomp, ncell]`` (or None) as torch tensors;
        ``tran_dt`` is option%tran_dt as the SOMDECOMP sandbox reads it"""
        import torch

        st = self._state
        if st is None:
            raise PfrxError("bind() a DeviceState first")
        n = self.cfg.c.naqcomp + self.cfg.c.nimcomp
        res = torch.empty((n, st.ncell), dtype=torch.float64, device=st.device)
        jac = torch.empty((n, n, st.ncell), dtype=torch.float64, device=st.device) if want_jacobian else None
        _check(lib().pfrx_reaction(self._h, float(tran_dt), int(bool(want_jacobian)), res.data_ptr(),
                                   jac.data_ptr() if jac is not None else None), "pfrx_reaction")
        return res, jac

    def update_auxvars(self, tran_xx=None, update_activity_coefs: bool = True) -> None:
        """RTUpdateAuxVars on the bound state (reactive_transport.F90:3525-3660): free-ion concentrations from
        the device block vector ``tran_xx`` [ncell, ncomp] (or the state's own), activity coefficients,
        RTAuxVarCompute"""
        _check(lib().pfrx_update_auxvars(self._h, tran_xx.data_ptr() if tran_xx is not None else None,
                                         int(bool(update_activity_coefs))), "pfrx_update_auxvars")

    def equilibrate_constraint(self, cons: abi.Constraint, conc):
        """batched ReactionEquilibrateConstraint on the bound state: ``conc[naqcomp, ncell]`` (device tensor)
        holds each cell's constraint values; the speciated rt_auxvar fields of the state are overwritten.
        Returns (num_iterations, ierror) as int32 device tensors"""
        import torch

        st = self._state
        if st is None:
            raise PfrxError("bind() a DeviceState first")
        if not conc.is_cuda or conc.dtype != torch.float64 or not conc.is_contiguous() or \
                tuple(conc.shape) != (self.cfg.c.naqcomp, st.ncell):
            raise PfrxError("conc must be a contiguous float64 device tensor [naqcomp, ncell]")
        its = torch.zeros(st.ncell, dtype=torch.int32, device=st.device)
        err = torch.zeros(st.ncell, dtype=torch.int32, device=st.device)
        _check(lib().pfrx_equilibrate_constraint(self._h, C.byref(cons.c), conc.data_ptr(), its.data_ptr(),
                                                 err.data_ptr()), "pfrx_equilibrate_constraint")
        return its, err

    # -- block vectors either side of the cell loop (pmc_subsurface_osrt.F90:260-274, 303-376) -- #
    def os_fixed_accum(self, fixed_accum) -> None:
        """fixed_accum[cell, i] = porosity*sat*1000*volume*total(i) for the aqueous components of
        the active cells; ``fixed_accum`` is a device tensor [ncell, ncomp] (PETSc block layout)"""
        _check(lib().pfrx_os_fixed_accum(self._h, fixed_accum.data_ptr()), "pfrx_os_fixed_accum")

    def os_load(self, solved_total=None, tran_xx=None) -> None:
        """total <- solved_total[:, :naq], immobile <- tran_xx[:, naq:] (block vectors [ncell, ncomp])"""
        _check(lib().pfrx_os_load(self._h, solved_total.data_ptr() if solved_total is not None else None,
                                  tran_xx.data_ptr() if tran_xx is not None else None), "pfrx_os_load")

    def os_store(self, tran_xx) -> None:
        """tran_xx[cell, :] <- (pri_molal, immobile) of the active cells"""
        _check(lib().pfrx_os_store(self._h, tran_xx.data_ptr()), "pfrx_os_store")

    def os_step_host(self, solved_total, tran_xx, tran_dt: float) -> abi.PfrxStepResult:
        """the operator-split step on the bound (device-resident) state with HOST block vectors
        [ncell, ncomp] (torch CPU tensors, pinned for overlap): upload, os_load, RStep, os_store,
        download of ``tran_xx`` (pmc_subsurface_osrt.F90:303-378)"""
        res = abi.PfrxStepResult()
        for t in (solved_total, tran_xx):
            if t is not None and (t.is_cuda or not t.is_contiguous()):
                raise PfrxError("os_step_host takes contiguous host tensors")
        _check(lib().pfrx_os_step_host(self._h, solved_total.data_ptr() if solved_total is not None else None,
                                       tran_xx.data_ptr(), float(tran_dt), C.byref(res)), "pfrx_os_step_host")
        return res

    # -- host-resident path (H2D + kernel + D2H inside the call) ---------------- #
    def rstep_host(self, host: abi.HostState, tran_dt: float) -> abi.PfrxStepResult:
        res = abi.PfrxStepResult()
        st = host.struct()
        _check(lib().pfrx_rstep_host(self._h, host.ncell, C.byref(st), float(tran_dt), C.byref(res)),
               "pfrx_rstep_host")
        return res

    def rstep_host_resident(self, fields) -> None:
        """keep these state fields (names of ``abi.STATE_DOUBLE_FIELDS``) resident in the device
        mirror of ``rstep_host``: uploaded once, not downloaded (``rstep_host_fetch`` on request)"""
        mask = 0
        for f in fields:
            mask |= 1 << abi.STATE_DOUBLE_FIELDS.index(f)
        _check(lib().pfrx_rstep_host_resident(self._h, mask), "pfrx_rstep_host_resident")

    def rstep_host_fetch(self, host: abi.HostState) -> None:
        st = host.struct()
        _check(lib().pfrx_rstep_host_fetch(self._h, host.ncell, C.byref(st)), "pfrx_rstep_host_fetch")

    # -- multi-GPU --------------------------------------------------------------- #
    def init_comm(self) -> None:
        """one rank per GPU; the NCCL id travels through torch.distributed"""
        import torch
        import torch.distributed as dist

        if not dist.is_initialized() or dist.get_world_size() == 1:
            return
        buf = (C.c_ubyte * 128)()
        if dist.get_rank() == 0:
            _check(lib().pfrx_comm_unique_id(buf), "pfrx_comm_unique_id")
        obj = [bytes(buf)]
        dist.broadcast_object_list(obj, src=0)
        ident = (C.c_ubyte * 128).from_buffer_copy(obj[0])
        _check(lib().pfrx_comm_init(self._h, dist.get_world_size(), dist.get_rank(), ident), "pfrx_comm_init")

    def allreduce(self, res: abi.PfrxStepResult) -> abi.PfrxStepResult:
        _check(lib().pfrx_allreduce(self._h, C.byref(res)), "pfrx_allreduce")
        return res

    # -- introspection ------------------------------------------------------------ #
    @property
    def stream_ptr(self) -> int:
        return int(lib().pfrx_stream(self._h))

    def cell_order(self, on: bool) -> None:
        """longest-first hand-out of the refill kernels (pfrx_cell_order); on by default"""
        _check(lib().pfrx_cell_order(self._h, int(bool(on))), "pfrx_cell_order")

    @property
    def launch_count(self) -> int:
        return int(lib().pfrx_launch_count(self._h))

    @property
    def bytes_per_cell(self) -> int:
        return int(lib().pfrx_bytes_per_cell(self._h))

    @property
    def signature(self) -> int:
        return int(lib().pfrx_config_signature(self._h))

    def load_specialized(self, cubin_path: Optional[str]) -> None:
        """attach a cubin written by :mod:`.specialize` (None detaches)"""
        arg = None if cubin_path is None else os.fsencode(cubin_path)
        _check(lib().pfrx_load_specialized(self._h, arg), "pfrx_load_specialized")

    def specialize(self, required: bool = False) -> bool:
        """Attach the network-specialised kernel for this configuration.

        The cubin must have been built beforehand (``specialize.build`` -- done by
        ``__graft_entry__.build()`` for the stock workloads; nvcc is not needed at
        run time).  Returns False when the network uses features the generator
        does not cover or no cubin exists, unless ``required``."""
        from . import specialize as _sp

        ok, why = _sp.supported(self.cfg)
        path = _sp.cubin_path(self.cfg) if ok else None
        if ok and not os.path.exists(path):
            ok, why = False, "no cubin at " + path
        if not ok:
            if required:
                raise PfrxError("specialised kernel unavailable: " + why)
            return False
        self.load_specialized(path)
        self.variant = "".join(k for k, v in _sp.VARIANT_STYLES.items() if v == _sp.default_variant(self.cfg)[1]) + \
            str(_sp.default_variant(self.cfg)[0])
        return True

    def last_transfer_bytes(self):
        """(H2D, D2H) bytes of the latest rstep_host"""
        a, b = C.c_int64(), C.c_int64()
        _check(lib().pfrx_last_transfer_bytes(self._h, C.byref(a), C.byref(b)), "pfrx_last_transfer_bytes")
        return a.value, b.value

    def autotune(self, state: DeviceState, tran_dt: float, variants=("s1", "k1", "q1", "p1", "w1"), sample: int = 303104,
                 repeats: int = 2) -> Dict[str, float]:
        """Pick the specialised-kernel variant that is fastest on THIS state.

        Which skeleton wins depends on the data, not only on the network: when
        the cells of a block need different numbers of Newton iterations the
        one-warp-per-block kernel's warps drift apart in its 280 KB instruction
        stream and stall on instruction fetch (C5: 2.4x slower than lock-step),
        when they need the same number the lock-step kernel only adds votes and
        register pressure (C3: 12 % slower).  So both are timed on a private copy
        of the first ``sample`` cells (a few milliseconds each) and the winner is
        attached.  ``state`` itself is not modified; it is bound on return.
        Returns {variant: seconds}; variants without a cubin are skipped."""
        import torch

        from . import specialize as _sp

        ok, why = _sp.supported(self.cfg)
        times: Dict[str, float] = {}
        if not ok:
            self.bind(state)
            return times
        stream = torch.cuda.ExternalStream(self.stream_ptr, device=state.device)
        ragged = [False]

        def time_variants(names, ncell):
            out: Dict[str, float] = {}
            trial = state.prefix(ncell)      # one private copy, restored before every run: the binding stays, so
            for v in names:                   # the refill kernels' longest-first order applies from the second run on
                warps, style = int(v[1:]), _sp.VARIANT_STYLES[v[0]]
                path = _sp.cubin_path(self.cfg, warps, style)
                if not os.path.exists(path):
                    continue
                self.load_specialized(path)
                self.bind(trial)
                best = None
                for r in range(repeats + 1):
                    for k in trial.t:
                        trial.t[k].copy_(state.t[k][..., :ncell])
                    torch.cuda.synchronize(state.device)
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(stream)
                    self.rstep_async(tran_dt)
                    e1.record(stream)
                    res = self.rstep_finish()
                    t = e0.elapsed_time(e1) * 1e-3
                    if r > 0:  # the first run pays module load and first touch
                        best = t if best is None else min(best, t)
                    if res.ncell_active > 0 and res.max_newton_iterations * res.ncell_active > 4 * res.sum_newton_iterations:
                        ragged[0] = True
                out[v] = best
            return out

        times = time_variants(variants, min(sample, state.ncell))
        big = min(state.ncell, 1 << 21)
        if times and ragged[0] and big > sample:
            # a ragged workload is as long as its slowest cells: on a small sample every skeleton measures that
            # tail and nothing else.  The candidates within 1.5x of the best run again on up to 2^21 cells.
            best = min(times.values())
            close = [v for v in times if times[v] <= 1.5 * best]
            if len(close) > 1:
                t2 = time_variants(close, big)
                scale = float(sample) / float(big)
                for v in times:
                    times[v] = t2[v] * scale if v in t2 else max(times[v], 2.0 * max(t2.values()) * scale)
        if times:
            win = min(times, key=times.get)
            self.load_specialized(_sp.cubin_path(self.cfg, int(win[1:]), _sp.VARIANT_STYLES[win[0]]))
            self.variant = win
        self.bind(state)
        return times

    def kernel_info(self) -> Dict[str, int]:
        a = (C.c_int * 5)()
        _check(lib().pfrx_kernel_info(self._h, a), "pfrx_kernel_info")
        return {"N": a[0], "lanes": a[1], "threads": a[2], "blocks_per_sm": a[3], "smem_bytes": a[4]}


def fp64_peak_tflops(device: int = 0):
    """measured DFMA peak of the device (TFLOP/s) and the SM clock it implies"""
    tf, mhz = C.c_double(), C.c_double()
    _check(lib().pfrx_diag_fp64_peak(int(device), C.byref(tf), C.byref(mhz)), "pfrx_diag_fp64_peak")
    return tf.value, mhz.value


def reduce_results(res: abi.PfrxStepResult, group=None) -> abi.PfrxStepResult:
    """The collective that replaces MPI_Allreduce(rstep_error, MAX) + MPI_Barrier
    (pmc_subsurface_osrt.F90:381-383), through torch.distributed: SUM of the
    cell / iteration / cut-cell counts, MAX of the error flag and of the per-cell
    maxima.  Works on any backend (gloo on CPU hosts, nccl on GPUs); the C-ABI
    twin is pfrx_allreduce()."""
    import torch
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return res
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    s = torch.tensor([res.ncell_active, res.sum_newton_iterations, res.num_cut_cells], dtype=torch.int64, device=dev)
    m = torch.tensor([res.max_newton_iterations, res.max_num_kinetic_state_updates, res.rstep_error,
                      res.max_sub_steps], dtype=torch.int64, device=dev)
    dist.all_reduce(s, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(m, op=dist.ReduceOp.MAX, group=group)
    out = abi.PfrxStepResult()
    out.ncell_active, out.sum_newton_iterations, out.num_cut_cells = (int(v) for v in s.tolist())
    (out.max_newton_iterations, out.max_num_kinetic_state_updates, out.rstep_error,
     out.max_sub_steps) = (int(v) for v in m.tolist())
    out.first_failed_cell = res.first_failed_cell  # shard-local by definition
    return out


def shard_range(ncell: int, rank: int, world: int):
    """contiguous ownership ranges, like PETSc DMDA local cells
    (pmc_subsurface_osrt.F90:349-350)"""
    lo = ncell * rank // world
    hi = ncell * (rank + 1) // world
    return lo, hi


def kinmr_checkpoint_rows(cfg: abi.ReactionConfig):
    """rows of ``kinmr_total_sorb`` in the order RTCheckpointKineticSorption* writes them
    (reactive_transport.F90:3968-4182); host-only"""
    import numpy as np

    n = C.c_int32(0)
    _check(lib().pfrx_kinmr_checkpoint_rows(C.byref(cfg.c), None, C.byref(n)), "pfrx_kinmr_checkpoint_rows")
    rows = np.zeros(max(n.value, 1), dtype=np.int32)
    _check(lib().pfrx_kinmr_checkpoint_rows(C.byref(cfg.c), rows.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(n)),
           "pfrx_kinmr_checkpoint_rows")
    return rows[:n.value]
