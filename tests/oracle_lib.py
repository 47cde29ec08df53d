"""ctypes loader of the CPU oracle (oracle/pfrx_oracle.c).  Test infrastructure
only: nothing in the product package imports this module."""
import ctypes as C
import os
import subprocess

import numpy as np

from pflotran_elm_interface_b200 import abi

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
SO = os.path.join(ROOT, "oracle", "_build", "libpfrx_oracle.so")

_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])


def lib():
    global _lib
    if _lib is not None:
        return _lib
    src = os.path.join(ROOT, "oracle", "pfrx_oracle.c")
    if (not os.path.exists(SO)) or (os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(SO)):
        build()
    L = C.CDLL(SO)
    cfgp, stp = C.POINTER(abi.PfrxConfig), C.POINTER(abi.PfrxState)
    dp = C.POINTER(C.c_double)
    L.pfrx_oracle_rstep.argtypes = [cfgp, C.c_int64, stp, C.c_double, C.POINTER(abi.PfrxStepResult), C.c_int]
    L.pfrx_oracle_rstep.restype = C.c_int
    L.pfrx_oracle_activity.argtypes = [cfgp, stp, C.c_int64]
    L.pfrx_oracle_auxvar_compute.argtypes = [cfgp, stp, C.c_int64]
    L.pfrx_oracle_girt_residual.argtypes = [cfgp, stp, C.c_int64, C.c_double, dp, dp, dp]
    L.pfrx_oracle_reaction.argtypes = [cfgp, stp, C.c_int64, C.c_double, dp, dp]
    L.pfrx_oracle_update_kinetic_state.argtypes = [cfgp, stp, C.c_int64, C.c_double]
    L.pfrx_oracle_rsolve.argtypes = [dp, dp, dp, dp, C.c_int, C.c_int]
    L.pfrx_oracle_lu_solve.argtypes = [dp, C.c_int, dp]
    L.pfrx_oracle_set_ref_bug_compat.argtypes = [C.c_int]
    L.pfrx_oracle_equilibrate_constraint.argtypes = [cfgp, C.POINTER(abi.PfrxConstraint), stp, C.c_int64, dp, C.c_int64,
                                                     C.POINTER(C.c_int)]
    _lib = L
    return L


def rstep(cfg: abi.ReactionConfig, state: abi.HostState, tran_dt: float, nthreads: int = 1):
    """the OS cell loop on the CPU; updates `state` in place"""
    res = abi.PfrxStepResult()
    st = state.struct()
    rc = lib().pfrx_oracle_rstep(C.byref(cfg.c), state.ncell, C.byref(st), float(tran_dt), C.byref(res), nthreads)
    assert rc == 0, rc
    return res


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def girt_residual(cfg, state, ic, dt):
    n = cfg.ncomp
    Res = np.zeros(n)
    Jac = np.zeros((n, n), order="F")
    acc = np.zeros(n)
    st = state.struct()
    e = lib().pfrx_oracle_girt_residual(C.byref(cfg.c), C.byref(st), ic, float(dt), _dp(Res), _dp(Jac), _dp(acc))
    return e, Res, Jac, acc


def reaction(cfg, state, ic, dt):
    """RReaction + RReactionDerivative alone on one cell: Res[n], Jac[n, n] (Jac[i, j] = dRes_i/dc_j)"""
    n = cfg.ncomp
    Res = np.zeros(n)
    Jac = np.zeros((n, n), order="F")
    st = state.struct()
    e = lib().pfrx_oracle_reaction(C.byref(cfg.c), C.byref(st), ic, float(dt), _dp(Res), _dp(Jac))
    assert e == 0
    return Res, Jac


def activity(cfg, state, ic=0):
    st = state.struct()
    return lib().pfrx_oracle_activity(C.byref(cfg.c), C.byref(st), ic)


def auxvar_compute(cfg, state, ic=0):
    st = state.struct()
    return lib().pfrx_oracle_auxvar_compute(C.byref(cfg.c), C.byref(st), ic)


def update_kinetic_state(cfg, state, ic, dt):
    st = state.struct()
    return lib().pfrx_oracle_update_kinetic_state(C.byref(cfg.c), C.byref(st), ic, float(dt))


def lu_solve(A, b):
    A = np.array(A, dtype=np.float64, order="F")
    b = np.array(b, dtype=np.float64)
    e = lib().pfrx_oracle_lu_solve(_dp(A), A.shape[0], _dp(b))
    return e, b


def rsolve(Res, Jac, conc, use_log):
    J = np.array(Jac, dtype=np.float64, order="F")
    r = np.array(Res, dtype=np.float64)
    c = np.array(conc, dtype=np.float64)
    u = np.zeros_like(r)
    e = lib().pfrx_oracle_rsolve(_dp(r), _dp(J), _dp(c), _dp(u), len(r), int(use_log))
    return e, u


# ---- op-counting build (oracle/pfrx_oracle_count.cpp): the same source with a counting scalar ----
SO_COUNT = os.path.join(ROOT, "oracle", "_build", "libpfrx_oracle_count.so")
_lib_count = None


def count_ops(cfg: abi.ReactionConfig, state: abi.HostState, tran_dt: float, nthreads: int = 1):
    """RStep over the cells of `state` (updated in place) with the operation counter on: returns
    (result, counted operations) -- add / sub / mul / div / compare = 1, exp / log / pow / sqrt / atan = 20,
    the convention of SURVEY.md section 8(d)"""
    global _lib_count
    if _lib_count is None:
        if not os.path.exists(SO_COUNT):
            build()
        L = C.CDLL(SO_COUNT)
        L.pfrx_oracle_rstep.argtypes = [C.POINTER(abi.PfrxConfig), C.c_int64, C.POINTER(abi.PfrxState), C.c_double,
                                        C.POINTER(abi.PfrxStepResult), C.c_int]
        L.pfrx_oracle_ops_get.restype = C.c_ulonglong
        _lib_count = L
    res = abi.PfrxStepResult()
    st = state.struct()
    _lib_count.pfrx_oracle_ops_reset()
    rc = _lib_count.pfrx_oracle_rstep(C.byref(cfg.c), state.ncell, C.byref(st), float(tran_dt), C.byref(res), nthreads)
    assert rc == 0, rc
    return res, int(_lib_count.pfrx_oracle_ops_get())


def equilibrate_constraint(cfg: abi.ReactionConfig, cons: abi.Constraint, state: abi.HostState, conc):
    """ReactionEquilibrateConstraint on every cell of ``state`` with conc[naqcomp, ncell]; returns (its, ierror)"""
    L = lib()
    conc = np.ascontiguousarray(conc, dtype=np.float64)
    n = state.ncell
    its = np.zeros(n, dtype=np.int32)
    err = np.zeros(n, dtype=np.int32)
    st = state.struct()
    for ic in range(n):
        it = C.c_int(0)
        err[ic] = L.pfrx_oracle_equilibrate_constraint(C.byref(cfg.c), C.byref(cons.c), C.byref(st), ic, _dp(conc), n,
                                                       C.byref(it))
        its[ic] = it.value
    return its, err
