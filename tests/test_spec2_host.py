"""Form 2 of the specialised kernels (csrc/pfrx_spec2.cuh + specialize2.py) WITHOUT a GPU: the
generated routines -- product-form speciation, symmetric ln-space Jacobian, sparse L D L^T or
the reference's LU -- are compiled for the host (same source, -DS2_HOST) and stepped through
RStep by tests/spec2_host_driver.cpp; results and Newton / sub-step counts are compared with
the oracle cell by cell."""
import ctypes as C
import os
import subprocess
import tempfile

import numpy as np
import pytest

import oracle_lib as orc
from pflotran_elm_interface_b200 import abi, specialize2, workloads as W

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "..", "pflotran_elm_interface_b200", "csrc")


class SpecParams(C.Structure):
    _fields_ = [("max_its", C.c_int), ("max_cuts", C.c_int), ("max_dlnC", C.c_double), ("tol_relchange", C.c_double),
                ("tol_res", C.c_double), ("tol_relres", C.c_double), ("min_sat", C.c_double)]


def _build(cfg, solver, tmp):
    src = specialize2.generate_source2(cfg, "lockstep", solver=solver)
    cu = os.path.join(tmp, f"gen_{solver}.cu")
    with open(cu, "w") as f:
        f.write(src)
    so = os.path.join(tmp, f"host_{solver}.so")
    cmd = ["g++", "-O1", "-ffp-contract=off", "-shared", "-fPIC", "-DS2_HOST", "-std=c++17", "-I", CSRC, "-x", "c++",
           "-include", cu, os.path.join(HERE, "spec2_host_driver.cpp"), "-o", so]
    subprocess.check_call(cmd)
    L = C.CDLL(so)
    L.s2_host_rstep.argtypes = [C.POINTER(abi.PfrxState), C.c_longlong, C.c_double, C.POINTER(SpecParams)]
    return L


def _params(cfg):
    c = cfg.c
    return SpecParams(c.maximum_reaction_iterations, c.maximum_reaction_cuts, c.max_dlnC_rreact,
                      c.max_relative_change_tolerance, c.max_residual_tolerance, c.max_rel_residual_tolerance,
                      c.rt_min_saturation)


def _compare(name, ref, got, tol=1e-10):
    assert np.array_equal(ref["num_iterations"], got["num_iterations"]), (
        name, np.flatnonzero(ref["num_iterations"][0] != got["num_iterations"][0])[:10])
    for f in ("num_sub_steps", "num_kinetic_state_updates", "ierror"):
        assert np.array_equal(ref[f], got[f]), (name, f)
    for f in ("total", "pri_molal", "immobile", "mnrl_volfrac", "sec_molal", "pri_act_coef", "sec_act_coef",
              "total_sorb_eq", "srfcplxrxn_free_site_conc", "eqsrfcplx_conc", "ln_act_h2o"):
        a, b = ref[f], got[f]
        if a.size == 0:
            continue
        scale = np.maximum(np.abs(a), np.abs(b))
        err = np.where(scale < 1e-30, 0.0, np.abs(a - b) / np.maximum(scale, 1e-300))
        assert err.max() < tol, (name, f, err.max())
    # mineral rates: relative to their natural scale (k A (1 - QK) cancels near equilibrium)
    a, b = ref["mnrl_rate"], got["mnrl_rate"]
    if a.size:
        scale = np.maximum(np.abs(a).max(axis=1, keepdims=True), 1e-300)
        assert (np.abs(a - b) / scale).max() < 1e-9, (name, "mnrl_rate")


@pytest.mark.parametrize("solver", ["sym", "lu"])
@pytest.mark.parametrize("name,ncell,dts", [("c2", 300, (3600.0, 86400.0)), ("c3", 160, (3600.0, 30 * 86400.0)),
                                            ("c5", 120, (86400.0,)), ("c3aw", 96, (3600.0,))])
def test_generated_code_on_the_host_matches_the_oracle(name, ncell, dts, solver):
    wl = W.by_name(name, ncell=ncell)
    ok, why = specialize2.supported2(wl.cfg)
    assert ok, why
    with tempfile.TemporaryDirectory() as tmp:
        L = _build(wl.cfg, solver, tmp)
        prm = _params(wl.cfg)
        for dt in dts:
            ref = wl.state.copy()
            orc.rstep(wl.cfg, ref, dt, 2)
            got = wl.state.copy()
            st = got.struct()
            assert L.s2_host_rstep(C.byref(st), got.ncell, float(dt), C.byref(prm)) == 0
            _compare(f"{name} dt={dt} {solver}", ref.a, got.a)
