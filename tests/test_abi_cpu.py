"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports
every symbol include/pfrx.h declares, and its struct layout matches the ctypes
mirror.  No compute call is made (there is no GPU in the build container)."""
import ctypes as C
import os
import re

import pytest

from pflotran_elm_interface_b200 import abi

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
SO = os.path.join(ROOT, "pflotran_elm_interface_b200", "libpfrx_b200.so")


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "pfrx.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pfrx_[a-z0-9_]+)\s*\(", src)))


def _lib():
    if not os.path.exists(SO):
        import __graft_entry__ as g

        g.build()
    return C.CDLL(SO)


def test_header_symbols_exported():
    L = _lib()
    names = _declared_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/pfrx.h but not exported"


def test_struct_layout_matches_header():
    L = _lib()
    L.pfrx_sizeof.restype = C.c_int64
    assert L.pfrx_sizeof(0) == C.sizeof(abi.PfrxConfig)
    assert L.pfrx_sizeof(1) == C.sizeof(abi.PfrxState)
    assert L.pfrx_sizeof(2) == C.sizeof(abi.PfrxStepResult)
    assert L.pfrx_abi_version() == abi.PFRX_ABI_VERSION


def test_no_cpu_fallback():
    """without a device pfrx_create must fail with PFRX_E_CUDA, not compute"""
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from pflotran_elm_interface_b200 import rstep, workloads

    wl = workloads.by_name("c1")
    with pytest.raises(rstep.PfrxError):
        rstep.ChemistryStep(wl.cfg, 0)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "pflotran_elm_interface_b200")
    for dp, _, fns in os.walk(pkg):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, fn)).read()
                assert "oracle_lib" not in txt and "pfrx_oracle" not in txt, fn


def _create_rc(cfg):
    """pfrx_create return code and message; configuration checks run before any CUDA call"""
    L = _lib()
    L.pfrx_create.argtypes = [C.POINTER(abi.PfrxConfig), C.c_int, C.POINTER(C.c_void_p)]
    L.pfrx_last_error.restype = C.c_char_p
    h = C.c_void_p()
    rc = L.pfrx_create(C.byref(cfg.c), 0, C.byref(h))
    return rc, L.pfrx_last_error().decode()


def test_unsupported_configurations_are_refused_not_approximated():
    """what the CUDA path does not cover comes back as PFRX_E_INVALID (1) with a reason --
    also on a machine without a GPU, because the checks precede device selection"""
    from pflotran_elm_interface_b200 import workloads

    # SOMDECOMP with its CO2 as a gas species (ITYPE_GAS = 1)
    cfg = workloads.by_name("c4s", ncell=1).cfg
    cfg.somdec.co2_itype = 1
    rc, msg = _create_rc(cfg)
    assert rc == 1 and "gas" in msg
    # ELM build that asks for the flow-coupled moisture response
    cfg = workloads.by_name("c4se", ncell=1).cfg
    cfg.arrays["somdec_moisture_response_function"][:] = 1
    rc, msg = _create_rc(cfg)
    assert rc == 1 and "MOISTURE_RESPONSE_FUNCTION" in msg
    # PLANTN without its PlantN pool
    cfg = workloads.by_name("c4fe", ncell=1).cfg
    cfg.plantn.plantn_id = -1
    rc, msg = _create_rc(cfg)
    assert rc == 1 and "PLANTN" in msg
    # ion exchange with a table missing
    cfg = workloads.by_name("c6", ncell=1).cfg
    cfg.c.eqionx_k = C.cast(None, abi.c_double_p)
    rc, msg = _create_rc(cfg)
    assert rc == 1 and "ion exchange" in msg
    # a general reaction naming a species that does not exist
    cfg = workloads.by_name("c7", ncell=1).cfg
    cfg.arrays["general_fwd_specid"][0] = 17
    rc, msg = _create_rc(cfg)
    assert rc == 1 and "general reaction" in msg
    # a microbial reaction with an inhibition type that does not exist
    cfg = workloads.by_name("c8", ncell=1).cfg
    cfg.arrays["microbial_inhibition_type"][0] = 2
    rc, msg = _create_rc(cfg)
    assert rc == 1 and "inhibition type" in msg
    # a sandbox list naming a sandbox that does not exist
    cfg = workloads.by_name("c4s", ncell=1).cfg
    cfg.arrays["sandbox_list"][0] = 99
    rc, msg = _create_rc(cfg)
    assert rc == 1 and "sandbox_list" in msg
    # ABI mismatch
    cfg = workloads.by_name("c2", ncell=1).cfg
    cfg.c.abi_version = 1
    rc, msg = _create_rc(cfg)
    assert rc == 1 and "abi_version" in msg
