"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports
every symbol include/pfrx.h declares, and its struct layout matches the ctypes
mirror.  No compute call is made (there is no GPU in the build container)."""
import ctypes as C
import os
import re

import numpy as np
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
    # an active gas species naming a component that does not exist
    cfg = workloads.by_name("c7g", ncell=1).cfg
    cfg.arrays["acteq_specid"][0] = 9
    rc, msg = _create_rc(cfg)
    assert rc == 1 and "active gas" in msg
    # the RADON sandbox without its mineral
    cfg = workloads.by_name("c7g", ncell=1).cfg
    cfg.radon.mineral_id = 3
    rc, msg = _create_rc(cfg)
    assert rc == 1 and "RADON" in msg
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


def test_kinetic_sorption_checkpoint_order():
    """pfrx_kinmr_checkpoint_rows against the loops of RTCheckpointKineticSorptionBinary
    (reactive_transport.F90:4006-4056) restated here on the multirate Hanford network: flagged components
    outermost, then reactions, then rates 1..nrate; every vector is a row of kinmr_total_sorb"""
    from pflotran_elm_interface_b200 import rstep, workloads as W

    wl = W.by_name("c3mr", ncell=2)
    c, a = wl.cfg.c, wl.cfg.arrays
    naq, nmr = c.naqcomp, c.nkinmrsrfcplxrxn
    assert nmr > 0
    flag = np.zeros(naq, dtype=bool)
    for q in range(nmr):
        irxn = int(a["kinmrsrfcplxrxn_to_srfcplxrxn"][q])
        for j in range(int(a["srfcplxrxn_ptr"][irxn]), int(a["srfcplxrxn_ptr"][irxn + 1])):
            icplx = int(a["srfcplxrxn_to_complex"][j])
            for p in range(int(a["srfcplx_ptr"][icplx]), int(a["srfcplx_ptr"][icplx + 1])):
                flag[int(a["srfcplx_specid"][p])] = True
    want = []
    for icomp in range(naq):
        if not flag[icomp]:
            continue
        for q in range(nmr):
            nrate = int(a["kinmr_rate_ptr"][q + 1] - a["kinmr_rate_ptr"][q])
            for irate in range(1, nrate + 1):
                want.append(naq * (int(a["kinmr_rate_ptr"][q]) + q + irate) + icomp)
    got = rstep.kinmr_checkpoint_rows(wl.cfg)
    assert list(got) == want and len(want) == int(flag.sum()) * int(a["kinmr_rate_ptr"][nmr])
    assert max(want) < wl.state.a["kinmr_total_sorb"].shape[0] and len(set(want)) == len(want)
    # a network without multirate sorption has nothing to checkpoint
    assert len(rstep.kinmr_checkpoint_rows(W.by_name("c2", ncell=2).cfg)) == 0
