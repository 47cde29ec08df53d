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
