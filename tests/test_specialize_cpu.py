"""The code generator for network-specialised kernels (no GPU needed): what it
accepts, that its output is deterministic and complete, and that the signature
tracks the tables the generated code bakes in."""
import copy
import re
import shutil

import numpy as np
import pytest

from pflotran_elm_interface_b200 import specialize, workloads as W


def test_supported_networks():
    for name, ok in (("c1", True), ("c2", True), ("c3", True), ("c5", True), ("c3mr", True), ("c4", True),
                     ("c4fe", True), ("c6", True), ("c7", True), ("c7s", True), ("c8", True), ("c4sw", True), ("c4st", True), ("c3t", False), ("c7g", False), ("c2pf", False),
                     ("c3an", False)):
        wl = W.by_name(name, ncell=2)
        got, why = specialize.supported(wl.cfg)
        assert got is ok, (name, why)
        if not ok:
            assert why
            with pytest.raises(ValueError):
                specialize.generate_source(wl.cfg)


def test_multirate_is_generated_for_the_one_warp_skeleton_only():
    wl = W.by_name("c3mr", ncell=2)
    src = specialize.generate_source(wl.cfg)
    assert "#define SPEC_NMR 1" in src and "void spec_mr_sorption(" in src and "spec_mr_rate_tab" in src
    with pytest.raises(ValueError):
        specialize.generate_source(wl.cfg, warps=1, style="refill")


def test_source_is_deterministic_and_covers_the_network():
    wl = W.by_name("c3", ncell=2)
    a = specialize.generate_source(wl.cfg)
    b = specialize.generate_source(W.by_name("c3", ncell=2).cfg)
    assert a == b
    c = wl.cfg.c
    # one product per secondary complex in the evaluation, and once more in the routine that
    # writes rt_auxvar%sec_molal when the cell is published
    assert len(re.findall(r"sp_\[\d+ \* ld\] = sk;", a)) == c.neqcplx
    assert len(re.findall(r"const double sk = s2_scale\(p, e, emax\);", a)) == 2 * c.neqcplx
    # tracers stay out of the matrix (two of the 15 Hanford primaries occur in no reaction)
    assert f"#define SPEC_N {c.naqcomp + c.nimcomp}" in a
    assert "#define SPEC_NC 13" in a
    assert f"#define SPEC_SIG {specialize.signature(wl.cfg)}ull" in a
    # every literal is an exact hexadecimal float or a small integer constant
    assert "0x1." in a and "nan" not in a.lower().replace("isnan", "")
    for routine in ("void spec2_eval(", "bool spec2_solve_sym(", "void spec2_store_totals(", "void spec2_store_act("):
        assert routine in a
    # form 1 of the same network (multirate variant): the reference's exp-of-sums formulation
    b1 = specialize.generate_source(W.by_name("c3mr", ncell=2).cfg)
    for routine in ("spec_activity", "spec_rtotal", "spec_sorption", "spec_minerals"):
        assert f"void {routine}(" in b1


@pytest.mark.parametrize("name", ["c1", "c2", "c3", "c3mr", "c5", "c4", "c4s", "c4se", "c4fe", "c6", "c7", "c7s", "c8"])
def test_library_and_generator_agree_on_the_signature(name, tmp_path):
    """config_signature() in csrc/pfrx_api.cu and specialize.signature() hash the same bytes, and the
    configuration written by pfrx_config_write generates the same kernel as the deck-built one"""
    import ctypes as C

    from pflotran_elm_interface_b200 import abi, rstep

    L = rstep.lib()
    cfg = W.by_name(name, ncell=2).cfg
    assert specialize.supported(cfg)[0]
    sig = specialize.signature(cfg)
    assert L.pfrx_config_signature_of(C.byref(cfg.c)) == sig
    f = str(tmp_path / "cfg.dump")
    assert L.pfrx_config_write(C.byref(cfg.c), f.encode()) == 0
    back = abi.ReactionConfig.from_dump(f)
    assert back.dump_signature == sig == specialize.signature(back)
    assert specialize.generate_source(back) == specialize.generate_source(cfg)


def test_signature_tracks_tables():
    wl = W.by_name("c2", ncell=2)
    s0 = specialize.signature(wl.cfg)
    assert s0 == specialize.signature(W.by_name("c2", ncell=7).cfg)  # cells do not matter
    assert s0 != specialize.signature(W.by_name("c3", ncell=2).cfg)
    cfg = W.by_name("c2", ncell=2).cfg
    cfg.arrays["eqcplx_logK"][0] += 1.0e-12  # in place: the ctypes struct points at this buffer
    assert specialize.signature(cfg) != s0


def test_sandbox_network():
    wl = W.by_name("c4", ncell=2)
    src = specialize.generate_source(wl.cfg)
    assert "#define SPEC_NCLM 7" in src and "void spec_sandbox(" in src
    # the aqueous tracer of the CLM-CN deck stays out of the matrix (no reaction), and so does the
    # respired C: a product only, its Jacobian column is the diagonal alone ("row-only")
    assert "#define SPEC_NC 11" in src and "#define SPEC_NROSPEC 1" in src


def test_row_only_species_leave_the_matrix(monkeypatch):
    """products and tracking species of the ELM-CN network: CO2, N2O, N2, PlantN, the uptake
    trackers and the litter N pools (the reference's Jacobian has no column for them)"""
    wl = W.by_name("c4fe", ncell=2)
    src = specialize.generate_source(wl.cfg)
    assert "#define SPEC_N 20" in src and "#define SPEC_NC 10" in src and "#define SPEC_NROSPEC 9" in src
    assert "bool spec_rowonly(" in src
    monkeypatch.setenv("PFRX_SPEC_NO_ROWONLY", "1")
    full = specialize.generate_source(wl.cfg)
    assert "#define SPEC_NC 19" in full and "#define SPEC_NROSPEC 0" in full
    # Hanford: every species sits in some complex, nothing is row-only
    src3 = specialize.generate_source(W.by_name("c3mr", ncell=2).cfg)
    assert "#define SPEC_NC 13" in src3 and "#define SPEC_NROSPEC 0" in src3


def test_variants_generate(monkeypatch):
    wl = W.by_name("c3", ncell=2)
    with pytest.raises(ValueError):
        specialize.generate_source(wl.cfg, warps=4, style="straight")
    # form 2 serves the one-warp styles of the Hanford / calcite class; the rolled dense solve and
    # everything with a sandbox or multirate sorption stay on form 1
    for style in ("straight", "lockstep", "refill", "refill_warp"):
        src = specialize.generate_source(wl.cfg, warps=1, style=style)
        assert "#define SPEC_FORM 2" in src and '#include "pfrx_spec2.cuh"' in src
        assert specialize.cubin_path(wl.cfg, 1, style).endswith("f2.cubin")
    assert '#include "pfrx_spec.cuh"' in specialize.generate_source(wl.cfg, warps=1, style="refill_looplu")
    assert '#include "pfrx_spec.cuh"' in specialize.generate_source(W.by_name("c4", ncell=2).cfg)
    assert '#include "pfrx_spec.cuh"' in specialize.generate_source(W.by_name("c3mr", ncell=2).cfg)
    monkeypatch.setenv("PFRX_SPEC_FORM", "1")
    assert '#include "pfrx_spec.cuh"' in specialize.generate_source(wl.cfg, warps=1, style="lockstep")


def test_form2_symbolic_factorisation():
    """the sparse L D L^T of form 2: elimination order, fill and the columns the reference's
    surface-complexation Jacobian leaves incomplete"""
    from pflotran_elm_interface_b200 import specialize2

    g = specialize2._Gen2(W.by_name("c3", ncell=2).cfg)
    assert g.nc == 13 and sorted(g.order) == g.coupled
    assert g.nl == 69                       # 58 structural entries of the lower triangle + 11 fill
    assert len(g.lslot) == g.nl and sorted(g.lslot.values()) == list(range(g.nl))
    # UO2++ / H+ / HCO3- share the surface site; only HCO3- (species 7) is missing from a complex
    assert g.sorb_species == [0, 4, 7] and g.ecols == [7]
    assert sorted(i for (i, j) in g.evar) == [0, 4]
    src = g.source()
    assert "#define SPEC_SYM 1" in src and "#define SPEC_NL 69" in src and "#define SPEC_THREADS 256" in src
    g5 = specialize2._Gen2(W.by_name("c5", ncell=2).cfg)
    assert g5.ecols == [] and g5.nl == 69
    lu = specialize2._Gen2(W.by_name("c3", ncell=2).cfg, solver="lu").source()
    assert "#define SPEC_SYM 0" in lu and "#define SPEC_THREADS 128" in lu


@pytest.mark.skipif(shutil.which("nvcc") is None, reason="needs nvcc")
def test_build_is_cached(tmp_path, monkeypatch):
    monkeypatch.setattr(specialize, "OUT", str(tmp_path))
    wl = W.by_name("c2", ncell=2)
    p = specialize.build(wl.cfg)
    assert p.endswith("_k1f2.cubin")
    import os

    t0 = os.path.getmtime(p)
    assert specialize.build(wl.cfg) == p and os.path.getmtime(p) == t0  # stamp hit, no recompile
    log = open(p[:-6] + ".log").read()
    assert "sm_100a" in log and re.search(r"Used \d+ registers", log)


@pytest.mark.parametrize("name", ["c2", "c3", "c3mr", "c5", "c4", "c4s", "c4se", "c4fe", "c6", "c7", "c7s", "c8"])
def test_structural_mask_covers_the_oracle_jacobian(name):
    """RSolve's row scaling in the generated kernels skips the entries outside spec_jrow_mask: every
    entry the oracle's Jacobian (accumulation + RReaction, reaction.F90:3868-3925) has on real cells
    must be inside it"""
    import oracle_lib as orc
    from pflotran_elm_interface_b200 import workloads

    wl = workloads.by_name(name, ncell=24)
    g = specialize._Gen(wl.cfg)
    g.source()
    assert 0.0 < g.jnz_density <= 1.0
    allowed = set(g._nz) | {(i, i) for i in range(wl.cfg.ncomp)}
    for cell in range(24):
        st = wl.state.copy()
        e, R, J, _ = orc.girt_residual(wl.cfg, st, cell, wl.tran_dt)
        assert e == 0
        nzi, nzj = np.nonzero(J)
        missing = {(int(i), int(j)) for i, j in zip(nzi, nzj)} - allowed
        # row-only species sit outside the dense core; their rows are not scaled by the mask
        missing = {(i, j) for (i, j) in missing if i not in g.roset}
        assert not missing, (name, cell, sorted(missing)[:8])
