"""Code generator for network-specialised kernels.

For one reaction network (:class:`~.abi.ReactionConfig`) this module writes a
CUDA translation unit in which every stoichiometric coefficient, logK, charge
and species index of the reference's per-cell routines is an immediate:

* ``spec_activity``  -- RActivityCoefficients, LAG branch   (reaction.F90:4553-4612)
* ``spec_rtotal``    -- RTotalAqueous + RTAccumulationDerivative (reaction.F90:4665-4759, 5775)
* ``spec_sorption``  -- RTotalSorbEqSurfCplx1, unit free-site stoichiometry (reaction_surf_complex.F90:641-900)
* ``spec_minerals``  -- RKineticMineral, TST without prefactors (reaction_mineral.F90:647-1078)

and includes ``csrc/pfrx_spec.cuh`` (RStep/RReact control flow, unrolled LU,
launch skeleton).  The result is compiled with nvcc for sm_100a into a cubin
that ``pfrx_load_specialized`` attaches to a handle; the cubin carries a
signature of the tables it was generated from and the library refuses a cubin
whose signature differs from the handle's configuration.

Networks the generator does not cover (multirate sorption, sandboxes, general
free-site stoichiometry, Temkin/affinity-power minerals, anisothermal logK) keep
running on the generic kernels.
"""
from __future__ import annotations

import math
import os
import struct
import subprocess
from typing import List, Optional, Tuple

import numpy as np

from . import abi, chem

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(CSRC, "_spec")

LOG_TO_LN = chem.LOG_TO_LN

# variant letter -> code style (PFRX_SPEC_VARIANT=<letter><warps per 32 cells>)
VARIANT_STYLES = {"s": "straight", "k": "lockstep", "l": "looplu", "m": "klooplu", "q": "refill",
                  "p": "refill_looplu", "w": "refill_warp"}


def _fnv1a(data: bytes) -> int:
    h = 0xCBF29CE484222325
    for b in data:
        h ^= b
        h = (h * 0x100000001B3) & 0xFFFFFFFFFFFFFFFF
    return h


def signature(cfg: abi.ReactionConfig) -> int:
    """FNV-1a over the tables the generated code bakes in; the byte sequence is
    the one config_signature() in csrc/pfrx_api.cu hashes"""
    c, a = cfg.c, cfg.arrays
    parts = [struct.pack("<12i3d", c.naqcomp, c.nimcomp, c.neqcplx, c.nkinmnrl, c.nsrfcplxrxn, c.nsrfcplx,
                         c.neqsrfcplxrxn, c.nkinmrsrfcplxrxn, c.clmcn_nrxn, c.use_log_formulation,
                         c.act_coef_update_frequency, c.use_activity_h2o, c.debyeA, c.debyeB, c.debyeBdot)]

    def add(name: str, count: int, dtype) -> None:
        if count > 0:
            arr = np.ascontiguousarray(a[name], dtype=dtype)
            assert arr.size == count, (name, arr.size, count)
            parts.append(arr.tobytes())

    f8, i4 = np.float64, np.int32
    add("primary_spec_Z", c.naqcomp, f8)
    add("primary_spec_a0", c.naqcomp, f8)
    if c.neqcplx > 0:
        n, nnz = c.neqcplx, int(a["eqcplx_ptr"][c.neqcplx])
        add("eqcplx_ptr", n + 1, i4)
        add("eqcplx_specid", nnz, i4)
        add("eqcplx_stoich", nnz, f8)
        for k in ("eqcplx_h2ostoich", "eqcplx_logK", "eqcplx_Z", "eqcplx_a0"):
            add(k, n, f8)
    if c.nkinmnrl > 0:
        n, nnz = c.nkinmnrl, int(a["kinmnrl_ptr"][c.nkinmnrl])
        add("kinmnrl_ptr", n + 1, i4)
        add("kinmnrl_specid", nnz, i4)
        add("kinmnrl_stoich", nnz, f8)
        for k in ("kinmnrl_h2ostoich", "kinmnrl_logK", "kinmnrl_molar_vol", "kinmnrl_rate_constant",
                  "kinmnrl_activation_energy", "kinmnrl_affinity_threshold", "kinmnrl_rate_limiter"):
            add(k, n, f8)
        add("kinmnrl_irreversible", n, i4)
    if c.nsrfcplxrxn > 0:
        nr, ns = c.nsrfcplxrxn, c.nsrfcplx
        nnz = int(a["srfcplx_ptr"][ns])
        add("srfcplxrxn_ptr", nr + 1, i4)
        add("srfcplxrxn_to_complex", int(a["srfcplxrxn_ptr"][nr]), i4)
        add("srfcplxrxn_surf_type", nr, i4)
        add("srfcplxrxn_to_surf", nr, i4)
        add("srfcplxrxn_site_density", nr, f8)
        add("srfcplx_ptr", ns + 1, i4)
        add("srfcplx_specid", nnz, i4)
        add("srfcplx_stoich", nnz, f8)
        for k in ("srfcplx_h2ostoich", "srfcplx_free_site_stoich", "srfcplx_logK"):
            add(k, ns, f8)
        add("eqsrfcplxrxn_to_srfcplxrxn", c.neqsrfcplxrxn, i4)
        if c.nkinmrsrfcplxrxn > 0:
            nm = c.nkinmrsrfcplxrxn
            add("kinmrsrfcplxrxn_to_srfcplxrxn", nm, i4)
            add("kinmr_rate_ptr", nm + 1, i4)
            add("kinmr_rate", int(a["kinmr_rate_ptr"][nm]), f8)
            add("kinmr_frac", int(a["kinmr_rate_ptr"][nm]), f8)
    if c.clmcn_nrxn > 0:
        parts.append(struct.pack("<3i", c.clmcn_npool, c.clmcn_C_species_id, c.clmcn_N_species_id))
        add("clmcn_CN_ratio", c.clmcn_npool, f8)
        for k in ("clmcn_pool_nspec", "clmcn_pool_C_id", "clmcn_pool_N_id"):
            add(k, c.clmcn_npool, i4)
        for k in ("clmcn_upstream_pool_id", "clmcn_downstream_pool_id"):
            add(k, c.clmcn_nrxn, i4)
        for k in ("clmcn_rate_constant", "clmcn_respiration_fraction", "clmcn_inhibition_constant"):
            add(k, c.clmcn_nrxn, f8)
    if c.somdec:
        sd = cfg.somdec
        parts.append(struct.pack("<14i3d", *[getattr(sd, k) for k in (
            "nrxn", "co2_id", "co2_itype", "o2_id", "o2_itype", "nh4_id", "no3_id", "n2o_id", "proton_id", "hr_id",
            "nmin_id", "nimm_id", "nimp_id", "ngasmin_id", "x0eps", "n2o_frac_mineralization",
            "inhibition_nh4_no3")]))
        for name, ctype in abi.PfrxSomdec._fields_[17:]:
            arr = a["somdec_" + name]
            parts.append(np.ascontiguousarray(arr, dtype=(f8 if ctype is abi.c_double_p else i4)).tobytes())
    if c.nitrif:
        nt = cfg.nitrif
        parts.append(struct.pack("<5i3d", nt.proton_id, nt.nh4_id, nt.no3_id, nt.n2o_id, nt.ngasnit_id,
                                 nt.k_nitr_max, nt.k_nitr_n2o, nt.x0eps))
    if c.denitr:
        dn = cfg.denitr
        parts.append(struct.pack("<4i3d", dn.no3_id, dn.n2_id, dn.n2o_id, dn.ngasdeni_id, dn.half_saturation,
                                 dn.k_deni_max, dn.x0eps))
    if c.plantn:
        pn = cfg.plantn
        parts.append(struct.pack("<6i5d", pn.nh4_id, pn.no3_id, pn.plantn_id, pn.plantndemand_id,
                                 pn.plantnh4uptake_id, pn.plantno3uptake_id, pn.half_saturation_nh4,
                                 pn.half_saturation_no3, pn.inhibition_nh4_no3, pn.x0eps_nh4, pn.x0eps_no3))
    if c.langmuir:
        lg = cfg.langmuir
        parts.append(struct.pack("<2i3d", lg.aq_id, lg.sorb_id, lg.k_kinetic, lg.k_equilibrium, lg.s_max))
    if c.cndegas:
        parts.append(bytes(cfg.cndegas))
    if c.calcite:
        parts.append(bytes(cfg.calcite))
    if c.somdec or c.nitrif or c.denitr or c.plantn or c.langmuir:
        parts.append(struct.pack("<i", (3 if c.elm_flow_coupled else 1) if c.elm_pflotran else 0))
        if c.nsandbox:
            parts.append(np.ascontiguousarray(a["sandbox_list"], dtype=i4).tobytes())
    # what the generator bakes in or refuses beyond the tables above (same bytes as config_signature())
    parts.append(struct.pack("<5i", c.h2o_aq_id, c.use_isothermal, c.act_coef_update_algorithm, c.use_total_as_guess,
                             c.use_full_geochemistry))
    for k in ("kinmnrl_Temkin_const", "kinmnrl_min_scale_factor", "kinmnrl_affinity_power"):
        if k in a:
            add(k, c.nkinmnrl, f8)
    if "kinmnrl_num_prefactors" in a:
        add("kinmnrl_num_prefactors", c.nkinmnrl, i4)
    if c.nsrfcplxrxn > 0 and "srfcplxrxn_stoich_flag" in a:
        add("srfcplxrxn_stoich_flag", c.nsrfcplxrxn, i4)
    # ion exchange, KD isotherms, dynamic KD (gen_sorption)
    parts.append(struct.pack("<4i", c.neqionxrxn, c.neqkdrxn, c.neqdynamickdrxn, c.ikd_units if c.neqkdrxn > 0 else 0))
    if c.neqionxrxn > 0:
        n, nnz = c.neqionxrxn, int(a["eqionx_ptr"][c.neqionxrxn])
        add("eqionx_ptr", n + 1, i4)
        add("eqionx_cationid", nnz, i4)
        add("eqionx_k", nnz, f8)
        add("eqionx_CEC", n, f8)
        add("eqionx_to_surf", n, i4)
        add("eqionx_Z_flag", n, i4)
    if c.neqkdrxn > 0:
        n = c.neqkdrxn
        for k in ("eqkd_specid", "eqkd_type", "eqkd_mineral"):
            add(k, n, i4)
        for k in ("eqkd_coeff", "eqkd_langmuir_b", "eqkd_freundlich_n"):
            add(k, n, f8)
    if c.neqdynamickdrxn > 0:
        n = c.neqdynamickdrxn
        for k in ("eqdynamickd_specid", "eqdynamickd_refspecid"):
            add(k, n, i4)
        for k in ("eqdynamickd_refspechigh", "eqdynamickd_low", "eqdynamickd_high", "eqdynamickd_power"):
            add(k, n, f8)
    # general / radioactive-decay / immobile-decay / microbial reactions (gen_kinetic)
    parts.append(struct.pack("<5i", c.ngeneral_rxn, c.nradiodecay_rxn, c.nimmobile_decay_rxn, c.nmicrobial_rxn,
                             c.microbial_concentration_units if c.nmicrobial_rxn > 0 else 0))
    if c.ngeneral_rxn > 0:
        n = c.ngeneral_rxn
        for pre in ("", "fwd_", "bwd_"):
            nnz = int(a[f"general_{pre}ptr"][n])
            add(f"general_{pre}ptr", n + 1, i4)
            add(f"general_{pre}specid", nnz, i4)
            add(f"general_{pre}stoich", nnz, f8)
        add("general_kf", n, f8)
        add("general_kr", n, f8)
    if c.nradiodecay_rxn > 0:
        n = c.nradiodecay_rxn
        nnz = int(a["radiodecay_ptr"][n])
        add("radiodecay_ptr", n + 1, i4)
        add("radiodecay_specid", nnz, i4)
        add("radiodecay_stoich", nnz, f8)
        add("radiodecay_forward_specid", n, i4)
        add("radiodecay_kf", n, f8)
    if c.nimmobile_decay_rxn > 0:
        add("immobile_decay_specid", c.nimmobile_decay_rxn, i4)
        add("immobile_decay_constant", c.nimmobile_decay_rxn, f8)
    if c.nmicrobial_rxn > 0:
        n = c.nmicrobial_rxn
        nnz, nm, nh = int(a["microbial_ptr"][n]), int(a["microbial_monod_ptr"][n]), int(a["microbial_inhibition_ptr"][n])
        add("microbial_ptr", n + 1, i4)
        add("microbial_specid", nnz, i4)
        add("microbial_stoich", nnz, f8)
        add("microbial_rate_constant", n, f8)
        if "microbial_activation_energy" in a:
            add("microbial_activation_energy", n, f8)
        add("microbial_monod_ptr", n + 1, i4)
        add("microbial_monod_specid", nm, i4)
        add("microbial_monod_K", nm, f8)
        add("microbial_monod_Cth", nm, f8)
        add("microbial_inhibition_ptr", n + 1, i4)
        add("microbial_inhibition_specid", nh, i4)
        add("microbial_inhibition_type", nh, i4)
        add("microbial_inhibition_C", nh, f8)
        add("microbial_inhibition_C2", nh, f8)
        add("microbial_biomassid", n, i4)
        add("microbial_biomass_yield", n, f8)
    return _fnv1a(b"".join(parts))


def _variant(cfg: abi.ReactionConfig, warps: Optional[int], style: Optional[str]) -> Tuple[int, str]:
    """(warps per 32 cells, code style) -- defaults from default_variant()"""
    dw, ds = default_variant(cfg)
    return (dw if warps is None else warps), (ds if style is None else style)


def uses_form2(cfg: abi.ReactionConfig, warps: int, style: str) -> bool:
    """form 2 (specialize2.py / csrc/pfrx_spec2.cuh: product-form speciation, symmetric ln-space
    Jacobian, sparse L D L^T) serves the one-warp-per-32-cells styles of the networks it covers"""
    from . import specialize2

    if warps != 1 or style not in specialize2.FORM2_STYLES:
        return False
    return supported(cfg)[0] and specialize2.supported2(cfg)[0]


def cubin_path(cfg: abi.ReactionConfig, warps: Optional[int] = None, style: Optional[str] = None) -> str:
    warps, style = _variant(cfg, warps, style)
    tag = {v: k for k, v in VARIANT_STYLES.items()}[style]
    form = "f2" if uses_form2(cfg, warps, style) else ""
    return os.path.join(OUT, f"spec_{signature(cfg):016x}_{tag}{warps}{form}.cubin")


def supported(cfg: abi.ReactionConfig) -> Tuple[bool, str]:
    c, a = cfg.c, cfg.arrays
    n = c.naqcomp + c.nimcomp
    if n > 20:
        return False, "more than 20 unknowns"
    if not c.use_full_geochemistry:
        return False, "use_full_geochemistry = 0 (RStep's tracer short cut)"
    if not c.use_isothermal and (c.neqcplx > 0 or c.nkinmnrl > 0 or c.nsrfcplx > 0):
        # without complexes, minerals and surface complexes nothing has a logK(T): the ELM-CN networks next to a
        # thermal flow mode run the same generated kernel
        return False, "anisothermal logK"
    if c.act_coef_update_algorithm != abi._chem.ACT_COEF_ALGORITHM_LAG:
        return False, "activity algorithm NEWTON"
    if c.nsrfcplxrxn != c.neqsrfcplxrxn + c.nkinmrsrfcplxrxn:
        return False, "kinetic surface complexation"
    if c.nsrfcplxrxn and np.any(a["srfcplxrxn_stoich_flag"] != 0):
        return False, "free-site stoichiometry other than 1"
    for k in ("kinmnrl_Temkin_const", "kinmnrl_min_scale_factor", "kinmnrl_affinity_power",
              "kinmnrl_num_prefactors"):
        if k in a:
            return False, k
    if c.use_total_as_guess:
        return False, "USE_TOTAL_CONCENTRATION_AS_GUESS"
    if (c.neqionxrxn > 0 or c.neqkdrxn > 0 or c.neqdynamickdrxn > 0) and (
            c.nkinmrsrfcplxrxn > 0 or c.clmcn_nrxn > 0 or c.somdec or c.nitrif or c.denitr or c.plantn or c.langmuir):
        return False, "ion exchange / KD isotherms next to multirate sorption or reaction sandboxes"
    if c.nradiodecay_rxn > 0 and (c.nsrfcplxrxn > 0 or c.neqionxrxn > 0):
        return False, "radioactive decay of an inventory sorbed by surface complexation or ion exchange"
    if has_kinetic3(cfg) and (c.nkinmrsrfcplxrxn > 0 or c.clmcn_nrxn > 0 or c.somdec or c.nitrif or c.denitr
                              or c.plantn or c.langmuir):
        return False, "general / decay / microbial reactions next to multirate sorption or reaction sandboxes"
    if c.nactive_gas > 0:
        return False, "active gas species"
    if c.radon:
        return False, "RADON sandbox"
    if c.cndegas:
        return False, "CNDEGAS sandbox"
    if c.calcite:
        return False, "CALCITE sandbox"
    if c.somdec or c.nitrif or c.denitr or c.plantn or c.langmuir:
        if os.environ.get("PFRX_SPEC_NO_ELMCN"):
            return False, "ELM-CN sandboxes disabled by PFRX_SPEC_NO_ELMCN"
        # the generated code uses dtotal = delta_ij * den/1000
        if c.neqcplx > 0:
            return False, "ELM-CN sandboxes together with secondary complexes"
    if c.somdec:
        sd, sa = cfg.somdec, {k[7:]: v for k, v in a.items() if k.startswith("somdec_")}
        if sd.x0eps <= 0.0:
            return False, "SOMDECOMP X0EPS <= 0"
        if sd.co2_itype != abi.SPEC_AQUEOUS or sd.o2_id >= 0 or sd.nh4_id < 0:
            return False, "SOMDECOMP CO2 must be aqueous, no O2 species, NH4+ present"
        if np.any(sa["upstream_is_aqueous"] != 0) or np.any(sa["downstream_is_aqueous"] != 0):
            return False, "SOMDECOMP aqueous pools"
        for k in ("temperature_response_function", "moisture_response_function", "ox_response_function", "q10", "ea",
                  "ox_half_saturation", "decomp_depth_efolding"):
            if np.any(sa[k] != sa[k][0]):
                return False, "SOMDECOMP reactions with different ABIOTIC_FACTORS"
        if np.any(sa["ox_specid"] >= 0):
            return False, "SOMDECOMP Ox species"
    return True, ""


def has_kinetic3(cfg: abi.ReactionConfig) -> bool:
    """RGeneral / RRadioactiveDecay / RImmobileDecay / RMicrobial present (generated by gen_kinetic)"""
    c = cfg.c
    return c.ngeneral_rxn > 0 or c.nradiodecay_rxn > 0 or c.nimmobile_decay_rxn > 0 or c.nmicrobial_rxn > 0


def _lit(x: float) -> str:
    """exact C++17 hexadecimal floating literal"""
    x = float(x)
    if x == 0.0:
        return "0.0"
    if x == int(x) and abs(x) < 1e6:
        return f"{x:.1f}"
    return float.hex(x)


def _term(st: float, expr: str) -> str:
    """' + st*expr' with exact simplifications for +-1"""
    if st == 1.0:
        return f" + {expr}"
    if st == -1.0:
        return f" - {expr}"
    return f" + {_lit(st)} * {expr}"


class _Gen:
    loop_lu = False  # dense solve as rolled loops (style "looplu")
    lockstep = False  # 128-thread blocks whose warps execute the same Newton iteration (style "lockstep")
    refill = False  # lock-step blocks whose finished lanes fetch the next cell (style "refill")
    onewarp = False  # refill in one-warp blocks: no block barriers, the warps of an SM drift (style "refill_warp")

    def __init__(self, cfg: abi.ReactionConfig):
        ok, why = supported(cfg)
        if not ok:
            raise ValueError("network not supported by the specialiser: " + why)
        self.cfg = cfg
        self.c = cfg.c
        self.a = cfg.arrays
        self.naq = cfg.c.naqcomp
        self.n = cfg.c.naqcomp + cfg.c.nimcomp
        self.ncx = cfg.c.neqcplx
        self.act_upd = cfg.c.act_coef_update_frequency == chem.ACT_COEF_FREQUENCY_NEWTON_ITER
        # activity classes (one Debye-Hueckel evaluation per distinct (Z, a0))
        self.cls: List[Tuple[float, float]] = []
        self.pri_cls = [self._class_of(z, a0) for z, a0 in zip(self.a["primary_spec_Z"], self.a["primary_spec_a0"])]
        self.cx_cls = ([self._class_of(z, a0) for z, a0 in zip(self.a["eqcplx_Z"], self.a["eqcplx_a0"])]
                       if self.ncx else [])
        self.out: List[str] = []
        # species that occur in some reaction form the matrix; the others are
        # diagonal rows/columns handled as scalars by pfrx_spec.cuh
        used = set()
        for ids in ("eqcplx_specid", "kinmnrl_specid", "srfcplx_specid"):
            if ids in self.a:
                used.update(int(v) for v in self.a[ids])
        if self.c.clmcn_nrxn > 0:
            naq = self.c.naqcomp
            used.add(naq + int(self.c.clmcn_C_species_id))
            used.add(naq + int(self.c.clmcn_N_species_id))
            for k in range(self.c.clmcn_npool):
                used.add(naq + int(self.a["clmcn_pool_C_id"][k]))
                if int(self.a["clmcn_pool_nspec"][k]) == 2:
                    used.add(naq + int(self.a["clmcn_pool_N_id"][k]))
        naq = self.c.naqcomp
        if self.c.somdec:
            sd = cfg.somdec
            sa = {k[7:]: v for k, v in self.a.items() if k.startswith("somdec_")}
            used.update(int(v) for v in (sd.co2_id, sd.nh4_id) if v >= 0)
            if sd.no3_id >= 0:
                used.add(int(sd.no3_id))
            if sd.n2o_id >= 0:
                used.add(int(sd.n2o_id))
            for v in (sd.hr_id, sd.nmin_id, sd.nimm_id, sd.nimp_id, sd.ngasmin_id):
                if v >= 0:
                    used.add(naq + int(v))
            for k in ("upstream_c_id", "upstream_n_id", "upstream_hr_id", "upstream_nmin_id", "upstream_nimp_id",
                      "upstream_nimm_id", "downstream_c_id", "downstream_n_id"):
                used.update(naq + int(v) for v in sa[k] if v >= 0)
        if self.c.nitrif:
            nt = cfg.nitrif
            used.update(int(v) for v in (nt.nh4_id, nt.no3_id, nt.n2o_id) if v >= 0)
            if nt.ngasnit_id >= 0:
                used.add(naq + int(nt.ngasnit_id))
        if self.c.plantn:
            pn = cfg.plantn
            used.update(int(v) for v in (pn.nh4_id, pn.no3_id) if v >= 0)
            used.update(naq + int(v) for v in (pn.plantn_id, pn.plantndemand_id, pn.plantnh4uptake_id,
                                               pn.plantno3uptake_id) if v >= 0)
        if self.c.langmuir:
            used.add(int(cfg.langmuir.aq_id))
            used.add(naq + int(cfg.langmuir.sorb_id))
        if self.c.denitr:
            dn = cfg.denitr
            used.update(int(v) for v in (dn.no3_id, dn.n2_id) if v >= 0)
            if dn.ngasdeni_id >= 0:
                used.add(naq + int(dn.ngasdeni_id))
        for ids in ("eqionx_cationid", "eqkd_specid", "eqdynamickd_specid", "eqdynamickd_refspecid"):
            if ids in self.a:
                used.update(int(v) for v in self.a[ids])
        self.dsp = {}
        for ids in ("general_specid", "general_fwd_specid", "general_bwd_specid", "radiodecay_specid",
                    "radiodecay_forward_specid", "microbial_specid", "microbial_monod_specid",
                    "microbial_inhibition_specid"):
            if ids in self.a:
                used.update(int(v) for v in self.a[ids])
        if self.c.nimmobile_decay_rxn > 0:
            used.update(naq + int(v) for v in self.a["immobile_decay_specid"])
        if self.c.nmicrobial_rxn > 0:
            for b in self.a["microbial_biomassid"]:
                if int(b) > 0:
                    used.add(int(b) - 1)
                elif int(b) < 0:
                    used.add(naq + (-int(b) - 1))
        # rows of d(total)/d(free) that RRadioactiveDecay reads: (parent, j) pairs kept beside the Jacobian
        self.dtp = {}
        self.used = sorted(used)
        self._recording = False
        self._pairs = set()
        self._nz = set()          # every (i, j) the generated code ever assigns something non-zero
        self._touch_all = False
        self._layout([])

    def _layout(self, rowonly) -> None:
        """core species form the dense matrix; ROW-ONLY species (their Jacobian column is the
        diagonal alone: reaction products and tracking species such as CO2(aq), N2O(aq), PlantN,
        Nmin) keep their few row entries in separate slots and are eliminated after the solve"""
        self.rowonly = list(rowonly)
        self.roset = set(rowonly)
        self.ropairs = {}
        self.coupled = [sp for sp in self.used if sp not in self.roset]
        self.cpos = {sp: ci for ci, sp in enumerate(self.coupled)}
        self.nc = len(self.coupled)

    def J(self, i: int, j: int) -> str:
        if self._recording:
            self._pairs.add((i, j))
        if not self._touch_all:
            self._nz.add((i, j))
        if i in self.roset:
            assert j == i or j not in self.roset, (i, j)
            k = self.ropairs.setdefault((i, j), len(self.ropairs))
            return f"SW(SPEC_OFF_RO + {k})"
        assert j not in self.roset, (i, j)
        return f"W[JX({self.cpos[i]}, {self.cpos[j]})]"

    def _class_of(self, z: float, a0: float) -> int:
        if not abs(z) > 1.0e-10:
            return -1
        key = (-z * z, float(a0))
        if key not in self.cls:
            self.cls.append(key)
        return self.cls.index(key)

    def w(self, s: str = "") -> None:
        self.out.append(s)

    # ------------------------------------------------------------------ pieces
    def gen_tables(self) -> None:
        z2 = [float(z) * float(z) for z in self.a["eqcplx_Z"]] if self.ncx else [0.0]
        cls = self.cx_cls if self.ncx else [-1]
        vol = [float(v) for v in self.a["kinmnrl_molar_vol"]] if self.c.nkinmnrl else [0.0]
        self.w("static __device__ const double spec_cx_z2_tab[] = {" + ", ".join(_lit(v) for v in z2) + "};")
        self.w("static __device__ const int spec_cx_cls_tab[] = {" + ", ".join(str(v) for v in cls) + "};")
        self.w("static __device__ const double spec_mn_vol_tab[] = {" + ", ".join(_lit(v) for v in vol) + "};")
        self.w("__device__ __forceinline__ double spec_cx_z2(int k) { return spec_cx_z2_tab[k]; }")
        self.w("__device__ __forceinline__ int spec_cx_cls(int k) { return spec_cx_cls_tab[k]; }")
        self.w("__device__ __forceinline__ double spec_mn_vol(int m) { return spec_mn_vol_tab[m]; }")
        self.w()

    def gen_activity(self) -> None:
        c, a = self.c, self.a
        self.w("__device__ __forceinline__ void spec_activity(const double (&c)[SPEC_N], SpecCell &s) {")
        terms = []
        for i in range(self.naq):
            z2 = float(a["primary_spec_Z"][i]) ** 2
            if z2 != 0.0:
                terms.append(f"c[{i}] * {_lit(z2)}")
        self.w("  double Ip = 0.0;")
        for t in terms:
            self.w(f"  Ip += {t};")
        self.w("  const double I = 0.5 * (Ip + s.Isec);")
        self.w("  const double sq = sqrt(I);")
        A, B, Bd = _lit(c.debyeA), _lit(c.debyeB), _lit(c.debyeBdot)
        for q, (negz2, a0) in enumerate(self.cls):
            self.w(f"  if (s.store) s.lgcls[{q}] = (sx_div({_lit(negz2)} * sq * {A}, 1.0 + {_lit(a0)} * {B} * sq) + {Bd} * I) * SPEC_LN;")
        if c.use_activity_h2o:
            mp = " + ".join(f"c[{i}]" for i in range(self.naq) if i != c.h2o_aq_id) or "0.0"
            self.w(f"  if (s.store) {{ double t = 1.0 - 0.017 * (({mp}) + s.msec); s.ln_act_h2o = t > 0.0 ? log(t) : 0.0; }}")
        self.w("}")
        self.w()

    def gen_rtotal(self) -> None:
        a, n, naq = self.a, self.n, self.naq
        self.w("__device__ __forceinline__ void spec_rtotal(const double (&c)[SPEC_N], double (&lna)[SPEC_N],")
        self.w("    double (&ic)[SPEC_N], double (&tot)[SPEC_N], SpecCell &s, double *W, double *sec_out,")
        self.w("    long long ld, double dt) {")
        self.w("  const double denL = s.den_kg * 1.e-3;")
        self.w("  const double psvd = s.por * s.sat * 1000.0 * s.vol / dt;")
        for i in range(n):
            if i < naq:
                if self.act_upd:
                    lg = "" if self.pri_cls[i] < 0 else f" + s.lgcls[{self.pri_cls[i]}]"
                else:
                    lg = f" + s.lngam[{i}]"
                self.w(f"  lna[{i}] = sx_log(c[{i}]){lg}; ic[{i}] = sx_rcp(c[{i}]); tot[{i}] = c[{i}];")
            else:
                self.w(f"  lna[{i}] = 0.0; ic[{i}] = 0.0; tot[{i}] = c[{i}];")
        self.w("  double Is = 0.0, ms = 0.0;")
        # how often each Jacobian entry is hit -> hot entries accumulate in registers
        hits = {}
        if self.ncx:
            ptr, ids, st = a["eqcplx_ptr"], a["eqcplx_specid"], a["eqcplx_stoich"]
            for k in range(self.ncx):
                sp = range(ptr[k], ptr[k + 1])
                for p2 in sp:
                    for p in sp:
                        e = (int(ids[p]), int(ids[p2]))
                        hits[e] = hits.get(e, 0) + 1
        budget = int(os.environ.get("PFRX_SPEC_HOT", 20 if n > 8 else 9))
        hot = sorted(hits, key=lambda e: -hits[e])[:budget]
        hot = [e for e in hot if hits[e] >= 4]
        self._nz |= set(hot)
        for (i, j) in hot:
            self.w(f"  double jh_{i}_{j} = {'1.0' if i == j else '0.0'};")
        written = set()
        if self.ncx:
            for k in range(self.ncx):
                sp = list(range(ptr[k], ptr[k + 1]))
                self.w("  {")
                lq = _lit(-float(a["eqcplx_logK"][k]) * LOG_TO_LN)
                expr = lq
                h2o = float(a["eqcplx_h2ostoich"][k])
                if h2o != 0.0:
                    expr += _term(h2o, "s.ln_act_h2o")
                for p in sp:
                    expr += _term(float(st[p]), f"lna[{int(ids[p])}]")
                if self.act_upd:
                    q = self.cx_cls[k]
                    arg = f"({expr})" if q < 0 else f"({expr}) - s.lgcls[{q}]"
                else:
                    arg = f"({expr}) - SW(SPEC_OFF_LNGSEC + {k})"
                self.w(f"    const double sk = sx_exp({arg});")
                self.w(f"    if (s.store) sec_out[{k} * ld] = sk;")
                z2 = float(a["eqcplx_Z"][k]) ** 2
                if z2 != 0.0:
                    self.w(f"    Is += sk * {_lit(z2)};")
                self.w("    ms += sk;")
                for p in sp:
                    i, s_i = int(ids[p]), float(st[p])
                    self.w(f"    tot[{i}] +={_term(s_i, 'sk')[2:]};" if s_i in (1.0,) else
                           (f"    tot[{i}] -= sk;" if s_i == -1.0 else f"    tot[{i}] += {_lit(s_i)} * sk;"))
                for p2 in sp:
                    j, s_j = int(ids[p2]), float(st[p2])
                    tj = f"(sk * ic[{j}])" if s_j == 1.0 else f"(({_lit(s_j)} * sk) * ic[{j}])"
                    self.w(f"    {{ const double t = {tj};")
                    for p in sp:
                        i, s_i = int(ids[p]), float(st[p])
                        val = "t" if s_i == 1.0 else f"{_lit(s_i)} * t"
                        if (i, j) in hot:
                            self.w(f"      jh_{i}_{j} += {val};")
                        else:
                            e = self.J(i, j)
                            if (i, j) in written:
                                self.w(f"      {e} += {val};")
                            else:
                                init = "1.0 + " if i == j else ""
                                self.w(f"      {e} = {init}{val};")
                                written.add((i, j))
                    self.w("    }")
                self.w("  }")
        self.w("  s.Isec = Is; s.msec = ms;")
        for i in range(naq):
            self.w(f"  tot[{i}] *= denL;")
        # finalise d(total)/d(free) * denL * psvd (RTAccumulationDerivative); only
        # aqueous species can be coupled (immobile ones need a sandbox)
        self.w("#pragma unroll")
        self.w("  for (int k = 0; k < SPEC_NRO; k++) SW(SPEC_OFF_RO + k) = 0.0;")
        if self.c.nradiodecay_rxn > 0:
            # rt_auxvar%aqueous%dtotal(parent, :) for RRadioactiveDecay (reaction.F90:5280-5290)
            self.dtp = {}
            for jc in sorted({int(v) for v in a["radiodecay_forward_specid"]}):
                for j in range(naq):
                    if (jc, j) in hot:
                        src = f"jh_{jc}_{j}"
                    elif (jc, j) in written:
                        src = self.J(jc, j)
                    elif jc == j:
                        src = "1.0"
                    else:
                        continue
                    k = self.dtp.setdefault((jc, j), len(self.dtp))
                    self.w(f"  s.dtp[{k}] = {src} * denL;")
        rec, self._recording = self._recording, False   # the loop below touches every pair: not structure
        self._touch_all = True
        for i in self.coupled:
            for j in self.coupled:
                e = self.J(i, j)
                if i >= naq or j >= naq:
                    # immobile species: accumulation V/dt on the diagonal (reaction.F90:5775-5848)
                    self.w(f"  {e} = {'s.vol / dt' if i == j else '0.0'};")
                elif (i, j) in hot:
                    self.w(f"  {e} = (jh_{i}_{j} * denL) * psvd;")
                elif (i, j) in written:
                    self.w(f"  {e} = ({e} * denL) * psvd;")
                elif i == j:
                    self.w(f"  {e} = (1.0 * denL) * psvd;")
                else:
                    self.w(f"  {e} = 0.0;")
        self._recording = rec
        self._touch_all = False
        if self.nc:
            self.w("  if (s.dry) {")
            self.w("#pragma unroll 1")
            self.w("    for (int e = 0; e < SPEC_NC * SPEC_JS; e++) SW(e) = 0.0;")
            self.w("#pragma unroll 1")
            self.w("    for (int i = 0; i < SPEC_NC; i++) SW(i * (SPEC_JS + 1)) = 1.0;")
            self.w("  }")
        self.w("}")
        self.w()

    def gen_sorption(self) -> None:
        c, a, n = self.c, self.a, self.n
        self.w("__device__ __forceinline__ void spec_sorption(const double (&c)[SPEC_N], const double (&lna)[SPEC_N],")
        self.w("    const double (&ic)[SPEC_N], double (&ts)[SPEC_N], SpecCell &s, double *W, const DevState &st,")
        self.w("    long long cell, double jscale) {")
        self.w("  (void)c;")
        for k in range(c.nsrfcplx):
            self.w(f"  if (s.store) s.scconc[{k}] = 0.0;")
        for e in range(c.neqsrfcplxrxn):
            self._emit_srfcplx_rxn(int(a["eqsrfcplxrxn_to_srfcplxrxn"][e]), "ts[%d]", True)
        # RTotalSorb's order (reaction.F90:4783-4835): surface complexation, ion exchange, dynamic KD, KD
        parents = sorted({int(v) for v in a["radiodecay_forward_specid"]}) if c.nradiodecay_rxn > 0 else []
        self.dsp = {}

        def ds_add(i: int, j: int, expr: str) -> None:
            # rt_auxvar%dtotal_sorb_eq(parent, j) for RRadioactiveDecay (reaction.F90:5257-5305)
            if i in parents:
                k = self.dsp.setdefault((i, j), len(self.dsp))
                self.w(f"    s.dsp[{k}] += {expr};")

        if parents and (c.neqkdrxn > 0 or c.neqdynamickdrxn > 0):
            self.w("#pragma unroll")
            self.w("  for (int k = 0; k < SPEC_NDSP; k++) s.dsp[k] = 0.0;")
        for r in range(c.neqionxrxn):
            self._emit_ionx(r)
        for r in range(c.neqdynamickdrxn):
            ikd, iref = int(a["eqdynamickd_specid"][r]), int(a["eqdynamickd_refspecid"][r])
            pw, lo = float(a["eqdynamickd_power"][r]), float(a["eqdynamickd_low"][r])
            hml = float(a["eqdynamickd_high"][r]) - lo
            self.w("  {  // RTotalSorbDynamicKD")
            self.w(f"    const double t = pow(c[{iref}] / {_lit(float(a['eqdynamickd_refspechigh'][r]))}, {_lit(pw)});")
            self.w(f"    const double KD = {_lit(lo)} + t * {_lit(hml)};")
            self.w(f"    const double dKD = {_lit(pw)} * t / c[{iref}] * {_lit(hml)};")
            self.w(f"    ts[{ikd}] = ts[{ikd}] + KD * c[{ikd}] * 250.0;")
            self.w(f"    {self.J(ikd, ikd)} += (KD * 250.0) * jscale;")
            self.w(f"    {self.J(ikd, iref)} += (dKD * c[{ikd}] * 250.0) * jscale;")
            ds_add(ikd, ikd, "KD * 250.0")
            ds_add(ikd, iref, f"dKD * c[{ikd}] * 250.0")
            self.w("  }")
        for r in range(c.neqkdrxn):
            ic_, ty, mn = int(a["eqkd_specid"][r]), int(a["eqkd_type"][r]), int(a["eqkd_mineral"][r])
            co = _lit(float(a["eqkd_coeff"][r]))
            self.w("  {  // RTotalSorbKD")
            if int(c.ikd_units) == 1:
                self.w(f"    double kd = {co} * s.den_kg * (1.0 - s.por) * s.spd * 1.e-3;")
            else:
                self.w(f"    double kd = {co};")
            if mn >= 0:
                self.w(f"    kd = kd * (st.mnrl_volfrac[{mn} * st.ld + cell]);")
            self.w(f"    const double m = c[{ic_}];")
            if ty == 1:      # PFRX_SORPTION_LINEAR
                self.w("    const double res = kd * m, dres = kd;")
            elif ty == 2:    # PFRX_SORPTION_LANGMUIR
                self.w("    const double t = kd * m;")
                self.w(f"    const double res = t * {_lit(float(a['eqkd_langmuir_b'][r]))} / (1.0 + t);")
                self.w("    const double dres = res / m - res / (1.0 + t) * t / m;")
            else:
                on = 1.0 / float(a["eqkd_freundlich_n"][r])
                self.w(f"    const double res = kd * pow(m, {_lit(on)});")
                self.w(f"    const double dres = res / m * {_lit(on)};")
            self.w(f"    ts[{ic_}] = ts[{ic_}] + res;")
            self.w(f"    {self.J(ic_, ic_)} += dres * jscale;")
            ds_add(ic_, ic_, "dres")
            self.w("  }")
        self.w("}")
        self.w()
        if c.nkinmrsrfcplxrxn > 0:
            # multirate reactions: the same equilibrium calculation per reaction, into s.mr_seq, with
            # V * A as the Jacobian scale (RMultiRateSorption, reaction_surf_complex.F90:552-637)
            nm, naq = c.nkinmrsrfcplxrxn, self.naq
            ptr = a["kinmr_rate_ptr"]
            self.w("__device__ __forceinline__ void spec_mr_sorption(const double (&lna)[SPEC_N], const double (&ic)[SPEC_N],")
            self.w("    SpecCell &s, double *W, const DevState &st, long long cell) {")
            for q in range(nm):
                self.w("  {")
                self.w(f"    const double jscale = s.vol * s.mr_A[{q}];")
                for i in range(naq):
                    self.w(f"    s.mr_seq[{q * naq + i}] = 0.0;")
                self._emit_srfcplx_rxn(int(a["kinmrsrfcplxrxn_to_srfcplxrxn"][q]), f"s.mr_seq[{q * naq} + %d]", False)
                self.w("  }")
            self.w("}")
            self.w()

    def _emit_ionx(self, r: int) -> None:
        """RTotalSorbEqIonx (reaction.F90:4906-5140) for exchange reaction r: Gaines-Thomas equivalent fractions
        (mixed valences: the scalar Newton iteration on KDj from the cell's previous answer), sorbed
        concentrations, jscale * d(total_sorb)/d(free) into the Jacobian.  The arithmetic of pfrx_tpc.cuh."""
        c, a = self.c, self.a
        p0, p1 = int(a["eqionx_ptr"][r]), int(a["eqionx_ptr"][r + 1])
        cats = [int(a["eqionx_cationid"][p]) for p in range(p0, p1)]
        ks = [float(a["eqionx_k"][p]) for p in range(p0, p1)]
        Z = [float(a["primary_spec_Z"][i]) for i in cats]
        nc = len(cats)
        surf = int(a["eqionx_to_surf"][r])
        cec = _lit(float(a["eqionx_CEC"][r]))
        self.w("  {  // RTotalSorbEqIonx")
        if surf >= 0:
            self.w(f"    const double omega = fmax({cec} * st.mnrl_volfrac[{surf} * st.ld + cell], 1.e-40);")
        else:
            self.w(f"    const double omega = {cec};")
        for j in range(nc):
            self.w(f"    double X{j};")
        if int(a["eqionx_Z_flag"][r]):
            self.w(f"    const double rkc = {_lit(ks[0])} * exp(lna[{cats[0]}]);")
            for j in range(1, nc):
                self.w(f"    const double ka{j} = {_lit(ks[j])} * exp(lna[{cats[j]}]);")
            self.w(f"    double rX = {_lit(Z[0])} * s.ixref[{r}] / omega;")
            self.w("    double KDj = rX / rkc;")
            self.w("    bool one_more = false;")
            self.w("    int it = 0;")
            self.w("    for (;;) {")
            self.w("      it++;")
            self.w("      if (it > 20000) break;")
            self.w("      rX = KDj * rkc;")
            self.w("      X0 = rX;")
            self.w("      double total = rX, dres = 0.0;")
            for j in range(1, nc):
                ratio = Z[j] / Z[0]
                self.w(f"      X{j} = ka{j} * pow(KDj, {_lit(ratio)});")
                self.w(f"      total = total + X{j};")
                self.w(f"      dres = dres + X{j} / KDj * {_lit(Z[j])};")
            self.w(f"      dres = dres / {_lit(Z[0])} + rkc;")
            self.w("      const double res = 1.0 - total;")
            self.w("      if (one_more) break;")
            self.w("      const double dK = res / dres;")
            self.w("      KDj = KDj + dK;")
            self.w("      KDj = fmax(KDj, 1.e-40);")
            self.w("      if (fabs(dK / KDj) < 1.e-12) one_more = true;")
            self.w("    }")
            self.w(f"    if (s.store) s.ixref[{r}] = rX * omega / {_lit(Z[0])};")
        else:
            self.w("    double sumkm = 0.0;")
            for j in range(nc):
                self.w(f"    X{j} = exp(lna[{cats[j]}]) * {_lit(ks[j])};")
                self.w(f"    sumkm = sumkm + X{j};")
            for j in range(nc):
                self.w(f"    X{j} = X{j} / sumkm;")
        self.w("    double sumZX = 0.0;")
        for j in range(nc):
            self.w(f"    sumZX = sumZX + {_lit(Z[j])} * X{j};")
        for i in range(nc):
            self.w(f"    {{ const double t1 = X{i} * omega / {_lit(Z[i])};")
            self.w(f"      if (s.store) s.ixconc[{p0 + i}] = t1;")
            self.w(f"      ts[{cats[i]}] = ts[{cats[i]}] + t1;")
            self.w(f"      const double t2 = {_lit(Z[i])} / sumZX;")
            for j in range(nc):
                if i == j:
                    d = f"t1 * (1.0 - (t2 * X{j})) * ic[{cats[j]}]"
                else:
                    d = f"(-t1) * t2 * X{j} * ic[{cats[j]}]"
                self.w(f"      {self.J(cats[i], cats[j])} += ({d}) * jscale;")
            self.w("    }")
        self.w("  }")

    def _emit_srfcplx_rxn(self, r: int, ts_fmt: str, store_conc: bool) -> None:
        """RTotalSorbEqSurfCplx1 for reaction r with unit free-site stoichiometry (closed form):
        sorbed totals into ts_fmt % species, jscale * d(total_sorb)/d(free) into the Jacobian"""
        c, a = self.c, self.a
        if True:
            cx = [int(v) for v in a["srfcplxrxn_to_complex"][a["srfcplxrxn_ptr"][r]:a["srfcplxrxn_ptr"][r + 1]]]
            ty = int(a["srfcplxrxn_surf_type"][r])
            dens = _lit(float(a["srfcplxrxn_site_density"][r]))
            self.w("  {")
            if ty == chem.MINERAL_SURFACE:
                self.w(f"    const double dens = {dens} * st.mnrl_volfrac[{int(a['srfcplxrxn_to_surf'][r])} * st.ld + cell];")
            elif ty == chem.ROCK_SURFACE:
                self.w(f"    const double dens = {dens} * s.spd * (1.0 - s.por);")
            else:
                self.w(f"    const double dens = {dens};")
            self.w("    if (dens < 1.e-40) {")
            self.w(f"      if (s.store) s.fsite[{r}] = 0.0;")
            self.w("    } else {")
            ptr, ids, st_ = a["srfcplx_ptr"], a["srfcplx_specid"], a["srfcplx_stoich"]
            for q, k in enumerate(cx):
                expr = _lit(-float(a["srfcplx_logK"][k]) * LOG_TO_LN)
                h2o = float(a["srfcplx_h2ostoich"][k])
                if h2o != 0.0:
                    expr += _term(h2o, "s.ln_act_h2o")
                for p in range(ptr[k], ptr[k + 1]):
                    expr += _term(float(st_[p]), f"lna[{int(ids[p])}]")
                self.w(f"      const double e{q} = sx_exp({expr});")
            self.w("      double esum = 0.0;")
            for q in range(len(cx)):
                self.w(f"      esum += e{q};")
            self.w("      const double fs = dens / (1.0 + esum);")
            self.w(f"      if (s.store) s.fsite[{r}] = fs;")
            for q, k in enumerate(cx):
                self.w(f"      const double S{q} = e{q} * fs;")
                if store_conc:
                    self.w(f"      if (s.store) s.scconc[{k}] += S{q};")
            self.w("      double den = 0.0;")
            for q in range(len(cx)):
                self.w(f"      den += S{q};")
            self.w("      den = den / fs + 1.0;")
            species = sorted({int(ids[p]) for k in cx for p in range(ptr[k], ptr[k + 1])})
            for i in species:
                self.w(f"      double tmp{i} = 0.0;")
            for q, k in enumerate(cx):
                for p in range(ptr[k], ptr[k + 1]):
                    i, nu = int(ids[p]), float(st_[p])
                    v = f"S{q}" if nu == 1.0 else f"{_lit(nu)} * S{q}"
                    self.w(f"      tmp{i} += {v}; {ts_fmt % i} += {v};")
            for i in species:
                self.w(f"      const double dsx{i} = (-tmp{i} / den) * ic[{i}];")
            for q, k in enumerate(cx):
                sp = list(range(ptr[k], ptr[k + 1]))
                self.w(f"      {{ const double nuiSx = S{q} / fs;")
                for p2 in sp:
                    j, nu_j = int(ids[p2]), float(st_[p2])
                    a1 = f"S{q} * ic[{j}]" if nu_j == 1.0 else f"{_lit(nu_j)} * S{q} * ic[{j}]"
                    self.w(f"        {{ const double t = {a1} + nuiSx * dsx{j};")
                    for p in sp:
                        i, nu_i = int(ids[p]), float(st_[p])
                        v = "t" if nu_i == 1.0 else f"({_lit(nu_i)} * t)"
                        self.w(f"          {self.J(i, j)} += jscale * {v};")
                    self.w("        }")
                self.w("      }")
            self.w("    }")
            self.w("  }")

    def gen_minerals(self) -> None:
        c, a, n = self.c, self.a, self.n
        self.w("__device__ __forceinline__ void spec_minerals(const double (&lna)[SPEC_N], const double (&ic)[SPEC_N],")
        self.w("    double (&res)[SPEC_N], SpecCell &s, double *W, const DevState &st, long long cell, bool apply) {")
        for m in range(c.nkinmnrl):
            ptr, ids, st_ = a["kinmnrl_ptr"], a["kinmnrl_specid"], a["kinmnrl_stoich"]
            sp = list(range(ptr[m], ptr[m + 1]))
            expr = _lit(-float(a["kinmnrl_logK"][m]) * LOG_TO_LN)
            h2o = float(a["kinmnrl_h2ostoich"][m])
            if h2o != 0.0:
                expr += _term(h2o, "s.ln_act_h2o")
            for p in sp:
                expr += _term(float(st_[p]), f"lna[{int(ids[p])}]")
            thr = float(a["kinmnrl_affinity_threshold"][m])
            lim = float(a["kinmnrl_rate_limiter"][m])
            eact = float(a["kinmnrl_activation_energy"][m])
            irr = int(a["kinmnrl_irreversible"][m])
            rate = _lit(float(a["kinmnrl_rate_constant"][m]))
            self.w("  {")
            self.w(f"    const double QK = sx_exp({expr});")
            self.w("    double aff = 1.0 - QK;")
            self.w("    const double sgn = copysign(1.0, aff);")
            self.w(f"    bool active = (st.mnrl_volfrac[{m} * st.ld + cell] > 0.0 || sgn < 0.0);")
            if irr == 1:
                self.w("    if (sgn < 0.0) active = false;")
            if thr > 0.0:
                self.w(f"    if (sgn < 0.0 && QK < {_lit(thr)}) active = false;")
            self.w("    double rate_vol = 0.0;")
            self.w("    if (active) {")
            if lim > 0.0:
                self.w(f"      aff = aff / (1.0 + (1.0 - aff) / {_lit(lim)});")
            if eact > 0.0:
                self.w(f"      const double spr = {rate} * exp({_lit(eact)} / 8.31446 * "
                       "(1.0 / (25.0 + 273.15) - 1.0 / (s.temp + 273.15)));")
            else:
                self.w(f"      const double spr = {rate} * 1.0;")
            self.w(f"      double Im_const = -st.mnrl_area[{m} * st.ld + cell];")
            self.w("      double Im = Im_const * sgn * fabs(aff) * spr;")
            self.w("      rate_vol = Im;")
            self.w("      if (apply) {")
            self.w("        Im_const = Im_const * s.vol;")
            self.w("        Im = Im * s.vol;")
            self.w("        const double dIm_dQK = -Im_const * spr;")
            if lim > 0.0:
                self.w(f"        const double den = 1.0 + (1.0 - aff) / {_lit(lim)};")
                self.w(f"        const double dfac = dIm_dQK * (1.0 + QK / {_lit(lim)} / den) * QK * (s.den_kg * 1.e-3) / den;")
            else:
                self.w("        const double dfac = dIm_dQK * QK * (s.den_kg * 1.e-3);")
            for p in sp:
                i, nu = int(ids[p]), float(st_[p])
                self.w(f"        res[{i}] +={' ' if nu == 1.0 else f' {_lit(nu)} *'} Im;")
            for p2 in sp:
                j, nu_j = int(ids[p2]), float(st_[p2])
                self.w(f"        {{ const double t = dfac * ({_lit(nu_j)} * ic[{j}]);")
                for p in sp:
                    i, nu_i = int(ids[p]), float(st_[p])
                    v = "t" if nu_i == 1.0 else f"{_lit(nu_i)} * t"
                    self.w(f"          {self.J(i, j)} += {v};")
                self.w("        }")
            self.w("      }")
            self.w("    }")
            self.w(f"    if (s.store) s.mrate[{m}] = rate_vol;")
            self.w("  }")
        self.w("}")
        self.w()

    def _lngam_expr(self, i: int) -> str:
        if self.act_upd:
            return "0.0" if self.pri_cls[i] < 0 else f"s.lgcls[{self.pri_cls[i]}]"
        return f"s.lngam[{i}]"

    def gen_kinetic(self) -> None:
        """RRadioactiveDecay, RGeneral, RMicrobial, RImmobileDecay in RReaction's order (reaction.F90:4095-4127) as
        straight-line code: species ids, stoichiometries, rate constants, Monod / inhibition constants as literals.
        Same arithmetic as the table-driven pfrx_tpc.cuh (which follows reaction.F90:5211-5460,
        reaction_microbial.F90:287-602, reaction_immobile.F90:244-296)."""
        c, a, naq = self.c, self.a, self.naq
        self.w("__device__ __forceinline__ void spec_kinetic(const double (&c)[SPEC_N], const double (&lna)[SPEC_N],")
        self.w("    const double (&ic)[SPEC_N], const double (&tot)[SPEC_N], const double (&ts)[SPEC_N], double (&res)[SPEC_N],")
        self.w("    SpecCell &s, double *W, double dt) {")
        self.w("  (void)lna; (void)ic; (void)tot; (void)ts; (void)dt; (void)W;")
        self.w("  const double L_water = s.por * s.vol * 1.e3 * s.sat;")
        self.w("  (void)L_water;")
        # ---- RRadioactiveDecay
        for r in range(c.nradiodecay_rxn):
            ptr, ids, st = a["radiodecay_ptr"], a["radiodecay_specid"], a["radiodecay_stoich"]
            jc = int(a["radiodecay_forward_specid"][r])
            kf = _lit(float(a["radiodecay_kf"][r]))
            self.w("  {  // RRadioactiveDecay")
            if c.neqkdrxn > 0 or c.neqdynamickdrxn > 0:
                self.w(f"    const double sum = tot[{jc}] * L_water + ts[{jc}] * s.vol;")
            else:
                self.w(f"    const double sum = tot[{jc}] * L_water;")
            self.w(f"    const double rate = sum * {kf};")
            self.w(f"    const double t = -1.0 * {kf};")
            for p in range(ptr[r], ptr[r + 1]):
                i, nu = int(ids[p]), float(st[p])
                self.w(f"    res[{i}] = res[{i}] - {_lit(nu)} * rate;")
                for (pj, j), k in self.dtp.items():
                    if pj == jc:
                        self._Jadd(i, j, f"t * {_lit(nu)} * s.dtp[{k}] * L_water")
            for p in range(ptr[r], ptr[r + 1]):
                i, nu = int(ids[p]), float(st[p])
                for (pj, j), k in self.dsp.items():
                    if pj == jc:
                        self._Jadd(i, j, f"t * {_lit(nu)} * s.dsp[{k}] * s.vol")
            self.w("  }")
        # ---- RGeneral
        if c.ngeneral_rxn > 0:
            self.w("  const double pdsv = s.por * s.den_kg * s.sat * s.vol;")
        for r in range(c.ngeneral_rxn):
            ptr, ids, st = a["general_ptr"], a["general_specid"], a["general_stoich"]
            fp, fi, fs = a["general_fwd_ptr"], a.get("general_fwd_specid"), a.get("general_fwd_stoich")
            bp, bi, bs = a["general_bwd_ptr"], a.get("general_bwd_specid"), a.get("general_bwd_stoich")
            kf, kr = float(a["general_kf"][r]), float(a["general_kr"][r])
            self.w("  {  // RGeneral")
            for side, kk, pp, ii, ss in (("f", kf, fp, fi, fs), ("r", kr, bp, bi, bs)):
                if kk > 0.0:
                    expr = _lit(math.log(kk))
                    for p in range(pp[r], pp[r + 1]):
                        expr += _term(float(ss[p]), f"lna[{int(ii[p])}]")
                    self.w(f"    const double lnQk{side} = {expr};")
                    self.w(f"    const double Qk{side} = sx_exp(lnQk{side});")
                else:
                    self.w(f"    const double Qk{side} = 0.0;")
            for p in range(ptr[r], ptr[r + 1]):
                i, nu = int(ids[p]), float(st[p])
                self.w(f"    res[{i}] = res[{i}] - {_lit(nu)} * (Qkf - Qkr) * pdsv;")
            if kf > 0.0:
                for q in range(fp[r], fp[r + 1]):
                    jc = int(fi[q])
                    self.w(f"    {{ const double t = -1.0 * {_lit(float(fs[q]))} * sx_exp(lnQkf - sx_log(c[{jc}])) * pdsv;")
                    for p in range(ptr[r], ptr[r + 1]):
                        self._Jadd(int(ids[p]), jc, f"{_lit(float(st[p]))} * t")
                    self.w("    }")
            if kr > 0.0:
                for q in range(bp[r], bp[r + 1]):
                    jc = int(bi[q])
                    self.w(f"    {{ const double t = {_lit(float(bs[q]))} * sx_exp(lnQkr - sx_log(c[{jc}])) * pdsv;")
                    for p in range(ptr[r], ptr[r + 1]):
                        self._Jadd(int(ids[p]), jc, f"{_lit(float(st[p]))} * t")
                    self.w("    }")
            self.w("  }")
        # ---- RMicrobial
        units = int(c.microbial_concentration_units) if c.nmicrobial_rxn > 0 else 0

        def dcdm(i: int) -> str:
            if units == 1:     # PFRX_MICROBIAL_MOLALITY / _ACTIVITY / _MOLARITY = 1 / 2 / 3
                return "1.0"
            if units == 2:
                lg = self._lngam_expr(i)
                return "1.0" if lg == "0.0" else f"exp({lg})"
            return "(s.den_kg * 1.e-3)"

        for r in range(c.nmicrobial_rxn):
            ptr, ids, st = a["microbial_ptr"], a["microbial_specid"], a["microbial_stoich"]
            mp, hp = a["microbial_monod_ptr"], a["microbial_inhibition_ptr"]
            m0, nm = int(mp[r]), int(mp[r + 1] - mp[r])
            h0, nh = int(hp[r]), int(hp[r + 1] - hp[r])
            self.w("  {  // RMicrobial")
            k = _lit(float(a["microbial_rate_constant"][r]))
            ea = float(a["microbial_activation_energy"][r]) if "microbial_activation_energy" in a else None
            if ea is not None:
                self.w(f"    const double k_eff = {k} * " + self.cellconst(f"exp({_lit(ea)} / 8.31446 * (1.0 / 298.15 - 1.0 / (s.temp + 273.15)))") + ";")
            else:
                self.w(f"    const double k_eff = {k};")
            # concentrations in the reaction's units, once per distinct species
            spec = []
            for ii in range(nm):
                spec.append(int(a["microbial_monod_specid"][m0 + ii]))
            for ii in range(nh):
                spec.append(int(a["microbial_inhibition_specid"][h0 + ii]))
            ib = int(a["microbial_biomassid"][r])
            if ib > 0:
                spec.append(ib - 1)
            for i in sorted(set(spec)):
                self.w(f"    const double dc{i} = {dcdm(i)};")
                self.w(f"    const double cc{i} = c[{i}] * dc{i};")
            for ii in range(nm):
                i = int(a["microbial_monod_specid"][m0 + ii])
                K, cth = _lit(float(a["microbial_monod_K"][m0 + ii])), _lit(float(a["microbial_monod_Cth"][m0 + ii]))
                self.w(f"    const double mden{ii} = {K} + cc{i} - {cth};")
                self.w(f"    const double monod{ii} = (cc{i} - {cth}) / mden{ii};")
                self.w(f"    const double dmon{ii} = dc{i} / mden{ii} - dc{i} * (cc{i} - {cth}) / (mden{ii} * mden{ii});")
            for ii in range(nh):
                i = int(a["microbial_inhibition_specid"][h0 + ii])
                ty = int(a["microbial_inhibition_type"][h0 + ii])
                C1, C2 = float(a["microbial_inhibition_C"][h0 + ii]), float(a["microbial_inhibition_C2"][h0 + ii])
                if ty == 3:       # PFRX_INHIBITION_MONOD
                    self.w(f"    const double hden{ii} = {_lit(C1)} + cc{i};")
                    self.w(f"    const double dinh{ii} = -1.0 * dc{i} * {_lit(C1)} / (hden{ii} * hden{ii});")
                    self.w(f"    const double inhib{ii} = {_lit(C1)} / ({_lit(C1)} + cc{i});")
                elif ty == 4:     # PFRX_INHIBITION_INVERSE_MONOD
                    self.w(f"    const double hden{ii} = {_lit(C1)} + cc{i};")
                    self.w(f"    const double dinh{ii} = dc{i} / hden{ii} - dc{i} * cc{i} / (hden{ii} * hden{ii});")
                    self.w(f"    const double inhib{ii} = cc{i} / ({_lit(C1)} + cc{i});")
                elif ty == 1:     # PFRX_INHIBITION_THRESHOLD
                    sg = "1.0" if math.copysign(1.0, C1) > 0 else "-1.0"
                    self.w(f"    const double ht{ii} = (cc{i} - {_lit(abs(C1))}) * {_lit(C2)};")
                    self.w(f"    const double dinh{ii} = {sg} * ({_lit(C2)} * dc{i} / (1.0 + ht{ii} * ht{ii})) / 3.14159265359;")
                    self.w(f"    const double inhib{ii} = 0.5 + {sg} * atan(ht{ii}) / 3.14159265359;")
                else:  # SMOOTHSTEP
                    lower = math.log10(C1) - 0.5 * C2
                    self.w(f"    const double hz{ii} = (log10(cc{i}) - {_lit(lower)}) / {_lit(C2)};")
                    self.w(f"    const double dinh{ii} = (hz{ii} < 0.0 || hz{ii} > 1.0) ? 0.0 : "
                           f"(6.0 * hz{ii} - 6.0 * (hz{ii} * hz{ii})) / ({_lit(C2)} * cc{i} * 2.30258509299) * dc{i};")
                    self.w(f"    const double inhib{ii} = hz{ii} < 0.0 ? 0.0 : (hz{ii} > 1.0 ? 1.0 : "
                           f"3.0 * (hz{ii} * hz{ii}) - 2.0 * (hz{ii} * hz{ii} * hz{ii}));")
            self.w("    double monod_terms = 1.0, inhib_terms = 1.0;")
            for ii in range(nm):
                self.w(f"    monod_terms = monod_terms * monod{ii};")
            for ii in range(nh):
                self.w(f"    inhib_terms = inhib_terms * inhib{ii};")
            brow, yld = -1, 0.0
            if ib > 0:
                brow, yld = ib - 1, float(a["microbial_biomass_yield"][r])
                self.w(f"    const double biomass_term = 1.0 * cc{brow} * L_water;")
                self.w(f"    const double dbio = dc{brow};")
            elif ib < 0:
                brow, yld = naq + (-ib - 1), float(a["microbial_biomass_yield"][r])
                self.w(f"    const double biomass_term = 1.0 * c[{brow}] * s.vol;")
                self.w("    const double dbio = 1.0;")
            else:
                self.w("    const double biomass_term = 1.0 * L_water;")
            self.w("    const double rate = k_eff * monod_terms * inhib_terms * biomass_term;")
            rows = [(int(ids[p]), float(st[p])) for p in range(ptr[r], ptr[r + 1])]
            for i, nu in rows:
                self.w(f"    res[{i}] = res[{i}] - {_lit(nu)} * rate;")
            if brow >= 0:
                self.w(f"    res[{brow}] = res[{brow}] - {_lit(yld)} * rate;")
            for ii in range(nm):
                jc = int(a["microbial_monod_specid"][m0 + ii])
                expr = "k_eff * inhib_terms * biomass_term"
                for jj in range(nm):
                    if jj != ii:
                        expr = f"({expr}) * monod{jj}"
                self.w(f"    {{ const double dR_dc = -1.0 * ({expr}) * dmon{ii};")
                for i, nu in rows:
                    self._Jadd(i, jc, f"{_lit(nu)} * dR_dc")
                if brow >= 0:
                    self._Jadd(brow, jc, f"{_lit(yld)} * dR_dc")
                self.w("    }")
            for ii in range(nh):
                jc = int(a["microbial_inhibition_specid"][h0 + ii])
                expr = "k_eff * monod_terms * biomass_term"
                for jj in range(nh):
                    if jj != ii:
                        expr = f"({expr}) * inhib{jj}"
                self.w(f"    {{ const double dR_dc = -1.0 * ({expr}) * dinh{ii};")
                for i, nu in rows:
                    self._Jadd(i, jc, f"{_lit(nu)} * dR_dc")
                if brow >= 0:
                    self._Jadd(brow, jc, f"{_lit(yld)} * dR_dc")
                self.w("    }")
            if brow >= 0:
                self.w("    { const double dRb = -1.0 * (k_eff * monod_terms * inhib_terms) * dbio;")
                for i, nu in rows:
                    self._Jadd(i, brow, f"{_lit(nu)} * dRb")
                self._Jadd(brow, brow, f"{_lit(yld)} * dRb")
                self.w("    }")
            self.w("  }")
        # ---- RImmobileDecay
        for r in range(c.nimmobile_decay_rxn):
            i = naq + int(a["immobile_decay_specid"][r])
            self.w("  {  // RImmobileDecay")
            self.w(f"    const double rc = {_lit(float(a['immobile_decay_constant'][r]))} * s.vol;")
            self.w(f"    res[{i}] = res[{i}] + rc * c[{i}];")
            self._Jadd(i, i, "rc")
            self.w("  }")
        self.w("}")
        self.w()

    def gen_sandbox(self) -> None:
        """RSandboxEvaluate (reaction_sandbox.F90:294-330): the sandboxes in the deck's order, each as
        straight-line code with the network's ids and constants as literals"""
        c = self.c
        order = [int(v) for v in self.a["sandbox_list"]] if c.nsandbox else [1, 2, 3, 4, 5, 6]
        self.w("__device__ __forceinline__ void spec_sandbox(const double (&c)[SPEC_N], const double (&lna)[SPEC_N],")
        self.w("    const double (&tot)[SPEC_N], double (&res)[SPEC_N], SpecCell &s, double *W, double dt) {")
        self.w("  const double denL = s.den_kg * 1.e-3;  // dtotal(i,i) of a network without complexes")
        self.w("  (void)denL; (void)lna; (void)tot; (void)dt;")
        for kind in order:
            if kind == abi.SANDBOX_CLM_CN and c.clmcn_nrxn > 0:
                self.w("  do {  // CLM-CN")
                self._emit_clm_cn()
                self.w("  } while (0);")
            elif kind == abi.SANDBOX_SOMDEC and c.somdec:
                self.w("  do {  // SOMDECOMP")
                self._emit_somdec()
                self.w("  } while (0);")
            elif kind == abi.SANDBOX_NITRIF and c.nitrif:
                self.w("  do {  // NITRIFICATION")
                self._emit_nitrif()
                self.w("  } while (0);")
            elif kind == abi.SANDBOX_DENITR and c.denitr:
                self.w("  do {  // DENITRIFICATION")
                self._emit_denitr()
                self.w("  } while (0);")
            elif kind == abi.SANDBOX_PLANTN and c.plantn:
                self.w("  do {  // PLANTN")
                self._emit_plantn()
                self.w("  } while (0);")
            elif kind == abi.SANDBOX_LANGMUIR and c.langmuir:
                self.w("  do {  // LANGMUIR")
                self._emit_langmuir()
                self.w("  } while (0);")
        self.w("}")
        self.w()

    # -- helpers of the ELM-CN emitters: dtotal = delta_ij * denL (no complexes) ------------
    def cellconst(self, expr: str) -> str:
        """a sub-expression of per-cell scalars only (temperature, saturation, ELM soil properties): evaluated once
        when the cell is loaded (spec_cell_constants) instead of in every Newton iteration of every sub-step"""
        if os.environ.get("PFRX_SPEC_NO_CELLCONST"):
            return "(" + expr + ")"
        if expr not in self.kc:
            self.kc.append(expr)
        return f"s.kc[{self.kc.index(expr)}]"

    def gen_cell_constants(self) -> None:
        self.w("__device__ __forceinline__ void spec_cell_constants(SpecCell &s) {")
        for k, e in enumerate(self.kc):
            self.w(f"  s.kc[{k}] = {e};")
        self.w("  (void)s;")
        self.w("}")
        self.w()

    def _Jsub(self, i: int, j: int, expr: str) -> None:
        self.w(f"    {self.J(i, j)} = {self.J(i, j)} - {expr};")

    def _Jadd(self, i: int, j: int, expr: str) -> None:
        self.w(f"    {self.J(i, j)} = {self.J(i, j)} + {expr};")

    def _conc(self, sid: int, itype: int) -> str:
        return f"tot[{sid}]" if itype == abi.SPEC_AQUEOUS else f"c[{self.naq + sid}]"

    def _emit_ph(self, proton_id: int) -> None:
        if proton_id >= 0:
            self.w(f"    const double ph = -lna[{proton_id}] * 0.43429448190325182765;")
        else:
            self.w("    const double ph = 6.5;")
        self.w("    double f_ph = 0.56 + atan(3.14159265358979323846 * 0.45 * (-5.0 + ph)) / 3.14159265358979323846;")

    def _emit_somdec(self) -> None:
        """SomDecReact / React1 / React2 / Nemission (reaction_sandbox_somdec.F90:1504-3640); the
        operation order of pfrx_sandbox.cuh, which is the reference's"""
        c, naq = self.c, self.naq
        sd = self.cfg.somdec
        sa = {k[7:]: v for k, v in self.a.items() if k.startswith("somdec_")}
        nrxn = int(sd.nrxn)
        x0 = _lit(sd.x0eps)
        x1 = _lit(sd.x0eps * 10.0)
        elm = bool(c.elm_pflotran)
        w = self.w
        w("    const double theta = s.sat * s.por;")
        # abiotic factors: identical for every reaction (supported())
        mf, of, tf = (int(sa[k][0]) for k in ("moisture_response_function", "ox_response_function",
                                              "temperature_response_function"))
        if elm and self.c.elm_flow_coupled and mf != 0:
            w(f"    double f_w = pfrx_sbx::elm_moisture_response(theta, {mf}, s.elm_sucsat, s.elm_bd_dry, s.elm_bsw, "
              "s.elm_watfc, s.elm_effpor);")
        elif elm:
            w("    double f_w = s.elm_w;")
        elif mf == 3:
            w("    double f_w;")
            w("    if (theta <= (double)0.08f) f_w = (double)0.01f;")
            w("    else f_w = log(theta / (double)0.08f) / 2.525728702545166015625;")
        else:
            w("    double f_w = 1.0;")
        if of == 2:
            w(f"    f_w = f_w * {self.cellconst('pfrx_sbx::wfps(s.sat)')};")
        elif elm:
            w("    f_w = f_w * s.elm_o;")
        if tf == 4:
            w("    const double f_t = " + self.cellconst(f"pfrx_sbx::temperature_response(s.temp, 4, {_lit(float(sa['ea'][0]))})") + ";")
        elif tf == 1:
            w("    const double f_t = " + self.cellconst("pfrx_sbx::temperature_response(s.temp, 1, 0.0)") + ";")
        elif tf in (2, 3):
            w("    const double f_t = " + self.cellconst(f"pfrx_sbx::temperature_response(s.temp, {tf}, {_lit(float(sa['q10'][0]))})") + ";")
        else:
            w("    const double f_t = " + ("s.elm_t;" if elm else "1.0;"))
        ef = float(sa["decomp_depth_efolding"][0])
        if elm and ef > 0.0:
            w("    const double f_depth = " + self.cellconst(f"fmin(1.0, fmax(1.e-20, exp(-s.elm_zsoil / {_lit(ef)})))") + ";")
        else:
            w("    const double f_depth = 1.0;")
        w("    if (f_t < 1.0e-20 || f_w < 1.0e-20 || f_depth < 1.0e-20) break;")
        w("    double net_nmin_rate = 0.0;")
        nh4, no3 = int(sd.nh4_id), int(sd.no3_id)
        co2 = int(sd.co2_id)
        dptr = sa["downstream_ptr"]
        for r in range(nrxn):
            uc = naq + int(sa["upstream_c_id"][r])
            ucid = int(sa["upstream_c_id"][r])
            un = naq + int(sa["upstream_n_id"][r]) if sa["upstream_n_id"][r] >= 0 else -1
            down = list(range(int(dptr[r]), int(dptr[r + 1])))
            w(f"    {{  // reaction {r}")
            rc, rd, ad = float(sa["rate_constant"][r]), float(sa["rate_decomposition"][r]), float(sa["rate_ad_factor"][r])
            if rc >= 0.0:
                w(f"    double k_decomp = {_lit(rc)};")
            elif rd >= 0.0:
                w(f"    double k_decomp = 1.0 - exp(-{_lit(rd)} * dt);")
                w("    k_decomp = k_decomp / dt;")
            else:
                w("    double k_decomp = 0.0;")
            w(f"    k_decomp = {_lit(ad)} * k_decomp;")
            if elm and ad > 1.0:
                w("    if (s.elm_kscalar > 0.0) k_decomp = k_decomp / s.elm_kscalar;")
            w("    k_decomp = fmin(k_decomp, 1.0 / dt);")
            w("    const double scaled = k_decomp * s.vol * f_t * f_w * f_depth;")
            w(f"    const double c_uc = c[{uc}];")
            w("    double feps0, dfeps0;")
            w(f"    pfrx_sbx::hsmooth(c_uc, {x1}, {x0}, feps0, dfeps0);")
            w("    double crate_uc = scaled * c_uc * feps0;")
            w("    double dcrate_uc_duc = scaled * (feps0 + c_uc * dfeps0);")
            # downstream N:C ratios
            ncd = {}
            for j in down:
                dn = int(sa["downstream_n_id"][j])
                dc = int(sa["downstream_c_id"][j])
                if dn >= 0 and dc >= 0:
                    w(f"    double ncd{j} = s.nc[{nrxn + j}];")
                    w(f"    if (c[{naq + dn}] >= {x0} && c[{naq + dc}] >= {x0}) {{ ncd{j} = c[{naq + dn}] / c[{naq + dc}]; "
                      f"if (s.store) s.nc[{nrxn + j}] = ncd{j}; }}")
                    ncd[j] = f"ncd{j}"
                else:
                    ncd[j] = _lit(float(sa["downstream_nc"][j]))
            if un >= 0:
                w(f"    double unc = s.nc[{r}];")
                w(f"    if (c[{un}] >= {x0} && c_uc >= {x0}) {{ unc = c[{un}] / c_uc; if (s.store) s.nc[{r}] = unc; }}")
                cst = 1.0
                for j in down:
                    cst = cst - float(sa["downstream_stoich"][j])
                w(f"    const double cst = {_lit(cst)};")
                w("    double nst = unc;")
                for j in down:
                    w(f"    nst = nst - {_lit(float(sa['downstream_stoich'][j]))} * {ncd[j]};")
                branches = ("both",)
            else:
                w(f"    const double unc = {_lit(float(sa['upstream_nc'][r]))};")
                w(f"    const double cst = {_lit(float(sa['mineral_c_stoich'][r]))};")
                w(f"    const double nst = {_lit(float(sa['mineral_n_stoich'][r]))};")
                branches = ("r1",) if float(sa["mineral_n_stoich"][r]) >= 0.0 else ("r2",)
            if branches == ("both",):
                w("    if (nst >= 0.0) {")
                self._emit_somdec_react(r, 1, sa, sd, uc, ucid, un, down, ncd)
                w("    } else {")
                self._emit_somdec_react(r, 2, sa, sd, uc, ucid, un, down, ncd)
                w("    }")
            else:
                w("    {")
                self._emit_somdec_react(r, 1 if branches == ("r1",) else 2, sa, sd, uc, ucid, un, down, ncd)
                w("    }")
            w("    }")
        # SomDecNemission (:3477-3640)
        if sd.n2o_id >= 0:
            n2o = int(sd.n2o_id)
            w(f"    if (net_nmin_rate > {x0}) {{")
            w(f"    const double c_nh4 = tot[{nh4}] * theta * 1000.0;")
            w(f"    double f_t2 = {self.cellconst('-0.06 + 0.13 * exp(0.07 * s.temp)')};")
            w(f"    double f_w2 = {self.cellconst('pfrx_sbx::wfps(s.sat)')};")
            self._emit_ph(int(sd.proton_id))
            w(f"    if (f_t2 > {x0} && f_w2 > {x0} && f_ph > {x0}) {{")
            w("    f_t2 = fmin(f_t2, 1.0); f_w2 = fmin(f_w2, 1.0); f_ph = fmin(f_ph, 1.0);")
            w("    const double temp_real = f_t2 * f_w2 * f_ph;")
            w("    double feps0, dfeps0;")
            w(f"    pfrx_sbx::hsmooth(c_nh4, {x1}, {x0}, feps0, dfeps0);")
            fr = _lit(sd.n2o_frac_mineralization)
            w(f"    const double nratecap = temp_real * {fr} * net_nmin_rate * dt;")
            w("    double fcap = 1.0, dfcap = 0.0;")
            w("    if (nratecap > c_nh4 * s.vol) {")
            w("      fcap = sx_monod(c_nh4 * s.vol, nratecap - c_nh4 * s.vol);")
            w("      dfcap = sx_dmonod(c_nh4 * s.vol, nratecap - c_nh4 * s.vol);")
            w("    }")
            w("    dfeps0 = dfeps0 * fcap + feps0 * dfcap;")
            w("    feps0 = feps0 * fcap;")
            w(f"    const double rate_n2o = temp_real * {fr} * net_nmin_rate * feps0;")
            w(f"    res[{nh4}] = res[{nh4}] + rate_n2o;")
            w(f"    res[{n2o}] = res[{n2o}] - 0.5 * rate_n2o;")
            if sd.ngasmin_id >= 0:
                w(f"    res[{naq + int(sd.ngasmin_id)}] -= rate_n2o;")
            w(f"    const double drate = temp_real * {fr} * net_nmin_rate * dfeps0;")
            self._Jadd(nh4, nh4, "drate * denL")
            if sd.ngasmin_id >= 0:
                self._Jsub(naq + int(sd.ngasmin_id), nh4, "drate")
            w("    }")
            w("    }")

    def _emit_somdec_react(self, r, which, sa, sd, uc, ucid, un, down, ncd) -> None:
        """SomDecReact1 (which == 1) or SomDecReact2 (which == 2) of reaction r; crate_uc, dcrate_uc_duc,
        cst, nst, unc are in scope"""
        naq, w = self.naq, self.w
        nh4, no3, co2 = int(sd.nh4_id), int(sd.no3_id), int(sd.co2_id)
        x0 = _lit(sd.x0eps)
        x1 = _lit(sd.x0eps * 10.0)
        react2 = which == 2
        if react2:
            w(f"    const double c_nh4 = tot[{nh4}] * theta * 1000.0;")
            w(f"    const double c_no3 = {('tot[%d] * theta * 1000.0' % no3) if no3 >= 0 else '0.0'};")
            w("    double finh = 1.0; bool skip = false;")
            if sd.inhibition_nh4_no3 > 0.0:
                w(f"    if (c_nh4 > {x0} && c_no3 > {x0}) finh = sx_monod(c_nh4 / c_no3, {_lit(1.0 / sd.inhibition_nh4_no3)});")
                w(f"    else if (c_nh4 > {x0} && c_no3 <= {x0}) finh = 1.0;")
                w(f"    else if (c_nh4 <= {x0} && c_no3 > {x0}) finh = 0.0;")
                w("    else skip = true;")
            w("    if (!skip) {")
            w("    double fnh4 = 1.0, dfnh4 = 0.0, fno3 = 1.0, dfno3 = 0.0;")
        # MONOD / INHIBITION lists
        w("    double fmb = 1.0, dfmb = 0.0;")
        for k in range(int(sa["monod_ptr"][r]), int(sa["monod_ptr"][r + 1])):
            sid, sty = int(sa["monod_specid"][k]), int(sa["monod_specitype"][k])
            mk, thr = _lit(float(sa["monod_half_saturation"][k])), _lit(float(sa["monod_threshold"][k]))
            pn = bool(sa["monod_pool_normalized"][k])
            w("    {")
            if react2 and sid == nh4:
                w(f"      double t = fmax(0.0, c_nh4 - {thr});")
                if pn:
                    w(f"      t = t / c[{uc}];")
                w(f"      fnh4 = sx_monod(t, {mk}); dfnh4 = sx_dmonod(t, {mk});")
            elif react2 and no3 >= 0 and sid == no3:
                w(f"      double t = fmax(0.0, c_no3 - {thr});")
                if pn:
                    w(f"      t = t / c[{uc}];")
                w(f"      fno3 = sx_monod(t, {mk}); dfno3 = sx_dmonod(t, {mk});")
            else:
                w(f"      double t = fmax(0.0, {self._conc(sid, sty)} - {thr});")
                if pn:
                    w(f"      t = t / c[{uc}];")
                    if sty == abi.SPEC_AQUEOUS:
                        w("      t = t * theta * 1000.0;" if react2 else "      t = t * s.por * s.sat * 1000.0;")
                w(f"      const double fx = sx_monod(t, {mk});")
                w(f"      const double dfx = {'sx_dmonod(t, ' + mk + ')' if ucid == sid else '0.0'};")
                w("      dfmb = dfmb * fx + fmb * dfx; fmb = fmb * fx;")
            w("    }")
        for k in range(int(sa["inhib_ptr"][r]), int(sa["inhib_ptr"][r + 1])):
            sid, sty, ity = int(sa["inhib_specid"][k]), int(sa["inhib_specitype"][k]), int(sa["inhib_itype"][k])
            ik, ik2 = float(sa["inhib_constant"][k]), float(sa["inhib_constant2"][k])
            w("    {")
            w(f"      const double t = {self._conc(sid, sty)};")
            if ik2 == -999.0 or ity != 1:
                if ity == 3:
                    w(f"      const double fx = {_lit(ik)} / (t + {_lit(ik)});")
                    w(f"      double dfx = -{_lit(ik)} / (t + {_lit(ik)}) / (t + {_lit(ik)});")
                elif ity == 4:
                    w(f"      const double fx = sx_monod(t, {_lit(ik)});")
                    w(f"      double dfx = sx_dmonod(t, {_lit(ik)});")
                else:
                    w("      const double fx = 1.0; double dfx = 0.0;")
            else:
                w(f"      const double fx = 0.5 + atan((t - {_lit(ik)}) * {_lit(ik2)}) / 3.14159265358979323846;")
                w(f"      const double u = (t - {_lit(ik)}) * {_lit(ik2)};")
                w(f"      double dfx = ({_lit(ik2)} / (1.0 + u * u)) / 3.14159265358979323846;")
            if ucid != sid:
                w("      dfx = 0.0;")
            w("      dfmb = dfmb * fx + fmb * dfx; fmb = fmb * fx;")
            w("    }")
        w("    dcrate_uc_duc = dcrate_uc_duc * fmb + crate_uc * dfmb;")
        w("    crate_uc = crate_uc * fmb;")
        # Ox Monod term: no Ox species (supported()) => f_ox = 1, df_ox = 0
        w("    dcrate_uc_duc = dcrate_uc_duc * 1.0 + crate_uc * 0.0;")
        w("    crate_uc = crate_uc * 1.0;")
        if react2:
            w("    { double feps0, dfeps0;")
            w(f"      pfrx_sbx::hsmooth(c_nh4, {x1}, {x0}, feps0, dfeps0);")
            w("      dfnh4 = dfnh4 * feps0 + fnh4 * dfeps0; fnh4 = fnh4 * feps0;")
            if no3 >= 0:
                w(f"      pfrx_sbx::hsmooth(c_no3, {x1}, {x0}, feps0, dfeps0);")
                w("      dfno3 = dfno3 * feps0 + fno3 * dfeps0; fno3 = fno3 * feps0;")
            w("    }")
            w("    const double nratecap = -crate_uc * nst * dt / 0.45;")
            w("    { double fcap = 1.0, dfcap = 0.0;")
            w("      if (nratecap * finh > c_nh4 * s.vol) {")
            w("        fcap = sx_monod(c_nh4 * s.vol, nratecap * finh - c_nh4 * s.vol);")
            w("        dfcap = sx_dmonod(c_nh4 * s.vol, nratecap * finh - c_nh4 * s.vol);")
            w("      }")
            w("      dfnh4 = dfnh4 * fcap + fnh4 * dfcap; fnh4 = fnh4 * fcap; }")
            if no3 >= 0:
                w("    { double fcap = 1.0, dfcap = 0.0;")
                w("      if (nratecap * (1.0 - finh) > c_no3 * s.vol) {")
                w("        fcap = sx_monod(c_no3 * s.vol, nratecap * (1.0 - finh) - c_no3 * s.vol);")
                w("        dfcap = sx_dmonod(c_no3 * s.vol, nratecap * (1.0 - finh) - c_no3 * s.vol);")
                w("      }")
                w("      dfno3 = dfno3 * fcap + fno3 * dfcap; fno3 = fno3 * fcap; }")
            w("    const double crate_nh4 = crate_uc * fnh4 * finh;")
            w("    const double crate_no3 = crate_uc * fno3 * (1.0 - finh);")
            w("    const double crate = crate_nh4 + crate_no3;")
        else:
            w("    const double crate = crate_uc;")
        # residual, common part
        uhr = int(sa["upstream_hr_id"][r])
        w(f"    res[{uc}] = res[{uc}] + crate;")
        w(f"    res[{co2}] = res[{co2}] - cst * crate;")
        if uhr >= 0:
            w(f"    res[{naq + uhr}] -= cst * crate;")
        if sd.hr_id >= 0:
            w(f"    res[{naq + int(sd.hr_id)}] -= cst * crate;")
        for j in down:
            dc = int(sa["downstream_c_id"][j])
            if dc >= 0:
                w(f"    res[{naq + dc}] = res[{naq + dc}] - {_lit(float(sa['downstream_stoich'][j]))} * crate;")
        if un >= 0:
            w(f"    res[{un}] = res[{un}] + unc * crate;")
        unmin, unimm, unimp = (int(sa[k][r]) for k in ("upstream_nmin_id", "upstream_nimm_id", "upstream_nimp_id"))
        if not react2:
            w(f"    res[{nh4}] = res[{nh4}] - nst * crate;")
            w("    net_nmin_rate = net_nmin_rate + nst * crate;")
            if unmin >= 0:
                w(f"    res[{naq + unmin}] -= nst * crate;")
            if sd.nmin_id >= 0:
                w(f"    res[{naq + int(sd.nmin_id)}] -= nst * crate;")
        else:
            w("    double nimm = 0.0;")
            w(f"    res[{nh4}] = res[{nh4}] - nst * crate_nh4; nimm = nimm + nst * crate_nh4;")
            if no3 >= 0:
                w(f"    res[{no3}] = res[{no3}] - nst * crate_no3; nimm = nimm + nst * crate_no3;")
            w("    net_nmin_rate = net_nmin_rate + nimm;")
            if unimm >= 0:
                w(f"    res[{naq + unimm}] += nst * crate;")
            if unimp >= 0:
                w(f"    res[{naq + unimp}] += nst * crate_uc;")
            if sd.nimm_id >= 0:
                w(f"    res[{naq + int(sd.nimm_id)}] += nst * crate;")
            if sd.nimp_id >= 0:
                w(f"    res[{naq + int(sd.nimp_id)}] += nst * crate_uc;")
        for j in down:
            dn = int(sa["downstream_n_id"][j])
            if dn >= 0:
                w(f"    res[{naq + dn}] = res[{naq + dn}] - {_lit(float(sa['downstream_stoich'][j]))} * crate * {ncd[j]};")

        # Jacobian columns; all pools immobile, CO2 aqueous, dtotal = delta * denL
        def column(jcol, dco2, duc, dun, wrt_uc):
            if wrt_uc:
                self._Jsub(co2, jcol, dco2)
            # else: CO2 row gets dco2 * dtotal(co2, nh4|no3) = 0
            if uhr >= 0:
                self._Jsub(naq + uhr, jcol, dco2)
            if sd.hr_id >= 0:
                self._Jsub(naq + int(sd.hr_id), jcol, dco2)
            self._Jsub(uc, jcol, duc)
            for j in down:
                dc = int(sa["downstream_c_id"][j])
                self._Jsub(naq + dc, jcol, f"{_lit(float(sa['downstream_stoich'][j]))} * (-1.0 * {duc})")

        def nrows(jcol, duc, dun):
            if un >= 0:
                self._Jsub(un, jcol, dun)
            for j in down:
                dn = int(sa["downstream_n_id"][j])
                if dn >= 0:
                    self._Jsub(naq + dn, jcol, f"{_lit(float(sa['downstream_stoich'][j]))} * (-1.0 * {duc}) * {ncd[j]}")

        if not react2:
            w("    const double dco2_duc = dcrate_uc_duc * cst;")
            w("    const double duc_duc = -1.0 * dcrate_uc_duc;")
            w("    const double dnh4_duc = dco2_duc * nst;")
            w("    const double dun_duc = unc * duc_duc;")
            column(uc, "dco2_duc", "duc_duc", "dun_duc", True)
            self._Jsub(nh4, uc, "dnh4_duc")
            if unmin >= 0:
                self._Jsub(naq + unmin, uc, "dnh4_duc")
            if sd.nmin_id >= 0:
                self._Jsub(naq + int(sd.nmin_id), uc, "dnh4_duc")
            nrows(uc, "duc_duc", "dun_duc")
        else:
            w("    double dcrate_dx = dcrate_uc_duc * (fnh4 * finh + fno3 - fno3 * finh);")
            w("    const double dco2_duc = dcrate_dx * cst;")
            w("    const double duc_duc = -1.0 * dcrate_dx;")
            w("    const double dnh4_duc = dcrate_uc_duc * nst * fnh4 * finh;")
            w("    const double dno3_duc = dcrate_uc_duc * nst * fno3 * (1.0 - finh);")
            w("    const double dun_duc = unc * duc_duc;")
            w("    dcrate_dx = (dfnh4 * finh + (fnh4 - fno3) * 0.0);")
            w("    dcrate_dx = dcrate_dx * crate_uc;")
            w("    const double duc_dnh4 = -1.0 * dcrate_dx;")
            w("    const double dco2_dnh4 = dcrate_dx * cst;")
            w("    double dnh4_dnh4 = fnh4 * 0.0 + dfnh4 * finh;")
            w("    dnh4_dnh4 = dnh4_dnh4 * crate_uc * nst;")
            w("    const double dun_dnh4 = unc * duc_dnh4;")
            w("    dcrate_dx = (fnh4 - fno3) * 0.0 + dfno3 * (1.0 - fnh4 * finh);")
            w("    dcrate_dx = dcrate_dx * crate_uc;")
            w("    const double duc_dno3 = -1.0 * dcrate_dx;")
            w("    const double dco2_dno3 = dcrate_dx * cst;")
            w("    double dno3_dno3 = -1.0 * fno3 * 0.0 + dfno3 * (1.0 - finh);")
            w("    dno3_dno3 = dno3_dno3 * crate_uc * nst;")
            w("    const double dun_dno3 = unc * duc_dno3;")
            # column uc
            column(uc, "dco2_duc", "duc_duc", "dun_duc", True)
            self._Jsub(nh4, uc, "dnh4_duc")
            if unimm >= 0:
                self._Jadd(naq + unimm, uc, "dnh4_duc")
            if sd.nimm_id >= 0:
                self._Jadd(naq + int(sd.nimm_id), uc, "dnh4_duc")
            if no3 >= 0:
                self._Jsub(no3, uc, "dno3_duc")
                if unimm >= 0:
                    self._Jadd(naq + unimm, uc, "dno3_duc")
                if sd.nimm_id >= 0:
                    self._Jadd(naq + int(sd.nimm_id), uc, "dno3_duc")
            nrows(uc, "duc_duc", "dun_duc")
            # column nh4
            column(nh4, "dco2_dnh4", "duc_dnh4", "dun_dnh4", False)
            self._Jsub(nh4, nh4, "dnh4_dnh4 * denL")
            if unimm >= 0:
                self._Jadd(naq + unimm, nh4, "dnh4_dnh4")
            if sd.nimm_id >= 0:
                self._Jadd(naq + int(sd.nimm_id), nh4, "dnh4_dnh4")
            nrows(nh4, "duc_dnh4", "dun_dnh4")
            # column no3
            if no3 >= 0:
                column(no3, "dco2_dno3", "duc_dno3", "dun_dno3", False)
                self._Jsub(no3, no3, "dno3_dno3 * denL")
                if unimm >= 0:
                    self._Jadd(naq + unimm, no3, "dno3_dno3")
                if sd.nimm_id >= 0:
                    self._Jadd(naq + int(sd.nimm_id), no3, "dno3_dno3")
                nrows(no3, "duc_dno3", "dun_dno3")
            w("    }")  # if (!skip)

    def _emit_nitrif(self) -> None:
        """NitrifReact (reaction_sandbox_nitrif.F90:234-502)"""
        naq, w = self.naq, self.w
        nt = self.cfg.nitrif
        nh4, no3, n2o = int(nt.nh4_id), int(nt.no3_id), int(nt.n2o_id)
        w("    double saturation = s.sat;")
        w("    const double L_water = saturation * s.por * 1.0e3;")
        w(f"    const double c_nh4 = tot[{nh4}] * L_water;")
        w("    double feps0, dfeps0;")
        if nt.x0eps > 0.0:
            w(f"    pfrx_sbx::hsmooth(c_nh4, {_lit(nt.x0eps * 10.0)}, {_lit(nt.x0eps)}, feps0, dfeps0);")
        else:
            w("    feps0 = 1.0; dfeps0 = 0.0;")
            w(f"    if (c_nh4 < {_lit(nt.x0eps)}) break;")
        if nh4 >= 0 and no3 >= 0:
            w("    {")
            w(f"    const double f_t = {self.cellconst('exp(0.08 * (s.temp - 25.0))')};")
            w("    saturation = fmax(0.0, fmin(saturation, 1.0));")
            w("    const double f_w = saturation * (1.0 - saturation) / 0.25;")
            w(f"    double t = fmin({_lit(nt.k_nitr_max)} * f_t * f_w * s.vol, 1.0);")
            w("    const double rate = t * (c_nh4 * feps0) * (c_nh4 / (c_nh4 + 4.0));")
            w(f"    res[{nh4}] = res[{nh4}] + rate;")
            w(f"    res[{no3}] = res[{no3}] - rate;")
            w("    t = c_nh4 * c_nh4 / (c_nh4 + 4.0) * dfeps0 + c_nh4 * (c_nh4 + 8.0) / (c_nh4 + 4.0) / (c_nh4 + 4.0) * feps0;")
            w(f"    const double drate = {_lit(nt.k_nitr_max)} * f_t * f_w * s.vol * t;")
            self._Jadd(nh4, nh4, "drate * denL")
            w("    }")
        rho_b = "s.elm_bd_dry" if self.c.elm_pflotran else "1.25e3"
        w(f"    const double rho_b = {rho_b};")
        w("    const double M_2_ug_per_g = " + self.cellconst(f"(14.0067 * 1.0e6) / (s.vol * {rho_b} * 1.e3)") + ";")
        w("    const double c_nh4_ugg = c_nh4 * s.vol * M_2_ug_per_g;")
        if n2o >= 0:
            w("    if (c_nh4_ugg > 3.0) {")
            w(f"    double f_t = {self.cellconst('-0.06 + 0.13 * exp(0.07 * s.temp)')};")
            sat_e = "fmax(0.0, fmin(s.sat, 1.0))" if (nh4 >= 0 and no3 >= 0) else "s.sat"   # `saturation` at this point
            w("    double f_w = " + self.cellconst(f"pfrx_sbx::wfps({sat_e})") + ";")
            self._emit_ph(int(nt.proton_id))
            w("    if (f_t > 0.0 && f_w > 0.0 && f_ph > 0.0) {")
            w("    f_t = fmin(f_t, 1.0); f_w = fmin(f_w, 1.0); f_ph = fmin(f_ph, 1.0);")
            w("    const double ex = exp(-0.0105 * c_nh4_ugg);")
            w(f"    double t = (1.0 - ex) * f_t * f_w * f_ph * {_lit(nt.k_nitr_n2o)};")
            w("    const double rate_n2o = t * (c_nh4 * feps0) * s.vol;")
            w(f"    res[{nh4}] = res[{nh4}] + rate_n2o;")
            w(f"    res[{n2o}] = res[{n2o}] - 0.5 * rate_n2o;")
            if nt.ngasnit_id >= 0:
                w(f"    res[{naq + int(nt.ngasnit_id)}] -= rate_n2o;")
            w("    t = (c_nh4 * dfeps0 + feps0) * (1.0 - ex);")
            w("    t = t + (c_nh4 * feps0) * 0.0105 * M_2_ug_per_g * ex;")
            w(f"    const double drate = t * {_lit(nt.k_nitr_n2o)} * f_t * f_w * f_ph * s.vol;")
            self._Jadd(nh4, nh4, "drate * denL")
            if nt.ngasnit_id >= 0:
                self._Jsub(naq + int(nt.ngasnit_id), nh4, "drate")
            w("    }")
            w("    }")

    def _emit_denitr(self) -> None:
        """DenitrReact (reaction_sandbox_denitr.F90:212-404)"""
        naq, w = self.naq, self.w
        dn = self.cfg.denitr
        no3, n2 = int(dn.no3_id), int(dn.n2_id)
        if n2 < 0:
            return
        w("    const double L_water = s.por * s.sat * 1.e3;")
        bsw = "s.elm_bsw" if self.c.elm_pflotran else "1.0"
        w(f"    const double f_t = {self.cellconst('exp(0.08 * (s.temp - 25.0))')};")
        w("    const double f_w = " + self.cellconst(f"((s.sat > 0.6) ? pow((s.sat - 0.6) / (1.0 - 0.6), {bsw}) : 0.0)") + ";")
        w(f"    const double c_no3 = tot[{no3}] * L_water;")
        w("    double feps0, dfeps0;")
        if dn.x0eps > 0.0:
            w(f"    pfrx_sbx::hsmooth(c_no3, {_lit(dn.x0eps * 10.0)}, {_lit(dn.x0eps)}, feps0, dfeps0);")
        else:
            w("    feps0 = 1.0; dfeps0 = 0.0;")
            w(f"    if (c_no3 <= {_lit(dn.x0eps)}) break;")
        if dn.half_saturation > 0.0:
            w(f"    const double fno3 = sx_monod(c_no3, {_lit(dn.half_saturation)});")
            w(f"    const double dfno3 = sx_dmonod(c_no3, {_lit(dn.half_saturation)});")
        else:
            w("    const double fno3 = 1.0, dfno3 = 0.0;")
        w("    if (f_t > 0.0 && f_w > 0.0) {")
        w(f"    const double rate = {_lit(dn.k_deni_max)} * f_t * f_w * fno3 * (c_no3 * s.vol * feps0);")
        w(f"    res[{no3}] = res[{no3}] + rate;")
        w(f"    res[{n2}] = res[{n2}] - 0.5 * rate;")
        if dn.ngasdeni_id >= 0:
            w(f"    res[{naq + int(dn.ngasdeni_id)}] -= rate;")
        w("    const double t = dfno3 * (c_no3 * s.vol * feps0) + fno3 * (c_no3 * s.vol * dfeps0 + feps0);")
        w(f"    const double drate = {_lit(dn.k_deni_max)} * f_t * f_w * t;")
        self._Jadd(no3, no3, "drate * denL")
        if dn.ngasdeni_id >= 0:
            self._Jsub(naq + int(dn.ngasdeni_id), no3, "drate")
        w("    }")

    def _emit_plantn(self) -> None:
        """PlantNReact (reaction_sandbox_plantn.F90:222-640)"""
        naq, w = self.naq, self.w
        pn = self.cfg.plantn
        nh4, no3, pl = int(pn.nh4_id), int(pn.no3_id), naq + int(pn.plantn_id)
        both = nh4 >= 0 and no3 >= 0
        w("    if (s.sat < 0.01) break;")
        w("    const double L_water = s.sat * s.por * 1.0e3;")
        w("    if (s.temp < -0.1) break;")
        w("    double c_nh4 = 0.0, c_no3 = 0.0, fnh4 = 1.0, dfnh4 = 0.0, fno3 = 1.0, dfno3 = 0.0, finh = 1.0;")
        if both:
            w(f"    c_nh4 = tot[{nh4}] * L_water; c_no3 = tot[{no3}] * L_water;")
            x4, x3 = _lit(pn.x0eps_nh4), _lit(pn.x0eps_no3)
            if pn.inhibition_nh4_no3 > 0.0:
                w(f"    if (c_nh4 > {x4} && c_no3 > {x3}) finh = sx_monod(c_nh4 / c_no3, {_lit(1.0 / pn.inhibition_nh4_no3)});")
                w(f"    else if (c_nh4 > {x4} && c_no3 <= {x3}) finh = 1.0;")
            else:
                w(f"    if (c_nh4 > {x4} && c_no3 <= {x3}) finh = 1.0;")
            w(f"    else if (c_nh4 <= {x4} && c_no3 > {x3}) finh = 0.0;")
            w("    else break;")
        w("    double feps0, dfeps0;")
        for sp, nm, hs, x0 in ((nh4, "nh4", pn.half_saturation_nh4, pn.x0eps_nh4),
                               (no3, "no3", pn.half_saturation_no3, pn.x0eps_no3)):
            if sp < 0:
                continue
            w(f"    c_{nm} = tot[{sp}] * L_water;")
            w(f"    f{nm} = sx_monod(c_{nm}, {_lit(hs)}); df{nm} = sx_dmonod(c_{nm}, {_lit(hs)});")
            if x0 > 0.0:
                w(f"    pfrx_sbx::hsmooth(c_{nm}, {_lit(x0 * 10.0)}, {_lit(x0)}, feps0, dfeps0);")
            else:
                w("    feps0 = 1.0; dfeps0 = 0.0;")
            w(f"    df{nm} = df{nm} * feps0 + f{nm} * dfeps0; f{nm} = f{nm} * feps0;")
        if self.c.elm_pflotran:
            w("    const double demand = fmax(0.0, s.elm_plantndemand * s.vol);")
            w("    if (demand <= 0.0) break;")
        else:
            w("    const double demand = 1.e-2 * s.vol;")
        if pn.plantndemand_id >= 0:
            w(f"    res[{naq + int(pn.plantndemand_id)}] -= demand;")
        w("    if (demand > 0.0) {")
        for sp, nm, fac in ((nh4, "nh4", "finh"), (no3, "no3", "(1.0 - finh)")):
            if sp < 0:
                continue
            cap = f"demand * {fac} * dt" if both else "demand * dt"
            w(f"    {{ const double cap = {cap}; double fcap = 1.0, dfcap = 0.0;")
            w(f"      if (cap > c_{nm} * s.vol) {{ fcap = sx_monod(c_{nm} * s.vol, cap - c_{nm} * s.vol); "
              f"dfcap = sx_dmonod(c_{nm} * s.vol, cap - c_{nm} * s.vol); }}")
            w(f"      df{nm} = df{nm} * fcap + f{nm} * dfcap; f{nm} = f{nm} * fcap; }}")
        w("    }")
        if nh4 >= 0:
            w("    { const double nrate = " + ("demand * fnh4 * finh;" if both else "demand * fnh4;"))
            w(f"      res[{nh4}] = res[{nh4}] + nrate; res[{pl}] = res[{pl}] - nrate;")
            if pn.plantnh4uptake_id >= 0:
                w(f"      res[{naq + int(pn.plantnh4uptake_id)}] -= nrate;")
            w("      const double dn = " + ("demand * (fnh4 * 0.0 + finh * dfnh4);" if both else "demand * dfnh4;"))
            self._Jadd(nh4, nh4, "dn * denL")
            self._Jsub(pl, nh4, "dn")
            if pn.plantnh4uptake_id >= 0:
                self._Jsub(naq + int(pn.plantnh4uptake_id), nh4, "dn")
            w("    }")
        if no3 >= 0:
            w("    { const double nrate = " + ("demand * fno3 * (1.0 - finh);" if both else "demand * fno3;"))
            w(f"      res[{no3}] = res[{no3}] + nrate; res[{pl}] = res[{pl}] - nrate;")
            if pn.plantno3uptake_id >= 0:
                w(f"      res[{naq + int(pn.plantno3uptake_id)}] -= nrate;")
            w("      const double dn = " + ("demand * (dfno3 * (1.0 - finh) + fno3 * (-1.0 * 0.0));" if both
                                            else "demand * dfno3;"))
            self._Jadd(no3, no3, "dn * denL")
            self._Jsub(pl, no3, "dn")
            if pn.plantno3uptake_id >= 0:
                self._Jsub(naq + int(pn.plantno3uptake_id), no3, "dn")
            w("    }")

    def _emit_langmuir(self) -> None:
        """LangmuirReact (reaction_sandbox_langmu.F90:183-330)"""
        naq, w = self.naq, self.w
        lg = self.cfg.langmuir
        aq, sb = int(lg.aq_id), naq + int(lg.sorb_id)
        smax, keq, kk = _lit(lg.s_max), _lit(lg.k_equilibrium), _lit(lg.k_kinetic)
        w("    const double Lwater = s.vol * 1000.0 * s.por * s.sat;")
        w(f"    const double c_aq = tot[{aq}], c_sorb = c[{sb}];")
        w("    double rate, drate_daq, drate_dsorb;")
        w(f"    if ({smax} < c_sorb) {{")
        w(f"      rate = ({smax} - c_sorb) * s.vol / dt; drate_dsorb = -1.0 / dt; drate_daq = 0.0;")
        w("    } else {")
        w(f"      const double c_aq_eq = 0.999 * c_sorb / ({smax} - 0.999 * c_sorb) / {keq};")
        w(f"      rate = {kk} * (c_aq - c_aq_eq) * Lwater;")
        w(f"      double t = -{kk} / {keq} * Lwater / s.vol;")
        w(f"      drate_dsorb = t * {smax} / ({smax} - 0.999 * c_sorb) / ({smax} - 0.999 * c_sorb);")
        w(f"      drate_daq = {kk};")
        w("      if (rate > 0.0) {")
        w(f"        double ratecap = 0.999 * ({smax} - c_sorb) * s.vol / dt;")
        w("        if (ratecap < rate) {")
        w("          const double fcap = ratecap / rate; t = -0.999 / dt;")
        w("          const double dfcap = (ratecap * drate_dsorb - rate * t) / rate / rate;")
        w("          drate_dsorb = fcap * drate_dsorb + rate * dfcap; rate = rate * fcap;")
        w("        }")
        w("        ratecap = 0.999 * (c_aq - c_aq_eq) * Lwater / dt;")
        w("        if (ratecap < rate) {")
        w(f"          const double fcap = ratecap / rate; t = -0.999 / {keq} * Lwater / s.vol / dt;")
        w(f"          t = t * {smax} / ({smax} - 0.999 * c_sorb) / ({smax} - 0.999 * c_sorb);")
        w("          double dfcap = (ratecap * drate_dsorb - rate * t) / rate / rate;")
        w("          drate_dsorb = fcap * drate_dsorb + rate * dfcap;")
        w("          t = 0.999 / dt; dfcap = (ratecap * drate_daq - rate * t) / rate / rate;")
        w("          drate_daq = fcap * drate_daq + rate * dfcap; rate = rate * fcap;")
        w("        }")
        w("      }")
        w("    }")
        w(f"    res[{aq}] = res[{aq}] + rate; res[{sb}] = res[{sb}] - rate;")
        self._Jadd(aq, aq, "drate_daq * denL")
        self._Jsub(sb, aq, "drate_daq")
        self._Jadd(aq, sb, "drate_dsorb")
        self._Jsub(sb, sb, "drate_dsorb")

    def _emit_clm_cn(self) -> None:
        """CLM_CN_React (reaction_sandbox_clm_cn.F90:468-787), one straight-line block per reaction"""
        c, a, naq = self.c, self.a, self.naq
        self.w("  const double temp_K = s.temp + 273.15;")
        self.w("  if (!(temp_K > 227.15)) break;")
        self.w("  const double F_t = " + self.cellconst("exp(308.56 * (1.408054069e-2 - 1.0 / ((s.temp + 273.15) - 227.13)))") + ";")
        self.w("  const double F_theta = " + self.cellconst("log(0.01 / fmax(0.01, s.sat)) * -2.17147241e-1") + ";")
        self.w("  const double cinh = F_t * F_theta;")
        iC = naq + int(c.clmcn_C_species_id)
        iN = naq + int(c.clmcn_N_species_id)
        J = self.J
        for r in range(c.clmcn_nrxn):
            up = int(a["clmcn_upstream_pool_id"][r])
            down = int(a["clmcn_downstream_pool_id"][r])
            resp = float(a["clmcn_respiration_fraction"][r])
            inhib = float(a["clmcn_inhibition_constant"][r])
            litter = int(a["clmcn_pool_nspec"][up]) != 1
            iCu = naq + int(a["clmcn_pool_C_id"][up])
            iNu = naq + int(a["clmcn_pool_N_id"][up]) if litter else -1
            self.w("  {")
            self.w(f"    const double k = {_lit(float(a['clmcn_rate_constant'][r]))} * s.vol * cinh;")
            if litter:
                self.w(f"    const double cn_up = c[{iCu}] / c[{iNu}];")
            else:
                self.w(f"    const double cn_up = {_lit(float(a['clmcn_CN_ratio'][up]))};")
            self.w("    const double st_upN = 1.0 / cn_up;")
            if down >= 0:
                iD = naq + int(a["clmcn_pool_C_id"][down])
                self.w(f"    const double st_dn = {_lit((1.0 - resp) * 1.0)};")
                self.w(f"    const double cn_dn = {_lit(float(a['clmcn_CN_ratio'][down]))};")
            else:
                iD = -1
                self.w("    const double st_dn = 0.0;")
                self.w("    const double cn_dn = 1.0;")
            self.w(f"    const double st_C = {_lit(resp * 1.0)};")
            self.w("    const double st_N = st_upN - st_dn / cn_dn;")
            if inhib > 1.0e-40:
                self.w("    const bool inh = st_N < 0.0;")
                self.w(f"    const double tr = c[{iN}] + {_lit(inhib)};")
                self.w(f"    const double Ninh = inh ? c[{iN}] / tr : 1.0;")
                self.w(f"    const double dNinh = inh ? {_lit(inhib)} / (tr * tr) : 0.0;")
            else:
                self.w("    const bool inh = false;")
                self.w("    const double Ninh = 1.0, dNinh = 0.0;")
            self.w(f"    const double rate = c[{iCu}] * k * Ninh;")
            self.w(f"    res[{iC}] = res[{iC}] - st_C * rate;")
            self.w(f"    res[{iN}] = res[{iN}] - st_N * rate;")
            self.w(f"    res[{iCu}] = res[{iCu}] - (-1.0) * 1.0 * rate;")
            if litter:
                self.w(f"    res[{iNu}] = res[{iNu}] - (-1.0) * st_upN * rate;")
            if iD >= 0:
                self.w(f"    res[{iD}] = res[{iD}] - st_dn * rate;")
            self.w("    const double drate = k * Ninh;")
            self.w(f"    const double dinh = c[{iCu}] * k * dNinh;")
            self.w(f"    {J(iCu, iCu)} = {J(iCu, iCu)} - (-1.0) * 1.0 * drate;")
            self.w(f"    if (inh) {J(iCu, iN)} = {J(iCu, iN)} - (-1.0) * 1.0 * dinh;")
            if iD >= 0:
                self.w(f"    {J(iD, iCu)} = {J(iD, iCu)} - st_dn * drate;")
                self.w(f"    if (inh) {J(iD, iN)} = {J(iD, iN)} - st_dn * dinh;")
            if litter:
                self.w(f"    {J(iNu, iCu)} = {J(iNu, iCu)} - (-1.0) * st_upN * drate;")
                self.w(f"    if (inh) {J(iNu, iN)} = {J(iNu, iN)} - (-1.0) * st_upN * dinh;")
                self.w(f"    {J(iNu, iCu)} = {J(iNu, iCu)} - (-1.0) * (-1.0) * c[{iNu}] / c[{iCu}] * k * Ninh;")
                self.w(f"    {J(iNu, iNu)} = {J(iNu, iNu)} - (-1.0) * k * Ninh;")
                self.w(f"    {J(iN, iCu)} = {J(iN, iCu)} - (-1.0) * c[{iNu}] / c[{iCu}] * k * Ninh;")
                self.w(f"    {J(iN, iNu)} = {J(iN, iNu)} - k * Ninh;")
            self.w(f"    {J(iC, iCu)} = {J(iC, iCu)} - st_C * drate;")
            self.w(f"    {J(iN, iCu)} = {J(iN, iCu)} - st_N * drate;")
            self.w("    if (inh) {")
            self.w(f"      {J(iC, iN)} = {J(iC, iN)} - st_C * dinh;")
            self.w(f"      {J(iN, iN)} = {J(iN, iN)} - st_N * dinh;")
            self.w("    }")
            self.w("  }")

    # ------------------------------------------------------------------ whole file
    def _gen_body(self, nsbx: int) -> List[str]:
        self.out = []
        self.kc = []
        self.gen_tables()
        self.gen_activity()
        self.gen_rtotal()
        self.gen_sorption()
        self.gen_minerals()
        if has_kinetic3(self.cfg):
            self.gen_kinetic()
        if nsbx > 0:
            self.gen_sandbox()
        self.gen_cell_constants()
        return self.out

    def gen_rowonly(self) -> None:
        """u_t = (res_t - sum_j J_tj u_j) / J_tt for the row-only species, after the dense solve
        (log formulation: columns scaled by c_j like RSolve does, reaction.F90:5493-5497).  The
        reference eliminates these rows inside its LU; the result is the same up to rounding."""
        naq = self.naq
        self.w("__device__ __forceinline__ bool spec_rowonly(double *W, double (&res)[SPEC_N], const double (&c)[SPEC_N],")
        self.w("    const SpecCell &s, double dt) {")
        self.w("  bool ok = true;")
        for t in self.rowonly:
            self.w("  {")
            if t < naq:
                self.w("    double Jd = (1.0 * (s.den_kg * 1.e-3)) * (s.por * s.sat * 1000.0 * s.vol / dt);")
            else:
                self.w("    double Jd = s.vol / dt;")
            self.w("    if (s.dry) Jd = 1.0;")
            if (t, t) in self.ropairs:
                self.w(f"    Jd = Jd + SW(SPEC_OFF_RO + {self.ropairs[(t, t)]});")
            self.w(f"    double r = res[{t}];")
            for (i, j), k in self.ropairs.items():
                if i == t and j != t:
                    cj = f" * c[{j}]" if self.c.use_log_formulation else ""
                    self.w(f"    r = r - (SW(SPEC_OFF_RO + {k}){cj}) * res[{j}];")
            self.w(f"    const double a = Jd{' * c[%d]' % t if self.c.use_log_formulation else ''};")
            self.w("    if (!(fabs(a) > 0.0)) ok = false;")
            self.w(f"    res[{t}] = sx_div(r, a);")
            self.w("  }")
        self.w("  (void)W; (void)c; (void)s; (void)dt;")
        self.w("  return ok;")
        self.w("}")
        self.w()

    def gen_sparse_solve(self) -> None:
        """The Newton system of the dense core by SPARSE elimination in a STATIC order: the network's
        structure (self._nz) is known when the kernel is generated, so the pivot order is chosen here
        (diagonal pivots, least fill first), every multiply-add is emitted with literal addresses and
        entries that are structurally zero cost nothing.  RSolve's row scaling and the implicit-scaled
        partial pivoting of LUDecomposition (reaction.F90:5485-5498, utility.F90:597-688) exist to make
        elimination safe for an arbitrary matrix; here safety is CHECKED instead: every multiplier must
        satisfy |l| <= 64 (threshold pivoting).  If one does not -- or is NaN -- the non-zeros are put
        back and the caller runs the reference's dense algorithm.  The solution is the same up to
        rounding (row scaling does not change it; the log formulation's column scaling by c_j is the
        substitution y = c * delta, undone at the end)."""
        nc = self.nc
        self.w("__device__ __forceinline__ bool spec_solve_sparse(double *W, double (&res)[SPEC_N], const double (&c)[SPEC_N]) {")
        if nc == 0:
            self.w("  (void)W; (void)res; (void)c; return true;\n}\n")
            return
        sp = {ci: s_ for s_, ci in self.cpos.items()}
        nz = {(self.cpos[i], self.cpos[j]) for (i, j) in self._nz if i in self.cpos and j in self.cpos}
        nz |= {(k, k) for k in range(nc)}
        cur = set(nz)
        left = list(range(nc))
        order = []
        fills = []
        while left:
            best = None
            for k in left:
                row = [j for j in left if j != k and (k, j) in cur]
                col = [i for i in left if i != k and (i, k) in cur]
                f = sum(1 for i in col for j in row if (i, j) not in cur)
                cost = (f, len(row) * len(col), k)
                if best is None or cost < best[0]:
                    best = (cost, k, row, col)
            _, k, row, col = best
            for i in col:
                for j in row:
                    if (i, j) not in cur:
                        cur.add((i, j))
                        fills.append((i, j))
            order.append((k, row, col))
            left.remove(k)
        self.sparse_order = [k for k, _, _ in order]
        self.sparse_fill = len(fills)
        nzl = sorted(nz)
        # PFRX_SPEC_SPARSE_THRESHOLD=0 makes every solve fall back: how the tests keep the dense path exercised
        thr = _lit(float(os.environ.get("PFRX_SPEC_SPARSE_THRESHOLD", "64.0")))
        self.w(f"  double sav[{len(nzl)}];  // the non-zeros as assembled, for the dense fall-back")
        for q, (i, j) in enumerate(nzl):
            self.w(f"  sav[{q}] = W[JX({i}, {j})];")
        for ci in range(nc):
            self.w(f"  double b{ci} = res[{sp[ci]}];")
        self.w("  bool ok = true;")
        fillset = set(fills)
        touched = set()
        nmul = 0
        for k, row, col in order:
            self.w(f"  const double ip{k} = 1.0 / W[JX({k}, {k})];")
            for j in row:
                self.w(f"  const double u{k}_{j} = W[JX({k}, {j})];")
            for i in col:
                self.w(f"  {{ const double l = W[JX({i}, {k})] * ip{k};")
                self.w(f"    ok = ok && (fabs(l) <= {thr});")
                for j in row:
                    if (i, j) in fillset and (i, j) not in touched:
                        self.w(f"    W[JX({i}, {j})] = -(l * u{k}_{j});")
                        touched.add((i, j))
                    else:
                        self.w(f"    W[JX({i}, {j})] -= l * u{k}_{j};")
                    nmul += 1
                self.w(f"    b{i} -= l * b{k};")
                self.w("  }")
        self.sparse_muladds = nmul
        self.w("  if (!ok) {")
        for q, (i, j) in enumerate(nzl):
            self.w(f"    W[JX({i}, {j})] = sav[{q}];")
        for (i, j) in fills:
            self.w(f"    W[JX({i}, {j})] = 0.0;")
        self.w("    return false;")
        self.w("  }")
        for k, row, col in reversed(order):
            expr = f"b{k}"
            for j in row:
                expr += f" - u{k}_{j} * y{j}"
            self.w(f"  const double y{k} = ({expr}) * ip{k};")
        for ci in range(nc):
            if self.c.use_log_formulation:
                self.w(f"  res[{sp[ci]}] = y{ci} / c[{sp[ci]}];")
            else:
                self.w(f"  res[{sp[ci]}] = y{ci};")
        self.w("  (void)c;")
        self.w("  return true;")
        self.w("}")
        self.w()

    def source(self) -> str:
        c = self.c
        n = self.n
        nsbx = (int(c.clmcn_nrxn > 0) + int(bool(c.somdec)) + int(bool(c.nitrif)) + int(bool(c.denitr))
                + int(bool(c.plantn)) + int(bool(c.langmuir)))
        # pass 1: every species that occurs in a reaction is in the matrix; record which Jacobian
        # entries the network really has
        self._layout([])
        self._pairs = set()
        self._recording = True
        self._gen_body(nsbx)
        self._recording = False
        rowonly = []
        if not os.environ.get("PFRX_SPEC_NO_ROWONLY"):
            cols = {j for (i, j) in self._pairs if i != j}
            rowonly = [t for t in self.used if t not in cols]
        # pass 2: the final layout
        self._layout(rowonly)
        body = self._gen_body(nsbx)
        self.out = []
        self.gen_rowonly()
        sparse = not self.loop_lu and os.environ.get("PFRX_SPEC_SOLVER", "sparse") != "dense"
        if sparse:
            self.gen_sparse_solve()
        body = body + self.out
        nro = len(self.ropairs)
        if self.loop_lu:
            slots = self.nc * (self.nc + 2) + self.nc
        else:
            slots = self.nc * (self.nc + 1) + 2 * n
        slots = max(1, slots + (0 if self.act_upd else self.ncx) + nro)
        per_warp = slots * 32 * 8 + 1024  # + the per-block reservation when a block is one warp
        if slots * 32 * 8 > 160 * 1024:
            threads = 32
            minblocks = 1
        elif n > 8:
            threads = 32
            # measured (profiles/r01_occupancy_sweep_spec.txt): one warp per scheduler is the
            # optimum for these 255-register kernels; a fifth warp per SM costs 12 %
            maxw = int(os.environ.get("PFRX_SPEC_MAXWARPS", "4"))
            minblocks = max(1, min(maxw, (228 * 1024) // per_warp))
        else:
            threads = 128
            minblocks = max(1, min(4, (228 * 1024) // (slots * 128 * 8 + 1024)))
        if self.lockstep and threads == 32 and not self.onewarp:
            # the warps that shared an SM as separate blocks become one block that votes
            # (PFRX_SPEC_LOCKSTEP_WARPS = 2: two blocks of two warps, for measurements)
            lw = int(os.environ.get("PFRX_SPEC_LOCKSTEP_WARPS", "0")) or minblocks
            lw = max(1, min(lw, minblocks))
            threads = 32 * lw
            minblocks = max(1, minblocks // lw)
        self.threads, self.minblocks, self.slots = threads, minblocks, slots
        self.out = []
        self.w("// generated by pflotran_elm_interface_b200/specialize.py -- do not edit")
        self.w(f"#define SPEC_N {n}")
        self.w(f"#define SPEC_NAQ {self.naq}")
        self.w(f"#define SPEC_NC {self.nc}")
        self.w(f"#define SPEC_NRO {nro}")
        self.w(f"#define SPEC_NROSPEC {len(self.rowonly)}")
        self.w(f"#define SPEC_NCX {self.ncx}")
        self.w(f"#define SPEC_NCLS {len(self.cls)}")
        self.w(f"#define SPEC_NKIN {c.nkinmnrl}")
        self.w(f"#define SPEC_NSRFRXN {c.nsrfcplxrxn}")
        self.w(f"#define SPEC_NSRFCPLX {c.nsrfcplx}")
        self.w(f"#define SPEC_NEQSR {c.neqsrfcplxrxn}")
        self.w(f"#define SPEC_NMR {c.nkinmrsrfcplxrxn}")
        if c.nkinmrsrfcplxrxn > 0:
            if self.lockstep:
                raise ValueError("multirate sorption is generated for the one-warp skeleton (style straight) only")
            mp = " : ".join(f"q == {q} ? {int(v)}" for q, v in enumerate(self.a["kinmr_rate_ptr"]))
            self.w("__host__ __device__ constexpr int spec_mr_ptr(int q) { return " + mp + " : 0; }")
            self.w("static __device__ const double spec_mr_rate_tab[] = {" + ", ".join(_lit(float(v)) for v in self.a["kinmr_rate"]) + "};")
            self.w("static __device__ const double spec_mr_frac_tab[] = {" + ", ".join(_lit(float(v)) for v in self.a["kinmr_frac"]) + "};")
        nnc = (len(self.a["somdec_upstream_nc"]) + len(self.a["somdec_downstream_nc"])) if c.somdec else 0
        self.w(f"#define SPEC_NCLM {c.clmcn_nrxn}")
        self.w(f"#define SPEC_NSBX {nsbx}")
        self.w(f"#define SPEC_NKIN3 {int(has_kinetic3(self.cfg))}")
        self.w(f"#define SPEC_NDTP {len(self.dtp)}")
        self.w(f"#define SPEC_NDSP {len(self.dsp)}")
        self.w(f"#define SPEC_NKC {len(self.kc)}")
        self.w(f"#define SPEC_NIONX {c.neqionxrxn}")
        self.w(f"#define SPEC_NIXCAT {int(self.a['eqionx_ptr'][c.neqionxrxn]) if c.neqionxrxn else 0}")
        self.w(f"#define SPEC_NSORB {c.neqsrfcplxrxn + c.neqionxrxn + c.neqkdrxn + c.neqdynamickdrxn}")
        self.w(f"#define SPEC_NNC {nnc}")
        self.w(f"#define SPEC_ELM {int(bool(c.elm_pflotran))}")
        if nnc:
            nc0 = list(self.a["somdec_upstream_nc"]) + list(self.a["somdec_downstream_nc"])
            self.w("static __device__ const double spec_nc0_tab[] = {" + ", ".join(_lit(float(v)) for v in nc0) + "};")
        self.w(f"#define SPEC_USE_LOG {int(c.use_log_formulation)}")
        self.w(f"#define SPEC_ACT_UPD {int(self.act_upd)}")
        self.w(f"#define SPEC_USE_ACT_H2O {int(c.use_activity_h2o)}")
        self.w(f"#define SPEC_SIG {signature(self.cfg)}ull")
        self.w(f"#define SPEC_THREADS {threads}")
        self.w(f"#define SPEC_FASTMATH {int(os.environ.get('PFRX_SPEC_FASTMATH', '1'))}")
        self.w(f"#define SPEC_LOOP_LU {int(self.loop_lu)}")
        if sparse:
            self.w("#define SPEC_SPARSE_LU 1")
        self.w(f"#define SPEC_LOCKSTEP {int(self.lockstep)}")
        self.w(f"#define SPEC_REFILL {int(self.refill)}")
        self.w(f"#define SPEC_MINBLOCKS {minblocks}")
        cm = " : ".join(f"i == {sp} ? {ci}" for sp, ci in self.cpos.items())
        so = " : ".join(f"ci == {ci} ? {sp}" for sp, ci in self.cpos.items())
        ro = " || ".join(f"i == {t}" for t in self.rowonly)
        self.w("__host__ __device__ constexpr int spec_cmap(int i) { return " + (cm + " : -1" if cm else "-1") + "; }")
        self.w("__host__ __device__ constexpr int spec_sp_of(int ci) { return " + (so + " : 0" if so else "0") + "; }")
        self.w("__host__ __device__ constexpr bool spec_is_rowonly(int i) { return " + (ro if ro else "false") + "; }")
        # structure of the dense core: bit cj of row ci is set when the generated code can make J(ci, cj)
        # non-zero (the diagonal always: accumulation, and the identity of a dry cell); RSolve's row
        # scaling skips the others, which hold exact zeros until the decomposition fills them in
        masks = []
        for sp_i, ci in self.cpos.items():
            m = 1 << ci
            for sp_j, cj in self.cpos.items():
                if (sp_i, sp_j) in self._nz:
                    m |= 1 << cj
            masks.append((ci, m))
        if os.environ.get("PFRX_SPEC_DENSE_SCALING") or self.nc > 60:
            masks = [(ci, (1 << max(self.nc, 1)) - 1) for ci, _ in masks]
        jm = " : ".join(f"ci == {ci} ? {m}ull" for ci, m in masks)
        self.w("__host__ __device__ constexpr unsigned long long spec_jrow_mask(int ci) { return "
               + (jm + " : 0ull" if jm else "0ull") + "; }")
        self.w("__host__ __device__ constexpr bool spec_jnz(int ci, int cj) { return (spec_jrow_mask(ci) >> cj) & 1ull; }")
        self.jnz_density = (sum(bin(m).count("1") for _, m in masks) / float(max(1, self.nc * self.nc)))
        pc = " : ".join(f"i == {i} ? {q}" for i, q in enumerate(self.pri_cls) if q >= 0)
        self.w("__host__ __device__ constexpr int spec_pri_cls(int i) { return " + (pc + " : -1" if pc else "-1") + "; }")
        self.w("__device__ __forceinline__ double spec_cx_z2(int k);")
        self.w("__device__ __forceinline__ int spec_cx_cls(int k);")
        self.w("__device__ __forceinline__ double spec_mn_vol(int m);")
        self.w('#include "pfrx_spec.cuh"')
        self.w()
        return "\n".join(self.out + body) + "\n"


def default_variant(cfg: abi.ReactionConfig) -> Tuple[int, str]:
    """(warps per 32 cells, style): PFRX_SPEC_VARIANT=<style letter>1 selects a skeleton by hand
    (s straight, k lock-step, q refill, w refill in independent warps, l / m / p with the rolled
    dense solve of form 1); the default is the lock-step skeleton for the networks of form 2 (two
    warps per scheduler share one instruction stream) and the nested-loop one otherwise.
    ChemistryStep.autotune times the built variants on the actual state."""
    env = os.environ.get("PFRX_SPEC_VARIANT")
    if env:
        return int(env[1:]), VARIANT_STYLES[env[0]]
    if uses_form2(cfg, 1, "lockstep"):
        return 1, "lockstep"
    return 1, "straight"


def generate_source(cfg: abi.ReactionConfig, warps: Optional[int] = None, style: Optional[str] = None) -> str:
    warps, style = _variant(cfg, warps, style)
    if uses_form2(cfg, warps, style):
        from . import specialize2

        return specialize2.generate_source2(cfg, style)
    if warps != 1:
        raise ValueError("one warp per 32 cells is the only layout (the multi-warp variants of round 1 measured "
                         "4-7x slower and were removed)")
    g = _Gen(cfg)
    g.loop_lu = style in ("looplu", "klooplu", "refill_looplu")
    g.lockstep = style in ("lockstep", "klooplu", "refill", "refill_looplu", "refill_warp")
    g.refill = style in ("refill", "refill_looplu", "refill_warp")
    g.onewarp = style == "refill_warp"
    return g.source()


def _stamp(src: str) -> str:
    import hashlib

    h = hashlib.sha1(src.encode())
    for d in ("pfrx_fastmath.cuh", "pfrx_spec.cuh", "pfrx_spec2.cuh", "pfrx_types.cuh",
              "pfrx_sandbox.cuh"):
        with open(os.path.join(CSRC, d), "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def build(cfg: abi.ReactionConfig, force: bool = False, verbose: bool = False, warps: Optional[int] = None,
          style: Optional[str] = None) -> str:
    """generate + ``nvcc -cubin``; returns the cubin path.  The cubin is cached
    under csrc/_spec/ by configuration signature, with a stamp of the generated
    source and the headers it includes."""
    os.makedirs(OUT, exist_ok=True)
    cubin = cubin_path(cfg, warps, style)
    base = cubin[:-len(".cubin")]
    cu, stamp_file = base + ".cu", base + ".stamp"
    src = generate_source(cfg, warps, style)
    stamp = _stamp(src)
    if (not force and os.path.exists(cubin) and os.path.exists(stamp_file)
            and open(stamp_file).read().strip() == stamp):
        return cubin
    with open(cu, "w") as f:
        f.write(src)
    cmd = [os.environ.get("NVCC", "nvcc"), "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3",
           "-std=c++17", "-cubin", "-I", CSRC, "-Xptxas", "-v", "-o", cubin, cu]
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    with open(base + ".log", "w") as f:
        f.write(" ".join(cmd) + "\n" + p.stdout)
    if p.returncode != 0:
        raise RuntimeError("nvcc failed for the specialised kernel:\n" + p.stdout[-4000:])
    with open(stamp_file, "w") as f:
        f.write(stamp + "\n")
    if verbose:
        print(p.stdout)
    return cubin


def main(argv=None) -> int:
    """``python -m pflotran_elm_interface_b200.specialize <dump> [--styles s,k,q,w] [--out DIR]``:
    generate and compile the specialised cubins for a configuration written by ``pfrx_config_dump`` /
    ``pfrx_config_write`` (include/pfrx.h).  Prints one cubin path per line; the host attaches one
    with ``pfrx_load_specialized``."""
    import argparse

    ap = argparse.ArgumentParser(prog="python -m pflotran_elm_interface_b200.specialize")
    ap.add_argument("dump")
    ap.add_argument("--styles", default="default", help="comma-separated: default, straight, lockstep, refill, refill_warp")
    ap.add_argument("--out", default=None, help="directory for the cubins (default: csrc/_spec of the package)")
    ap.add_argument("--force", action="store_true")
    a = ap.parse_args(argv)
    cfg = abi.ReactionConfig.from_dump(a.dump)
    ok, why = supported(cfg)
    if not ok:
        print("not covered by the code generator: " + why)
        return 2
    sig = signature(cfg)
    if cfg.dump_signature is not None and cfg.dump_signature != sig:
        print(f"signature mismatch: file {cfg.dump_signature:016x}, computed {sig:016x}")
        return 3
    global OUT
    if a.out:
        OUT = os.path.abspath(a.out)
    for st in a.styles.split(","):
        path = build(cfg, force=a.force) if st == "default" else build(cfg, force=a.force, warps=1, style=st)
        print(path)
    return 0


if __name__ == "__main__":
    import sys

    sys.exit(main())
