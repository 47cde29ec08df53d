"""ctypes mirror of ``include/pfrx.h`` and the flattening of a
:class:`~.chem.ReactionNetwork` into ``pfrx_config``.

The structures here are the C ABI itself (same field order as the header); the
same definitions are used by the tests to drive the CPU oracle, which takes the
identical structs.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional

import numpy as np

from . import chem as _chem

PFRX_ABI_VERSION = 9
PFRX_MAX_NCOMP = 32

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)


class PfrxConfig(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32),
        ("naqcomp", C.c_int32),
        ("nimcomp", C.c_int32),
        ("use_full_geochemistry", C.c_int32),
        ("use_log_formulation", C.c_int32),
        ("use_total_as_guess", C.c_int32),
        ("use_isothermal", C.c_int32),
        ("act_coef_update_frequency", C.c_int32),
        ("act_coef_update_algorithm", C.c_int32),
        ("use_activity_h2o", C.c_int32),
        ("h2o_aq_id", C.c_int32),
        ("maximum_reaction_iterations", C.c_int32),
        ("maximum_reaction_cuts", C.c_int32),
        ("max_dlnC_rreact", C.c_double),
        ("max_relative_change_tolerance", C.c_double),
        ("max_residual_tolerance", C.c_double),
        ("max_rel_residual_tolerance", C.c_double),
        ("rt_min_saturation", C.c_double),
        ("debyeA", C.c_double),
        ("debyeB", C.c_double),
        ("debyeBdot", C.c_double),
        ("primary_spec_Z", c_double_p),
        ("primary_spec_a0", c_double_p),
        ("neqcplx", C.c_int32),
        ("eqcplx_ptr", c_int32_p),
        ("eqcplx_specid", c_int32_p),
        ("eqcplx_stoich", c_double_p),
        ("eqcplx_h2ostoich", c_double_p),
        ("eqcplx_logK", c_double_p),
        ("eqcplx_logKcoef", c_double_p),
        ("eqcplx_Z", c_double_p),
        ("eqcplx_a0", c_double_p),
        ("nkinmnrl", C.c_int32),
        ("kinmnrl_ptr", c_int32_p),
        ("kinmnrl_specid", c_int32_p),
        ("kinmnrl_stoich", c_double_p),
        ("kinmnrl_h2ostoich", c_double_p),
        ("kinmnrl_logK", c_double_p),
        ("kinmnrl_logKcoef", c_double_p),
        ("kinmnrl_molar_vol", c_double_p),
        ("kinmnrl_rate_constant", c_double_p),
        ("kinmnrl_activation_energy", c_double_p),
        ("kinmnrl_affinity_threshold", c_double_p),
        ("kinmnrl_rate_limiter", c_double_p),
        ("kinmnrl_irreversible", c_int32_p),
        ("kinmnrl_Temkin_const", c_double_p),
        ("kinmnrl_min_scale_factor", c_double_p),
        ("kinmnrl_affinity_power", c_double_p),
        ("kinmnrl_num_prefactors", c_int32_p),
        ("kinmnrl_pref_nspec", c_int32_p),
        ("kinmnrl_prefactor_id", c_int32_p),
        ("kinmnrl_pref_alpha", c_double_p),
        ("kinmnrl_pref_beta", c_double_p),
        ("kinmnrl_pref_atten_coef", c_double_p),
        ("kinmnrl_pref_rate", c_double_p),
        ("kinmnrl_pref_activation_energy", c_double_p),
        ("nsrfcplxrxn", C.c_int32),
        ("nsrfcplx", C.c_int32),
        ("srfcplxrxn_ptr", c_int32_p),
        ("srfcplxrxn_to_complex", c_int32_p),
        ("srfcplxrxn_surf_type", c_int32_p),
        ("srfcplxrxn_to_surf", c_int32_p),
        ("srfcplxrxn_site_density", c_double_p),
        ("srfcplxrxn_stoich_flag", c_int32_p),
        ("srfcplx_ptr", c_int32_p),
        ("srfcplx_specid", c_int32_p),
        ("srfcplx_stoich", c_double_p),
        ("srfcplx_h2ostoich", c_double_p),
        ("srfcplx_free_site_stoich", c_double_p),
        ("srfcplx_logK", c_double_p),
        ("srfcplx_logKcoef", c_double_p),
        ("neqsrfcplxrxn", C.c_int32),
        ("eqsrfcplxrxn_to_srfcplxrxn", c_int32_p),
        ("nkinmrsrfcplxrxn", C.c_int32),
        ("kinmrsrfcplxrxn_to_srfcplxrxn", c_int32_p),
        ("kinmr_rate_ptr", c_int32_p),
        ("kinmr_rate", c_double_p),
        ("kinmr_frac", c_double_p),
        ("neqionxrxn", C.c_int32),
        ("eqionx_ptr", c_int32_p),
        ("eqionx_cationid", c_int32_p),
        ("eqionx_k", c_double_p),
        ("eqionx_CEC", c_double_p),
        ("eqionx_to_surf", c_int32_p),
        ("eqionx_Z_flag", c_int32_p),
        ("neqkdrxn", C.c_int32),
        ("ikd_units", C.c_int32),
        ("eqkd_specid", c_int32_p),
        ("eqkd_type", c_int32_p),
        ("eqkd_mineral", c_int32_p),
        ("eqkd_coeff", c_double_p),
        ("eqkd_langmuir_b", c_double_p),
        ("eqkd_freundlich_n", c_double_p),
        ("neqdynamickdrxn", C.c_int32),
        ("eqdynamickd_specid", c_int32_p),
        ("eqdynamickd_refspecid", c_int32_p),
        ("eqdynamickd_refspechigh", c_double_p),
        ("eqdynamickd_low", c_double_p),
        ("eqdynamickd_high", c_double_p),
        ("eqdynamickd_power", c_double_p),
        ("clmcn_nrxn", C.c_int32),
        ("clmcn_npool", C.c_int32),
        ("clmcn_C_species_id", C.c_int32),
        ("clmcn_N_species_id", C.c_int32),
        ("clmcn_CN_ratio", c_double_p),
        ("clmcn_pool_nspec", c_int32_p),
        ("clmcn_pool_C_id", c_int32_p),
        ("clmcn_pool_N_id", c_int32_p),
        ("clmcn_upstream_pool_id", c_int32_p),
        ("clmcn_downstream_pool_id", c_int32_p),
        ("clmcn_rate_constant", c_double_p),
        ("clmcn_respiration_fraction", c_double_p),
        ("clmcn_inhibition_constant", c_double_p),
        ("somdec", C.c_void_p),
        ("nitrif", C.c_void_p),
        ("denitr", C.c_void_p),
        ("plantn", C.c_void_p),
        ("langmuir", C.c_void_p),
        ("cndegas", C.c_void_p),
        ("calcite", C.c_void_p),
        ("radon", C.c_void_p),
        ("nsandbox", C.c_int32),
        ("sandbox_list", c_int32_p),
        ("elm_pflotran", C.c_int32),
        ("ngeneral_rxn", C.c_int32),
        ("general_ptr", c_int32_p),
        ("general_specid", c_int32_p),
        ("general_stoich", c_double_p),
        ("general_fwd_ptr", c_int32_p),
        ("general_fwd_specid", c_int32_p),
        ("general_fwd_stoich", c_double_p),
        ("general_bwd_ptr", c_int32_p),
        ("general_bwd_specid", c_int32_p),
        ("general_bwd_stoich", c_double_p),
        ("general_kf", c_double_p),
        ("general_kr", c_double_p),
        ("nradiodecay_rxn", C.c_int32),
        ("radiodecay_ptr", c_int32_p),
        ("radiodecay_specid", c_int32_p),
        ("radiodecay_stoich", c_double_p),
        ("radiodecay_forward_specid", c_int32_p),
        ("radiodecay_kf", c_double_p),
        ("nimmobile_decay_rxn", C.c_int32),
        ("immobile_decay_specid", c_int32_p),
        ("immobile_decay_constant", c_double_p),
        ("nmicrobial_rxn", C.c_int32),
        ("microbial_concentration_units", C.c_int32),
        ("microbial_ptr", c_int32_p),
        ("microbial_specid", c_int32_p),
        ("microbial_stoich", c_double_p),
        ("microbial_rate_constant", c_double_p),
        ("microbial_activation_energy", c_double_p),
        ("microbial_monod_ptr", c_int32_p),
        ("microbial_monod_specid", c_int32_p),
        ("microbial_monod_K", c_double_p),
        ("microbial_monod_Cth", c_double_p),
        ("microbial_inhibition_ptr", c_int32_p),
        ("microbial_inhibition_specid", c_int32_p),
        ("microbial_inhibition_type", c_int32_p),
        ("microbial_inhibition_C", c_double_p),
        ("microbial_inhibition_C2", c_double_p),
        ("microbial_biomassid", c_int32_p),
        ("microbial_biomass_yield", c_double_p),
        ("nactive_gas", C.c_int32),
        ("pad_gas_", C.c_int32),
        ("acteq_ptr", c_int32_p),
        ("acteq_specid", c_int32_p),
        ("acteq_stoich", c_double_p),
        ("acteq_h2ostoich", c_double_p),
        ("acteq_logK", c_double_p),
        ("acteq_logK_coef", c_double_p),
        ("elm_flow_coupled", C.c_int32),
        ("pad_elm_", C.c_int32),
    ]


SANDBOX_CLM_CN, SANDBOX_SOMDEC, SANDBOX_NITRIF, SANDBOX_DENITR, SANDBOX_PLANTN, SANDBOX_LANGMUIR = 1, 2, 3, 4, 5, 6
SANDBOX_CNDEGAS = 7
SANDBOX_CALCITE = 8
SANDBOX_RADON = 9
SPEC_AQUEOUS, SPEC_IMMOBILE = 0, 2


class PfrxSomdec(C.Structure):
    _fields_ = (
        [(f, C.c_int32) for f in (
            "nrxn", "co2_id", "co2_itype", "o2_id", "o2_itype", "nh4_id", "no3_id", "n2o_id", "proton_id",
            "hr_id", "nmin_id", "nimm_id", "nimp_id", "ngasmin_id")]
        + [(f, C.c_double) for f in ("x0eps", "n2o_frac_mineralization", "inhibition_nh4_no3")]
        + [(f, c_double_p) for f in ("rate_constant", "rate_decomposition", "rate_ad_factor")]
        + [(f, c_int32_p) for f in ("upstream_c_id", "upstream_n_id", "upstream_is_aqueous", "upstream_hr_id",
                                    "upstream_nmin_id", "upstream_nimp_id", "upstream_nimm_id")]
        + [(f, c_double_p) for f in ("upstream_nc", "mineral_c_stoich", "mineral_n_stoich")]
        + [(f, c_int32_p) for f in ("downstream_ptr", "downstream_c_id", "downstream_n_id",
                                    "downstream_is_aqueous")]
        + [(f, c_double_p) for f in ("downstream_stoich", "downstream_nc")]
        + [(f, c_int32_p) for f in ("temperature_response_function", "moisture_response_function",
                                    "ox_response_function")]
        + [(f, c_double_p) for f in ("q10", "ea", "ox_half_saturation", "decomp_depth_efolding")]
        + [(f, c_int32_p) for f in ("ox_specid", "ox_specitype")]
        + [(f, c_int32_p) for f in ("monod_ptr", "monod_specid", "monod_specitype", "monod_pool_normalized")]
        + [(f, c_double_p) for f in ("monod_half_saturation", "monod_threshold")]
        + [(f, c_int32_p) for f in ("inhib_ptr", "inhib_itype", "inhib_specid", "inhib_specitype")]
        + [(f, c_double_p) for f in ("inhib_constant", "inhib_constant2")]
    )


class PfrxNitrif(C.Structure):
    _fields_ = ([(f, C.c_int32) for f in ("proton_id", "nh4_id", "no3_id", "n2o_id", "ngasnit_id")]
                + [(f, C.c_double) for f in ("k_nitr_max", "k_nitr_n2o", "x0eps")])


class PfrxPlantn(C.Structure):
    _fields_ = ([(f, C.c_int32) for f in ("nh4_id", "no3_id", "plantn_id", "plantndemand_id", "plantnh4uptake_id",
                                          "plantno3uptake_id")]
                + [(f, C.c_double) for f in ("half_saturation_nh4", "half_saturation_no3", "inhibition_nh4_no3",
                                             "x0eps_nh4", "x0eps_no3")])


class PfrxLangmuir(C.Structure):
    _fields_ = ([(f, C.c_int32) for f in ("aq_id", "sorb_id")]
                + [(f, C.c_double) for f in ("k_kinetic", "k_equilibrium", "s_max")])


class PfrxCndegas(C.Structure):
    _fields_ = ([(f, C.c_int32) for f in ("co2a_id", "n2oa_id", "n2a_id", "co2g_id", "n2og_id", "n2g_id", "proton_id",
                                           "himm_id", "fixph_on", "initialize_with_molality", "cell_state_mode", "pad_")]
                + [(f, C.c_double) for f in ("k_kinetic_co2", "k_kinetic_n2o", "k_kinetic_n2", "k_kinetic_h", "fixph",
                                             "reference_temperature", "reference_pressure")])


# constraint types of pfrx_equilibrate_constraint (transport_constraint_rt.F90:22-34)
CONSTRAINT_NULL, CONSTRAINT_FREE, CONSTRAINT_TOTAL, CONSTRAINT_LOG, CONSTRAINT_PH = 0, 1, 2, 3, 4
CONSTRAINT_MINERAL, CONSTRAINT_GAS, CONSTRAINT_CHARGE_BAL = 7, 8, 9


class PfrxConstraint(C.Structure):
    _fields_ = [("naqcomp", C.c_int32), ("initialize_with_molality", C.c_int32), ("max_iterations", C.c_int32),
                ("reserved", C.c_int32),
                ("type", C.POINTER(C.c_int32)),
                ("eq_logK", C.POINTER(C.c_double)), ("eq_logK_coef", C.POINTER(C.c_double)),
                ("eq_h2o_stoich", C.POINTER(C.c_double)),
                ("eq_ptr", C.POINTER(C.c_int32)), ("eq_spec", C.POINTER(C.c_int32)),
                ("eq_stoich", C.POINTER(C.c_double))]


class Constraint:
    """pfrx_constraint with the numpy arrays its pointers refer to kept alive"""

    def __init__(self, naqcomp: int, types, eq_logK=None, eq_h2o_stoich=None, eq_ptr=None, eq_spec=None,
                 eq_stoich=None, eq_logK_coef=None, initialize_with_molality: bool = False, max_iterations: int = 0):
        self.c = PfrxConstraint()
        self.c.naqcomp = int(naqcomp)
        self.c.initialize_with_molality = int(bool(initialize_with_molality))
        self.c.max_iterations = int(max_iterations)
        self.a = {}

        def put(name, arr, dt, ct):
            if arr is None:
                return
            v = np.ascontiguousarray(arr, dtype=dt)
            self.a[name] = v
            setattr(self.c, name, v.ctypes.data_as(C.POINTER(ct)))

        put("type", types, np.int32, C.c_int32)
        put("eq_logK", eq_logK, np.float64, C.c_double)
        put("eq_logK_coef", eq_logK_coef, np.float64, C.c_double)
        put("eq_h2o_stoich", eq_h2o_stoich, np.float64, C.c_double)
        put("eq_ptr", eq_ptr, np.int32, C.c_int32)
        put("eq_spec", eq_spec if eq_spec is not None and len(eq_spec) else (None if eq_spec is None else [0]),
            np.int32, C.c_int32)
        put("eq_stoich", eq_stoich if eq_stoich is not None and len(eq_stoich) else (None if eq_stoich is None else [0.0]),
            np.float64, C.c_double)


class PfrxCalciteSandbox(C.Structure):
    _fields_ = ([(f, C.c_int32) for f in ("mineral_id", "h_ion_id", "calcium_id", "bicarbonate_id")]
                + [(f, C.c_double) for f in ("rate_constant1", "rate_constant2")])


class PfrxRadon(C.Structure):
    _fields_ = [("species_id", C.c_int32), ("mineral_id", C.c_int32), ("radon_generation_rate", C.c_double)]


class PfrxDenitr(C.Structure):
    _fields_ = ([(f, C.c_int32) for f in ("no3_id", "n2_id", "n2o_id", "ngasdeni_id")]
                + [(f, C.c_double) for f in ("half_saturation", "k_deni_max", "x0eps")])


STATE_DOUBLE_FIELDS = [
    "total", "pri_molal", "immobile", "pri_act_coef", "sec_act_coef", "sec_molal", "ln_act_h2o",
    "mnrl_volfrac", "mnrl_area", "mnrl_rate", "srfcplxrxn_free_site_conc", "eqsrfcplx_conc",
    "total_sorb_eq", "kinmr_total_sorb", "den_kg", "sat", "temp", "porosity", "volume",
    "soil_particle_density",
]
# ELM per-cell scalars (pfrx_state.elm_*): present when the configuration sets
# elm_pflotran, NULL otherwise
STATE_ELM_FIELDS = ["elm_w_scalar", "elm_o_scalar", "elm_t_scalar", "elm_zsoil", "elm_kscalar_decomp_c",
                    "elm_bulkdensity_dry", "elm_bsw", "elm_rate_plantndemand", "somdec_nc",
                    "eqionx_ref_cation_sorbed_conc", "eqionx_conc", "pres", "sandbox_aux",
                    "sat_gas", "total_gas", "gas_pp", "elm_sucsat", "elm_watfc", "elm_effporosity"]
STATE_INT_FIELDS = ["imat", "num_sub_steps", "num_iterations", "num_kinetic_state_updates", "ierror"]
# fields the step updates ("io" in pfrx.h) and per-cell results
STATE_IO_FIELDS = [
    "total", "pri_molal", "immobile", "pri_act_coef", "sec_act_coef", "sec_molal", "ln_act_h2o",
    "mnrl_volfrac", "mnrl_rate", "srfcplxrxn_free_site_conc", "eqsrfcplx_conc", "total_sorb_eq",
    "kinmr_total_sorb", "somdec_nc", "eqionx_ref_cation_sorbed_conc", "eqionx_conc", "sandbox_aux",
    "total_gas", "gas_pp",
]
STATE_RESULT_FIELDS = ["num_sub_steps", "num_iterations", "num_kinetic_state_updates", "ierror"]


class PfrxState(C.Structure):
    _fields_ = (
        [("ld", C.c_int64)]
        + [(f, c_double_p) for f in STATE_DOUBLE_FIELDS]
        + [("imat", c_int32_p)]
        + [(f, c_double_p) for f in STATE_ELM_FIELDS]
        + [(f, c_int32_p) for f in STATE_INT_FIELDS if f != "imat"]
    )


class PfrxStepResult(C.Structure):
    _fields_ = [
        ("ncell_active", C.c_int64),
        ("sum_newton_iterations", C.c_int64),
        ("max_newton_iterations", C.c_int32),
        ("max_num_kinetic_state_updates", C.c_int32),
        ("rstep_error", C.c_int32),
        ("max_sub_steps", C.c_int32),
        ("num_cut_cells", C.c_int64),
        ("first_failed_cell", C.c_int64),
    ]

    def as_dict(self) -> Dict[str, int]:
        return {f: int(getattr(self, f)) for f, _ in self._fields_}


def _dp(a: Optional[np.ndarray]):
    if a is None or a.size == 0:
        return C.cast(None, c_double_p)
    return a.ctypes.data_as(c_double_p)


def _ip(a: Optional[np.ndarray]):
    if a is None or a.size == 0:
        return C.cast(None, c_int32_p)
    return a.ctypes.data_as(c_int32_p)


def _f64(x) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(x, dtype=np.float64))


def _i32(x) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(x, dtype=np.int32))


class ReactionConfig:
    """Owns the numpy tables and the ``pfrx_config`` that points into them."""

    def __init__(self, net: _chem.ReactionNetwork):
        self.net = net
        self.arrays: Dict[str, np.ndarray] = {}
        self.c = PfrxConfig()
        self._build()

    def _keep(self, name: str, arr: np.ndarray) -> np.ndarray:
        self.arrays[name] = arr
        return arr

    def _build(self) -> None:
        net, ch, c = self.net, self.net.chem, self.c
        c.abi_version = PFRX_ABI_VERSION
        c.naqcomp = net.naqcomp
        c.nimcomp = net.nimcomp
        c.use_full_geochemistry = 1
        c.use_log_formulation = int(ch.use_log_formulation)
        c.use_total_as_guess = int(ch.use_total_as_guess)
        c.use_isothermal = int(net.use_isothermal)
        c.act_coef_update_frequency = ch.act_coef_update_frequency
        c.act_coef_update_algorithm = ch.act_coef_update_algorithm
        c.use_activity_h2o = int(ch.use_activity_h2o)
        c.h2o_aq_id = net.primary_names.index("H2O") if "H2O" in net.primary_names else -1
        c.maximum_reaction_iterations = ch.maximum_reaction_iterations
        c.maximum_reaction_cuts = ch.maximum_reaction_cuts
        c.max_dlnC_rreact = ch.max_dlnC_rreact
        c.max_relative_change_tolerance = ch.max_relative_change_tolerance
        c.max_residual_tolerance = ch.max_residual_tolerance
        c.max_rel_residual_tolerance = ch.max_rel_residual_tolerance
        c.rt_min_saturation = 1.0e-40
        c.debyeA, c.debyeB, c.debyeBdot = net.debyeA, net.debyeB, net.debyeBdot
        c.primary_spec_Z = _dp(self._keep("primary_spec_Z", _f64(net.primary_Z)))
        c.primary_spec_a0 = _dp(self._keep("primary_spec_a0", _f64(net.primary_a0)))

        # secondary complexes
        c.neqcplx = net.neqcplx
        ptr, ids, st = net.csr(net.sec_rxn)
        c.eqcplx_ptr = _ip(self._keep("eqcplx_ptr", ptr))
        c.eqcplx_specid = _ip(self._keep("eqcplx_specid", ids))
        c.eqcplx_stoich = _dp(self._keep("eqcplx_stoich", st))
        c.eqcplx_h2ostoich = _dp(self._keep("eqcplx_h2ostoich", _f64([r.h2o_stoich for r in net.sec_rxn])))
        c.eqcplx_logK = _dp(self._keep("eqcplx_logK", net.logKs(net.sec_rxn)))
        if not net.use_isothermal and net.neqcplx:
            c.eqcplx_logKcoef = _dp(self._keep("eqcplx_logKcoef", _f64(net.logK_coefs(net.sec_rxn))))
        c.eqcplx_Z = _dp(self._keep("eqcplx_Z", _f64(net.eqcplx_Z)))
        c.eqcplx_a0 = _dp(self._keep("eqcplx_a0", _f64(net.eqcplx_a0)))

        # kinetic minerals
        c.nkinmnrl = net.nkinmnrl
        if net.nkinmnrl:
            rx = [net.mnrl_rxn[n] for n in net.kinmnrl_names]
            ptr, ids, st = net.csr(rx)
            km = net.kinmnrl
            c.kinmnrl_ptr = _ip(self._keep("kinmnrl_ptr", ptr))
            c.kinmnrl_specid = _ip(self._keep("kinmnrl_specid", ids))
            c.kinmnrl_stoich = _dp(self._keep("kinmnrl_stoich", st))
            c.kinmnrl_h2ostoich = _dp(self._keep("kinmnrl_h2ostoich", _f64([r.h2o_stoich for r in rx])))
            c.kinmnrl_logK = _dp(self._keep("kinmnrl_logK", net.logKs(rx)))
            if not net.use_isothermal:
                c.kinmnrl_logKcoef = _dp(self._keep("kinmnrl_logKcoef", _f64(net.logK_coefs(rx))))
            c.kinmnrl_molar_vol = _dp(self._keep("kinmnrl_molar_vol",
                                                  _f64([net.mnrl_molar_vol[n] for n in net.kinmnrl_names])))
            c.kinmnrl_rate_constant = _dp(self._keep("kinmnrl_rate_constant", _f64([m.rate_constant for m in km])))
            c.kinmnrl_activation_energy = _dp(self._keep("kinmnrl_activation_energy",
                                                          _f64([m.activation_energy for m in km])))
            c.kinmnrl_affinity_threshold = _dp(self._keep("kinmnrl_affinity_threshold",
                                                           _f64([m.affinity_threshold for m in km])))
            c.kinmnrl_rate_limiter = _dp(self._keep("kinmnrl_rate_limiter", _f64([m.rate_limiter for m in km])))
            c.kinmnrl_irreversible = _ip(self._keep("kinmnrl_irreversible", _i32([m.irreversible for m in km])))
            # optional arrays exist iff any mineral sets them
            # (reaction_database.F90:2047-2092)
            if any(m.temkin is not None for m in km):
                c.kinmnrl_Temkin_const = _dp(self._keep(
                    "kinmnrl_Temkin_const", _f64([1.0 if m.temkin is None else m.temkin for m in km])))
            if any(m.min_scale_factor is not None for m in km):
                c.kinmnrl_min_scale_factor = _dp(self._keep(
                    "kinmnrl_min_scale_factor",
                    _f64([1.0 if m.min_scale_factor is None else m.min_scale_factor for m in km])))
            if any(m.affinity_power is not None for m in km):
                c.kinmnrl_affinity_power = _dp(self._keep(
                    "kinmnrl_affinity_power",
                    _f64([1.0 if m.affinity_power is None else m.affinity_power for m in km])))
            if any(m.prefactors for m in km):
                MP, MS = _chem.MAX_PREFACTORS, _chem.MAX_PREFACTOR_SPECIES
                n = net.nkinmnrl
                npref = np.zeros(n, dtype=np.int32)
                nspec = np.zeros(n * MP, dtype=np.int32)
                pid = np.zeros(n * MP * MS, dtype=np.int32)
                alpha = np.zeros(n * MP * MS)
                beta = np.zeros(n * MP * MS)
                atten = np.zeros(n * MP * MS)
                prate = np.zeros(n * MP)
                peact = np.zeros(n * MP)
                for im, m in enumerate(km):
                    npref[im] = len(m.prefactors)
                    for ip, pf in enumerate(m.prefactors):
                        nspec[im * MP + ip] = len(pf["species"])
                        prate[im * MP + ip] = pf["rate"]
                        peact[im * MP + ip] = pf["activation_energy"]
                        for isp, sp in enumerate(pf["species"]):
                            q = (im * MP + ip) * MS + isp
                            if sp["name"] in net.primary_names:
                                pid[q] = net.primary_names.index(sp["name"])
                            else:
                                pid[q] = -(net.secondary_names.index(sp["name"]) + 1)
                            alpha[q], beta[q], atten[q] = sp["alpha"], sp["beta"], sp["atten"]
                c.kinmnrl_num_prefactors = _ip(self._keep("kinmnrl_num_prefactors", npref))
                c.kinmnrl_pref_nspec = _ip(self._keep("kinmnrl_pref_nspec", nspec))
                c.kinmnrl_prefactor_id = _ip(self._keep("kinmnrl_prefactor_id", pid))
                c.kinmnrl_pref_alpha = _dp(self._keep("kinmnrl_pref_alpha", alpha))
                c.kinmnrl_pref_beta = _dp(self._keep("kinmnrl_pref_beta", beta))
                c.kinmnrl_pref_atten_coef = _dp(self._keep("kinmnrl_pref_atten_coef", atten))
                c.kinmnrl_pref_rate = _dp(self._keep("kinmnrl_pref_rate", prate))
                c.kinmnrl_pref_activation_energy = _dp(self._keep("kinmnrl_pref_activation_energy", peact))

        # surface complexation
        nrxn = len(net.srfcplxrxn)
        c.nsrfcplxrxn = nrxn
        c.nsrfcplx = len(net.srfcplx_names)
        if nrxn:
            rptr = np.zeros(nrxn + 1, dtype=np.int32)
            r2c: List[int] = []
            for i, rx in enumerate(net.srfcplxrxn):
                r2c += [net.srfcplx_names.index(n) for n in rx.complexes]
                rptr[i + 1] = len(r2c)
            c.srfcplxrxn_ptr = _ip(self._keep("srfcplxrxn_ptr", rptr))
            c.srfcplxrxn_to_complex = _ip(self._keep("srfcplxrxn_to_complex", _i32(r2c)))
            c.srfcplxrxn_surf_type = _ip(self._keep("srfcplxrxn_surf_type",
                                                     _i32([r.surface_type for r in net.srfcplxrxn])))
            to_surf = []
            for r in net.srfcplxrxn:
                if r.surface_type == _chem.MINERAL_SURFACE:
                    # srfcplxrxn_to_surf indexes the kinetic-mineral list
                    # (reaction_database.F90: mapping by kinmnrl_names)
                    to_surf.append(net.kinmnrl_names.index(r.surface_name))
                else:
                    to_surf.append(-1)
            c.srfcplxrxn_to_surf = _ip(self._keep("srfcplxrxn_to_surf", _i32(to_surf)))
            c.srfcplxrxn_site_density = _dp(self._keep("srfcplxrxn_site_density",
                                                        _f64([r.site_density for r in net.srfcplxrxn])))
            flags = []
            for r in net.srfcplxrxn:
                fs = [net.srfcplx_free_site_stoich[net.srfcplx_names.index(n)] for n in r.complexes]
                # reaction_database.F90: stoich_flag set when any free-site
                # stoichiometry differs from 1
                flags.append(int(any(abs(v - 1.0) > 1.0e-40 for v in fs)))
            c.srfcplxrxn_stoich_flag = _ip(self._keep("srfcplxrxn_stoich_flag", _i32(flags)))
            ptr, ids, st = net.csr(net.srfcplx_rxn)
            c.srfcplx_ptr = _ip(self._keep("srfcplx_ptr", ptr))
            c.srfcplx_specid = _ip(self._keep("srfcplx_specid", ids))
            c.srfcplx_stoich = _dp(self._keep("srfcplx_stoich", st))
            c.srfcplx_h2ostoich = _dp(self._keep("srfcplx_h2ostoich",
                                                  _f64([r.h2o_stoich for r in net.srfcplx_rxn])))
            c.srfcplx_free_site_stoich = _dp(self._keep("srfcplx_free_site_stoich",
                                                         _f64(net.srfcplx_free_site_stoich)))
            c.srfcplx_logK = _dp(self._keep("srfcplx_logK", net.logKs(net.srfcplx_rxn)))
            if not net.use_isothermal:
                c.srfcplx_logKcoef = _dp(self._keep("srfcplx_logKcoef", _f64(net.logK_coefs(net.srfcplx_rxn))))
            c.neqsrfcplxrxn = len(net.eq_rxn_ids)
            c.eqsrfcplxrxn_to_srfcplxrxn = _ip(self._keep("eqsrfcplxrxn_to_srfcplxrxn", _i32(net.eq_rxn_ids)))
            c.nkinmrsrfcplxrxn = len(net.mr_rxn_ids)
            if net.mr_rxn_ids:
                c.kinmrsrfcplxrxn_to_srfcplxrxn = _ip(self._keep("kinmrsrfcplxrxn_to_srfcplxrxn",
                                                                  _i32(net.mr_rxn_ids)))
                mptr = np.zeros(len(net.mr_rxn_ids) + 1, dtype=np.int32)
                rates: List[float] = []
                fracs: List[float] = []
                for k, i in enumerate(net.mr_rxn_ids):
                    rates += net.srfcplxrxn[i].rates
                    fracs += net.srfcplxrxn[i].site_fractions
                    mptr[k + 1] = len(rates)
                c.kinmr_rate_ptr = _ip(self._keep("kinmr_rate_ptr", mptr))
                c.kinmr_rate = _dp(self._keep("kinmr_rate", _f64(rates)))
                c.kinmr_frac = _dp(self._keep("kinmr_frac", _f64(fracs)))

        # ion exchange, KD isotherms, dynamic KD
        ix = getattr(net, "ionx", None)
        if ix:
            c.neqionxrxn = len(ix["CEC"])
            c.eqionx_ptr = _ip(self._keep("eqionx_ptr", _i32(ix["ptr"])))
            c.eqionx_cationid = _ip(self._keep("eqionx_cationid", _i32(ix["cationid"])))
            c.eqionx_k = _dp(self._keep("eqionx_k", _f64(ix["k"])))
            c.eqionx_CEC = _dp(self._keep("eqionx_CEC", _f64(ix["CEC"])))
            c.eqionx_to_surf = _ip(self._keep("eqionx_to_surf", _i32(ix["to_surf"])))
            c.eqionx_Z_flag = _ip(self._keep("eqionx_Z_flag", _i32(ix["Z_flag"])))
        kd = getattr(net, "kd", None)
        if kd:
            c.neqkdrxn = len(kd["specid"])
            c.ikd_units = kd["ikd_units"]
            for k in ("specid", "type", "mineral"):
                setattr(c, "eqkd_" + k, _ip(self._keep("eqkd_" + k, _i32(kd[k]))))
            for k in ("coeff", "langmuir_b", "freundlich_n"):
                setattr(c, "eqkd_" + k, _dp(self._keep("eqkd_" + k, _f64(kd[k]))))
        dk = getattr(net, "dynkd", None)
        if dk:
            c.neqdynamickdrxn = len(dk["specid"])
            for k in ("specid", "refspecid"):
                setattr(c, "eqdynamickd_" + k, _ip(self._keep("eqdynamickd_" + k, _i32(dk[k]))))
            for k in ("refspechigh", "low", "high", "power"):
                setattr(c, "eqdynamickd_" + k, _dp(self._keep("eqdynamickd_" + k, _f64(dk[k]))))

        # general / radioactive decay / immobile decay
        g = getattr(net, "general", None)
        if g:
            c.ngeneral_rxn = len(g["kf"])
            for k in ("ptr", "specid", "fwd_ptr", "fwd_specid", "bwd_ptr", "bwd_specid"):
                setattr(c, "general_" + k, _ip(self._keep("general_" + k, _i32(g[k]))))
            for k in ("stoich", "fwd_stoich", "bwd_stoich", "kf", "kr"):
                setattr(c, "general_" + k, _dp(self._keep("general_" + k, _f64(g[k]))))
        rd = getattr(net, "radiodecay", None)
        if rd:
            c.nradiodecay_rxn = len(rd["kf"])
            for k in ("ptr", "specid", "forward_specid"):
                setattr(c, "radiodecay_" + k, _ip(self._keep("radiodecay_" + k, _i32(rd[k]))))
            for k in ("stoich", "kf"):
                setattr(c, "radiodecay_" + k, _dp(self._keep("radiodecay_" + k, _f64(rd[k]))))
        idc = getattr(net, "immdecay", None)
        if idc:
            c.nimmobile_decay_rxn = len(idc["k"])
            c.immobile_decay_specid = _ip(self._keep("immobile_decay_specid", _i32(idc["specid"])))
            c.immobile_decay_constant = _dp(self._keep("immobile_decay_constant", _f64(idc["k"])))

        mb = getattr(net, "microbial", None)
        if mb:
            c.nmicrobial_rxn = len(mb["rate_constant"])
            c.microbial_concentration_units = mb["units"]
            for k in ("ptr", "specid", "monod_ptr", "monod_specid", "inhibition_ptr", "inhibition_specid",
                      "inhibition_type", "biomassid"):
                setattr(c, "microbial_" + k, _ip(self._keep("microbial_" + k, _i32(mb[k]))))
            for k in ("stoich", "rate_constant", "monod_K", "monod_Cth", "inhibition_C", "inhibition_C2",
                      "biomass_yield"):
                setattr(c, "microbial_" + k, _dp(self._keep("microbial_" + k, _f64(mb[k]))))
            if any(e > 0.0 for e in mb["activation_energy"]):
                c.microbial_activation_energy = _dp(self._keep("microbial_activation_energy",
                                                               _f64(mb["activation_energy"])))

        # active gas species (RTotalGas)
        ag = getattr(net, "active_gas", None)
        if ag:
            c.nactive_gas = len(ag["logK"])
            c.acteq_ptr = _ip(self._keep("acteq_ptr", _i32(ag["ptr"])))
            c.acteq_specid = _ip(self._keep("acteq_specid", _i32(ag["specid"])))
            c.acteq_stoich = _dp(self._keep("acteq_stoich", _f64(ag["stoich"])))
            c.acteq_h2ostoich = _dp(self._keep("acteq_h2ostoich", _f64(ag["h2ostoich"])))
            c.acteq_logK = _dp(self._keep("acteq_logK", _f64(ag["logK"])))
            if not net.use_isothermal:
                c.acteq_logK_coef = _dp(self._keep("acteq_logK_coef", _f64(ag["logK_coef"])))

        # CLM-CN
        cc = net.clmcn
        if cc is not None:
            c.clmcn_nrxn = cc["nrxn"]
            c.clmcn_npool = cc["npool"]
            c.clmcn_C_species_id = cc["C_id"]
            c.clmcn_N_species_id = cc["N_id"]
            c.clmcn_CN_ratio = _dp(self._keep("clmcn_CN_ratio", _f64(cc["CN_ratio"])))
            c.clmcn_pool_nspec = _ip(self._keep("clmcn_pool_nspec", _i32(cc["pool_nspec"])))
            c.clmcn_pool_C_id = _ip(self._keep("clmcn_pool_C_id", _i32(cc["pool_C_id"])))
            c.clmcn_pool_N_id = _ip(self._keep("clmcn_pool_N_id", _i32(cc["pool_N_id"])))
            c.clmcn_upstream_pool_id = _ip(self._keep("clmcn_upstream_pool_id", _i32(cc["up"])))
            c.clmcn_downstream_pool_id = _ip(self._keep("clmcn_downstream_pool_id", _i32(cc["down"])))
            c.clmcn_rate_constant = _dp(self._keep("clmcn_rate_constant", _f64(cc["rate_constant"])))
            c.clmcn_respiration_fraction = _dp(self._keep("clmcn_respiration_fraction", _f64(cc["resp"])))
            c.clmcn_inhibition_constant = _dp(self._keep("clmcn_inhibition_constant", _f64(cc["inhib"])))

        self._sandboxes()

    def _sandboxes(self) -> None:
        """SOMDECOMP / NITRIFICATION / DENITRIFICATION tables (pfrx_somdec etc.)"""
        net, c = self.net, self.c
        order: List[int] = []
        for kind in getattr(net, "sandbox_order", []):
            order.append({"CLM-CN": SANDBOX_CLM_CN, "SOMDECOMP": SANDBOX_SOMDEC, "NITRIFICATION": SANDBOX_NITRIF,
                          "DENITRIFICATION": SANDBOX_DENITR, "PLANTN": SANDBOX_PLANTN,
                          "LANGMUIR": SANDBOX_LANGMUIR, "CNDEGAS": SANDBOX_CNDEGAS, "CALCITE": SANDBOX_CALCITE,
                          "RADON": SANDBOX_RADON}[kind])
        if order:
            c.nsandbox = len(order)
            c.sandbox_list = _ip(self._keep("sandbox_list", _i32(order)))
        c.elm_pflotran = int(getattr(net, "elm_pflotran", False))
        c.elm_flow_coupled = int(getattr(net, "elm_flow_coupled", False))
        sd = getattr(net, "somdec", None)
        if sd is not None:
            o = PfrxSomdec()
            for k, v in sd["scalars"].items():
                setattr(o, k, v)
            for k, v in sd["int_arrays"].items():
                setattr(o, k, _ip(self._keep("somdec_" + k, _i32(v))))
            for k, v in sd["real_arrays"].items():
                setattr(o, k, _dp(self._keep("somdec_" + k, _f64(v))))
            self.somdec = o
            c.somdec = C.cast(C.pointer(o), C.c_void_p)
        nt = getattr(net, "nitrif", None)
        if nt is not None:
            o = PfrxNitrif()
            for k, v in nt.items():
                setattr(o, k, v)
            self.nitrif = o
            c.nitrif = C.cast(C.pointer(o), C.c_void_p)
        dn = getattr(net, "denitr", None)
        if dn is not None:
            o = PfrxDenitr()
            for k, v in dn.items():
                setattr(o, k, v)
            self.denitr = o
            c.denitr = C.cast(C.pointer(o), C.c_void_p)
        pn = getattr(net, "plantn", None)
        if pn is not None:
            o = PfrxPlantn()
            for k, v in pn.items():
                setattr(o, k, v)
            self.plantn = o
            c.plantn = C.cast(C.pointer(o), C.c_void_p)
        lg = getattr(net, "langmuir", None)
        if lg is not None:
            o = PfrxLangmuir()
            for k, v in lg.items():
                setattr(o, k, v)
            self.langmuir = o
            c.langmuir = C.cast(C.pointer(o), C.c_void_p)
        cd = getattr(net, "cndegas", None)
        if cd is not None:
            o = PfrxCndegas()
            for k, v in cd.items():
                setattr(o, k, v)
            self.cndegas = o
            c.cndegas = C.cast(C.pointer(o), C.c_void_p)
        cs = getattr(net, "calcite_sandbox", None)
        if cs is not None:
            o = PfrxCalciteSandbox()
            for k, v in cs.items():
                setattr(o, k, v)
            self.calcite = o
            c.calcite = C.cast(C.pointer(o), C.c_void_p)
        rn = getattr(net, "radon", None)
        if rn is not None:
            o = PfrxRadon()
            for k, v in rn.items():
                setattr(o, k, v)
            self.radon = o
            c.radon = C.cast(C.pointer(o), C.c_void_p)

    # ------------------------------------------------------------------ #
    @classmethod
    def from_dump(cls, path: str) -> "ReactionConfig":
        """a configuration written by ``pfrx_config_dump`` / ``pfrx_config_write`` (the C side of the
        boundary): the scalars of the structs and every table the code generator reads, with exact bit
        patterns.  What set-up needs beyond the generator (names, the reaction network object) is not
        in the file: the result serves :mod:`.specialize`, not the deck-level helpers."""
        self = cls.__new__(cls)
        self.net = None
        self.arrays = {}
        structs = {"config": PfrxConfig, "somdec": PfrxSomdec, "nitrif": PfrxNitrif, "denitr": PfrxDenitr,
                   "plantn": PfrxPlantn, "langmuir": PfrxLangmuir, "cndegas": PfrxCndegas,
                   "calcite": PfrxCalciteSandbox, "radon": PfrxRadon}
        prefix = {"c": ("", None), "sd": ("somdec_", "somdec"), "nt": ("nitrif_", "nitrif"), "dn": ("denitr_", "denitr"),
                  "pn": ("plantn_", "plantn"), "lg": ("langmuir_", "langmuir")}
        objs = {}
        sig = None
        with open(path) as f:
            head = f.readline().split()
            if head[:2] != ["pfrx_config_dump", "1"] or int(head[3]) != PFRX_ABI_VERSION:
                raise ValueError(f"{path}: not a pfrx_config_dump of ABI {PFRX_ABI_VERSION}")
            tables = []
            for ln in f:
                w = ln.split()
                if not w:
                    continue
                if w[0] == "S":
                    ty = structs[w[1]]
                    raw = bytes.fromhex(w[3])
                    if len(raw) != int(w[2]) or len(raw) != C.sizeof(ty):
                        raise ValueError(f"{path}: struct {w[1]} has {len(raw)} bytes, this build expects {C.sizeof(ty)}")
                    o = ty.from_buffer_copy(raw)
                    for name, ct in ty._fields_:          # pointers of the writer's address space mean nothing here
                        if ct in (c_double_p, c_int32_p, C.c_void_p):
                            setattr(o, name, None)
                    objs[w[1]] = o
                elif w[0] == "T":
                    tables.append((w[1], int(w[2]), int(w[3]), w[4]))
                elif w[0] == "signature":
                    sig = int(w[1], 16)
        self.c = objs["config"]
        for k in ("somdec", "nitrif", "denitr", "plantn", "langmuir", "cndegas", "calcite", "radon"):
            if k in objs:
                setattr(self, k, objs[k])
                setattr(self.c, k, C.cast(C.pointer(objs[k]), C.c_void_p))
        for name, elem, count, hx in tables:
            owner, field = name.split("->")
            pre, attr = prefix[owner]
            obj = self.c if attr is None else getattr(self, attr)
            ct = dict(type(obj)._fields_)[field]
            dt = np.float64 if ct is c_double_p else np.int32
            arr = np.frombuffer(bytes.fromhex(hx), dtype=dt).copy()
            if arr.size != count or arr.itemsize != elem:
                raise ValueError(f"{path}: table {name} is malformed")
            self._keep(pre + field, arr)
            setattr(obj, field, _dp(arr) if ct is c_double_p else _ip(arr))
        if "somdec" in objs:   # empty tables (e.g. no inhibition terms) are not written: the generator expects the keys
            for field, ct in PfrxSomdec._fields_:
                if ct in (c_double_p, c_int32_p) and "somdec_" + field not in self.arrays:
                    self.arrays["somdec_" + field] = np.zeros(0, dtype=np.float64 if ct is c_double_p else np.int32)
        self.dump_signature = sig
        return self

    @property
    def ncomp(self) -> int:
        return self.c.naqcomp + self.c.nimcomp

    def field_rows(self) -> Dict[str, int]:
        """number of components (rows) of every per-cell field"""
        c = self.c
        nmr = c.nkinmrsrfcplxrxn
        mr_rows = 0
        if nmr:
            mr_rows = c.naqcomp * (int(self.arrays["kinmr_rate_ptr"][nmr]) + nmr)
        return {
            "total": c.naqcomp, "pri_molal": c.naqcomp, "immobile": c.nimcomp,
            "pri_act_coef": c.naqcomp, "sec_act_coef": c.neqcplx, "sec_molal": c.neqcplx,
            "ln_act_h2o": 1, "mnrl_volfrac": c.nkinmnrl, "mnrl_area": c.nkinmnrl, "mnrl_rate": c.nkinmnrl,
            "srfcplxrxn_free_site_conc": c.nsrfcplxrxn, "eqsrfcplx_conc": c.nsrfcplx,
            "total_sorb_eq": (c.naqcomp if (c.neqsrfcplxrxn + c.neqionxrxn + c.neqkdrxn + c.neqdynamickdrxn) > 0
                              else 0),
            "kinmr_total_sorb": mr_rows,
            "den_kg": 1, "sat": 1, "temp": 1, "porosity": 1, "volume": 1, "soil_particle_density": 1,
            **{f: (1 if c.elm_pflotran else 0) for f in STATE_ELM_FIELDS},
            "eqionx_ref_cation_sorbed_conc": c.neqionxrxn,
            "eqionx_conc": int(self.arrays["eqionx_ptr"][c.neqionxrxn]) if c.neqionxrxn else 0,
            "pres": 1 if (getattr(self, "cndegas", None) is not None and self.cndegas.cell_state_mode >= 1) else 0,
            "sandbox_aux": 1 if getattr(self, "calcite", None) is not None else 0,
            "sat_gas": 1 if c.nactive_gas > 0 else 0,
            "total_gas": c.naqcomp if c.nactive_gas > 0 else 0,
            "gas_pp": max(c.nactive_gas, 0),
            "elm_sucsat": 1 if c.elm_flow_coupled else 0, "elm_watfc": 1 if c.elm_flow_coupled else 0,
            "elm_effporosity": 1 if c.elm_flow_coupled else 0,
            "somdec_nc": (len(self.arrays["somdec_upstream_nc"]) + len(self.arrays.get("somdec_downstream_nc", []))
                          if c.somdec else 0),
            "imat": 1, "num_sub_steps": 1, "num_iterations": 1, "num_kinetic_state_updates": 1, "ierror": 1,
        }


class HostState:
    """Cell-major SoA state in host (numpy) memory: ``field[k, cell]``."""

    def __init__(self, cfg: ReactionConfig, ncell: int):
        self.cfg = cfg
        self.ncell = int(ncell)
        self.a: Dict[str, np.ndarray] = {}
        rows = cfg.field_rows()
        for f in STATE_DOUBLE_FIELDS + STATE_ELM_FIELDS:
            self.a[f] = np.zeros((rows[f], self.ncell), dtype=np.float64)
        for f in STATE_INT_FIELDS:
            self.a[f] = np.zeros((rows[f], self.ncell), dtype=np.int32)
        # neutral defaults
        self.a["pri_act_coef"][:] = 1.0
        self.a["sec_act_coef"][:] = 1.0
        self.a["den_kg"][:] = 1000.0
        self.a["sat"][:] = 1.0
        self.a["temp"][:] = 25.0
        self.a["porosity"][:] = 0.25
        self.a["volume"][:] = 1.0
        self.a["soil_particle_density"][:] = 2650.0
        self.a["imat"][:] = 1
        self.a["srfcplxrxn_free_site_conc"][:] = 1.0e-9
        self.a["eqionx_ref_cation_sorbed_conc"][:] = 1.0e-9   # reactive_transport_aux.F90:282
        self.a["eqionx_conc"][:] = 1.0e-9
        self.a["elm_rate_plantndemand"][:] = 1.0e-2
        for f in ("elm_w_scalar", "elm_o_scalar", "elm_t_scalar", "elm_kscalar_decomp_c", "elm_bsw"):
            self.a[f][:] = 1.0
        self.a["elm_bulkdensity_dry"][:] = 1.25e3
        self.a["pres"][:] = 101325.0
        self.a["elm_sucsat"][:] = 200.0
        self.a["elm_watfc"][:] = 0.1
        self.a["elm_effporosity"][:] = 0.4
        if rows["somdec_nc"]:
            nc0 = np.concatenate([cfg.arrays["somdec_upstream_nc"],
                                  cfg.arrays.get("somdec_downstream_nc", np.zeros(0))])
            self.a["somdec_nc"][:] = nc0[:, None]

    def __getitem__(self, k: str) -> np.ndarray:
        return self.a[k]

    def copy(self) -> "HostState":
        o = HostState.__new__(HostState)
        o.cfg, o.ncell = self.cfg, self.ncell
        o.a = {k: v.copy() for k, v in self.a.items()}
        return o

    def struct(self) -> PfrxState:
        s = PfrxState()
        s.ld = self.ncell
        for f in STATE_DOUBLE_FIELDS + STATE_ELM_FIELDS:
            setattr(s, f, _dp(self.a[f]))
        for f in STATE_INT_FIELDS:
            setattr(s, f, _ip(self.a[f]))
        return s

    def tile(self, reps: int) -> "HostState":
        o = HostState.__new__(HostState)
        o.cfg, o.ncell = self.cfg, self.ncell * reps
        o.a = {k: np.ascontiguousarray(np.tile(v, (1, reps))) for k, v in self.a.items()}
        return o
