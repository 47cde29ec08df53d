"""Synthetic per-cell states of the BASELINE.json configurations (SURVEY.md
section 8(d)), seeded with ``numpy.random.default_rng(20261017)``.

Every workload starts from waters speciated by
:func:`~.constraint.equilibrate_constraint` (the constraints of the reference's
own decks) and mixes them per cell the way a transport step would: totals mix
linearly (transport is linear in totals), the free-ion guess is what the cell
held before the step.  The decks/databases read here are the committed
fixtures under ``tests/golden`` (verbatim reference test data) plus two small
decks written for C2/C5 in the reference's input format.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import numpy as np

from . import abi, chem, constraint, eos

SEED = 20261017
DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def _read(name: str) -> str:
    with open(os.path.join(DATA, name)) as f:
        return f.read()


@dataclass
class Workload:
    name: str
    cfg: abi.ReactionConfig
    state: abi.HostState
    tran_dt: float
    net: chem.ReactionNetwork
    note: str = ""


# --------------------------------------------------------------------------- #
C2_DECK = """
# C2: 1D calcite column chemistry (H+/HCO3-/Ca++, 6 complexes, kinetic Calcite)
CHEMISTRY
  PRIMARY_SPECIES
    H+
    HCO3-
    Ca++
  /
  SECONDARY_SPECIES
    OH-
    CO3--
    CO2(aq)
    CaOH+
    CaHCO3+
    CaCO3(aq)
  /
  MINERALS
    Calcite
  /
  MINERAL_KINETICS
    Calcite
      RATE_CONSTANT 1.d-6 mol/m^2-sec
    /
  /
  DATABASE ./calcite.dat
  LOG_FORMULATION
  ACTIVITY_COEFFICIENTS TIMESTEP
END
CONSTRAINT background
  CONCENTRATIONS
    H+     8.0     P
    HCO3-  1.d-3   T
    Ca++   5.d-4   M Calcite
  /
  MINERALS
    Calcite 1.d-5 1.d0 m^2/m^3
  /
END
CONSTRAINT inlet
  CONCENTRATIONS
    H+     5.0     P
    HCO3-  1.d-3   T
    Ca++   1.d-6   Z
  /
  MINERALS
    Calcite 1.d-5 1.d0 m^2/m^3
  /
END
"""


def _mix_fill(state: abi.HostState, waters, weights: np.ndarray, rng, jitter: float = 0.01) -> None:
    """cells = convex mixes of speciated waters.  weights [nwater, ncell]."""
    a = state.a
    ncell = state.ncell

    def lin(field):
        vals = [np.asarray(getattr(w, field), dtype=np.float64) for w in waters]
        return sum(v[:, None] * weights[i][None, :] for i, v in enumerate(vals))

    def geo(field):
        vals = [np.log(np.maximum(np.asarray(getattr(w, field), dtype=np.float64), 1.0e-300)) for w in waters]
        return np.exp(sum(v[:, None] * weights[i][None, :] for i, v in enumerate(vals)))

    tot = lin("total")
    if jitter > 0.0:
        tot = tot * np.exp(jitter * rng.standard_normal(tot.shape))
    a["total"][...] = tot
    a["pri_molal"][...] = geo("pri_molal")
    a["pri_act_coef"][...] = lin("pri_act_coef")
    if a["sec_molal"].shape[0]:
        a["sec_molal"][...] = geo("sec_molal")
        a["sec_act_coef"][...] = lin("sec_act_coef")
    assert ncell == tot.shape[1]


def calcite_column(ncell: int = 10000, tran_dt: float = 3600.0, seed: int = SEED, prefactors: bool = False,
                   anisothermal: bool = False, no_geochemistry: bool = False, sandbox: bool = False) -> Workload:
    """C2: 10k-cell calcite column, post-transport totals on a logistic front.

    ``prefactors``: the calcite rate as the sum of three parallel mechanisms in the
    PREFACTOR form of reaction_mineral.F90:838-890 (acid, carbonic-acid and neutral, the
    shape of the Plummer/Chou rate laws): an H+ term, a term in the SECONDARY species
    CO2(aq) with an attenuation denominator, and a species-free term with an activation
    energy."""
    rng = np.random.default_rng(seed)
    deck = C2_DECK
    if sandbox:
        # the chemistry of regression_tests/default/reaction_sandbox/reaction_sandbox_calcite.in: the mineral's
        # own rate constant zero, the CALCITE sandbox's two parallel pathways in its place
        deck = deck.replace("      RATE_CONSTANT 1.d-6 mol/m^2-sec\n", "      RATE_CONSTANT 0.d0\n")
        deck = deck.replace("  DATABASE ./calcite.dat\n", "  REACTION_SANDBOX\n    CALCITE\n      RATE_CONSTANT1 5.d-7\n"
                            "      RATE_CONSTANT2 5.d-7\n    /\n  /\n  DATABASE ./calcite.dat\n")
        assert "CALCITE" in deck and "RATE_CONSTANT 0.d0" in deck
    dk = chem.read_deck(deck)
    if prefactors:
        # prefactors == "primary": the second mechanism in primary species only (its analytic
        # Jacobian is then a true derivative and can be checked by finite differences)
        second = "HCO3-" if prefactors == "primary" else "CO2(aq)"
        dk.chemistry.mineral_kinetics[0].prefactors = [
            {"rate": 8.9e-1 * 1.0e-4, "activation_energy": 0.0,
             "species": [{"name": "H+", "alpha": 1.0, "beta": 0.0, "atten": 0.0}]},
            {"rate": 5.0e-4 * 1.0e-4, "activation_energy": 0.0,
             "species": [{"name": second, "alpha": 1.0, "beta": 0.5, "atten": 40.0},
                         {"name": "Ca++", "alpha": 0.25, "beta": 1.0, "atten": 3.0}]},
            {"rate": 6.5e-7 * 1.0e-4, "activation_energy": 23500.0, "species": []},
        ]
    net = chem.ReactionNetwork(dk.chemistry, chem.Database(_read("calcite.dat")), use_isothermal=not anisothermal)
    cfg = abi.ReactionConfig(net)
    if no_geochemistry:
        cfg.c.use_full_geochemistry = 0   # RStep's tracer short cut: pri_molal = total / den * 1000 (reaction.F90:3600)
    den = eos.water_density_ifc67()
    bg = constraint.equilibrate_constraint(net, dk.constraints["background"], den_kg=den)
    inl = constraint.equilibrate_constraint(net, dk.constraints["inlet"], den_kg=den)
    st = abi.HostState(cfg, ncell)
    x = (np.arange(ncell) + 0.5) / ncell
    w = 1.0 / (1.0 + np.exp((x - 0.35) / 0.05))
    _mix_fill(st, [bg, inl], np.stack([1.0 - w, w]), rng)
    st["mnrl_volfrac"][...] = 1.0e-5
    st["mnrl_area"][...] = 1.0
    st["den_kg"][...] = den
    st["porosity"][...] = 0.25
    if prefactors:
        st["temp"][...] = rng.uniform(10.0, 40.0, ncell)
        return Workload("c2_calcite_prefactors", cfg, st, tran_dt, net,
                        "C2 with a three-mechanism PREFACTOR rate law for Calcite")
    if anisothermal:
        st["temp"][...] = rng.uniform(5.0, 60.0, ncell)
        return Workload("c2_calcite_anisothermal", cfg, st, tran_dt, net, "C2 with logK(T), cells at 5-60 C")
    if no_geochemistry:
        return Workload("c2_calcite_no_geochemistry", cfg, st, tran_dt, net, "C2 with use_full_geochemistry = 0")
    if sandbox:
        st["mnrl_volfrac"][0, ::5] = 0.0      # cells without the mineral: dissolution switched off, precipitation not
        return Workload("c2_calcite_sandbox", cfg, st, tran_dt, net, "C2 with the CALCITE reaction sandbox")
    return Workload("c2_calcite_column", cfg, st, tran_dt, net,
                    "H+/HCO3-/Ca++ + 6 complexes + kinetic Calcite, logistic inlet/background front")


def calcite_batch() -> Workload:
    """C1: the single cell of regression_tests/ascem/batch/calcite-kinetics."""
    dk, net = chem.load_network(_read("calcite-kinetics.in"), _read("calcite.dat"))
    cfg = abi.ReactionConfig(net)
    den = eos.water_density_ifc67()
    sp = constraint.equilibrate_constraint(net, dk.constraints["initial"], den_kg=den, porosity=0.5)
    st = abi.HostState(cfg, 1)
    constraint.fill_cells(st, sp)
    st["den_kg"][...] = den
    st["porosity"][...] = 0.5
    return Workload("c1_calcite_batch", cfg, st, 0.5, net, "ascem/batch/calcite-kinetics, one cell, dt = 0.5 s")


# --------------------------------------------------------------------------- #
def _hanford_network(variant: str = "base", activity_newton: bool = False, anisothermal: bool = False,
                     activity_h2o: bool = False, total_as_guess: bool = False):
    deck = _read("543_hanford_srfcplx_base.in")
    dk = chem.read_deck(deck)
    ch = dk.chemistry
    ch.use_activity_h2o = ch.use_activity_h2o or activity_h2o          # ACTIVITY_WATER
    ch.use_total_as_guess = ch.use_total_as_guess or total_as_guess    # reaction.F90:3640 (guess = total)
    if activity_newton:
        # ACTIVITY_COEFFICIENTS NEWTON NEWTON_ITERATION: ionic strength iterated to 1e-6
        ch.act_coef_update_algorithm = chem.ACT_COEF_ALGORITHM_NEWTON
    if variant == "mr":
        # 50-rate multirate on rock density (543_hanford_srfcplx_mr.in /
        # column/surface_complexation_mr_os.in)
        dk_mr = chem.read_deck(_read("surface_complexation_mr_os.in"))
        ch.srfcplx_rxns = dk_mr.chemistry.srfcplx_rxns
    elif variant == "minerals":
        # C5: Hanford basis, six kinetic minerals, no sorption
        ch.srfcplx_rxns = []
        rates = {"Calcite": 1.0e-8, "Metatorbernite": 2.0e-13, "Dolomite": 1.0e-9, "Gypsum": 1.0e-7,
                 "Fluorite": 1.0e-9, "Schoepite": 1.0e-10}
        ch.mineral_kinetics = [chem.MineralKinetics(n, rate_constant=r) for n, r in rates.items()]
    net = chem.ReactionNetwork(ch, chem.Database(_read("hanford_subset.dat")), use_isothermal=not anisothermal)
    return dk, net


def hanford(ncell: int = 256 * 256 * 64, tran_dt: float = 3600.0, variant: str = "base",
            seed: int = SEED, activity_newton: bool = False, anisothermal: bool = False, activity_h2o: bool = False,
            total_as_guess: bool = False, inner_newton_sites: bool = False) -> Workload:
    """C3 (variant base|mr) and C5 (variant minerals): Hanford 15 primary / 88
    secondary; cells are Dirichlet(1) mixes of the three deck waters.

    Options for the parity tests of the less-travelled branches: ``anisothermal`` (logK(T) from the
    database's temperature fit, cells at 5-60 C: RUpdateTempDependentCoefs, reaction.F90:5976-6067),
    ``activity_h2o`` (ACTIVITY_WATER, reaction.F90:4580-4612), ``total_as_guess``
    (reaction.F90:3640) and ``inner_newton_sites`` (the free-site Newton iteration of
    RTotalSorbEqSurfCplx1, reaction_surf_complex.F90:700-800, forced through srfcplxrxn_stoich_flag)."""
    rng = np.random.default_rng(seed)
    dk, net = _hanford_network(variant, activity_newton, anisothermal, activity_h2o, total_as_guess)
    cfg = abi.ReactionConfig(net)
    if inner_newton_sites and cfg.c.nsrfcplxrxn > 0:
        cfg.arrays["srfcplxrxn_stoich_flag"][...] = 1   # in place: the ctypes struct points at this buffer
    den = eos.water_density_ifc67()
    waters = []
    for nm in ("groundwater", "U_source", "river_water"):
        cons = dk.constraints[nm]
        if variant == "minerals":
            cons.minerals = {k: (0.05, 100.0) for k in net.kinmnrl_names}
        waters.append(constraint.equilibrate_constraint(net, cons, den_kg=den, porosity=0.25,
                                                        soil_particle_density=2500.0))
    st = abi.HostState(cfg, ncell)
    wts = rng.dirichlet(np.ones(3), size=ncell).T
    _mix_fill(st, waters, wts, rng, jitter=0.0)
    st["den_kg"][...] = den
    st["porosity"][...] = rng.uniform(0.2, 0.3, ncell)
    st["soil_particle_density"][...] = rng.choice([2000.0, 2500.0], ncell)
    nk = net.nkinmnrl
    if variant == "minerals":
        st["mnrl_volfrac"][...] = rng.uniform(0.0, 0.1, (nk, ncell))
        st["mnrl_area"][...] = 100.0
    else:
        st["mnrl_volfrac"][0, :] = rng.uniform(0.05, 0.15, ncell)          # Calcite
        st["mnrl_volfrac"][1, :] = rng.choice([0.0, 1.0e-4], ncell)        # Metatorbernite
        st["mnrl_area"][...] = 100.0                                       # 1 cm^2/cm^3
    nr = len(net.srfcplxrxn)
    if nr:
        st["srfcplxrxn_free_site_conc"][...] = 1.0e-9
        # sorbed state consistent with the mixed water: one RTotalSorb pass on
        # the host would cost minutes at 4M cells; instead start the multirate
        # sites from the mixed equilibrium sorbed totals of the three waters
        if net.eq_rxn_ids:
            vals = [w.total_sorb_eq for w in waters]
            st["total_sorb_eq"][...] = sum(v[:, None] * wts[i][None, :] for i, v in enumerate(vals))
        if net.mr_rxn_ids:
            vals = [w.kinmr_total_sorb for w in waters]
            st["kinmr_total_sorb"][...] = sum(v[:, None] * wts[i][None, :] for i, v in enumerate(vals))
    if anisothermal:
        st["temp"][...] = rng.uniform(5.0, 60.0, ncell)
    name = {"base": "c3_hanford_srfcplx", "mr": "c3_hanford_multirate", "minerals": "c5_hanford_minerals"}[variant]
    for flag, tag in ((anisothermal, "_anisothermal"), (activity_h2o, "_activity_h2o"), (total_as_guess, "_total_guess"),
                      (inner_newton_sites, "_inner_newton")):
        if flag:
            name += tag
    return Workload(name, cfg, st, tran_dt, net, f"Hanford 15/88, variant {variant}, Dirichlet mix of 3 deck waters")


# --------------------------------------------------------------------------- #
def clm_cn(ncell: int = 2 * 1024 * 1024, tran_dt: float = 1800.0, seed: int = SEED) -> Workload:
    """C4(a): regression_tests/ngee/CLM-CN network (1 aq + 12 immobile, 7 rxns)."""
    rng = np.random.default_rng(seed)
    dk, net = chem.load_network(_read("CLM-CN.in"), _read("CLM-CN_database.dat"))
    cfg = abi.ReactionConfig(net)
    sp = constraint.equilibrate_constraint(net, dk.constraints["initial"], den_kg=1000.0)
    st = abi.HostState(cfg, ncell)
    constraint.fill_cells(st, sp)
    imm0 = sp.immobile
    pools = imm0[:, None] * np.exp(rng.standard_normal((net.nimcomp, ncell)))
    iN = net.immobile_names.index("N")
    pools[iN, :] = 10.0 ** rng.uniform(-8.0, -4.0, ncell)
    st["immobile"][...] = pools
    st["temp"][...] = rng.uniform(-5.0, 30.0, ncell)
    st["sat"][...] = rng.uniform(0.05, 1.0, ncell)
    st["den_kg"][...] = 1000.0
    st["porosity"][...] = 0.25
    st["volume"][...] = 1.0
    return Workload("c4_clm_cn", cfg, st, tran_dt, net, "ngee/CLM-CN 13-dof sandbox, lognormal pools")


# --------------------------------------------------------------------------- #
C4S_DECK = """
# C4(b): ngee-style ELM-CN network -- SOM decomposition (the CLM-CN cascade of
# ngee/CLMCNplus/clm_cn1.in with N immobilisation from NH4+/NO3- as in clm_nimm4.in
# and the N tracking species of clm_nmit.in), nitrification and denitrification
CHEMISTRY
  PRIMARY_SPECIES
    CO2(aq)
    N2O(aq)
    NH4+
    NO3-
    N2(aq)
    H+
  /
  IMMOBILE_SPECIES
    SOM1
    SOM2
    SOM3
    SOM4
    Lit1C
    Lit2C
    Lit3C
    Lit1N
    Lit2N
    Lit3N
  /
  REACTION_SANDBOX
    SOMDECOMP
      __ABIOTIC__
      POOLS
        SOM1 12.d0
        SOM2 12.d0
        SOM3 10.d0
        SOM4 10.d0
        Lit1
        Lit2
        Lit3
      /
      REACTION
        UPSTREAM_POOL Lit1
        DOWNSTREAM_POOL SOM1 0.61d0
        TURNOVER_TIME 20. h
        MONOD
          SPECIES_NAME NH4+
          HALF_SATURATION_CONSTANT 1.d-5
        /
        MONOD
          SPECIES_NAME NO3-
          HALF_SATURATION_CONSTANT 1.d-5
        /
      /
      REACTION
        UPSTREAM_POOL Lit2
        DOWNSTREAM_POOL SOM2 0.45
        TURNOVER_TIME 14. d
        MONOD
          SPECIES_NAME NH4+
          HALF_SATURATION_CONSTANT 1.d-5
        /
        MONOD
          SPECIES_NAME NO3-
          HALF_SATURATION_CONSTANT 1.d-5
        /
      /
      REACTION
        UPSTREAM_POOL Lit3
        DOWNSTREAM_POOL SOM3 0.71d0
        TURNOVER_TIME 71. d
        MONOD
          SPECIES_NAME NH4+
          HALF_SATURATION_CONSTANT 1.d-5
        /
        MONOD
          SPECIES_NAME NO3-
          HALF_SATURATION_CONSTANT 1.d-5
        /
      /
      REACTION
        UPSTREAM_POOL SOM1
        DOWNSTREAM_POOL SOM2 0.72d0
        TURNOVER_TIME 14. d
      /
      REACTION
        UPSTREAM_POOL SOM2
        DOWNSTREAM_POOL SOM3 0.54d0
        TURNOVER_TIME 71. d
      /
      REACTION
        UPSTREAM_POOL SOM3
        DOWNSTREAM_POOL SOM4 0.45d0
        TURNOVER_TIME 2. y
      /
      REACTION
        UPSTREAM_POOL SOM4
        TURNOVER_TIME 27.4 y
      /
    /
    NITRIFICATION
      NITRIFICATION_RATE_COEF 1.d-6
      N2O_RATE_COEF_NITRIFICATION 3.5d-8
    /
    DENITRIFICATION
      DENITRIFICATION_RATE_COEF 2.5d-6
      NITRATE_HALF_SATURATION 1.d-9
    /
  /
  DATABASE ./CLM-CN_database.dat
END
CONSTRAINT initial
  CONCENTRATIONS
    NH4+    4.d-5      T
    NO3-    2.d-5      T
    CO2(aq) 1.d-10     T
    N2O(aq) 1.d-12     T
    N2(aq)  1.d-10     T
    H+      6.0d0      pH
  /
  IMMOBILE
    SOM1  1.d-2
    SOM2  1.d-2
    SOM3  1.d-1
    SOM4  1.d-1
    Lit1C 0.1852d-0
    Lit2C 0.4578d-0
    Lit3C 0.2662d-0
    Lit1N 0.00508954d-0
    Lit2N 0.01258096d-0
    Lit3N 0.00731553d-0
  /
END
"""


def elm_cn(ncell: int = 2 * 1024 * 1024, tran_dt: float = 1800.0, seed: int = SEED, elm: bool = False,
           full: bool = False, degas: bool = False, flow: Optional[str] = None, anisothermal: bool = False) -> Workload:
    """C4(b): SOMDECOMP + NITRIFICATION + DENITRIFICATION on 6 aqueous + 10
    immobile species.  ``elm=True`` is the ELM_PFLOTRAN build in BGC-only
    coupling: the moisture / oxygen / temperature scalars, soil depth,
    decomposition scalar, dry bulk density and Clapp-Hornberger b are per-cell
    ELM inputs; otherwise the CLM-CN temperature response and the log-theta
    moisture response of the stand-alone build are evaluated from T and theta.
    ``flow`` ("CLMCN" or "DLEM", with ``elm``): the ELM build next to a flow mode -- SOMDECOMP's moisture factor
    comes from GetMoistureResponse (elm_rspfuncs.F90:124-237) on the cell's sucsat / bsw / bulk density or field
    capacity / effective porosity instead of ELM's w_scalar."""
    rng = np.random.default_rng(seed)
    if flow:
        assert elm and flow in ("CLMCN", "DLEM")
    abiotic = ("""ABIOTIC_FACTORS
        MOISTURE_RESPONSE_FUNCTION
          %s
        /
      /""" % flow) if flow else "" if elm else """ABIOTIC_FACTORS
        TEMPERATURE_RESPONSE_FUNCTION
          CLMCN
        /
        MOISTURE_RESPONSE_FUNCTION
          LOGTHETA
        /
      /"""
    deck = C4S_DECK.replace("__ABIOTIC__", abiotic)
    if full:
        # the whole ELM-CN network: + plant N uptake and kinetic NH4+ sorption, with their
        # tracking species (ngee/CLMCNplus/clm_nuptake3.in, clm_nh4absorption.in)
        deck = deck.replace("    Lit3N\n  /", "    Lit3N\n    PlantN\n    Plantnh4uptake\n    Plantno3uptake\n"
                            "    NH4sorb\n  /")
        deck = deck.replace("    DENITRIFICATION\n", "    PLANTN\n      AMMONIUM_HALF_SATURATION 1.0d-5\n"
                            "      NITRATE_HALF_SATURATION 1.0d-5\n      AMMONIUM_INHIBITION_NITRATE 1.0d0\n    /\n"
                            "    LANGMUIR\n      NAME_AQ NH4+\n      NAME_SORB NH4sorb\n"
                            "      EQUILIBRIUM_CONSTANT 1.0d4\n      KINETIC_CONSTANT 1.0d-5\n      S_MAX 1.0d-1\n    /\n"
                            "    DENITRIFICATION\n")
        deck = deck.replace("    Lit3N 0.00731553d-0\n", "    Lit3N 0.00731553d-0\n    PlantN 1.d-20\n"
                            "    Plantnh4uptake 1.d-20\n    Plantno3uptake 1.d-20\n"
                            "    NH4sorb 1.d-3\n")
    if degas:
        # + the CNDEGAS sandbox (reaction_sandbox_cndegas.F90): CO2 / N2O / N2 exchange with gas reservoirs
        # kept as immobile species, and the pH-stat.  No deck of the reference's test suite runs it.
        deck = deck.replace("    Lit3N\n  /", "    Lit3N\n    CO2imm\n    N2Oimm\n    N2imm\n    Himm\n  /")
        deck = deck.replace("    DENITRIFICATION\n", "    CNDEGAS\n      KINETIC_CONSTANT_CO2 2.0d-5\n"
                            "      KINETIC_CONSTANT_N2O 1.0d-5\n      KINETIC_CONSTANT_N2 5.0d-6\n"
                            "      KINETIC_CONSTANT_H+ 1.0d-5\n      FIXPH 6.5\n    /\n    DENITRIFICATION\n", 1)
        deck = deck.replace("    Lit3N 0.00731553d-0\n", "    Lit3N 0.00731553d-0\n    CO2imm 1.66d-2\n"
                            "    N2Oimm 1.3d-5\n    N2imm 32.5d0\n    Himm 1.d-20\n")
    dk = chem.read_deck(deck)
    if degas:
        names = list(dk.chemistry.immobile)
        dk.chemistry.cndegas["gas_ids"] = {"CO2(g)": names.index("CO2imm"), "N2O(g)": names.index("N2Oimm"),
                                           "N2(g)": names.index("N2imm")}
        dk.chemistry.cndegas["cell_state_mode"] = 2 if elm else 0
    net = chem.ReactionNetwork(dk.chemistry, chem.Database(_read("clmcnplus_CLM-CN_database.dat")),
                               dk.reference_temperature, not anisothermal)
    assert dk.chemistry.unsupported == [], dk.chemistry.unsupported
    net.elm_pflotran = bool(elm)
    net.elm_flow_coupled = bool(flow)
    cfg = abi.ReactionConfig(net)
    sp = constraint.equilibrate_constraint(net, dk.constraints["initial"], den_kg=1000.0)
    st = abi.HostState(cfg, ncell)
    constraint.fill_cells(st, sp)
    st["immobile"][...] = sp.immobile[:, None] * np.exp(rng.standard_normal((net.nimcomp, ncell)))
    if full:
        for nm in ("PlantN", "Plantnh4uptake", "Plantno3uptake"):
            st["immobile"][net.immobile_names.index(nm)] = 1.0e-20   # trackers are reset by ELM every step
        st["immobile"][net.immobile_names.index("NH4sorb")] = 10.0 ** rng.uniform(-4.0, -1.2, ncell)
    # mineral N from depleted to fertilised, independent of the pH
    scale = np.ones((net.naqcomp, ncell))
    for nm, lo, hi in (("NH4+", -7.0, -2.3), ("NO3-", -8.0, -3.5), ("CO2(aq)", -10.0, -4.0)):
        i = net.primary_names.index(nm)
        v = 10.0 ** rng.uniform(lo, hi, ncell)
        scale[i] = v / st["total"][i]
    st["total"][...] *= scale
    st["pri_molal"][...] *= scale
    st["temp"][...] = rng.uniform(-5.0, 30.0, ncell)
    st["sat"][...] = rng.uniform(0.3, 1.0, ncell)
    st["den_kg"][...] = 1000.0
    st["porosity"][...] = rng.uniform(0.25, 0.5, ncell)
    st["volume"][...] = 1.0
    if elm:
        st["elm_w_scalar"][...] = rng.uniform(0.05, 1.0, ncell)
        st["elm_o_scalar"][...] = rng.uniform(0.2, 1.0, ncell)
        st["elm_t_scalar"][...] = rng.uniform(0.05, 1.5, ncell)
        st["elm_zsoil"][...] = rng.uniform(0.01, 3.0, ncell)
        st["elm_kscalar_decomp_c"][...] = 1.0
        st["elm_bulkdensity_dry"][...] = rng.uniform(900.0, 1600.0, ncell)
        st["elm_bsw"][...] = rng.uniform(2.0, 10.0, ncell)
        # plant N demand: zero at night / in winter for a third of the cells
        st["elm_rate_plantndemand"][...] = np.where(rng.random(ncell) < 0.33, 0.0, 10.0 ** rng.uniform(-9.0, -6.5, ncell))
    if flow:
        st["elm_sucsat"][...] = rng.uniform(50.0, 600.0, ncell)          # mm H2O (ELM's Clapp-Hornberger table)
        st["elm_effporosity"][...] = st["porosity"] * rng.uniform(0.8, 1.0, ncell)
        st["elm_watfc"][...] = st["elm_effporosity"] * rng.uniform(0.2, 0.6, ncell)
    if degas:
        for nm, lo, hi in (("CO2imm", 1.0e-2, 8.0e-2), ("N2Oimm", 5.0e-6, 1.0e-4), ("N2imm", 30.0, 35.0)):
            st["immobile"][net.immobile_names.index(nm)] = rng.uniform(lo, hi, ncell)
        st["immobile"][net.immobile_names.index("Himm")] = 1.0e-20
        if elm:
            st["pres"][...] = rng.uniform(0.9e5, 1.3e5, ncell)
    name = (("c4f_elm_cn_full" if full else "c4s_elm_cn") + ("_elmscalars" if elm else "") + ("_degas" if degas else "")
            + (f"_flow_{flow.lower()}" if flow else ""))
    note = "SOMDECOMP (7 rxns, N immobilisation from NH4+/NO3-) + NITRIFICATION + DENITRIFICATION"
    if full:
        note += " + PLANTN + LANGMUIR"
    if degas:
        note += " + CNDEGAS"
    return Workload(name, cfg, st, tran_dt, net, note + f", {net.ncomp} dof")


# --------------------------------------------------------------------------- #
C6_DECK = """
# C6: sorption without surface complexes -- the ion exchange network of
# ascem/batch/ion-exchange-valocchi.in (Na+ reference, Ca++, Mg++ on a CEC tied to nothing /
# to Halite), plus one KD isotherm of every type and a dynamic KD
CHEMISTRY
  PRIMARY_SPECIES
    Na+
    Ca++
    Mg++
    Cl-
    K+
    Tracer
    Tracer2
    NO3-
  /
  MINERALS
    Halite
  /
  MINERAL_KINETICS
    Halite
      RATE_CONSTANT 1.d-40 mol/cm^2-sec
    /
  /
  SORPTION
    ION_EXCHANGE_RXN
      CEC 750. eq/m^3
      CATIONS
        Ca++  3.38638672536d0
        Na+   1.d0 REFERENCE
        Mg++  6.00240096038d0
      /
    /
    ION_EXCHANGE_RXN
      MINERAL Halite
      CEC 5.d4
      CATIONS
        Na+   1.d0 REFERENCE
        K+    2.5d0
      /
    /
    ISOTHERM_REACTIONS
      Tracer
        TYPE LINEAR
        DISTRIBUTION_COEFFICIENT 500.
      /
      Tracer2
        TYPE FREUNDLICH
        DISTRIBUTION_COEFFICIENT 20.
        FREUNDLICH_N 1.5
      /
      K+
        TYPE LANGMUIR
        DISTRIBUTION_COEFFICIENT 3.d3
        LANGMUIR_B 0.2
      /
    /
    DYNAMIC_KD_REACTIONS
      NO3-
        REFERENCE_SPECIES Cl-
        REFERENCE_SPECIES_HIGH 0.2
        KD_LOW 0.05
        KD_HIGH 2.0
        KD_POWER 1.5
      /
    /
  /
  DATABASE ./hanford_subset.dat
  LOG_FORMULATION
END
CONSTRAINT initial
  CONCENTRATIONS
    Na+     8.65d-2 T
    Ca++    1.82d-2 T
    Mg++    1.11d-2 T
    Cl-     2.d-3   Z
    K+      1.d-3   T
    Tracer  1.d-6   T
    Tracer2 1.d-5   T
    NO3-    1.d-4   T
  /
  MINERALS
    Halite 1.d-5 1.d0 cm^2/cm^3
  /
END
CONSTRAINT inlet
  CONCENTRATIONS
    Na+     9.4d-3  T
    Ca++    5.d-4   T
    Mg++    2.13d-3 T
    Cl-     1.d-2   Z
    K+      1.d-4   T
    Tracer  1.d-3   T
    Tracer2 1.d-3   T
    NO3-    1.d-3   T
  /
  MINERALS
    Halite 1.d-5 1.d0 cm^2/cm^3
  /
END
"""


def _sorbed_totals(net: chem.ReactionNetwork, water, volfrac: float, den_kg: float, porosity: float,
                   particle_density: float = 2650.0) -> np.ndarray:
    """total_sorb_eq of one speciated water (ion exchange by bisection on the exchanger's
    charge balance, KD isotherms, dynamic KD): the resident sorbed state the C6 cells start from.
    Set-up helper with its own arithmetic -- the step re-equilibrates anyway."""
    m = np.asarray(water.pri_molal, dtype=np.float64)
    act = m * np.asarray(water.pri_act_coef, dtype=np.float64)
    Z = np.asarray(net.primary_Z, dtype=np.float64)
    out = np.zeros(net.naqcomp)
    ix = net.ionx
    if ix:
        for r in range(len(ix["CEC"])):
            ids = ix["cationid"][ix["ptr"][r]:ix["ptr"][r + 1]]
            ks = np.asarray(ix["k"][ix["ptr"][r]:ix["ptr"][r + 1]])
            omega = ix["CEC"][r] * (volfrac if ix["to_surf"][r] >= 0 else 1.0)
            zr = Z[ids[0]]

            def total(kd):
                return sum(ks[j] * act[ids[j]] * kd ** (Z[ids[j]] / zr) for j in range(len(ids)))

            lo, hi = 1.0e-30, 1.0e30
            for _ in range(400):
                mid = np.sqrt(lo * hi)
                if total(mid) > 1.0:
                    hi = mid
                else:
                    lo = mid
            kd = np.sqrt(lo * hi)
            for j in range(len(ids)):
                out[ids[j]] += ks[j] * act[ids[j]] * kd ** (Z[ids[j]] / zr) * omega / Z[ids[j]]
    if net.dynkd:
        d = net.dynkd
        for r in range(len(d["specid"])):
            kd = d["low"][r] + (m[d["refspecid"][r]] / d["refspechigh"][r]) ** d["power"][r] * (d["high"][r] - d["low"][r])
            out[d["specid"][r]] += kd * m[d["specid"][r]] * 250.0
    if net.kd:
        k = net.kd
        for r in range(len(k["specid"])):
            i = k["specid"][r]
            kd = k["coeff"][r]
            if k["ikd_units"] == 1:
                kd = kd * den_kg * (1.0 - porosity) * particle_density * 1.0e-3
            if k["type"][r] == 1:
                out[i] += kd * m[i]
            elif k["type"][r] == 2:
                out[i] += kd * m[i] * k["langmuir_b"][r] / (1.0 + kd * m[i])
            else:
                out[i] += kd * m[i] ** (1.0 / k["freundlich_n"][r])
    return out


def ion_exchange(ncell: int = 1 << 20, tran_dt: float = 86400.0, seed: int = SEED) -> Workload:
    """C6: ion exchange (mixed valences: inner Newton; equal valences: closed form; CEC absolute and
    tied to a mineral), linear / Langmuir / Freundlich KD isotherms and a dynamic KD.  Cells are mixes
    of the resident water and the inlet water of the Valocchi problem; the sorbed state is the resident
    one, so the step re-partitions every cation."""
    rng = np.random.default_rng(seed)
    dk, net = chem.load_network(C6_DECK, _read("hanford_subset.dat"))
    assert dk.chemistry.unsupported == [], dk.chemistry.unsupported
    cfg = abi.ReactionConfig(net)
    den = eos.water_density_ifc67(25.0)
    waters = [constraint.equilibrate_constraint(net, dk.constraints[k], den_kg=den, porosity=0.25)
              for k in ("initial", "inlet")]
    st = abi.HostState(cfg, ncell)
    f = rng.random(ncell)
    wts = np.stack([1.0 - f, f])
    _mix_fill(st, waters, wts, rng, jitter=0.0)
    st["den_kg"][...] = den
    st["porosity"][...] = 0.25
    st["volume"][...] = rng.uniform(0.5, 2.0, ncell)
    st["sat"][...] = rng.uniform(0.4, 1.0, ncell)
    st["temp"][...] = 25.0
    vf = 10.0 ** rng.uniform(-6.0, -4.0, ncell)
    st["mnrl_volfrac"][...] = vf[None, :]
    st["mnrl_area"][...] = 100.0
    # the exchanger and the KD sites hold what the resident water left there; the mixed water of
    # the cell then re-partitions every cation (sorbed totals are linear in the mineral-bound CEC)
    s_lo = _sorbed_totals(net, waters[0], 1.0e-6, den, 0.25)
    s_hi = _sorbed_totals(net, waters[0], 1.0e-4, den, 0.25)
    st["total_sorb_eq"][...] = s_lo[:, None] + (s_hi - s_lo)[:, None] * ((vf - 1.0e-6) / (1.0e-4 - 1.0e-6))[None, :]
    return Workload("c6_ion_exchange_kd", cfg, st, tran_dt, net,
                    "2 ion-exchange reactions (5 cations), 3 KD isotherms, 1 dynamic KD, kinetic Halite")


C7_DECK = """
# C7: the other kinetic terms of RReaction -- two general reactions (a third-order forward /
# first-order backward pair and an irreversible one), a radioactive decay with a daughter, an immobile
# decay -- next to a secondary complex, so that d(total)/d(free) is not the identity
CHEMISTRY
  PRIMARY_SPECIES
    A(aq)
    B(aq)
    C(aq)
  /
  SECONDARY_SPECIES
    AB(aq)
  /
  IMMOBILE_SPECIES
    Xim
  /
  GENERAL_REACTION
    REACTION A(aq) + 2 B(aq) <-> C(aq)
    FORWARD_RATE 4.d1
    BACKWARD_RATE 2.d-7
  /
  GENERAL_REACTION
    REACTION C(aq) <-> 0.5 A(aq)
    FORWARD_RATE 1.d-7
    BACKWARD_RATE 0.d0
  /
  RADIOACTIVE_DECAY_REACTION
    REACTION B(aq) <-> 0.5 C(aq)
    HALF_LIFE 30. d
  /
  IMMOBILE_DECAY_REACTION
    SPECIES_NAME Xim
    HALF_LIFE 10. d
  /
  DATABASE ./hanford_subset.dat
  LOG_FORMULATION
  ACTIVITY_COEFFICIENTS TIMESTEP
END
CONSTRAINT initial
  CONCENTRATIONS
    A(aq)  1.d-3 T
    B(aq)  2.d-3 T
    C(aq)  1.d-5 T
  /
  IMMOBILE
    Xim 1.d-2
  /
END
CONSTRAINT inlet
  CONCENTRATIONS
    A(aq)  1.d-5 T
    B(aq)  5.d-4 T
    C(aq)  3.d-3 T
  /
  IMMOBILE
    Xim 1.d-4
  /
END
"""


C7_SORPTION = """  SORPTION
    ISOTHERM_REACTIONS
      B(aq)
        TYPE FREUNDLICH
        DISTRIBUTION_COEFFICIENT 40.
        FREUNDLICH_N 1.3
      /
      C(aq)
        TYPE LINEAR
        DISTRIBUTION_COEFFICIENT 150.
      /
    /
    DYNAMIC_KD_REACTIONS
      B(aq)
        REFERENCE_SPECIES A(aq)
        REFERENCE_SPECIES_HIGH 3.d-3
        KD_LOW 0.2
        KD_HIGH 1.5
        KD_POWER 1.2
      /
    /
  /
"""


def general_decay(ncell: int = 1 << 20, tran_dt: float = 86400.0, seed: int = SEED, sorbing: bool = False) -> Workload:
    """C7: RGeneral, RRadioactiveDecay and RImmobileDecay on mixes of two waters.  ``sorbing``: the
    decaying parent B(aq) and its daughter also sorb (Freundlich + dynamic KD, linear KD), so the sorbed
    inventory decays too and the Jacobian needs d(total_sorb)/d(free) (reaction.F90:5257-5305)"""
    rng = np.random.default_rng(seed)
    deck = C7_DECK.replace("  DATABASE ./hanford_subset.dat", C7_SORPTION + "  DATABASE ./hanford_subset.dat") if sorbing \
        else C7_DECK
    dk, net = chem.load_network(deck, _read("hanford_subset.dat"))
    assert dk.chemistry.unsupported == [], dk.chemistry.unsupported
    cfg = abi.ReactionConfig(net)
    den = eos.water_density_ifc67(25.0)
    waters = [constraint.equilibrate_constraint(net, dk.constraints[k], den_kg=den, porosity=0.3)
              for k in ("initial", "inlet")]
    st = abi.HostState(cfg, ncell)
    f = rng.random(ncell)
    _mix_fill(st, waters, np.stack([1.0 - f, f]), rng, jitter=0.0)
    st["immobile"][...] = (10.0 ** rng.uniform(-4.0, -2.0, ncell))[None, :]
    st["den_kg"][...] = den
    st["porosity"][...] = 0.3
    st["volume"][...] = rng.uniform(0.5, 2.0, ncell)
    st["sat"][...] = rng.uniform(0.4, 1.0, ncell)
    st["temp"][...] = 25.0
    if sorbing:
        # sorbed totals in equilibrium with the first water, so the step re-partitions towards the mix
        s0 = _sorbed_totals(net, waters[0], 0.0, den, 0.3)
        st["total_sorb_eq"][...] = s0[:, None]
    return Workload("c7s_general_decay_sorbing" if sorbing else "c7_general_decay", cfg, st, tran_dt, net,
                    "2 general reactions, 1 radioactive decay with daughter, 1 immobile decay, 1 complex"
                    + (", parent and daughter sorbing (2 KD isotherms, 1 dynamic KD)" if sorbing else ""))


C7G_DECK = """
# C7g: an ACTIVE gas phase (RTotalGas) -- the chemistry of default/batch/radon.in (Rn(g) over Rn(aq), zero-order
# generation by the RADON sandbox, radioactive decay of the aqueous + gaseous inventory) next to the
# carbonate system with CO2(g) active, so that one gas couples two components and carries water
CHEMISTRY
  PRIMARY_SPECIES
    Rn(aq)
    SiO2(aq)
    H+
    HCO3-
  /
  SECONDARY_SPECIES
    OH-
    CO3--
    CO2(aq)
  /
  ACTIVE_GAS_SPECIES
    GAS_TRANSPORT_IS_UNVETTED
    Rn(g)
    CO2(g)
  /
  MINERALS
    Quartz
  /
  MINERAL_KINETICS
    Quartz
      RATE_CONSTANT 1.d-13
    /
  /
  REACTION_SANDBOX
    RADON
      SPECIES_NAME Rn(aq)
      MINERAL_NAME Quartz
      RADON_GENERATION_RATE 1.1627850420873736e-19
    /
  /
  RADIOACTIVE_DECAY_REACTION
    REACTION Rn(aq) <->
    HALF_LIFE 3.8235 d
  /
  DATABASE ./hanford_subset.dat
  LOG_FORMULATION
  ACTIVITY_COEFFICIENTS TIMESTEP
END
CONSTRAINT initial
  CONCENTRATIONS
    Rn(aq)    1.d-18  T
    SiO2(aq)  1.d-4   T
    H+        7.5     P
    HCO3-     2.d-3   T
  /
  MINERALS
    Quartz  0.5d0 1.d2 m^2/m^3
  /
END
CONSTRAINT inlet
  CONCENTRATIONS
    Rn(aq)    1.d-15  T
    SiO2(aq)  3.d-5   T
    H+        5.5     P
    HCO3-     1.d-4   T
  /
  MINERALS
    Quartz  0.5d0 1.d2 m^2/m^3
  /
END
"""


def active_gas(ncell: int = 1 << 20, tran_dt: float = 86400.0, seed: int = SEED, anisothermal: bool = False) -> Workload:
    """C7g: two active gas species (reaction_gas.F90:87-174), the RADON sandbox and radioactive decay of an
    inventory that sits mostly in the gas phase; liquid saturations from 0.05 to 0.9"""
    rng = np.random.default_rng(seed)
    dk = chem.read_deck(C7G_DECK)
    net = chem.ReactionNetwork(dk.chemistry, chem.Database(_read("hanford_subset.dat")), use_isothermal=not anisothermal)
    assert dk.chemistry.unsupported == [], dk.chemistry.unsupported
    cfg = abi.ReactionConfig(net)
    den = eos.water_density_ifc67(25.0)
    waters = [constraint.equilibrate_constraint(net, dk.constraints[k], den_kg=den, porosity=0.3)
              for k in ("initial", "inlet")]
    st = abi.HostState(cfg, ncell)
    f = rng.random(ncell)
    _mix_fill(st, waters, np.stack([1.0 - f, f]), rng, jitter=0.0)
    st["den_kg"][...] = den
    st["porosity"][...] = rng.uniform(0.2, 1.0, ncell)
    st["volume"][...] = rng.uniform(0.5, 2.0, ncell)
    # a quarter of the cells nearly dry.  (Not the radon deck's 1e-5: with CO2(g) active the aqueous carbonate
    # system of such a cell is a 1e-5 share of the inventory and H+ is conditioned like 1/sat -- rounding-level
    # differences between two correct solvers would reach 1e-10.  The deck itself runs in the tests.)
    sat = np.where(rng.random(ncell) < 0.25, 0.05, rng.uniform(0.05, 0.9, ncell))
    st["sat"][...] = sat
    st["sat_gas"][...] = 1.0 - sat
    st["temp"][...] = rng.uniform(5.0, 60.0, ncell) if anisothermal else 25.0
    st["mnrl_volfrac"][0, :] = rng.uniform(0.1, 1.0, ncell)
    st["mnrl_area"][0, :] = 1.0e2
    # gas-phase totals of the previous step: in equilibrium with the first water
    # (the fixed accumulation reads them, reaction.F90:5761-5769)
    pp_rn = 10.0 ** rng.uniform(-22.0, -16.0, ncell)
    pp_co2 = 10.0 ** rng.uniform(-4.0, -2.0, ncell)
    rt = 8.31446 * (st["temp"][0] + 273.15)
    names = net.primary_names
    st["total_gas"][names.index("Rn(aq)"), :] = pp_rn * 1.0e5 / rt * 1.0e-3
    st["total_gas"][names.index("H+"), :] = pp_co2 * 1.0e5 / rt * 1.0e-3
    st["total_gas"][names.index("HCO3-"), :] = pp_co2 * 1.0e5 / rt * 1.0e-3
    st["gas_pp"][0, :] = pp_rn
    st["gas_pp"][1, :] = pp_co2
    return Workload("c7gt_active_gas_anisothermal" if anisothermal else "c7g_active_gas", cfg, st, tran_dt, net,
                    "2 active gas species (Rn, CO2), RADON sandbox, radioactive decay of the aqueous + gaseous "
                    "inventory, 3 complexes, kinetic Quartz")


C8_DECK = """
# C8: Monod-type microbial reactions (the ABCD network of default/batch/ABCD_microbial*.in):
# immobile biomass with yield and decay, a second reaction on aqueous biomass, all four inhibition
# types, activation energy, concentrations as activities with Debye-Hueckel coefficients
CHEMISTRY
  PRIMARY_SPECIES
    A(aq)
    B(aq)
    C(aq)
    D(aq)
    Na+
    Cl-
  /
  IMMOBILE_SPECIES
    D(im)
  /
  MICROBIAL_REACTION
    CONCENTRATION_UNITS ACTIVITY
    REACTION A(aq) + 2 B(aq) <-> 1.5 C(aq)
    RATE_CONSTANT 1.d-6
    ACTIVATION_ENERGY 6.5d0 kJ/mol
    MONOD
      SPECIES_NAME A(aq)
      HALF_SATURATION_CONSTANT 1.d-5
      THRESHOLD_CONCENTRATION 1.d-20
    /
    MONOD
      SPECIES_NAME B(aq)
      HALF_SATURATION_CONSTANT 1.d-4
      THRESHOLD_CONCENTRATION 1.d-11
    /
    INHIBITION
      SPECIES_NAME C(aq)
      TYPE THRESHOLD 1.d5
      INHIBITION_CONSTANT -1.d-4
    /
    INHIBITION
      SPECIES_NAME Cl-
      TYPE MONOD
      INHIBITION_CONSTANT 5.d-2
    /
    BIOMASS
      SPECIES_NAME D(im)
      YIELD 0.01d0
    /
  /
  MICROBIAL_REACTION
    CONCENTRATION_UNITS ACTIVITY
    REACTION C(aq) <-> 0.5 A(aq)
    RATE_CONSTANT 2.d-4
    MONOD
      SPECIES_NAME C(aq)
      HALF_SATURATION_CONSTANT 2.d-4
    /
    INHIBITION
      SPECIES_NAME A(aq)
      TYPE INVERSE_MONOD
      INHIBITION_CONSTANT 6.d-4
    /
    INHIBITION
      SPECIES_NAME B(aq)
      TYPE SMOOTHSTEP 1.5
      INHIBITION_CONSTANT 3.d-4
    /
    BIOMASS
      SPECIES_NAME D(aq)
      YIELD 0.05d0
    /
  /
  GENERAL_REACTION
    REACTION D(aq) <->
    FORWARD_RATE 60.d-9 1/min
    BACKWARD_RATE 0.d0
  /
  IMMOBILE_DECAY_REACTION
    SPECIES_NAME D(im)
    RATE_CONSTANT 1.d-9
  /
  DATABASE ./hanford_subset.dat
  LOG_FORMULATION
  ACTIVITY_COEFFICIENTS TIMESTEP
END
CONSTRAINT initial
  CONCENTRATIONS
    A(aq)  1.d-4 T
    B(aq)  1.d-3 T
    C(aq)  1.d-7 T
    D(aq)  1.d-5 T
    Na+    5.d-2 T
    Cl-    5.d-2 Z
  /
  IMMOBILE
    D(im) 1.d-4
  /
END
CONSTRAINT inlet
  CONCENTRATIONS
    A(aq)  2.d-3 T
    B(aq)  3.d-4 T
    C(aq)  4.d-4 T
    D(aq)  1.d-7 T
    Na+    1.d-3 T
    Cl-    1.d-3 Z
  /
  IMMOBILE
    D(im) 1.d-6
  /
END
"""


def microbial(ncell: int = 1 << 20, tran_dt: float = 86400.0, seed: int = SEED) -> Workload:
    """C8: RMicrobial with every option, next to a general reaction and an immobile decay"""
    rng = np.random.default_rng(seed)
    dk, net = chem.load_network(C8_DECK, _read("hanford_subset.dat"))
    assert dk.chemistry.unsupported == [], dk.chemistry.unsupported
    cfg = abi.ReactionConfig(net)
    den = eos.water_density_ifc67(25.0)
    waters = [constraint.equilibrate_constraint(net, dk.constraints[k], den_kg=den, porosity=0.3)
              for k in ("initial", "inlet")]
    st = abi.HostState(cfg, ncell)
    f = rng.random(ncell)
    _mix_fill(st, waters, np.stack([1.0 - f, f]), rng, jitter=0.0)
    st["immobile"][...] = (10.0 ** rng.uniform(-6.0, -3.0, ncell))[None, :]
    st["den_kg"][...] = den
    st["porosity"][...] = 0.3
    st["volume"][...] = rng.uniform(0.5, 2.0, ncell)
    st["sat"][...] = rng.uniform(0.4, 1.0, ncell)
    st["temp"][...] = rng.uniform(5.0, 35.0, ncell)
    return Workload("c8_microbial", cfg, st, tran_dt, net,
                    "2 microbial reactions (4 inhibition types, immobile and aqueous biomass, activation energy), "
                    "1 general reaction, 1 immobile decay")


def by_name(name: str, ncell: Optional[int] = None, tran_dt: Optional[float] = None) -> Workload:
    table = {
        "c1": (calcite_batch, {}),
        "c2": (calcite_column, {}),
        "c2pf": (calcite_column, {"prefactors": "full"}),
        "c2pfp": (calcite_column, {"prefactors": "primary"}),
        "c3": (hanford, {"variant": "base"}),
        "c3mr": (hanford, {"variant": "mr"}),
        "c3an": (hanford, {"variant": "base", "activity_newton": True}),
        "c3t": (hanford, {"variant": "base", "anisothermal": True}),
        "c3aw": (hanford, {"variant": "base", "activity_h2o": True}),
        "c3tg": (hanford, {"variant": "base", "total_as_guess": True}),
        "c3sf": (hanford, {"variant": "base", "inner_newton_sites": True}),
        "c2t": (calcite_column, {"anisothermal": True}),
        "c2ng": (calcite_column, {"no_geochemistry": True}),
        "c2sb": (calcite_column, {"sandbox": True}),
        "c4": (clm_cn, {}),
        "c4s": (elm_cn, {}),
        "c4se": (elm_cn, {"elm": True}),
        "c4fe": (elm_cn, {"full": True, "elm": True}),
        "c4g": (elm_cn, {"degas": True}),
        "c4ge": (elm_cn, {"degas": True, "elm": True}),
        "c4st": (elm_cn, {"elm": True, "flow": "DLEM", "anisothermal": True}),   # next to a thermal flow mode
        "c4sw": (elm_cn, {"elm": True, "flow": "CLMCN"}),
        "c4sd": (elm_cn, {"elm": True, "flow": "DLEM"}),
        "c4fw": (elm_cn, {"full": True, "elm": True, "flow": "CLMCN"}),
        "c5": (hanford, {"variant": "minerals"}),
        "c6": (ion_exchange, {}),
        "c7": (general_decay, {}),
        "c7s": (general_decay, {"sorbing": True}),
        "c7g": (active_gas, {}),
        "c7gt": (active_gas, {"anisothermal": True}),
        "c8": (microbial, {}),
    }
    fn, kw = table[name]
    kw = dict(kw)
    if ncell is not None and name != "c1":
        kw["ncell"] = ncell
    if tran_dt is not None and name != "c1":
        kw["tran_dt"] = tran_dt
    return fn(**kw)


# --------------------------------------------------------------------------- #
def flops_model(net: chem.ReactionNetwork) -> Tuple[float, float]:
    """Algorithmic flops of one Newton iteration, split into
    (residual/Jacobian evaluation, linear solve): the closed form of SURVEY.md
    section 8(d) with add/sub/mul/div/compare = 1, FMA = 2 and
    exp/log/pow/sqrt = 20.  A cell-substep that converges in k iterations
    evaluates k times and solves k-1 times."""
    naq, n = net.naqcomp, net.ncomp
    ncx = net.neqcplx
    nnz = sum(len(r.ids) for r in net.sec_rxn)
    nnz2 = sum(len(r.ids) ** 2 for r in net.sec_rxn)
    f = 8.0 * nnz + 2.0 * nnz2 + 3.0 * ncx + 2.0 * naq * naq
    f += 20.0 * (2.0 * naq + ncx + nnz)
    if net.chem.act_coef_update_frequency == chem.ACT_COEF_FREQUENCY_NEWTON_ITER:
        f += 9.0 * (naq + ncx) + 20.0 * (naq + ncx + 1)
    for nm in net.kinmnrl_names:
        m = len(net.mnrl_rxn[nm].ids)
        f += 60.0 + 20.0 * (m + 1)
    for rx in net.srfcplxrxn:
        f += 20.0 * (2 * len(rx.complexes) + 1) + 12.0 * len(rx.complexes) * 9
        if rx.itype == "MULTIRATE_KINETIC":
            f += 6.0 * len(rx.rates) * naq
    if net.clmcn is not None:
        f += 45.0 * net.clmcn["nrxn"] + 40.0
    if getattr(net, "somdec", None) is not None:
        # per reaction: abiotic factors, smoothing, Monod terms, rate caps, ~25 residual /
        # Jacobian updates per column (3 columns in the immobilisation branch)
        f += 150.0 * net.somdec["scalars"]["nrxn"] + 20.0 * 6
    if getattr(net, "nitrif", None) is not None:
        f += 60.0 + 20.0 * 7
    if getattr(net, "denitr", None) is not None:
        f += 40.0 + 20.0 * 2
    if getattr(net, "plantn", None) is not None:
        f += 90.0
    if getattr(net, "langmuir", None) is not None:
        f += 50.0
    if getattr(net, "cndegas", None) is not None:
        f += 3.0 * (40.0 + 20.0 * 4) + 30.0
    g = getattr(net, "general", None)
    if g:
        # two exponentials (~25 flops each) and a log per rate law, residual and Jacobian updates
        f += sum(80.0 + 4.0 * (g["ptr"][k + 1] - g["ptr"][k]) * (g["fwd_ptr"][k + 1] - g["fwd_ptr"][k] +
                                                                g["bwd_ptr"][k + 1] - g["bwd_ptr"][k] + 1)
                 for k in range(len(g["kf"])))
    rd = getattr(net, "radiodecay", None)
    if rd:
        f += sum(4.0 + (rd["ptr"][k + 1] - rd["ptr"][k]) * (2.0 + 4.0 * naq) for k in range(len(rd["kf"])))
    if getattr(net, "immdecay", None):
        f += 4.0 * len(net.immdecay["k"])
    mb = getattr(net, "microbial", None)
    if mb:
        for k in range(len(mb["rate_constant"])):
            nm = mb["monod_ptr"][k + 1] - mb["monod_ptr"][k]
            nh = mb["inhibition_ptr"][k + 1] - mb["inhibition_ptr"][k]
            ns = mb["ptr"][k + 1] - mb["ptr"][k]
            f += 30.0 + 12.0 * nm + 40.0 * nh + (nm + nh + 1) * (2.0 * ns + 12.0)
    f += 10.0 * n
    solve = (2.0 / 3.0) * n ** 3 + 5.0 * n * n + 20.0 * n
    return f, solve
