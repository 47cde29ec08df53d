"""Initial speciation of transport constraints on the host (numpy).

Mirrors ``ReactionEquilibrateConstraint`` (src/pflotran/reaction.F90:1328-2117)
for the constraint types the hot-path decks use: T (total aqueous), F (free),
L (log free), P (pH), Z (charge balance), M (mineral equilibrium), G (gas
partial pressure).  Runs once per constraint at set-up; per-cell state for the
GPU step is then assembled from the speciated constraints.

Speciation primitives (totals, activity coefficients, surface complexation)
are independent numpy restatements of reaction.F90:4368-4759 and
reaction_surf_complex.F90:641-900; they are host set-up code and share nothing
with ``oracle/``.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np

from . import abi as _abi
from . import chem as _chem
from .chem import LOG_TO_LN


@dataclass
class Speciation:
    """Everything RTAuxVarCompute leaves in an rt_auxvar for one water."""
    pri_molal: np.ndarray
    total: np.ndarray                 # mol/L
    pri_act_coef: np.ndarray
    sec_act_coef: np.ndarray
    sec_molal: np.ndarray
    ln_act_h2o: float = 0.0
    immobile: Optional[np.ndarray] = None
    mnrl_volfrac: Optional[np.ndarray] = None
    mnrl_area: Optional[np.ndarray] = None
    free_site: Optional[np.ndarray] = None
    eqsrfcplx_conc: Optional[np.ndarray] = None
    total_sorb_eq: Optional[np.ndarray] = None
    kinmr_total_sorb: Optional[np.ndarray] = None   # [naq*(nrate+1) rows] flattened like pfrx_state
    num_iterations: int = 0
    extras: Dict[str, float] = field(default_factory=dict)


def total_aqueous(net: _chem.ReactionNetwork, pri_molal, pri_act_coef, sec_act_coef, ln_act_h2o, den_kg):
    """RTotalAqueous, reaction.F90:4665-4759 -> (total, sec_molal, dtotal)"""
    naq = net.naqcomp
    ln_conc = np.log(pri_molal)
    ln_act = ln_conc + np.log(pri_act_coef)
    total = pri_molal.astype(np.float64).copy()
    dtotal = np.eye(naq)
    sec = np.zeros(net.neqcplx)
    for k, rx in enumerate(net.sec_rxn):
        lnQK = -net.logK_at_tref(rx.logK_T) * LOG_TO_LN
        if rx.h2o_stoich != 0.0:
            lnQK += rx.h2o_stoich * ln_act_h2o
        for i, s in zip(rx.ids, rx.stoich):
            lnQK += s * ln_act[i]
        sec[k] = np.exp(lnQK) / sec_act_coef[k]
        for i, s in zip(rx.ids, rx.stoich):
            total[i] += s * sec[k]
        for j, sj in zip(rx.ids, rx.stoich):
            t = sj * np.exp(lnQK - ln_conc[j]) / sec_act_coef[k]
            for i, si in zip(rx.ids, rx.stoich):
                dtotal[i, j] += si * t
    f = den_kg * 1.0e-3
    return total * f, sec, dtotal * f


def activity_coefficients(net: _chem.ReactionNetwork, pri_molal, sec_molal):
    """LAG branch of RActivityCoefficients, reaction.F90:4553-4612"""
    Z, a0 = net.primary_Z, net.primary_a0
    cZ, ca0 = net.eqcplx_Z, net.eqcplx_a0
    I = 0.5 * (np.sum(pri_molal * Z * Z) + np.sum(sec_molal * cZ * cZ))
    sq = np.sqrt(I)
    A, B, Bd = net.debyeA, net.debyeB, net.debyeBdot

    def gam(z, a):
        g = np.exp((-z * z * sq * A / (1.0 + a * B * sq) + Bd * I) * LOG_TO_LN)
        return np.where(np.abs(z) > 1.0e-10, g, 1.0)

    return gam(Z, a0), gam(cZ, ca0)


def surface_complexation_eq(net: _chem.ReactionNetwork, irxn: int, pri_molal, pri_act_coef, ln_act_h2o,
                            free_site_guess, mnrl_volfrac, porosity, soil_particle_density):
    """RTotalSorbEqSurfCplx1 without derivatives,
    reaction_surf_complex.F90:641-800 -> (free_site, srfcplx_conc[nsrfcplx], total_sorb[naq])"""
    rx = net.srfcplxrxn[irxn]
    ln_act = np.log(pri_molal) + np.log(pri_act_coef)
    if rx.surface_type == _chem.MINERAL_SURFACE:
        dens = rx.site_density * mnrl_volfrac[net.kinmnrl_names.index(rx.surface_name)]
    elif rx.surface_type == _chem.ROCK_SURFACE:
        dens = rx.site_density * soil_particle_density * (1.0 - porosity)
    else:
        dens = rx.site_density
    nsc = len(net.srfcplx_names)
    conc = np.zeros(nsc)
    tot = np.zeros(net.naqcomp)
    if dens < 1.0e-40:
        return 0.0, conc, tot
    ids = [net.srfcplx_names.index(n) for n in rx.complexes]
    fs = max(free_site_guess, 1.0e-40)
    nonunit = any(abs(net.srfcplx_free_site_stoich[i] - 1.0) > 1.0e-40 for i in ids)
    one_more = False
    it = 0
    while True:
        it += 1
        total = fs
        lnfs = np.log(fs)
        for i in ids:
            r = net.srfcplx_rxn[i]
            lnQK = -net.logK_at_tref(r.logK_T) * LOG_TO_LN
            if r.h2o_stoich != 0.0:
                lnQK += r.h2o_stoich * ln_act_h2o
            lnQK += net.srfcplx_free_site_stoich[i] * lnfs
            for j, s in zip(r.ids, r.stoich):
                lnQK += s * ln_act[j]
            conc[i] = np.exp(lnQK)
            total += net.srfcplx_free_site_stoich[i] * conc[i]
        if one_more:
            break
        if nonunit:
            res = dens - total
            d = 1.0 + sum(net.srfcplx_free_site_stoich[i] * conc[i] / fs for i in ids)
            dfs = res / d
            fs = fs + (0.5 if it > 1000 else 1.0) * dfs
            if abs(dfs / fs) < 1.0e-12 or it > 100000:
                one_more = True
        else:
            total = total / fs
            fs = dens / total
            one_more = True
    for i in ids:
        r = net.srfcplx_rxn[i]
        for j, s in zip(r.ids, r.stoich):
            tot[j] += s * conc[i]
    return fs, conc, tot


def _solve_scaled(Res, Jac, conc, use_log):
    """RSolve (reaction.F90:5457-5516): row scaling, optional ln-scaling, solve"""
    J = Jac.copy()
    r = Res.copy()
    for i in range(len(r)):
        norm = 1.0 / max(1.0, np.max(np.abs(J[i, :])))
        r[i] *= norm
        J[i, :] *= norm
    if use_log:
        J = J * conc[None, :]
    return np.linalg.solve(J, r)


def equilibrate_constraint(net: _chem.ReactionNetwork, cons: _chem.Constraint, den_kg: float = 997.16,
                           porosity: float = 0.25, soil_particle_density: float = 2650.0,
                           max_iterations: int = 10000) -> Speciation:
    """ReactionEquilibrateConstraint, reaction.F90:1328-2117"""
    ch = net.chem
    naq = net.naqcomp
    names = net.primary_names
    by_name = {c[0]: c for c in cons.conc}
    if ch.initialize_with_molality:
        molal_to_molar = den_kg / 1000.0
        molar_to_molal = 1.0
    else:
        molal_to_molar = 1.0
        molar_to_molal = 1000.0 / den_kg
    ctype: List[str] = []
    conc = np.zeros(naq)
    aux: List[str] = []
    for nm in names:
        if nm not in by_name:
            raise KeyError(f"constraint {cons.name}: no concentration for {nm}")
        _, v, t, a = by_name[nm]
        ctype.append(t.upper())
        conc[naq - naq + names.index(nm)] = v
        aux.append(a)
    free = np.full(naq, 1.0e-9)
    total_conc = np.zeros(naq)
    for i in range(naq):
        t = ctype[i]
        if t in ("T", "TOTAL"):
            total_conc[i] = conc[i] * molal_to_molar
        elif t in ("F", "FREE"):
            free[i] = conc[i] * molar_to_molal
        elif t in ("L", "LOG"):
            free[i] = (10.0 ** conc[i]) * molar_to_molal
        elif t in ("Z", "CHG", "M", "MINERAL", "MNRL"):
            free[i] = conc[i] * molar_to_molal
        elif t in ("P", "PH"):
            free[i] = 10.0 ** (-conc[i])
        elif t in ("G", "GAS"):
            if conc[i] <= 0.0:
                conc[i] = 10.0 ** conc[i]
        else:
            raise ValueError(f"constraint type {t} not supported")
    # mineral state
    nk = net.nkinmnrl
    volfrac = np.zeros(nk)
    area = np.zeros(nk)
    for k, nm in enumerate(net.kinmnrl_names):
        if nm in cons.minerals:
            volfrac[k], area[k] = cons.minerals[nm]
    pri = free.copy()
    gam_p = np.ones(naq)
    gam_s = np.ones(net.neqcplx)
    sec = np.zeros(net.neqcplx)
    ln_act_h2o = 0.0
    it = 0
    it_act_on = 0
    compute_act = False
    Z = net.primary_Z
    while True:
        for i in range(naq):
            if ctype[i] in ("F", "FREE", "L", "LOG"):
                pri[i] = free[i]
        if ch.act_coef_update_frequency != _chem.ACT_COEF_FREQUENCY_OFF and compute_act:
            gam_p, gam_s = activity_coefficients(net, pri, sec)
        total, sec, dtotal = total_aqueous(net, pri, gam_p, gam_s, ln_act_h2o, den_kg)
        Res = np.zeros(naq)
        Jac = np.zeros((naq, naq))
        for i in range(naq):
            t = ctype[i]
            if t in ("T", "TOTAL"):
                Res[i] = total[i] - total_conc[i]
                Jac[i, :] = dtotal[i, :]
            elif t in ("F", "FREE", "L", "LOG"):
                Jac[i, i] = 1.0
            elif t in ("Z", "CHG"):
                Res[i] = np.sum(Z * total)
                Jac[i, :] = Z @ dtotal
            elif t in ("P", "PH"):
                pri[i] = 10.0 ** (-conc[i]) / gam_p[i]
                Jac[i, i] = 1.0
            elif t in ("M", "MINERAL", "MNRL"):
                rx = net.mnrl_rxn[aux[i]]
                lnQK = -net.logK_at_tref(rx.logK_T) * LOG_TO_LN
                if rx.h2o_stoich != 0.0:
                    lnQK += rx.h2o_stoich * ln_act_h2o
                for j, s in zip(rx.ids, rx.stoich):
                    lnQK += s * np.log(pri[j] * gam_p[j])
                    Jac[i, j] = s / pri[j]
                Res[i] = lnQK
            elif t in ("G", "GAS"):
                rx = net.gas_rxn[aux[i]]
                lnQK = -net.logK_at_tref(rx.logK_T) * LOG_TO_LN
                if rx.h2o_stoich != 0.0:
                    lnQK += rx.h2o_stoich * ln_act_h2o
                for j, s in zip(rx.ids, rx.stoich):
                    lnQK += s * np.log(pri[j] * gam_p[j])
                    Jac[i, j] = s / pri[j]
                Res[i] = lnQK - np.log(conc[i])
        max_res = np.max(np.abs(Res))
        if ch.use_log_formulation:
            use_log = (it % 2 == 0) if 3 < it < 9 else True
        else:
            use_log = False
        update = _solve_scaled(Res, Jac, pri, use_log)
        prev = pri.copy()
        if use_log:
            update = np.sign(update) * np.minimum(np.abs(update), ch.max_dlnC)
            pri = pri * np.exp(-update)
        else:
            mask = prev <= update
            if np.any(mask):
                mr = np.min(np.abs(prev[mask] / update[mask]))
                if mr <= 1.0:
                    update = update * mr * 0.99
            pri = prev - update
        if np.min(pri) <= 0.0:
            raise FloatingPointError(f"constraint {cons.name}: non-positive free-ion concentration")
        max_rel = np.max(np.abs((pri - prev) / prev))
        it += 1
        if it >= max_iterations:
            raise RuntimeError(f"constraint {cons.name}: no convergence in {it} iterations")
        if max_res < ch.max_residual_tolerance and max_rel < ch.max_relative_change_tolerance:
            if compute_act and it - it_act_on > 1:
                break
            if not compute_act:
                it_act_on = it
            compute_act = True
    sp = Speciation(pri_molal=pri, total=total, pri_act_coef=gam_p, sec_act_coef=gam_s, sec_molal=sec,
                    ln_act_h2o=ln_act_h2o, mnrl_volfrac=volfrac, mnrl_area=area, num_iterations=it)
    # NB: like the reference, total/sec_molal are those of the last RTotal call
    # (before the final update); the first RTAuxVarCompute of a run refreshes them.
    if net.nimcomp:
        sp.immobile = np.array([cons.immobile.get(n, 1.0e-40) for n in net.immobile_names], dtype=np.float64)
    # sorbed state (reaction.F90:2036-2060)
    nrxn = len(net.srfcplxrxn)
    if nrxn:
        sp.free_site = np.full(nrxn, 1.0e-9)
        sp.eqsrfcplx_conc = np.zeros(len(net.srfcplx_names))
        sp.total_sorb_eq = np.zeros(naq)
        for irxn in net.eq_rxn_ids:
            fs, cc, tot = surface_complexation_eq(net, irxn, pri, gam_p, ln_act_h2o, sp.free_site[irxn], volfrac,
                                                  porosity, soil_particle_density)
            sp.free_site[irxn] = fs
            sp.eqsrfcplx_conc += cc
            sp.total_sorb_eq += tot
        if net.mr_rxn_ids:
            rows: List[np.ndarray] = []
            for irxn in net.mr_rxn_ids:
                fs, cc, tot = surface_complexation_eq(net, irxn, pri, gam_p, ln_act_h2o, sp.free_site[irxn], volfrac,
                                                      porosity, soil_particle_density)
                sp.free_site[irxn] = fs
                rows.append(tot)
                for fr in net.srfcplxrxn[irxn].site_fractions:
                    rows.append(fr * tot)
            sp.kinmr_total_sorb = np.concatenate(rows)
    return sp


def fill_cells(state, sp: Speciation, cells=slice(None)) -> None:
    """write one speciated water into cells of a HostState"""
    a = state.a

    def put(name, val):
        if val is not None and a[name].shape[0] > 0:
            a[name][:, cells] = np.asarray(val, dtype=np.float64).reshape(-1, 1)

    put("pri_molal", sp.pri_molal)
    put("total", sp.total)
    put("pri_act_coef", sp.pri_act_coef)
    put("sec_act_coef", sp.sec_act_coef)
    put("sec_molal", sp.sec_molal)
    put("ln_act_h2o", [sp.ln_act_h2o])
    put("immobile", sp.immobile)
    put("mnrl_volfrac", sp.mnrl_volfrac)
    put("mnrl_area", sp.mnrl_area)
    put("srfcplxrxn_free_site_conc", sp.free_site)
    put("eqsrfcplx_conc", sp.eqsrfcplx_conc)
    put("total_sorb_eq", sp.total_sorb_eq)
    put("kinmr_total_sorb", sp.kinmr_total_sorb)


_TYPE_CODES = {"T": _abi.CONSTRAINT_TOTAL, "TOTAL": _abi.CONSTRAINT_TOTAL, "F": _abi.CONSTRAINT_FREE,
               "FREE": _abi.CONSTRAINT_FREE, "L": _abi.CONSTRAINT_LOG, "LOG": _abi.CONSTRAINT_LOG,
               "P": _abi.CONSTRAINT_PH, "PH": _abi.CONSTRAINT_PH, "Z": _abi.CONSTRAINT_CHARGE_BAL,
               "CHG": _abi.CONSTRAINT_CHARGE_BAL, "M": _abi.CONSTRAINT_MINERAL, "MINERAL": _abi.CONSTRAINT_MINERAL,
               "MNRL": _abi.CONSTRAINT_MINERAL, "G": _abi.CONSTRAINT_GAS, "GAS": _abi.CONSTRAINT_GAS}


def to_abi(net: _chem.ReactionNetwork, cons: _chem.Constraint):
    """the CONSTRAINT block as ``pfrx_equilibrate_constraint`` takes it: (abi.Constraint with the types and the
    mineral / gas reactions in the network's basis, values[naqcomp] in the units of the block)"""
    naq = net.naqcomp
    by_name = {c[0]: c for c in cons.conc}
    types = np.zeros(naq, dtype=np.int32)
    vals = np.zeros(naq)
    logK = np.zeros(naq)
    coef = np.zeros((naq, 5))
    h2o = np.zeros(naq)
    ptr = [0]
    spec: List[int] = []
    st: List[float] = []
    for i, nm in enumerate(net.primary_names):
        if nm not in by_name:
            raise KeyError(f"constraint {cons.name}: no concentration for {nm}")
        _, v, t, aux = by_name[nm]
        code = _TYPE_CODES.get(t.upper())
        if code is None:
            raise ValueError(f"constraint type {t} not supported")
        types[i] = code
        vals[i] = v
        if code in (_abi.CONSTRAINT_MINERAL, _abi.CONSTRAINT_GAS):
            rx = net.mnrl_rxn[aux] if code == _abi.CONSTRAINT_MINERAL else net.gas_rxn[aux]
            logK[i] = net.logK_at_tref(rx.logK_T)
            coef[i] = _chem.fit_logK_coefs(net.db.temperatures, rx.logK_T)
            h2o[i] = rx.h2o_stoich
            spec += [int(j) for j in rx.ids]
            st += [float(x) for x in rx.stoich]
        ptr.append(len(spec))
    k = _abi.Constraint(naq, types, logK, h2o, ptr, spec, st, eq_logK_coef=coef,
                        initialize_with_molality=net.chem.initialize_with_molality)
    return k, vals
