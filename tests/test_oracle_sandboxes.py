"""CPU checks of the oracle's ELM-CN sandboxes beyond the golds.

NITRIFICATION and DENITRIFICATION have no regression gold in the reference
(ngee/CLMCNplus/TAI names them only in comments and uses MICROBIAL_REACTION), so
for NitrifReact / DenitrReact the statement is "parity unpinned by reference
tests; pinned by source restatement" -- plus the checks here: the analytic
Jacobian against finite differences (the reference's own perturbation_tolerance,
reaction.F90:41), N mass balance of the rates, and the whole workload through
RStep.  SOMDECOMP is pinned by 16 golds (test_oracle_golden.py); its Jacobian is
checked here only where the reference's is a true derivative.
"""
import numpy as np
import pytest

import oracle_lib as orc
from pflotran_elm_interface_b200 import abi, workloads as W


def _fd_jacobian(wl, cell, dt, rows, cols, pert=1.0e-6):
    """d Res[rows] / d c[cols] by central differences on the free concentrations
    (totals recomputed by the oracle's RTAuxVarCompute inside `reaction`)"""
    naq = wl.cfg.c.naqcomp
    out = np.zeros((len(rows), len(cols)))
    for b, j in enumerate(cols):
        res = []
        for sgn in (+1.0, -1.0):
            st = wl.state.copy()
            f, k = ("pri_molal", j) if j < naq else ("immobile", j - naq)
            st.a[f][k, cell] *= 1.0 + sgn * pert
            r, _ = orc.reaction(wl.cfg, st, cell, dt)
            res.append((r, st.a[f][k, cell]))
        out[:, b] = (res[0][0][rows] - res[1][0][rows]) / (res[0][1] - res[1][1])
    return out


@pytest.mark.parametrize("name", ["c4s", "c4se"])
def test_nitrif_denitr_jacobian_vs_finite_differences(name):
    """NitrifReact / DenitrReact differentiate their rates with respect to the bulk
    concentration c = total * theta * 1000 (mol/m^3) and multiply by dtotal, i.e. the
    reference's Jacobian entries are d(rate)/d(c) -- a factor theta*1000 short of
    d(rate)/d(molality) -- and the product rows use dtotal(product, reactant), which is
    zero without complexes (reaction_sandbox_nitrif.F90:330-337,
    reaction_sandbox_denitr.F90:352-362).  The restatement keeps both; this test pins the
    derivative formulas themselves: analytic * theta * 1000 == finite difference on the
    reactant's own row, and the N balance of the rates."""
    wl = W.by_name(name, ncell=40)
    net = wl.net
    nh4, no3 = net.primary_names.index("NH4+"), net.primary_names.index("NO3-")
    n2, n2o = net.primary_names.index("N2(aq)"), net.primary_names.index("N2O(aq)")
    # SOMDECOMP off for this check
    cfg = abi.ReactionConfig(net)
    cfg.c.somdec = None
    wl2 = W.Workload(wl.name, cfg, wl.state, wl.tran_dt, net)
    wl2.state.cfg = cfg
    checked = n2o_cells = 0
    for cell in range(40):
        st0 = wl2.state.copy()
        r, J = orc.reaction(cfg, st0, cell, wl.tran_dt)
        if np.abs(r).max() == 0.0:
            continue
        rows = np.array([nh4, no3, n2, n2o])
        fd = _fd_jacobian(wl2, cell, wl.tran_dt, rows, [nh4, no3])
        lw = wl.state["porosity"][0, cell] * wl.state["sat"][0, cell] * 1000.0
        for k, sp in ((0, nh4), (1, no3)):
            if fd[k, k] != 0.0:
                assert abs(J[sp, sp] * lw - fd[k, k]) <= 1.0e-4 * abs(fd[k, k]), (cell, sp, J[sp, sp] * lw, fd[k, k])
        # the true derivative of the product rows is minus (half) the reactant's
        assert abs(fd[1, 0] + 2.0 * fd[3, 0] + fd[0, 0]) <= 1e-3 * abs(fd[0, 0])
        assert abs(2.0 * fd[2, 1] + fd[1, 1]) <= 1e-3 * abs(fd[1, 1])
        assert J[no3, nh4] == 0.0 and J[n2, no3] == 0.0     # dtotal(product, reactant) = 0 here
        # N balance: NH4+ -> NO3- (1:1) or 1/2 N2O; NO3- -> 1/2 N2
        assert abs(r[nh4] + r[no3] + 2.0 * r[n2] + 2.0 * r[n2o]) <= 1e-12 * np.abs(r).max()
        n2o_cells += int(r[n2o] != 0.0)
        checked += 1
    assert checked >= 30 and n2o_cells >= 3


def test_somdec_conserves_carbon_and_nitrogen():
    """every SOMDECOMP reaction moves C from the upstream pool to downstream pools + CO2 and
    N from the upstream pool to downstream pools + mineral N (+ 1/2 N2O per N emitted)"""
    wl = W.by_name("c4s", ncell=60)
    net = wl.net
    cfg = abi.ReactionConfig(net)
    cfg.c.nitrif = None
    cfg.c.denitr = None
    st = wl.state
    st.cfg = cfg
    naq = net.naqcomp
    imm = {n: naq + i for i, n in enumerate(net.immobile_names)}
    pri = {n: i for i, n in enumerate(net.primary_names)}
    nc_som = {p: r for p, r in net.chem.somdec.pools if r is not None}
    for cell in range(60):
        r, _ = orc.reaction(cfg, st.copy(), cell, wl.tran_dt)
        # residual sign: + sink, - source
        dC = -(r[pri["CO2(aq)"]] + sum(r[imm[k]] for k in ("SOM1", "SOM2", "SOM3", "SOM4", "Lit1C", "Lit2C", "Lit3C")))
        dN = -(r[pri["NH4+"]] + r[pri["NO3-"]] + 2.0 * r[pri["N2O(aq)"]]
               + sum(r[imm[k]] for k in ("Lit1N", "Lit2N", "Lit3N"))
               + sum(nc_som[k] * r[imm[k]] for k in ("SOM1", "SOM2", "SOM3", "SOM4")))
        scale = np.abs(r).max()
        assert abs(dC) <= 1e-12 * scale and abs(dN) <= 1e-12 * scale, (cell, dC, dN, scale)


@pytest.mark.parametrize("name,dt", [("c4s", 1800.0), ("c4se", 86400.0)])
def test_rstep_on_elm_cn_workload(name, dt):
    wl = W.by_name(name, ncell=1500, tran_dt=dt)
    before = wl.state.copy()
    st = wl.state.copy()
    res = orc.rstep(wl.cfg, st, dt, 2)
    assert res.rstep_error == 0 and res.ncell_active == 1500
    assert np.all(st["total"] > 0.0) and np.all(st["immobile"] > 0.0)
    assert np.all(st["num_kinetic_state_updates"] >= 1)   # any sandbox forces the flag (reaction.F90:5965)
    # total N (aqueous per m^3 bulk + immobile) is conserved up to the N2O/N2 bookkeeping
    net = wl.net
    theta = before["porosity"] * before["sat"] * 1000.0
    nc_som = {p: r for p, r in net.chem.somdec.pools if r is not None}

    def total_n(s):
        t = 0.0
        for nm, w in (("NH4+", 1.0), ("NO3-", 1.0), ("N2O(aq)", 2.0), ("N2(aq)", 2.0)):
            t = t + w * s["total"][net.primary_names.index(nm)] * theta[0]
        for i, nm in enumerate(net.immobile_names):
            if nm.endswith("N") and nm.startswith("Lit"):
                t = t + s["immobile"][i]
            elif nm in nc_som:
                t = t + nc_som[nm] * s["immobile"][i]
        return t

    n0, n1 = total_n(before), total_n(st)
    assert np.all(np.abs(n1 - n0) <= 1e-5 * np.abs(n0))  # Newton stops at 1e-6 relative change


def test_sorption_jacobian_vs_finite_differences():
    """ion exchange (inner Newton and closed form, absolute and mineral-bound CEC), linear /
    Langmuir / Freundlich KD and dynamic KD: d(total_sorb_eq)/d(free) of the oracle against
    central differences of its own residual (reaction.F90:41 perturbation_tolerance scale)"""
    wl = W.by_name("c6", ncell=12)
    cfg, dt = wl.cfg, wl.tran_dt
    naq = cfg.c.naqcomp
    for cell in range(12):
        st0 = wl.state.copy()
        e, R0, J, _ = orc.girt_residual(cfg, st0, cell, dt)
        assert e == 0
        for j in range(naq):
            if wl.net.primary_names[j] == "SiO2(aq)":
                continue   # only in the Quartz rate law, whose Jacobian is the reference's approximate one
            cols = []
            for sgn in (1.0, -1.0):
                st = wl.state.copy()
                st.a["pri_molal"][j, cell] *= 1.0 + sgn * 1.0e-6
                _, R, _, _ = orc.girt_residual(cfg, st, cell, dt)
                cols.append((R, st.a["pri_molal"][j, cell]))
            fd = (cols[0][0] - cols[1][0]) / (cols[0][1] - cols[1][1])
            scale = np.abs(J[:, j]).max()
            assert np.abs(J[:, j] - fd).max() <= 2.0e-5 * scale, (cell, j, J[:, j], fd)
    # the exchanger is full: sum Z_i S_i = CEC for the absolute-CEC reaction
    st = wl.state.copy()
    orc.auxvar_compute(cfg, st, 0)
    z = cfg.arrays["primary_spec_Z"]
    ptr, cat = cfg.arrays["eqionx_ptr"], cfg.arrays["eqionx_cationid"]
    s0 = sum(z[cat[k]] * st["eqionx_conc"][k, 0] for k in range(ptr[0], ptr[1]))
    assert abs(s0 - 750.0) < 1e-9


@pytest.mark.parametrize("name", ["c7", "c7s"])
def test_general_decay_jacobian_vs_finite_differences(name):
    """RGeneral, RRadioactiveDecay (through dtotal of a network with a complex; c7s: of a sorbing
    parent, through dtotal_sorb_eq too) and RImmobileDecay: the oracle's analytic Jacobian against
    central differences of its own residual, all unknowns"""
    wl = W.by_name(name, ncell=8)
    cfg, dt = wl.cfg, wl.tran_dt
    naq, n = cfg.c.naqcomp, cfg.ncomp
    assert cfg.c.ngeneral_rxn == 2 and cfg.c.nradiodecay_rxn == 1 and cfg.c.nimmobile_decay_rxn == 1
    for cell in range(8):
        st0 = wl.state.copy()
        e, R0, J, _ = orc.girt_residual(cfg, st0, cell, dt)
        assert e == 0
        for j in range(n):
            fld, k = ("pri_molal", j) if j < naq else ("immobile", j - naq)
            cols = []
            for sgn in (1.0, -1.0):
                st = wl.state.copy()
                st.a[fld][k, cell] *= 1.0 + sgn * 1.0e-6
                _, R, _, _ = orc.girt_residual(cfg, st, cell, dt)
                cols.append((R, st.a[fld][k, cell]))
            fd = (cols[0][0] - cols[1][0]) / (cols[0][1] - cols[1][1])
            scale = np.abs(J[:, j]).max()
            assert np.abs(J[:, j] - fd).max() <= 2.0e-5 * scale, (cell, j, J[:, j], fd)


def test_active_gas_jacobian_vs_finite_differences():
    """RTotalGas (reaction_gas.F90:87-174) -- two gases, one over two components and water -- in the
    accumulation and in the decaying inventory: the oracle's analytic Jacobian against central differences
    of its own residual.  The RADON sandbox has no derivative and needs none."""
    wl = W.by_name("c7g", ncell=8)
    cfg, dt = wl.cfg, wl.tran_dt
    naq = cfg.c.naqcomp
    assert cfg.c.nactive_gas == 2 and cfg.c.nradiodecay_rxn == 1 and cfg.c.radon
    for cell in range(8):
        st0 = wl.state.copy()
        e, R0, J, _ = orc.girt_residual(cfg, st0, cell, dt)
        assert e == 0
        for j in range(naq):
            if wl.net.primary_names[j] == "SiO2(aq)":
                continue   # only in the Quartz rate law, whose Jacobian is the reference's approximate one
            cols = []
            for sgn in (1.0, -1.0):
                st = wl.state.copy()
                st.a["pri_molal"][j, cell] *= 1.0 + sgn * 1.0e-6
                _, R, _, _ = orc.girt_residual(cfg, st, cell, dt)
                cols.append((R, st.a["pri_molal"][j, cell]))
            fd = (cols[0][0] - cols[1][0]) / (cols[0][1] - cols[1][1])
            scale = np.abs(J[:, j]).max()
            assert np.abs(J[:, j] - fd).max() <= 2.0e-5 * scale, (cell, j, J[:, j], fd)
    # the gas share is not negligible: without the gas phase the Rn column is smaller
    sat_g = wl.state["sat_gas"][0, 0]
    assert sat_g > 0.05


def test_radon_secular_equilibrium_through_rstep():
    """the radon deck's chemistry through the operator-split RStep: after many half-lives the inventory
    (water + gas) equals generation / decay constant"""
    wl = W.by_name("c7g", ncell=16, tran_dt=3.8235 * 86400.0)
    st = wl.state.copy()
    for _ in range(60):
        res = orc.rstep(wl.cfg, st, wl.tran_dt, 2)
        assert res.rstep_error == 0
    i = wl.net.primary_names.index("Rn(aq)")
    lam = -np.log(0.5) / (3.8235 * 86400.0)
    inv = (st["total"][i] * st["sat"][0] + st["total_gas"][i] * st["sat_gas"][0]) * st["porosity"][0] * 1.0e3 * st["volume"][0]
    want = 1.1627850420873736e-19 * st["mnrl_volfrac"][0] * st["volume"][0] / lam
    # backward Euler with dt = one half-life converges to the same fixed point
    assert np.allclose(inv, want, rtol=1.0e-6), (inv, want)


@pytest.mark.parametrize("kind", ["CLMCN", "DLEM"])
def test_elm_flow_coupled_moisture_response(kind):
    """GetMoistureResponse of the ELM_PFLOTRAN build (elm_rspfuncs.F90:124-237): with a flow mode SOMDECOMP's f_w is
    the Clapp-Hornberger (CLMCN) or DLEM curve of the cell's soil properties.  f_w multiplies every SOMDECOMP rate, so
    the flow-coupled residual must equal the BGC-only one evaluated with w_scalar := the curve, restated here."""
    n = 64
    wl = W.by_name("c4sw" if kind == "CLMCN" else "c4sd", ncell=n)
    ref = W.by_name("c4se", ncell=n)
    assert wl.cfg.c.elm_flow_coupled == 1 and ref.cfg.c.elm_flow_coupled == 0
    st = wl.state
    theta = st["sat"][0] * st["porosity"][0]
    if kind == "CLMCN":
        g, minpsi = 9.8068, -10.0e6
        maxpsi = st["elm_sucsat"][0] * (-g)
        lsat = theta / np.minimum(1.0, 1.0 - np.minimum(0.9999, st["elm_bulkdensity_dry"][0] / 2.70e3))
        psi = np.minimum(st["elm_sucsat"][0] * (-g) * lsat ** (-st["elm_bsw"][0]), maxpsi)
        f = np.where(psi > minpsi, np.log(minpsi / psi) / np.log(minpsi / maxpsi), 0.0)
        f = np.where(psi > maxpsi - 100.0, f * 0.10, f)
    else:
        ts, tr = st["elm_effporosity"][0], st["elm_watfc"][0]
        se = (theta - tr) / (ts - tr)
        f = np.clip(float(np.float32(1.0)) - se * se * float(np.float32(0.368)) * np.exp(se), 0.0, 1.0)
        f = np.where(theta >= ts, 1.0, np.where(theta <= tr, 0.0, f))
    assert 0.0 <= f.min() and f.max() <= 1.0 and np.ptp(f) > 0.3      # the cells cover the curve
    # the same state in the BGC-only configuration with w_scalar = f
    for k in ref.state.a:
        if k in st.a and ref.state.a[k].shape == st.a[k].shape:
            ref.state.a[k][...] = st.a[k]
    ref.state.a["elm_w_scalar"][0] = f
    for c in range(n):
        r1, j1 = orc.reaction(wl.cfg, wl.state.copy(), c, wl.tran_dt)
        r0, j0 = orc.reaction(ref.cfg, ref.state.copy(), c, ref.tran_dt)
        assert np.abs(r1 - r0).max() <= 1.0e-13 * max(np.abs(r0).max(), 1e-300), (c, f[c], r1, r0)
        assert np.abs(j1 - j0).max() <= 1.0e-13 * max(np.abs(j0).max(), 1e-300)


def test_general_decay_rstep_mass_balance():
    """C7 through RStep: every cell converges without cuts, and the immobile species decays by the
    backward-Euler factor 1/(1 + k dt) of its half-life"""
    wl = W.by_name("c7", ncell=500)
    st = wl.state.copy()
    res = orc.rstep(wl.cfg, st, wl.tran_dt, 2)
    assert res.rstep_error == 0 and res.num_cut_cells == 0
    k = -np.log(0.5) / (10.0 * 86400.0)
    want = wl.state["immobile"][0] / (1.0 + k * wl.tran_dt)
    assert np.allclose(st["immobile"][0], want, rtol=1.0e-9)


def test_microbial_jacobian_vs_finite_differences():
    """RMicrobial: Monod terms with thresholds, the four inhibition types, activation energy,
    activities as concentrations -- the oracle's Jacobian against central differences of its own
    residual.  The biomass columns are left out: the reference's d(rate)/d(biomass) omits the
    L_water / volume factor (reaction_microbial.F90:578-583), restated as written and pinned by the
    seven ABCD_microbial golds through the Newton iteration counts."""
    wl = W.by_name("c8", ncell=10)
    cfg, dt = wl.cfg, wl.tran_dt
    naq, n = cfg.c.naqcomp, cfg.ncomp
    assert cfg.c.nmicrobial_rxn == 2 and set(cfg.arrays["microbial_inhibition_type"]) == {1, 3, 4, 5}
    bio = {wl.net.primary_names.index("D(aq)"), naq + wl.net.immobile_names.index("D(im)")}
    for cell in range(10):
        st0 = wl.state.copy()
        e, R0, J, _ = orc.girt_residual(cfg, st0, cell, dt)
        assert e == 0
        for j in range(n):
            if j in bio:
                continue
            fld, k = ("pri_molal", j) if j < naq else ("immobile", j - naq)
            cols = []
            for sgn in (1.0, -1.0):
                st = wl.state.copy()
                st.a[fld][k, cell] *= 1.0 + sgn * 1.0e-6
                _, R, _, _ = orc.girt_residual(cfg, st, cell, dt)
                cols.append((R, st.a[fld][k, cell]))
            fd = (cols[0][0] - cols[1][0]) / (cols[0][1] - cols[1][1])
            scale = np.abs(J[:, j]).max()
            assert np.abs(J[:, j] - fd).max() <= 5.0e-5 * scale, (cell, j, J[:, j], fd)


def test_cndegas_solubilities_and_rates():
    """CNDEGAS has no gold in the reference's test suite (no deck under regression_tests names the
    sandbox): parity unpinned by reference tests, pinned by source restatement -- plus the linear form
    of the rate, rate = k (c - c_eq) L_water, and a plausibility band for the solubility fits (the Henry
    constants of CO2 and N2O in fresh water at 25 C are 0.034 and 0.024-0.025 mol/(L atm) to two digits)."""
    wl = W.by_name("c4g", ncell=4)
    a = wl.state.a
    a["temp"][...] = 25.0
    net = wl.net
    ico2, in2o, in2 = (net.primary_names.index(n) for n in ("CO2(aq)", "N2O(aq)", "N2(aq)"))
    cd = wl.cfg.cndegas
    # switch the other sandboxes' contributions to these rows off by evaluating twice: only the aqueous
    # concentration of the gas changes, so the difference of the residuals is k * L_water * dtotal * dc
    cell, dt = 0, wl.tran_dt
    r0, j0 = orc.reaction(wl.cfg, wl.state.copy(), cell, dt)
    lw = a["volume"][0, cell] * 1000.0 * a["porosity"][0, cell] * 0.5      # stand-alone build: liquid saturation 0.5
    for idx, k in ((ico2, cd.k_kinetic_co2), (in2o, cd.k_kinetic_n2o), (in2, cd.k_kinetic_n2)):
        st = wl.state.copy()
        st.a["pri_molal"][idx, cell] *= 2.0
        r1, _ = orc.reaction(wl.cfg, st, cell, dt)
        dc = wl.state.a["pri_molal"][idx, cell] * a["den_kg"][0, cell] * 1e-3
        gas_row = net.naqcomp + (cd.co2g_id if idx == ico2 else cd.n2og_id if idx == in2o else cd.n2g_id)
        assert r1[gas_row] - r0[gas_row] == pytest.approx(-k * lw * dc, rel=1e-9)
        assert j0[gas_row, idx] == pytest.approx(-k * lw, rel=1e-12)      # d/d(total), as written (:335)
    # equilibrium concentrations: a cell whose dissolved gas sits exactly at c_eq has no exchange
    # -> solve rate = 0 for c by two evaluations (the rate is linear in c)
    def c_eq(idx, gas_row):
        out = []
        for f in (1.0, 2.0):
            st = wl.state.copy()
            st.a["pri_molal"][idx, cell] = f * 1.0e-6
            st.a["total"][idx, cell] = f * 1.0e-6
            r, _ = orc.reaction(wl.cfg, st, cell, dt)
            out.append(r[gas_row])
        # r = -k lw (c - ceq) + other(c-independent): slope known
        slope = (out[1] - out[0]) / 1.0e-6
        return slope
    # slopes equal -k lw dtotal: consistency of the linear form
    assert c_eq(ico2, net.naqcomp + cd.co2g_id) == pytest.approx(-cd.k_kinetic_co2 * lw, rel=1e-6)
    # the coefficients of the fits as the reference has them (:568-573, :652-657)
    tk = 298.15
    k0_co2 = np.exp(-58.0931 + 90.5069 * (100.0 / tk) + 22.2940 * np.log(tk / 100.0))
    k0_n2o = np.exp(-62.7076 + 97.3066 * (100.0 / tk) + 24.1406 * np.log(tk / 100.0))
    assert k0_co2 == pytest.approx(3.4e-2, rel=2e-2) and k0_n2o == pytest.approx(2.45e-2, rel=3e-2)


@pytest.mark.parametrize("name", ["c4g", "c4ge"])
def test_cndegas_equilibrium_is_a_fixed_point(name):
    """mass balance of the exchange (what leaves the water enters the reservoir row, mol/s) and the sign
    of the rate on either side of the solubility"""
    wl = W.by_name(name, ncell=16)
    net, cd = wl.net, wl.cfg.cndegas
    for cell in range(16):
        if wl.state.a["sat"][0, cell] < 0.05:
            continue
        r, j = orc.reaction(wl.cfg, wl.state.copy(), cell, wl.tran_dt)
        # N2: no other sandbox touches the N2 reservoir row, and N2(aq) only gains from denitrification
        row_g = net.naqcomp + cd.n2g_id
        st = wl.state.copy()
        i = net.primary_names.index("N2(aq)")
        st.a["pri_molal"][i, cell] = 1.0e-2   # far above any atmospheric solubility: degassing
        st.a["total"][i, cell] = 1.0e-2
        r_hi, _ = orc.reaction(wl.cfg, st, cell, wl.tran_dt)
        assert r_hi[row_g] < 0.0 < r_hi[i] + 1e-30 or r_hi[row_g] < 0.0   # residual = +sink / -source
        st.a["pri_molal"][i, cell] = 1.0e-12
        st.a["total"][i, cell] = 1.0e-12
        r_lo, _ = orc.reaction(wl.cfg, st, cell, wl.tran_dt)
        assert r_lo[row_g] > 0.0                                          # dissolving: the reservoir loses
        # pH-stat rows are equal and opposite
        ip, ih = cd.proton_id, net.naqcomp + cd.himm_id
        st = wl.state.copy()
        r2, _ = orc.reaction(wl.cfg, st, cell, wl.tran_dt)
        assert r2[ih] != 0.0


def test_calcite_sandbox_is_the_mineral_rate_law_it_splits():
    """CalciteEvaluate (reaction_sandbox_calcite.F90:177-365): the deck of the reference's regression test
    (default/reaction_sandbox/reaction_sandbox_calcite.in) sets RATE_CONSTANT1 = RATE_CONSTANT2 = 5e-7 "so that
    the parallel rate pathways sum to 1.d-6", the rate constant of the mineral in the calcite decks whose golds
    pin RKineticMineral (test_oracle_golden.py::test_calcite_kinetics_gold).  With pathway 2's hard-wired
    pKeq 1.8487 equal to the database's, a step through the sandbox must land where the step through
    MINERAL_KINETICS lands -- concentrations, volume fractions (CalciteUpdateKineticState vs
    MineralUpdateKineticState) and iteration counts -- and the rate it leaves in rt_auxvar%auxiliary_data
    must be the mineral rate of the other run."""
    n = 1500
    for dt in (3600.0, 30 * 86400.0):
        sb, mk = W.by_name("c2sb", ncell=n, tran_dt=dt), W.by_name("c2", ncell=n, tran_dt=dt)
        assert sb.cfg.c.calcite and sb.cfg.arrays["kinmnrl_rate_constant"][0] == 0.0
        mk.state.a["mnrl_volfrac"][...] = sb.state.a["mnrl_volfrac"]      # every fifth cell has no mineral
        ra, rb = orc.rstep(sb.cfg, sb.state, dt, 4), orc.rstep(mk.cfg, mk.state, dt, 4)
        assert ra.as_dict() == rb.as_dict()
        for f in ("pri_molal", "total", "sec_molal"):
            np.testing.assert_allclose(sb.state.a[f], mk.state.a[f], rtol=1e-12, atol=1e-300, err_msg=f)
        # (a volume fraction that has all but dissolved is a difference of nearly equal numbers)
        np.testing.assert_allclose(sb.state.a["mnrl_volfrac"], mk.state.a["mnrl_volfrac"], rtol=1e-12, atol=1e-18)
        np.testing.assert_allclose(sb.state.a["sandbox_aux"][0], mk.state.a["mnrl_rate"][0], rtol=1e-9, atol=1e-22)
        assert np.abs(sb.state.a["mnrl_rate"]).max() == 0.0 and np.abs(sb.state.a["sandbox_aux"]).max() > 0.0


def test_calcite_sandbox_jacobian_vs_finite_differences():
    wl = W.by_name("c2sb", ncell=16)
    cfg, dt = wl.cfg, wl.tran_dt
    n = cfg.ncomp
    for cell in range(16):
        r0, J = orc.reaction(cfg, wl.state.copy(), cell, dt)
        # the reference's derivative carries the factor kg water / L water (dQK_dmj * molality_to_molarity,
        # reaction_sandbox_calcite.F90:262-264, as RKineticMineral's does): d Res / d molality times den_kg * 1e-3
        fd = _fd_jacobian(wl, cell, dt, np.arange(n), np.arange(n)) * (wl.state.a["den_kg"][0, cell] * 1.0e-3)
        scale = np.abs(J).max()
        if scale == 0.0:       # no mineral and undersaturated: the sandbox is switched off
            assert np.abs(fd).max() == 0.0
            continue
        assert np.abs(J - fd).max() <= 2.0e-5 * scale, (cell, J, fd)
