"""Checks of the oracle's RStep/RReact restatement that the reference's own
tests cannot provide: LU known-answer tests,
analytic vs finite-difference Jacobians, a tracer through RReact, sub-stepping
properties, and the REF_BUG_COMPAT self-check."""
import numpy as np
import pytest

import oracle_lib as orc
from pflotran_elm_interface_b200 import abi, workloads as W


# ---- LU (utility.F90:597-735): unpinned by any reference test -> our own KATs ---
@pytest.mark.parametrize("n", [1, 2, 3, 5, 13, 15, 16])
def test_lu_against_numpy(n):
    rng = np.random.default_rng(n)
    for _ in range(20):
        A = rng.standard_normal((n, n)) + n * np.eye(n)
        b = rng.standard_normal(n)
        e, x = orc.lu_solve(A, b)
        assert e == 0
        np.testing.assert_allclose(x, np.linalg.solve(A, b), rtol=1e-11, atol=1e-13)


def test_lu_pivoting_and_singular_rows():
    # needs a row swap: zero on the diagonal
    A = np.array([[0.0, 2.0, 1.0], [1.0, 0.0, 3.0], [4.0, 1.0, 0.0]])
    b = np.array([1.0, 2.0, 3.0])
    e, x = orc.lu_solve(A, b)
    assert e == 0
    np.testing.assert_allclose(A @ x, b, atol=1e-13)
    # implicit scaling: a badly scaled row must not steal the pivot
    A = np.array([[1e-10, 1.0], [1.0, 1.0]])
    e, x = orc.lu_solve(A, np.array([1.0, 2.0]))
    np.testing.assert_allclose(x, np.linalg.solve(A, [1.0, 2.0]), rtol=1e-12)
    # an all-zero row is reported, not divided by (stop_on_error = false)
    e, _ = orc.lu_solve(np.array([[1.0, 2.0], [0.0, 0.0]]), np.array([1.0, 1.0]))
    assert e == 1
    # zero pivot after elimination is replaced by tiny = 1e-20 (no crash)
    e, x = orc.lu_solve(np.array([[1.0, 1.0], [1.0, 1.0]]), np.array([1.0, 1.0]))
    assert e == 0 and np.all(np.isfinite(x))


def test_rsolve_scaling_and_log_formulation():
    rng = np.random.default_rng(7)
    n = 6
    J = rng.standard_normal((n, n)) * 10.0 ** rng.integers(-3, 6, (n, 1)) + np.diag(10.0 ** rng.integers(0, 6, n))
    r = rng.standard_normal(n)
    c = 10.0 ** rng.uniform(-9, -2, n)
    e, u = orc.rsolve(r, J, c, 0)
    np.testing.assert_allclose(u, np.linalg.solve(J, r), rtol=1e-9)
    e, u = orc.rsolve(r, J, c, 1)  # unknown = ln c: columns scaled by c
    np.testing.assert_allclose(u, np.linalg.solve(J * c[None, :], r), rtol=1e-9)


# ---- analytic Jacobian vs finite differences (perturbation 1e-5..1e-7) -------------
@pytest.mark.parametrize("name,n", [("c2", 16), ("c2pfp", 16), ("c4", 16), ("c3", 6), ("c3mr", 4), ("c5", 4)])
def test_jacobian_matches_finite_differences(name, n):
    wl = W.by_name(name, ncell=n)
    cfg, dt = wl.cfg, wl.tran_dt
    naq, nim = cfg.c.naqcomp, cfg.c.nimcomp
    for ic in range(min(n, 3)):
        st = wl.state.copy()
        if cfg.c.act_coef_update_frequency == 2:
            orc.activity(cfg, st, ic)
        e, R0, J, _ = orc.girt_residual(cfg, st, ic, dt)
        assert e == 0
        x0 = np.concatenate([st["pri_molal"][:, ic], st["immobile"][:, ic]])
        for j in range(naq + nim):
            Rpm = []
            h = 1.0e-4 * x0[j]
            for sgn in (+1.0, -1.0):
                st2 = wl.state.copy()
                for k in ("pri_act_coef", "sec_act_coef", "srfcplxrxn_free_site_conc", "kinmr_total_sorb"):
                    st2.a[k][...] = st.a[k]
                if j < naq:
                    st2["pri_molal"][j, ic] += sgn * h
                else:
                    st2["immobile"][j - naq, ic] += sgn * h
                Rpm.append(orc.girt_residual(cfg, st2, ic, dt)[1])
            fd = (Rpm[0] - Rpm[1]) / (2.0 * h)
            # compare in the scaled form RSolve uses: column j times c_j, row by its max
            Js = J * x0[None, :]
            rown = np.maximum(np.abs(Js).max(axis=1), 1e-300)
            err = np.abs(fd * x0[j] - Js[:, j]) / rown
            # CLM-CN's "revision to avoid division by 0" Jacobian entries are
            # deliberately approximate (reaction_sandbox_clm_cn.F90:737-769)
            # RKineticMineral multiplies dQK/dm_j by den_kg*1e-3 (= 0.997,
            # reaction_mineral.F90:985-987): its analytic mineral term is 0.3 % off by
            # construction, i.e. ~1e-3 of a row at most for these decks
            tol = 2e-2 if name == "c4" else (1e-3 if wl.net.nkinmnrl else 1e-6)
            assert err.max() < tol, (name, ic, j, err.max())


# ---- RReact as a whole ----------------------------------------------------------------
def test_tracer_through_rreact_returns_transported_total():
    """a non-reacting component leaves RStep with
    total == transported total and free = total / density"""
    wl = W.by_name("c3", ncell=8)
    st = wl.state.copy()
    it = wl.net.primary_names.index("Tracer")
    st["total"][it, :] = np.linspace(1e-7, 1e-3, 8)
    before = st["total"][it].copy()
    res = orc.rstep(wl.cfg, st, wl.tran_dt)
    assert res.rstep_error == 0
    np.testing.assert_allclose(st["total"][it], before, rtol=1e-6)  # Newton stops at 1e-6 relative change
    np.testing.assert_allclose(st["pri_molal"][it] * st["den_kg"][0] / 1000.0, before, rtol=1e-6)


def test_ref_bug_compat_reproduces_the_noop():
    """with the fork's RSolve (no back-substitution) RReact leaves the free-ion
    guess untouched and throws the transported totals away"""
    wl = W.by_name("c2", ncell=64)
    st = wl.state.copy()
    guess = st["pri_molal"].copy()
    orc.lib().pfrx_oracle_set_ref_bug_compat(1)
    try:
        res = orc.rstep(wl.cfg, st, wl.tran_dt)
    finally:
        orc.lib().pfrx_oracle_set_ref_bug_compat(0)
    assert res.max_newton_iterations == 1
    assert np.array_equal(st["pri_molal"], guess)
    # the fixed oracle moves the solution
    st2 = wl.state.copy()
    orc.rstep(wl.cfg, st2, wl.tran_dt)
    assert not np.allclose(st2["pri_molal"], guess, rtol=1e-3)


def test_mass_balance_and_substep_consistency():
    wl = W.by_name("c2", ncell=256, tran_dt=86400.0)
    a = wl.state.copy()
    orc.rstep(wl.cfg, a, wl.tran_dt)
    d = a["total"] - wl.state["total"]
    scale = np.abs(wl.state["total"]).max()
    assert np.abs(d[2] - d[1]).max() < 1e-5 * scale    # CaCO3 + H+ = Ca++ + HCO3-
    assert np.abs(d[2] + d[0]).max() < 1e-5 * scale
    # forcing cuts: more, smaller backward-Euler sub-steps; stoichiometric mass
    # balance must hold on that path too
    wl2 = W.by_name("c2", ncell=256, tran_dt=86400.0)
    wl2.cfg.c.maximum_reaction_iterations = 9
    b = wl2.state.copy()
    res = orc.rstep(wl2.cfg, b, wl2.tran_dt)
    assert res.num_cut_cells > 0 and res.rstep_error == 0 and res.max_sub_steps > 1
    d = b["total"] - wl2.state["total"]
    assert np.abs(d[2] - d[1]).max() < 1e-5 * scale
    assert np.abs(d[2] + d[0]).max() < 1e-5 * scale


def test_too_many_cuts_sets_ierror_and_keeps_going():
    wl = W.by_name("c2", ncell=128)
    wl.cfg.c.maximum_reaction_iterations = 1
    wl.cfg.c.maximum_reaction_cuts = 2
    st = wl.state.copy()
    res = orc.rstep(wl.cfg, st, wl.tran_dt)
    assert res.rstep_error == 1
    assert res.first_failed_cell == int(np.flatnonzero(st["ierror"][0])[0])
    assert res.ncell_active == 128            # every cell is attempted


def test_threads_do_not_change_results():
    wl = W.by_name("c4", ncell=999)
    a, b = wl.state.copy(), wl.state.copy()
    ra = orc.rstep(wl.cfg, a, wl.tran_dt, 1)
    rb = orc.rstep(wl.cfg, b, wl.tran_dt, 5)
    assert ra.as_dict() == rb.as_dict()
    for k in a.a:
        assert np.array_equal(a.a[k], b.a[k])


@pytest.mark.parametrize("name", ["c2", "c3", "c4", "c4s"])
def test_op_counting_build_is_the_same_oracle(name):
    """oracle/pfrx_oracle_count.cpp compiles pfrx_oracle.c with a counting scalar in place of double: the
    results are bit-identical to the plain build, and the count per Newton iteration brackets the closed
    form of SURVEY 8(d) that round 1 used as the roofline numerator (the closed form misses the
    convergence tests, the update and part of the sorption arithmetic: 9-42 % low)"""
    from pflotran_elm_interface_b200 import workloads as W

    wl = W.by_name(name, ncell=256)
    a, b = wl.state.copy(), wl.state.copy()
    r0 = orc.rstep(wl.cfg, a, wl.tran_dt, 2)
    r1, ops = orc.count_ops(wl.cfg, b, wl.tran_dt, 2)
    assert r0.as_dict() == r1.as_dict()
    for f in a.a:
        assert np.array_equal(a.a[f], b.a[f], equal_nan=True), f
    f_eval, f_solve = W.flops_model(wl.net)
    its = r0.sum_newton_iterations
    closed = its * f_eval + max(0, its - r0.ncell_active) * f_solve
    assert 1.0 < ops / closed < 1.6, ops / closed
