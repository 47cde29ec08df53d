"""Single-cell GIRT batch driver used to replay the reference's regression
golds with the oracle's residual/Jacobian functions.

Per step (pm_rt.F90, reactive_transport.F90):
  PMRTInitializeTimestep :508   act. coefs if TIMESTEP frequency
  RTUpdateFixedAccumulation :931 fixed = A(c^k)
  SNES Newton (basic line search, PMRTCheckUpdatePre :857 clamps) on
      r = (A(c) - fixed)/dt + R(c)
  PMRTUpdateSolution2 :1177     RTUpdateEquilibriumState + RTUpdateKineticState
  PMRTUpdateTimestep :718       dt growth by tfac(newton its)
"""
import numpy as np

import oracle_lib as orc
from pflotran_elm_interface_b200 import chem

TFAC = [2.0, 2.0, 2.0, 2.0, 2.0, 1.8, 1.6, 1.4, 1.2, 1.0, 1.0, 1.0, 1.0]


class GirtBatch:
    def __init__(self, cfg, state, deck, use_log=None, max_dlnC=5.0):
        self.cfg, self.state, self.deck = cfg, state, deck
        self.use_log = bool(cfg.c.use_log_formulation) if use_log is None else use_log
        self.max_dlnC = max_dlnC
        self.rtol = deck.newton.get("RTOL", 1.0e-8)
        self.atol = deck.newton.get("ATOL", 1.0e-50)
        self.stol = deck.newton.get("STOL", 1.0e-8)
        self.maxit = int(deck.newton.get("MAXIMUM_NUMBER_OF_ITERATIONS", deck.newton.get("MAXIT", 50)))
        self.divtol = 1.0e4
        self.steps = 0
        self.cuts = 0
        self.last_dt = 0.0
        self.newton_its = 0
        self.time = 0.0
        self.naq = cfg.c.naqcomp
        self.nim = cfg.c.nimcomp

    # unknowns <-> state
    def _get_c(self):
        a = self.state.a
        return np.concatenate([a["pri_molal"][:, 0], a["immobile"][:, 0]])

    def _set_c(self, c):
        a = self.state.a
        a["pri_molal"][:, 0] = c[: self.naq]
        if self.nim:
            a["immobile"][:, 0] = c[self.naq:]

    def _F(self, dt, fixed):
        e, Res, Jac, acc = orc.girt_residual(self.cfg, self.state, 0, dt)
        assert e == 0
        F = Res - fixed / dt
        if getattr(self.deck, "numerical_jacobian", False):
            # NUMERICAL_JACOBIAN in the deck (SNES finite differences): forward differences
            # of the same residual; only the Newton path depends on it, not the converged state
            c0 = self._get_c().copy()
            Jac = np.zeros((len(c0), len(c0)))
            for j in range(len(c0)):
                h = 1.0e-8 * max(abs(c0[j]), 1.0e-6)   # PETSc MatFD "ds": error_rel * max(|x|, umin)
                cj = c0.copy()
                cj[j] += h
                self._set_c(cj)
                _, Rj, _, _ = orc.girt_residual(self.cfg, self.state, 0, dt)
                Jac[:, j] = ((Rj - fixed / dt) - F) / h
            self._set_c(c0)
            orc.girt_residual(self.cfg, self.state, 0, dt)
        return F, Jac

    def _newton(self, dt, fixed):
        """one SNESSolve: (iterations, converged)"""
        cfg = self.cfg
        c = self._get_c()
        F, J = self._F(dt, fixed)
        fnorm0 = np.linalg.norm(F)
        its = 0
        if fnorm0 < self.atol:
            return 0, True
        while its < self.maxit:
            Jm = np.array(J)
            if self.use_log:
                Jm = Jm * c[None, :]
            try:
                dx = np.linalg.solve(Jm, F)
            except np.linalg.LinAlgError:
                return its, False
            its += 1
            if self.use_log:
                dx = np.sign(dx) * np.minimum(np.abs(dx), self.max_dlnC)
                x = np.log(c)
                x_new = x - dx
                c_new = np.exp(x_new)
                xnorm, snorm = np.linalg.norm(x_new), np.linalg.norm(dx)
            else:
                mask = c <= dx
                if np.any(mask):
                    mr = np.min(np.abs(c[mask] / dx[mask]))
                    if mr < 1.0:
                        dx = dx * mr * 0.99
                c_new = c - dx
                xnorm, snorm = np.linalg.norm(c_new), np.linalg.norm(dx)
            c = c_new
            self._set_c(c)
            F, J = self._F(dt, fixed)
            fnorm = np.linalg.norm(F)
            if not np.isfinite(fnorm):
                return its, False          # SNES_DIVERGED_FNORM_NAN
            if fnorm < self.atol or fnorm <= self.rtol * fnorm0 or snorm < self.stol * xnorm:
                return its, True
            if fnorm > self.divtol * fnorm0:
                return its, False          # SNES_DIVERGED_DTOL (SNESConvergedDefault, divtol = 1e4)
        return its, False                  # SNES_DIVERGED_MAX_IT

    def step(self, dt):
        """TimestepperSNESStepDT (timestepper_SNES.F90:365-450): on a failed solve the
        solution is put back (RTTimeCut, reactive_transport.F90:51-83) and dt is halved
        (TimestepperBaseCutDT, timestepper_base.F90:707-791).  Returns the iterations of
        the successful solve and the dt it used; wasted iterations count in newton_its."""
        cfg = self.cfg
        # fixed accumulation at time level k -- with the OLD activity
        # coefficients (reactive_transport.F90:1012-1014), then the update
        _, _, _, fixed = orc.girt_residual(cfg, self.state, 0, dt)
        if cfg.c.act_coef_update_frequency == chem.ACT_COEF_FREQUENCY_TIMESTEP:
            orc.activity(cfg, self.state, 0)
        c_old = self._get_c().copy()
        while True:
            its, ok = self._newton(dt, fixed)
            self.newton_its += its
            if ok:
                break
            self.cuts += 1
            if self.cuts > 10000:
                raise RuntimeError("time step cut criteria exceeded")
            dt = 0.5 * dt
            self._set_c(c_old.copy())
        # RTUpdateEquilibriumState: totals at the converged free-ion values
        orc.auxvar_compute(cfg, self.state, 0)
        orc.update_kinetic_state(cfg, self.state, 0, dt)
        self.steps += 1
        self.time += dt
        self.last_dt = dt
        return its

    def run(self):
        dk = self.deck
        dt = dk.initial_dt
        if dk.max_steps is not None and dk.max_steps < 0:
            # MAX_STEPS -1: the run stops after PMRTInitializeRun (pm_rt.F90:470-475)
            orc.auxvar_compute(self.cfg, self.state, 0)
            return self
        tol = 1.0e-10
        while self.time < dk.final_time * (1.0 - 1.0e-14):
            if self.time + dt * (1.0 + tol) >= dk.final_time:
                dt_step = dk.final_time - self.time
            else:
                dt_step = dt
            its = self.step(dt_step)
            dt_step = self.last_dt          # smaller than requested after a cut
            # PMRTUpdateTimestep "original implementation", pm_rt.F90:760-775
            if dk.ts_acceleration == 0:
                dt = dt_step              # iacceleration == 0: the step size is left alone (:734)
                continue
            if its <= dk.ts_acceleration:
                # its == 0 (residual below ATOL at once): the reference indexes tfac(0); growth
                # like a one-iteration step is the benign reading
                fac = TFAC[max(its, 1) - 1] if its <= len(TFAC) else 0.5
            else:
                fac = 0.5
            dt = min(min(2.0 * dt_step, fac * dt_step), dk.maximum_dt)
            dt = max(dt, dk.minimum_dt)   # pm_rt.F90:780
        return self


def read_gold(path):
    """.regression.gold -> {section title: {key: value}}"""
    out = {}
    cur = None
    with open(path) as f:
        for ln in f:
            ln = ln.rstrip("\n")
            if ln.startswith("--"):
                cur = ln.strip()
                if cur.startswith("-- "):
                    cur = cur[3:]
                if cur.endswith(" --"):
                    cur = cur[:-3]
                out[cur] = {}
            elif ":" in ln and cur is not None:
                k, v = ln.split(":", 1)
                try:
                    out[cur][k.strip()] = float(v)
                except ValueError:
                    out[cur][k.strip()] = v.strip()
    return out
