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
        self.steps = 0
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
        return Res - fixed / dt, Jac

    def step(self, dt):
        cfg = self.cfg
        # fixed accumulation at time level k -- with the OLD activity
        # coefficients (reactive_transport.F90:1012-1014), then the update
        _, _, _, fixed = orc.girt_residual(cfg, self.state, 0, dt)
        if cfg.c.act_coef_update_frequency == chem.ACT_COEF_FREQUENCY_TIMESTEP:
            orc.activity(cfg, self.state, 0)
        c = self._get_c()
        F, J = self._F(dt, fixed)
        fnorm0 = np.linalg.norm(F)
        its = 0
        if not fnorm0 < self.atol:
            while its < self.maxit:
                if cfg.c.act_coef_update_frequency == chem.ACT_COEF_FREQUENCY_NEWTON_ITER and its > 0:
                    pass  # GIRT updates act. coefs in RTUpdateAuxVars; handled by caller decks we replay
                Jm = np.array(J)
                if self.use_log:
                    Jm = Jm * c[None, :]
                dx = np.linalg.solve(Jm, F)
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
                if fnorm < self.atol or fnorm <= self.rtol * fnorm0 or snorm < self.stol * xnorm:
                    break
        # RTUpdateEquilibriumState: totals at the converged free-ion values
        orc.auxvar_compute(cfg, self.state, 0)
        orc.update_kinetic_state(cfg, self.state, 0, dt)
        self.steps += 1
        self.newton_its += its
        self.time += dt
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
            # PMRTUpdateTimestep "original implementation", pm_rt.F90:760-775
            if its <= dk.ts_acceleration:
                fac = TFAC[its - 1] if 1 <= its <= len(TFAC) else 0.5
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
