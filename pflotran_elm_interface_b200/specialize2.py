"""Code generator for the second form of the network-specialised kernels
(``csrc/pfrx_spec2.cuh``): secondary molalities as products of mantissa powers,
the Jacobian assembled symmetric in ln-space.

Applies to networks in the LOG formulation with aqueous complexes of integer
stoichiometry, kinetic minerals (TST) and equilibrium surface complexation with
unit free-site stoichiometry -- the Hanford / calcite class (BASELINE configs C2,
C3, C5).  Everything else keeps form 1 (``specialize._Gen``).

Routines restated (all citations into /root/reference/src/pflotran):

* RActivityCoefficients, LAG branch           reaction.F90:4553-4612
* RTotalAqueous + RTAccumulation[Derivative]   reaction.F90:4665-4759, 5710-5848
* RTotalSorbEqSurfCplx1 (closed form)          reaction_surf_complex.F90:641-900
* RAccumulationSorb[Derivative]                reaction.F90:5144-5207
* RKineticMineral (TST, no prefactors)         reaction_mineral.F90:647-1078
"""
from __future__ import annotations

import math
import os
import re
from typing import Dict, List, Optional, Tuple

import numpy as np

from . import abi, chem

LOG_TO_LN = chem.LOG_TO_LN


def _lit(x: float) -> str:
    x = float(x)
    if x == 0.0:
        return "0.0"
    if x == int(x) and abs(x) < 1e6:
        return f"{x:.1f}"
    return float.hex(x)


def _is_int(v: float, lim: int = 8) -> bool:
    return float(v) == int(v) and abs(int(v)) <= lim


def supported2(cfg: abi.ReactionConfig) -> Tuple[bool, str]:
    """can the network use form 2?  (specialize.supported() must hold as well)"""
    c, a = cfg.c, cfg.arrays
    if os.environ.get("PFRX_SPEC_FORM", "2") == "1":
        return False, "PFRX_SPEC_FORM=1"
    if not c.use_log_formulation:
        return False, "linear formulation"
    if c.neqcplx <= 0:
        return False, "no aqueous complexes"
    if c.nkinmrsrfcplxrxn > 0:
        return False, "multirate sorption"
    if c.clmcn_nrxn > 0 or c.somdec or c.nitrif or c.denitr or c.plantn or c.langmuir:
        return False, "reaction sandbox"
    if c.ngeneral_rxn > 0 or c.nradiodecay_rxn > 0 or c.nimmobile_decay_rxn > 0 or c.nmicrobial_rxn > 0:
        return False, "general / decay / microbial reactions (non-symmetric Jacobian)"
    for name in ("eqcplx_stoich", "eqcplx_h2ostoich", "kinmnrl_stoich", "kinmnrl_h2ostoich", "srfcplx_stoich",
                 "srfcplx_h2ostoich"):
        if name in a and not all(_is_int(v) for v in a[name]):
            return False, "non-integer stoichiometry in " + name
    if c.act_coef_update_frequency not in (chem.ACT_COEF_FREQUENCY_NEWTON_ITER, chem.ACT_COEF_FREQUENCY_TIMESTEP,
                                           chem.ACT_COEF_FREQUENCY_OFF):
        return False, "activity update frequency"
    if c.naqcomp + c.nimcomp > 24:
        return False, "more than 24 unknowns"
    return True, ""


def _split(v: float) -> Tuple[float, int]:
    """v = m * 2^e with m in [1, 2)"""
    m, e = math.frexp(v)
    return m * 2.0, e - 1


def _kinv(logk: float) -> Tuple[float, int]:
    """mantissa and exponent of exp(-logK * LOG_TO_LN) (the reference's truncated constant)"""
    lnv = -float(logk) * LOG_TO_LN
    if abs(lnv) < 700.0:
        return _split(math.exp(lnv))
    e = math.floor(lnv / math.log(2.0))
    return _split(math.exp(lnv - e * math.log(2.0)))[0], int(e)


class _Gen2:
    def __init__(self, cfg: abi.ReactionConfig, threads_sync: bool = True, refill: bool = False,
                 solver: Optional[str] = None):
        from . import specialize as sp1

        ok, why = sp1.supported(cfg)
        if ok:
            ok, why = supported2(cfg)
        if not ok:
            raise ValueError("network not supported by form 2 of the specialiser: " + why)
        self.cfg, self.c, self.a = cfg, cfg.c, cfg.arrays
        self.naq = cfg.c.naqcomp
        self.n = cfg.c.naqcomp + cfg.c.nimcomp
        self.ncx = cfg.c.neqcplx
        self.sync, self.refill = threads_sync, refill
        self.solver = solver or os.environ.get("PFRX_SPEC2_SOLVER", "sym")
        assert self.solver in ("sym", "lu")
        self.sym = self.solver == "sym"
        self.act_upd = cfg.c.act_coef_update_frequency == chem.ACT_COEF_FREQUENCY_NEWTON_ITER
        a = self.a
        self.cls: List[Tuple[float, float]] = []
        self.pri_cls = [self._class_of(z, a0) for z, a0 in zip(a["primary_spec_Z"], a["primary_spec_a0"])]
        self.cx_cls = [self._class_of(z, a0) for z, a0 in zip(a["eqcplx_Z"], a["eqcplx_a0"])]
        used = set()
        for ids in ("eqcplx_specid", "kinmnrl_specid", "srfcplx_specid"):
            if ids in a:
                used.update(int(v) for v in a[ids])
        assert all(i < self.naq for i in used)
        self.coupled = sorted(used)
        self.cpos = {sp: ci for ci, sp in enumerate(self.coupled)}
        self.nc = len(self.coupled)
        self.out: List[str] = []
        # reactions as factor lists
        self.cx = [self._factors("eqcplx", k) for k in range(self.ncx)]
        self.mn = [self._factors("kinmnrl", m) for m in range(self.c.nkinmnrl)]
        self.sc = [self._factors("srfcplx", k) for k in range(self.c.nsrfcplx)]
        self.eqsr = [int(v) for v in a["eqsrfcplxrxn_to_srfcplxrxn"]] if self.c.neqsrfcplxrxn else []
        self.sr_cx = {r: [int(v) for v in a["srfcplxrxn_to_complex"][a["srfcplxrxn_ptr"][r]:a["srfcplxrxn_ptr"][r + 1]]]
                      for r in self.eqsr}
        self.sorb_species = sorted({i for r in self.eqsr for k in self.sr_cx[r] for i, _ in self.sc[k][0]})
        # structure of Jt (symmetric, species indices)
        self.struct = set()
        for sp, _h in self.cx + self.mn:
            for i, _ in sp:
                for j, _ in sp:
                    self.struct.add((i, j))
        for r in self.eqsr:
            spc = sorted({i for k in self.sr_cx[r] for i, _ in self.sc[k][0]})
            for i in spc:
                for j in spc:
                    self.struct.add((i, j))
        for i in self.coupled:
            self.struct.add((i, i))
        # powers of the activities (and of the activity of water) the products need
        self.pw: Dict[int, set] = {}
        self.wpw: set = set()
        for sp, h in self.cx + self.mn + self.sc:
            for i, nu in sp:
                self.pw.setdefault(i, set()).add(int(nu))
            if h:
                self.wpw.add(int(h))
        # The reference's d(total_sorb)/d(free) loops, per complex, over the species OF THAT COMPLEX
        # only (reaction_surf_complex.F90:860-890): the effect of species j on complex q through the
        # free-site concentration is dropped when q does not contain j.  Jt_ref = Jt_exact + E with
        # E(i, j) = (V/dt) nu_qi (S_q / Sx) tmp_j / den summed over such q -- a few columns, kept apart
        # from the symmetric matrix (sym: Sherman-Morrison; lu: added to the full matrix).
        self.ecol: Dict[int, List[Tuple[int, int, int, float]]] = {}   # j -> [(i, reaction, complex slot q, nu_qi)]
        for r in self.eqsr:
            cx = self.sr_cx[r]
            spc = sorted({i for k in cx for i, _ in self.sc[k][0]})
            for j in spc:
                for q, k in enumerate(cx):
                    d = dict(self.sc[k][0])
                    if j not in d:
                        for i, nu in d.items():
                            self.ecol.setdefault(j, []).append((i, r, q, nu))
        self.ecols = sorted(self.ecol)
        self.evar: Dict[Tuple[int, int], int] = {}                      # (i, j) -> index in ev[]
        for j in self.ecols:
            for i in sorted({t[0] for t in self.ecol[j]}):
                self.evar[(i, j)] = len(self.evar)
        if self.sym and len(self.ecols) > 2:
            raise ValueError("form 2 / sym: more than two incomplete columns in the reference's sorption Jacobian")
        self._symbolic()

    # ------------------------------------------------------------------ helpers
    def _class_of(self, z: float, a0: float) -> int:
        if not abs(z) > 1.0e-10:
            return -1
        key = (-z * z, float(a0))
        if key not in self.cls:
            self.cls.append(key)
        return self.cls.index(key)

    def _factors(self, kind: str, k: int):
        a = self.a
        ptr, ids, st = a[kind + "_ptr"], a[kind + "_specid"], a[kind + "_stoich"]
        sp = [(int(ids[p]), float(st[p])) for p in range(ptr[k], ptr[k + 1])]
        return sp, float(a[kind + "_h2ostoich"][k])

    def _symbolic(self) -> None:
        """elimination order (greedy minimum fill) and the structure of L for the sparse L D L^T"""
        nodes = list(self.coupled)
        adj = {i: {j for j in nodes if j != i and (i, j) in self.struct} for i in nodes}
        order: List[int] = []
        work = {k: set(v) for k, v in adj.items()}
        while work:
            best = None
            for v, nb in work.items():
                nbs = sorted(nb)
                fill = sum(1 for x in range(len(nbs)) for y in range(x + 1, len(nbs)) if nbs[y] not in work[nbs[x]])
                key = (fill, len(nb), v)
                if best is None or key < best[0]:
                    best = (key, v)
            v = best[1]
            nb = work.pop(v)
            nbs = sorted(nb)
            for x in range(len(nbs)):
                for y in range(x + 1, len(nbs)):
                    work[nbs[x]].add(nbs[y])
                    work[nbs[y]].add(nbs[x])
            for x in nb:
                work[x].discard(v)
            order.append(v)
        self.order = order                      # position -> species
        self.epos = {sp: p for p, sp in enumerate(order)}
        # column structure of L by positions: col[j] = sorted rows i > j
        n = len(order)
        low = {j: set() for j in range(n)}
        for (a_, b_) in self.struct:
            pa, pb = self.epos[a_], self.epos[b_]
            if pa > pb:
                low[pb].add(pa)
        for j in range(n):
            rows = sorted(low[j])
            for x in range(len(rows)):
                for y in range(x + 1, len(rows)):
                    low[rows[x]].add(rows[y])   # fill: rows[y] > rows[x]
        self.lcol = {j: sorted(low[j]) for j in range(n)}
        self.lrow = {i: sorted(j for j in range(n) if i in low[j]) for i in range(n)}
        # slots: diagonal first, then the columns
        self.lslot: Dict[Tuple[int, int], int] = {}
        k = 0
        for j in range(n):
            self.lslot[(j, j)] = k
            k += 1
        for j in range(n):
            for i in self.lcol[j]:
                self.lslot[(i, j)] = k
                k += 1
        self.nl = k

    def w(self, s: str = "") -> None:
        self.out.append(s)

    @staticmethod
    def _pname(base: str, p: int) -> str:
        return f"{base}p{p}" if p > 0 else f"{base}m{-p}"

    def _emit_powers(self, base: str, powers, indent: str = "  ") -> None:
        """base + 'p1' holds the mantissa f; the other powers as multiplication chains
        (q = ceil(q/2) + floor(q/2); only powers on the way to a needed one are emitted)"""
        pos = sorted(p for p in powers if p > 1)
        neg = sorted(-p for p in powers if p < -1)

        def chain(prefix: str, need: List[int]) -> None:
            made = {1}
            for p in need:
                todo = [p]
                while todo:
                    q = todo[-1]
                    if q in made:
                        todo.pop()
                        continue
                    aa, bb = (q + 1) // 2, q // 2
                    miss = [t for t in (aa, bb) if t not in made]
                    if miss:
                        todo.extend(miss)
                        continue
                    self.w(f"{indent}const double {base}{prefix}{q} = {base}{prefix}{aa} * {base}{prefix}{bb};")
                    made.add(q)
                    todo.pop()

        chain("p", pos)
        if any(p < 0 for p in powers):
            self.w(f"{indent}const double {base}m1 = sx_rcp({base}p1);")
            chain("m", neg)

    def _product(self, kf: float, ke: int, sp, h2o: float, extra: Optional[str], pvar: str, evar: str,
                 indent: str = "    ") -> None:
        """pvar = kf * prod f_i^nu (* fw^h2o) (* extra); evar = ke + sum nu e_i (+ h2o ew)"""
        if os.environ.get("PFRX_S2_DBG_EXPFORM"):
            # debugging aid: the reference's formulation exp(-lnK + sum nu ln a) for every product
            lnv = math.log(kf) + ke * math.log(2.0)
            expr = _lit(lnv)
            for i, nu in sp:
                g = "" if (not self.act_upd or self.pri_cls[i] < 0) else f" + log(g{self.pri_cls[i]})"
                if not self.act_upd:
                    g = f" + log(SW(S2_OFF_FRZ + {i}))"
                expr += f" + {_lit(nu)} * (log(c{i}){g})"
            if h2o:
                expr += f" + {_lit(h2o)} * log(s.aw)"
            ex_ = f" * {extra}" if extra else ""
            self.w(f"{indent}const double {pvar} = exp({expr}){ex_};")
            self.w(f"{indent}const int {evar} = 0;")
            return
        fac = [self._pname(f"f{i}", int(nu)) for i, nu in sp]
        ex = [(int(nu), f"e{i}") for i, nu in sp]
        if h2o:
            fac.append(self._pname("fw", int(h2o)))
            ex.append((int(h2o), "ew"))
        if extra:
            fac.append(extra)
        # (f_a * f_b) * (K * f_c) ...: a shallow product tree instead of a chain
        terms = [self.kc(kf)] + fac
        while len(terms) > 1:
            nxt = []
            for x in range(0, len(terms) - 1, 2):
                nxt.append(f"({terms[x]} * {terms[x + 1]})")
            if len(terms) % 2:
                nxt.append(terms[-1])
            terms = nxt
        self.w(f"{indent}const double {pvar} = {terms[0]};")
        terms = [str(ke)] if ke else []
        for nu, e in ex:
            if nu == 1:
                terms.append(f"+ {e}")
            elif nu == -1:
                terms.append(f"- {e}")
            else:
                terms.append(f"{'+' if nu > 0 else '-'} {abs(nu)} * {e}")
        s = " ".join(terms) if terms else "0"
        if s.startswith("+ "):
            s = s[2:]
        self.w(f"{indent}const int {evar} = {s};")

    def kc(self, v: float) -> str:
        """a double constant as an operand from the constant bank (DMUL R, R, c[3][..]) instead of an
        immediate, which costs two UMOVs per use when its low word is not zero"""
        if not int(os.environ.get("PFRX_SPEC2_KTAB", "1")):
            return _lit(v)
        key = float(v).hex()
        if key not in self.ktab:
            self.ktab[key] = len(self.ktab)
        return f"SK({self.ktab[key]})"

    def H(self, i: int, j: int) -> Tuple[int, int]:
        return (i, j) if i <= j else (j, i)

    def slot(self, i: int, j: int) -> str:
        """where Jt(i, j) (species indices, symmetric) is accumulated in the slice"""
        if self.sym:
            pi, pj = self.epos[i], self.epos[j]
            if pi < pj:
                pi, pj = pj, pi
            return f"SW({self.lslot[(pi, pj)]})"
        ci, cj = self.cpos[i], self.cpos[j]
        if ci > cj:
            ci, cj = cj, ci
        return f"W[JX({ci}, {cj})]"

    def _gen_activities(self, update_aw: bool) -> None:
        """activity coefficients at the ionic strength I (LAG: RActivityCoefficients) or the frozen
        ones of the slice; mantissa / exponent of the activities and the powers the products need.
        Emitted twice with identical arithmetic: in spec2_eval and in spec2_store_sec."""
        c, a, naq, w = self.c, self.a, self.naq, self.w
        if self.act_upd:
            w("  const double sq = sqrt(I);")
            A, B, Bd = _lit(c.debyeA), _lit(c.debyeB), _lit(c.debyeBdot)
            for q, (negz2, a0) in enumerate(self.cls):
                w(f"  const double g{q} = sx_exp((sx_div({_lit(negz2)} * sq * {A}, 1.0 + {_lit(a0)} * {B} * sq) + {Bd} * I) * SPEC_LN);")
            if c.use_activity_h2o and update_aw:
                mp = " + ".join(f"c{i}" for i in range(naq) if i != c.h2o_aq_id) or "0.0"
                w(f"  if (s.store) {{ const double t = 1.0 - 0.017 * (({mp}) + s.msec); s.aw = t > 0.0 ? t : 1.0; }}")
            used_rg = sorted({q for q in self.cx_cls if q >= 0})
            for q in used_rg:
                w(f"  const double rg{q} = sx_rcp(g{q});")
        # ---- mantissa / exponent of the activities
        for i in sorted(self.pw):
            if self.act_upd:
                g = "" if self.pri_cls[i] < 0 else f" * g{self.pri_cls[i]}"
            else:
                g = f" * SW(S2_OFF_FRZ + {i})"
            w(f"  int e{i}; const double f{i}p1 = s2_split(c{i}{g}, e{i});")
            self._emit_powers(f"f{i}", self.pw[i])
        if self.wpw:
            w("  int ew; const double fwp1 = s2_split(s.aw, ew);")
            self._emit_powers("fw", self.wpw)

    def _gen_complex(self, k: int) -> None:
        """sk = molality of aqueous complex k (RTotalAqueous, reaction_aux / reaction.F90 product form)"""
        a, naq = self.a, self.naq
        sp, h2o = self.cx[k]
        kf, ke = _kinv(float(a["eqcplx_logK"][k]))
        if self.act_upd:
            extra = f"rg{self.cx_cls[k]}" if self.cx_cls[k] >= 0 else None
        else:
            extra = f"SW(S2_OFF_FRZ + {naq + k})"
        self._product(kf, ke, sp, h2o, extra, "p", "e")
        self.w("    const double sk = s2_scale(p, e, emax);")

    # ------------------------------------------------------------------ spec2_eval
    def gen_eval(self) -> None:
        c, a, n, naq = self.c, self.a, self.n, self.naq
        w = self.w
        start = len(self.out)
        w("S2_FN void spec2_eval(double (&res)[SPEC_N], double (&tv)[S2_NTV], double (&ev)[S2_NEV], Spec2Cell &s, double *W,")
        w("                      const DevState &st, long long cell) {")
        for x in range(max(1, len(self.evar))):
            w(f"  ev[{x}] = 0.0;")
        w("  const long long ld = st.ld;")
        w("  const double psv = s.dry ? 0.0 : s.psv;")
        w("  int emax = 0;")
        for i in range(n):
            w(f"  const double c{i} = SW(S2_OFF_C + {i});")
        # ---- activity coefficients (RActivityCoefficients, LAG branch)
        if self.act_upd:
            terms = [f"c{i} * {_lit(float(a['primary_spec_Z'][i]) ** 2)}" for i in range(naq)
                     if float(a["primary_spec_Z"][i]) != 0.0]
            w("  double Ip = 0.0;")
            for t in terms:
                w(f"  Ip += {t};")
            w("  const double I = 0.5 * (Ip + s.Isec);")
            w("  if (s.store) s.Iact = I;")
        self._gen_activities(update_aw=True)
        # ---- RTotalAqueous
        for i in self.coupled:
            w(f"  double tot{i} = c{i};")
        w("  double Is = 0.0, ms = 0.0;")
        hits: Dict[Tuple[int, int], int] = {}
        for sp, _h in self.cx:
            ids = [i for i, _ in sp]
            for x in ids:
                for y in ids:
                    if x <= y:
                        hits[(x, y)] = hits.get((x, y), 0) + 1
        budget = int(os.environ.get("PFRX_SPEC2_HOT", "32"))
        hot = set(sorted(hits, key=lambda e: (-hits[e], e))[:budget])
        for (i, j) in sorted(hot):
            w(f"  double h_{i}_{j} = 0.0;")
        written = set()
        self.aq_struct = set(hits)
        for k in range(self.ncx):
            sp, h2o = self.cx[k]
            w("  {")
            self._gen_complex(k)
            z2 = float(a["eqcplx_Z"][k]) ** 2
            if z2 == 1.0:
                w("    Is += sk;")
            elif z2 != 0.0:
                w(f"    Is += sk * {_lit(z2)};")
            w("    ms += sk;")
            for i, nu in sp:
                w(f"    tot{i} += sk;" if nu == 1.0 else (f"    tot{i} -= sk;" if nu == -1.0 else f"    tot{i} += {_lit(nu)} * sk;"))
            for x, (i, nui) in enumerate(sp):
                for (j, nuj) in sp[x:]:
                    e_ = self.H(i, j)
                    wgt = nui * nuj
                    val = "sk" if wgt == 1.0 else f"{_lit(wgt)} * sk"
                    if e_ in hot:
                        w(f"    h_{e_[0]}_{e_[1]} += {val};")
                    else:
                        tgt = self.slot(*e_)
                        if e_ in written:
                            w(f"    {tgt} += {val};")
                        else:
                            w(f"    {tgt} = {val};")
                            written.add(e_)
            w("  }")
        w("  s.Isec = Is; s.msec = ms;")
        for i in self.coupled:
            w(f"  tot{i} *= s.denL;")
            w(f"  tv[{self.cpos[i]}] = tot{i};")
        # ---- equilibrium surface complexation (closed form, unit free-site stoichiometry)
        xs: Dict[Tuple[int, int], List[str]] = {}   # symmetric extra terms of Jt
        if self.eqsr:
            for k in range(c.nsrfcplx):
                w(f"  double scc{k} = 0.0;")
            for i in self.sorb_species:
                w(f"  double ts{i} = 0.0;")
        for r in self.eqsr:
            cx = self.sr_cx[r]
            ty = int(a["srfcplxrxn_surf_type"][r])
            dens = _lit(float(a["srfcplxrxn_site_density"][r]))
            spc = sorted({i for k in cx for i, _ in self.sc[k][0]})
            for x, i in enumerate(spc):
                for j in spc[x:]:
                    w(f"  double x{r}_{i}_{j} = 0.0;")
                    xs.setdefault((i, j), []).append(f"x{r}_{i}_{j}")
            w("  {")
            if ty == chem.MINERAL_SURFACE:
                w(f"    const double dens = {dens} * SW(S2_OFF_MN + {2 * int(a['srfcplxrxn_to_surf'][r])});")
            elif ty == chem.ROCK_SURFACE:
                w(f"    const double dens = {dens} * s.rock;")
            else:
                w(f"    const double dens = {dens};")
            w("    double fs = 0.0;")
            w("    if (!(dens < 1.e-40)) {")
            for q, k in enumerate(cx):
                sp, h2o = self.sc[k]
                kf, ke = _kinv(float(a["srfcplx_logK"][k]))
                self._product(kf, ke, sp, h2o, None, f"p{q}", f"pe{q}", indent="      ")
                w(f"      const double q{q} = s2_scale(p{q}, pe{q}, emax);")
            w("      double esum = 0.0;")
            for q in range(len(cx)):
                w(f"      esum += q{q};")
            w("      fs = sx_div(dens, 1.0 + esum);")
            for q, k in enumerate(cx):
                w(f"      const double S{q} = q{q} * fs;")
                w(f"      scc{k} += S{q};")
            w("      double den = 0.0;")
            for q in range(len(cx)):
                w(f"      den += S{q};")
            w("      const double rfs = sx_rcp(fs);")
            w("      den = den * rfs + 1.0;")
            w("      const double rden = sx_rcp(den);")
            for i in spc:
                w(f"      double tmp{i} = 0.0;")
            for q, k in enumerate(cx):
                for i, nu in self.sc[k][0]:
                    v = f"S{q}" if nu == 1.0 else f"{_lit(nu)} * S{q}"
                    w(f"      tmp{i} += {v};")
            for i in spc:
                w(f"      ts{i} += tmp{i};")
            w("      const double jscale = s.vol * s.rdt;")
            for x, i in enumerate(spc):
                for j in spc[x:]:
                    terms = []
                    for q, k in enumerate(cx):
                        d = dict(self.sc[k][0])
                        if i in d and j in d:
                            wgt = d[i] * d[j]
                            terms.append(f"S{q}" if wgt == 1.0 else f"{_lit(wgt)} * S{q}")
                    aij = " + ".join(terms) if terms else "0.0"
                    w(f"      x{r}_{i}_{j} = jscale * (({aij}) - (tmp{i} * rfs) * (tmp{j} * rden));")
            for j in self.ecols:
                for (i, rr, q, nu) in self.ecol[j]:
                    if rr == r:
                        w(f"      ev[{self.evar[(i, j)]}] += jscale * (({_lit(nu)} * S{q}) * rfs) * (tmp{j} * rden);")
            w("    }")
            w(f"    if (s.store) st.free_site[{r} * ld + cell] = fs;")
            w("  }")
        if self.eqsr:
            w("  if (s.store && st.eqsrfcplx_conc) {")
            for k in range(c.nsrfcplx):
                w(f"    st.eqsrfcplx_conc[{k} * ld + cell] = scc{k};")
            w("  }")
            for q, i in enumerate(self.sorb_species):
                w(f"  tv[{self.nc + q}] = ts{i};")
        # ---- residual: (accumulation - fixed accumulation) / dt (RReact, reaction.F90:3880-3900)
        for i in range(n):
            if i < naq:
                tot = f"tot{i}" if i in self.cpos else f"(SW(S2_OFF_C + {i}) * s.denL)"
                acc = f"psv * {tot}"
                if self.eqsr:
                    acc = f"{acc} + {'ts%d' % i if i in self.sorb_species else '0.0'} * s.vol"
                w(f"  res[{i}] = (({acc}) - SW(S2_OFF_FIX + {i})) * s.rdt;")
            else:
                w(f"  res[{i}] = s.dry ? 0.0 : ((0.0 + SW(S2_OFF_C + {i}) * s.vol) - SW(S2_OFF_FIX + {i})) * s.rdt;")
        # ---- kinetic minerals (TST)
        for m in range(c.nkinmnrl):
            sp, h2o = self.mn[m]
            kf, ke = _kinv(float(a["kinmnrl_logK"][m]))
            thr = float(a["kinmnrl_affinity_threshold"][m])
            lim = float(a["kinmnrl_rate_limiter"][m])
            eact = float(a["kinmnrl_activation_energy"][m])
            irr = int(a["kinmnrl_irreversible"][m])
            rate = _lit(float(a["kinmnrl_rate_constant"][m]))
            w(f"  double y{m} = 0.0;")
            w("  {")
            self._product(kf, ke, sp, h2o, None, "p", "e")
            w("    const double QK = s2_scale(p, e, emax);")
            w("    double aff = 1.0 - QK;")
            w("    const double sgn = copysign(1.0, aff);")
            w(f"    bool active = (SW(S2_OFF_MN + {2 * m}) > 0.0 || sgn < 0.0);")
            if irr == 1:
                w("    if (sgn < 0.0) active = false;")
            if thr > 0.0:
                w(f"    if (sgn < 0.0 && QK < {_lit(thr)}) active = false;")
            if lim > 0.0:
                w(f"    aff = aff / (1.0 + (1.0 - aff) / {_lit(lim)});")
            if eact > 0.0:
                w(f"    const double spr = {rate} * exp({_lit(eact)} / 8.31446 * (1.0 / (25.0 + 273.15) - 1.0 / (s.temp + 273.15)));")
            else:
                w(f"    const double spr = {rate} * 1.0;")
            w(f"    const double Im_const = -SW(S2_OFF_MN + {2 * m + 1});")
            w("    const double rate_vol = active ? Im_const * sgn * fabs(aff) * spr : 0.0;")
            w(f"    if (s.store && s.rates) st.mnrl_rate[{m} * ld + cell] = rate_vol;")
            w("    const bool apply = active && !s.dry;")
            w("    const double Im = apply ? rate_vol * s.vol : 0.0;")
            w("    const double dIm_dQK = -(Im_const * s.vol) * spr;")
            if lim > 0.0:
                w(f"    const double den = 1.0 + (1.0 - aff) / {_lit(lim)};")
                w(f"    const double dfac = dIm_dQK * (1.0 + QK / {_lit(lim)} / den) * QK * s.denL / den;")
            else:
                w("    const double dfac = dIm_dQK * QK * s.denL;")
            w(f"    y{m} = apply ? dfac : 0.0;")
            for i, nu in sp:
                w(f"    res[{i}] +={' ' if nu == 1.0 else f' {_lit(nu)} *'} Im;")
            w("  }")
            for x, (i, nui) in enumerate(sp):
                for (j, nuj) in sp[x:]:
                    wgt = nui * nuj
                    xs.setdefault(self.H(i, j), []).append(f"y{m}" if wgt == 1.0 else f"{_lit(wgt)} * y{m}")
        # an exponent beyond the double range: NaN, as exp() -> +Inf ends in the reference
        w("  if (emax > 960) res[0] = res[0] * S2_INF * 0.0;")
        # ---- Jt: K1 (delta_ij c_i + S_ij) + sorption + minerals
        w("  const double K1 = s.denL * (psv * s.rdt);")
        w("  const double dg = s.dry ? 1.0 : K1;")
        for x, i in enumerate(self.coupled):
            for j in self.coupled[x:]:
                if (i, j) not in self.struct:
                    if not self.sym:
                        ci, cj = self.cpos[i], self.cpos[j]
                        w(f"  W[JX({ci}, {cj})] = 0.0; W[JX({cj}, {ci})] = 0.0;")
                    continue
                terms = []
                if (i, j) in self.aq_struct:
                    src = f"h_{i}_{j}" if (i, j) in hot else self.slot(i, j)
                    terms.append(f"K1 * {src}")
                if i == j:
                    terms.append(f"dg * SW(S2_OFF_C + {i})")
                terms += xs.get((i, j), [])
                expr = " + ".join(terms) if terms else "0.0"
                if self.sym:
                    w(f"  {self.slot(i, j)} = {expr};")
                elif i == j:
                    e_ = f" + ev[{self.evar[(i, i)]}]" if (i, i) in self.evar else ""
                    w(f"  {self.slot(i, j)} = {expr}{e_};")
                else:
                    ci, cj = self.cpos[i], self.cpos[j]
                    eu = f" + ev[{self.evar[(i, j)]}]" if (i, j) in self.evar else ""
                    el = f" + ev[{self.evar[(j, i)]}]" if (j, i) in self.evar else ""
                    w(f"  {{ const double v = {expr}; W[JX({ci}, {cj})] = v{eu}; W[JX({cj}, {ci})] = v{el}; }}")
        w("}")
        w()
        self._inline_rare_powers(start)

    def _inline_rare_powers(self, start: int) -> None:
        """a power f<i>[pm]<n> that spec2_eval uses at most PFRX_SPEC2_INLINE times is multiplied out
        where it is used instead of being held in a register pair from the top of the routine
        (255 registers per thread: every pair that stays live through the 88 complexes is a spill)"""
        thresh = int(os.environ.get("PFRX_SPEC2_INLINE", "2"))
        if thresh <= 0:
            return
        body = self.out[start:]
        pat = re.compile(r"^\s*const double (f\w+[pm]\d+) = (f\w+[pm]\d+ \* f\w+[pm]\d+);$")
        changed = True
        while changed:
            changed = False
            for x in range(len(body) - 1, -1, -1):
                m = pat.match(body[x])
                if not m:
                    continue
                name, expr = m.group(1), m.group(2)
                word = re.compile(r"\b" + name + r"\b")
                uses = sum(len(word.findall(l)) for l in body[x + 1:])
                if 0 < uses <= thresh:
                    body[x + 1:] = [word.sub(f"({expr})", l) for l in body[x + 1:]]
                    del body[x]
                    changed = True
                    break
        self.out[start:] = body

    # ------------------------------------------------------------------ sparse L D L^T
    def gen_solve_sym(self) -> None:
        """left-looking L D L^T in the elimination order of _symbolic(), forward substitution fused
        (the right-hand sides are updated with each finished column), then D^-1 and L^-T.  Right-hand
        sides: the residual, and one vector per incomplete column of the reference's sorption
        Jacobian (Sherman-Morrison: (A + sum_r w_r e_jr^T) u = b)."""
        w = self.w
        n = self.nc
        nr = len(self.ecols)
        w("S2_FN bool spec2_solve_sym(double *W, double (&res)[SPEC_N], const double (&ev)[S2_NEV]) {")
        w("  bool ok = true;")
        for p in range(n):
            w(f"  double b{p} = res[{self.order[p]}];")
        # extra right-hand sides; structural non-zeros tracked per vector so that zeros cost nothing
        znz: List[set] = []
        for r, j in enumerate(self.ecols):
            nz = set()
            init = {self.epos[i]: x for (i, jj), x in self.evar.items() if jj == j}
            for p_ in range(n):
                w(f"  double z{r}_{p_} = {'ev[%d]' % init[p_] if p_ in init else '0.0'};")
            nz.update(init)
            znz.append(nz)
        struct_pos = set()
        for (a_, b_) in self.struct:
            pa, pb = self.epos[a_], self.epos[b_]
            struct_pos.add((max(pa, pb), min(pa, pb)))
        for j in range(n):
            w(f"  // column {j} (species {self.order[j]})")
            row = self.lrow[j]
            for k in row:
                w(f"  const double l{j}_{k} = SW({self.lslot[(j, k)]});")
                w(f"  const double v{j}_{k} = l{j}_{k} * d{k};")
            w(f"  double d{j} = SW({self.lslot[(j, j)]});")
            for k in row:
                w(f"  d{j} -= l{j}_{k} * v{j}_{k};")
            w(f"  if (!(d{j} > 0.0)) ok = false;")
            w(f"  const double r{j} = sx_rcp(d{j});")
            for i in self.lcol[j]:
                common = [k for k in self.lrow[i] if k in row]
                start_ = f"SW({self.lslot[(i, j)]})" if (i, j) in struct_pos else "0.0"
                w(f"  {{ double t = {start_};")
                for k in common:
                    w(f"    t -= SW({self.lslot[(i, k)]}) * v{j}_{k};")
                w(f"    t *= r{j};")
                w(f"    SW({self.lslot[(i, j)]}) = t;")
                w(f"    b{i} -= t * b{j};")
                for r in range(nr):
                    if j in znz[r]:
                        w(f"    z{r}_{i} -= t * z{r}_{j};")
                w("  }")
                for r in range(nr):
                    if j in znz[r]:
                        znz[r].add(i)
        for j in range(n):
            w(f"  b{j} *= r{j};")
            for r in range(nr):
                if j in znz[r]:
                    w(f"  z{r}_{j} *= r{j};")
        for j in range(n - 1, -1, -1):
            for i in self.lcol[j]:
                use = [r for r in range(nr) if i in znz[r]]
                w(f"  {{ const double l = SW({self.lslot[(i, j)]}); b{j} -= l * b{i};"
                  + "".join(f" z{r}_{j} -= l * z{r}_{i};" for r in use) + " }")
                for r in use:
                    znz[r].add(j)
        # (A + sum_r w_r e_jr^T) u = b:  u = y - sum_r z_r t_r,  (I + Z[j, :]) t = y[j]
        if nr == 1:
            pj = self.epos[self.ecols[0]]
            w(f"  const double t0 = sx_div(b{pj}, 1.0 + z0_{pj});")
            for p_ in range(n):
                if p_ in znz[0]:
                    w(f"  b{p_} -= z0_{p_} * t0;")
        elif nr == 2:
            pa, pb = self.epos[self.ecols[0]], self.epos[self.ecols[1]]
            w(f"  const double m00 = 1.0 + z0_{pa}, m01 = z1_{pa}, m10 = z0_{pb}, m11 = 1.0 + z1_{pb};")
            w("  const double rdet = 1.0 / (m00 * m11 - m01 * m10);")
            w(f"  const double t0 = (b{pa} * m11 - m01 * b{pb}) * rdet, t1 = (m00 * b{pb} - m10 * b{pa}) * rdet;")
            for p_ in range(n):
                terms = []
                if p_ in znz[0]:
                    terms.append(f"z0_{p_} * t0")
                if p_ in znz[1]:
                    terms.append(f"z1_{p_} * t1")
                if terms:
                    w(f"  b{p_} -= {' + '.join(terms)};")
        for p_ in range(n):
            w(f"  res[{self.order[p_]}] = b{p_};")
        w("  (void)ev;")
        w("  return ok;")
        w("}")
        w()

    def gen_begin(self) -> None:
        """RReact entry (reaction.F90:3829-3850): the fixed accumulation of the sub-step and what the
        kinetic minerals / mineral-bound sites read, into the slice; and, at RStep entry, the ionic
        strength sums of the state's secondary species"""
        c, a, n, naq = self.c, self.a, self.n, self.naq
        w = self.w
        w("S2_FN void spec2_begin(double *W, const Spec2Cell &s, const DevState &st, long long cell) {")
        w("  const long long ld = st.ld;")
        w("  const double psv = s.dry ? 0.0 : s.psv;")
        for i in range(n):
            if i < naq:
                fix = f"psv * st.total[{i} * ld + cell]"
                if self.eqsr:
                    fix = f"{fix} + st.total_sorb_eq[{i} * ld + cell] * s.vol"
                w(f"  SW(S2_OFF_FIX + {i}) = {fix};")
            else:
                w(f"  SW(S2_OFF_FIX + {i}) = 0.0 + st.immobile[{i - naq} * ld + cell] * s.vol;")
        for m in range(c.nkinmnrl):
            w(f"  SW(S2_OFF_MN + {2 * m}) = st.mnrl_volfrac[{m} * ld + cell];")
            w(f"  SW(S2_OFF_MN + {2 * m + 1}) = st.mnrl_area[{m} * ld + cell];")
        w("  (void)psv;")
        w("}")
        w()
        w("S2_FN void spec2_isec(Spec2Cell &s, const DevState &st, long long cell) {")
        w("  const long long ld = st.ld;")
        w("  const double *sp_ = st.sec_molal + cell;")
        for k in range(self.ncx):
            w(f"  const double m{k} = sp_[{k} * ld];")
        w("  double Is = 0.0, ms = 0.0;")
        for k in range(self.ncx):
            z2 = float(a["eqcplx_Z"][k]) ** 2
            if z2 == 1.0:
                w(f"  Is += m{k};")
            elif z2 != 0.0:
                w(f"  Is += m{k} * {_lit(z2)};")
            w(f"  ms += m{k};")
        w("  s.Isec = Is; s.msec = ms;")
        w("}")
        w()

    def gen_stores(self) -> None:
        c, a, naq = self.c, self.a, self.naq
        w = self.w
        w("S2_FN void spec2_store_totals(const double (&tv)[S2_NTV], const double *W, const Spec2Cell &s, const DevState &st,")
        w("                              long long cell, bool tot, bool sorb) {")
        w("  const long long ld = st.ld;")
        w("  if (tot) {")
        for i in range(naq):
            if i in self.cpos:
                w(f"    st.total[{i} * ld + cell] = tv[{self.cpos[i]}];")
            else:
                w(f"    st.total[{i} * ld + cell] = SW(S2_OFF_C + {i}) * s.denL;")
        w("  }")
        if self.eqsr:
            w("  if (sorb) {")
            for i in range(naq):
                if i in self.sorb_species:
                    w(f"    st.total_sorb_eq[{i} * ld + cell] = tv[{self.nc + self.sorb_species.index(i)}];")
                else:
                    w(f"    st.total_sorb_eq[{i} * ld + cell] = 0.0;")
            w("  }")
        w("  (void)sorb; (void)s; (void)W;")
        w("}")
        w()
        # rt_auxvar%sec_molal of the cell's last evaluation, recomputed once when the cell is published
        # (same source text as spec2_eval's products: products of the same operands in the same
        # order, no additions to contract, so the values are the ones that evaluation had)
        w("S2_FN void spec2_store_sec(const double *W, const Spec2Cell &s, const DevState &st, long long cell) {")
        if self.ncx:
            w("  const long long ld = st.ld;")
            w("  int emax = 0;")
            for i in sorted(self.pw):
                w(f"  const double c{i} = SW(S2_OFF_C + {i});")
            if self.act_upd:
                w("  const double I = s.Iact;")
            self._gen_activities(update_aw=False)
            w("  double *sp_ = st.sec_molal + cell;")
            for k in range(self.ncx):
                w("  {")
                self._gen_complex(k)
                w(f"    sp_[{k} * ld] = sk;")
                w("  }")
            w("  (void)emax;")
        w("  (void)W; (void)s; (void)st; (void)cell;")
        w("}")
        w()
        w("S2_FN void spec2_store_act(const Spec2Cell &s, const DevState &st, long long cell) {")
        if self.act_upd:
            w("  const long long ld = st.ld;")
            w("  const double I = s.Iact, sq = sqrt(I);")
            A, B, Bd = _lit(c.debyeA), _lit(c.debyeB), _lit(c.debyeBdot)
            for q, (negz2, a0) in enumerate(self.cls):
                w(f"  const double g{q} = sx_exp((sx_div({_lit(negz2)} * sq * {A}, 1.0 + {_lit(a0)} * {B} * sq) + {Bd} * I) * SPEC_LN);")
            for i in range(naq):
                q = self.pri_cls[i]
                w(f"  st.pri_act_coef[{i} * ld + cell] = {'1.0' if q < 0 else 'g%d' % q};")
            w("  double *sp_ = st.sec_act_coef + cell;")
            for k in range(self.ncx):
                q = self.cx_cls[k]
                w(f"  *sp_ = {'1.0' if q < 0 else 'g%d' % q}; sp_ += ld;")
        w("  (void)s; (void)st; (void)cell;")
        w("}")
        w()
        w("S2_FN void spec2_load_frozen(double *W, const DevState &st, long long cell) {")
        if not self.act_upd:
            w("  const long long ld = st.ld;")
            for i in range(naq):
                w(f"  SW(S2_OFF_FRZ + {i}) = st.pri_act_coef[{i} * ld + cell];")
            w("#pragma unroll 4")
            w(f"  for (int k = 0; k < {self.ncx}; k++) SW(S2_OFF_FRZ + {naq} + k) = 1.0 / st.sec_act_coef[k * ld + cell];")
        w("  (void)W; (void)st; (void)cell;")
        w("}")
        w()

    # ------------------------------------------------------------------ whole file
    def layout(self) -> Tuple[int, int, int]:
        """(slots per thread, threads per block, min blocks per SM)"""
        self.nstash = self.nc + len(self.sorb_species)
        nj = self.nl if self.sym else self.nc * (self.nc + 1)
        slots = nj + 2 * self.n + 2 * self.c.nkinmnrl + (0 if self.act_upd else self.naq + self.ncx)
        smem_max = 227 * 1024
        if self.n <= 8:
            threads = 128
            minblocks = max(1, min(6, (smem_max + 1024) // (slots * threads * 8 + 1024)))
            return slots, threads, minblocks
        # large systems: as many warps as the slices admit, at most two per scheduler (the register
        # file holds 8 warps of 255 registers); one block, so that they share one instruction stream
        warps = min(8, (smem_max) // (slots * 32 * 8))
        warps = 8 if warps >= 8 else (4 if warps >= 4 else max(1, warps))
        env = os.environ.get("PFRX_SPEC2_WARPS")
        if env:
            warps = int(env)
        return slots, 32 * warps, 1

    def source(self) -> str:
        from . import specialize as sp1

        c = self.c
        slots, threads, minblocks = self.layout()
        self.out = []
        self.ktab: Dict[str, int] = {}
        self.gen_eval()
        if self.sym:
            self.gen_solve_sym()
        self.gen_begin()
        self.gen_stores()
        body = self.out
        self.out = []
        w = self.w
        w("// generated by pflotran_elm_interface_b200/specialize2.py -- do not edit")
        w("#define SPEC_FORM 2")
        w(f"#define SPEC_N {self.n}")
        w(f"#define SPEC_NAQ {self.naq}")
        w(f"#define SPEC_NC {self.nc}")
        w(f"#define SPEC_NCX {self.ncx}")
        w(f"#define SPEC_NCLS {len(self.cls)}")
        w(f"#define SPEC_NKIN {c.nkinmnrl}")
        w(f"#define SPEC_NSRFRXN {c.nsrfcplxrxn}")
        w(f"#define SPEC_NSRFCPLX {c.nsrfcplx}")
        w(f"#define SPEC_NEQSR {c.neqsrfcplxrxn}")
        w(f"#define SPEC_NSTASH {self.nstash}")
        w(f"#define SPEC_SYM {int(self.sym)}")
        w(f"#define SPEC_NL {self.nl}")
        w(f"#define SPEC_NEV {len(self.evar)}")
        w("#define SPEC_USE_LOG 1")
        w(f"#define SPEC_ACT_UPD {int(self.act_upd)}")
        w(f"#define SPEC_USE_ACT_H2O {int(c.use_activity_h2o)}")
        w(f"#define SPEC_SIG {sp1.signature(self.cfg)}ull")
        w(f"#define SPEC_THREADS {threads}")
        w(f"#define SPEC_MINBLOCKS {minblocks}")
        w(f"#define SPEC_FASTMATH {int(os.environ.get('PFRX_SPEC_FASTMATH', '1'))}")
        w(f"#define SPEC_REFILL {int(self.refill)}")
        w(f"#define SPEC_SYNC {int(self.sync)}")
        cm = " : ".join(f"i == {sp} ? {ci}" for sp, ci in self.cpos.items())
        so = " : ".join(f"ci == {ci} ? {sp}" for sp, ci in self.cpos.items())
        w("#ifndef S2_HOST")
        w("#define S2_CE __host__ __device__ constexpr")
        w("#else")
        w("#define S2_CE constexpr")
        w("#endif")
        w("S2_CE int spec_cmap(int i) { return " + (cm + " : -1" if cm else "-1") + "; }")
        w("S2_CE int spec_sp_of(int ci) { return " + (so + " : 0" if so else "0") + "; }")
        masks = []
        for sp_i, ci in self.cpos.items():
            m = 0
            for sp_j, cj in self.cpos.items():
                if (sp_i, sp_j) in self.struct:
                    m |= 1 << cj
            masks.append((ci, m))
        jm = " : ".join(f"ci == {ci} ? {m}ull" for ci, m in masks)
        w("S2_CE unsigned long long spec_jrow_mask(int ci) { return " + (jm + " : 0ull" if jm else "0ull") + "; }")
        w("S2_CE bool spec_jnz(int ci, int cj) { return (spec_jrow_mask(ci) >> cj) & 1ull; }")
        z2 = [float(z) * float(z) for z in self.a["eqcplx_Z"]]
        vol = [float(v) for v in self.a["kinmnrl_molar_vol"]] if c.nkinmnrl else [0.0]
        w("#ifndef S2_HOST")
        w("static __device__ const double spec_cx_z2_tab[] = {" + ", ".join(_lit(v) for v in z2) + "};")
        w("__device__ __forceinline__ double spec_cx_z2(int k) { return spec_cx_z2_tab[k]; }")
        w("#else")
        w("static const double spec_cx_z2_tab[] = {" + ", ".join(_lit(v) for v in z2) + "};")
        w("static inline double spec_cx_z2(int k) { return spec_cx_z2_tab[k]; }")
        w("#endif")
        mv = " : ".join(f"m == {m} ? {_lit(v)}" for m, v in enumerate(vol))
        w("S2_CE double spec_mn_vol(int m) { return " + mv + " : 0.0; }")
        kt = ", ".join(_lit(float.fromhex(k)) for k in self.ktab) or "0.0"
        w("#ifndef S2_HOST")
        w("static __constant__ double spec_k_tab[] = {" + kt + "};")
        w("#else")
        w("static const double spec_k_tab[] = {" + kt + "};")
        w("#endif")
        w("#define SK(i) spec_k_tab[i]")
        w('#include "pfrx_spec2.cuh"')
        w()
        self.slots, self.threads, self.minblocks = slots, threads, minblocks
        return "\n".join(self.out + body) + "\n"


def generate_source2(cfg: abi.ReactionConfig, style: str, solver: Optional[str] = None) -> str:
    """style: straight (warps of a block independent), lockstep, refill, refill_warp"""
    sync = style in ("lockstep", "refill")
    refill = style in ("refill", "refill_warp")
    return _Gen2(cfg, threads_sync=sync, refill=refill, solver=solver).source()


FORM2_STYLES = ("straight", "lockstep", "refill", "refill_warp")
