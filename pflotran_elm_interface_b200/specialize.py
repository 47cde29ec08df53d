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
VARIANT_STYLES = {"s": "straight", "k": "lockstep", "l": "looplu", "m": "klooplu", "r": "rolled"}


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
    if c.clmcn_nrxn > 0:
        parts.append(struct.pack("<3i", c.clmcn_npool, c.clmcn_C_species_id, c.clmcn_N_species_id))
        add("clmcn_CN_ratio", c.clmcn_npool, f8)
        for k in ("clmcn_pool_nspec", "clmcn_pool_C_id", "clmcn_pool_N_id"):
            add(k, c.clmcn_npool, i4)
        for k in ("clmcn_upstream_pool_id", "clmcn_downstream_pool_id"):
            add(k, c.clmcn_nrxn, i4)
        for k in ("clmcn_rate_constant", "clmcn_respiration_fraction", "clmcn_inhibition_constant"):
            add(k, c.clmcn_nrxn, f8)
    return _fnv1a(b"".join(parts))


def _variant(cfg: abi.ReactionConfig, warps: Optional[int], style: Optional[str]) -> Tuple[int, str]:
    """(warps per 32 cells, code style) -- defaults from default_variant()"""
    dw, ds = default_variant(cfg)
    return (dw if warps is None else warps), (ds if style is None else style)


def cubin_path(cfg: abi.ReactionConfig, warps: Optional[int] = None, style: Optional[str] = None) -> str:
    warps, style = _variant(cfg, warps, style)
    tag = {v: k for k, v in VARIANT_STYLES.items()}[style]
    return os.path.join(OUT, f"spec_{signature(cfg):016x}_{tag}{warps}.cubin")


def supported(cfg: abi.ReactionConfig) -> Tuple[bool, str]:
    c, a = cfg.c, cfg.arrays
    n = c.naqcomp + c.nimcomp
    if n > 20:
        return False, "more than 20 unknowns"
    if not c.use_isothermal:
        return False, "anisothermal logK"
    if c.act_coef_update_algorithm != abi._chem.ACT_COEF_ALGORITHM_LAG:
        return False, "activity algorithm NEWTON"
    if c.nkinmrsrfcplxrxn > 0:
        return False, "multirate sorption"
    if c.nsrfcplxrxn != c.neqsrfcplxrxn:
        return False, "non-equilibrium surface complexation"
    if c.nsrfcplxrxn and np.any(a["srfcplxrxn_stoich_flag"] != 0):
        return False, "free-site stoichiometry other than 1"
    for k in ("kinmnrl_Temkin_const", "kinmnrl_min_scale_factor", "kinmnrl_affinity_power",
              "kinmnrl_num_prefactors"):
        if k in a:
            return False, k
    if c.use_total_as_guess:
        return False, "USE_TOTAL_CONCENTRATION_AS_GUESS"
    if c.somdec or c.nitrif or c.denitr:
        return False, "SOMDECOMP / NITRIFICATION / DENITRIFICATION sandbox"
    return True, ""


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
        self.coupled = sorted(used)
        self.cpos = {sp: ci for ci, sp in enumerate(self.coupled)}
        self.nc = len(self.coupled)

    def J(self, i: int, j: int) -> str:
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
        self.w("__device__ __forceinline__ void spec_sorption(const double (&lna)[SPEC_N], const double (&ic)[SPEC_N],")
        self.w("    double (&ts)[SPEC_N], SpecCell &s, double *W, const DevState &st, long long cell, double jscale) {")
        for k in range(c.nsrfcplx):
            self.w(f"  if (s.store) s.scconc[{k}] = 0.0;")
        for e in range(c.neqsrfcplxrxn):
            r = int(a["eqsrfcplxrxn_to_srfcplxrxn"][e])
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
                    self.w(f"      tmp{i} += {v}; ts[{i}] += {v};")
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
        self.w("}")
        self.w()

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

    def gen_sandbox(self) -> None:
        """CLM_CN_React (reaction_sandbox_clm_cn.F90:468-787), one straight-line block per reaction"""
        c, a, naq = self.c, self.a, self.naq
        self.w("__device__ __forceinline__ void spec_sandbox(const double (&c)[SPEC_N], double (&res)[SPEC_N],")
        self.w("    const SpecCell &s, double *W) {")
        self.w("  const double temp_K = s.temp + 273.15;")
        self.w("  if (!(temp_K > 227.15)) return;")
        self.w("  const double F_t = exp(308.56 * (1.408054069e-2 - 1.0 / (temp_K - 227.13)));")
        self.w("  const double F_theta = log(0.01 / fmax(0.01, s.sat)) * -2.17147241e-1;")
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
        self.w("}")
        self.w()

    # ------------------------------------------------------------------ whole file
    def source(self) -> str:
        c = self.c
        n = self.n
        if self.loop_lu:
            slots = self.nc * (self.nc + 2) + self.nc
        else:
            slots = self.nc * (self.nc + 1) + 2 * n
        slots = max(1, slots + (0 if self.act_upd else self.ncx))
        per_warp = slots * 32 * 8 + 1024  # + the per-block reservation when a block is one warp
        if slots * 32 * 8 > 160 * 1024:
            threads = 32
            minblocks = 1
        elif n > 8:
            threads = 32
            # measured (profiles/r01_occupancy_sweep_spec.txt): one warp per scheduler is the
            # optimum for these 255-register kernels; a fifth warp per SM costs 12 %
            minblocks = max(1, min(4, (228 * 1024) // per_warp))
        else:
            threads = 128
            minblocks = max(1, min(4, (228 * 1024) // (slots * 128 * 8 + 1024)))
        if self.lockstep and threads == 32:
            # the warps that shared an SM as separate blocks become one block that votes
            threads = 32 * minblocks
            minblocks = 1
        self.threads, self.minblocks, self.slots = threads, minblocks, slots
        o = self.out
        o.clear()
        self.w("// generated by pflotran_elm_interface_b200/specialize.py -- do not edit")
        self.w(f"#define SPEC_N {n}")
        self.w(f"#define SPEC_NAQ {self.naq}")
        self.w(f"#define SPEC_NC {self.nc}")
        self.w(f"#define SPEC_NCX {self.ncx}")
        self.w(f"#define SPEC_NCLS {len(self.cls)}")
        self.w(f"#define SPEC_NKIN {c.nkinmnrl}")
        self.w(f"#define SPEC_NSRFRXN {c.nsrfcplxrxn}")
        self.w(f"#define SPEC_NSRFCPLX {c.nsrfcplx}")
        self.w(f"#define SPEC_NEQSR {c.neqsrfcplxrxn}")
        self.w(f"#define SPEC_NCLM {c.clmcn_nrxn}")
        self.w(f"#define SPEC_USE_LOG {int(c.use_log_formulation)}")
        self.w(f"#define SPEC_ACT_UPD {int(self.act_upd)}")
        self.w(f"#define SPEC_USE_ACT_H2O {int(c.use_activity_h2o)}")
        self.w(f"#define SPEC_SIG {signature(self.cfg)}ull")
        self.w(f"#define SPEC_THREADS {threads}")
        self.w(f"#define SPEC_FASTMATH {int(os.environ.get('PFRX_SPEC_FASTMATH', '1'))}")
        self.w(f"#define SPEC_LOOP_LU {int(self.loop_lu)}")
        self.w(f"#define SPEC_LOCKSTEP {int(self.lockstep)}")
        self.w(f"#define SPEC_MINBLOCKS {minblocks}")
        cm = " : ".join(f"i == {sp} ? {ci}" for sp, ci in self.cpos.items())
        so = " : ".join(f"ci == {ci} ? {sp}" for sp, ci in self.cpos.items())
        self.w("__host__ __device__ constexpr int spec_cmap(int i) { return " + (cm + " : -1" if cm else "-1") + "; }")
        self.w("__host__ __device__ constexpr int spec_sp_of(int ci) { return " + (so + " : 0" if so else "0") + "; }")
        pc = " : ".join(f"i == {i} ? {q}" for i, q in enumerate(self.pri_cls) if q >= 0)
        self.w("__host__ __device__ constexpr int spec_pri_cls(int i) { return " + (pc + " : -1" if pc else "-1") + "; }")
        self.w("__device__ __forceinline__ double spec_cx_z2(int k);")
        self.w("__device__ __forceinline__ int spec_cx_cls(int k);")
        self.w("__device__ __forceinline__ double spec_mn_vol(int m);")
        self.w('#include "pfrx_spec.cuh"')
        self.w()
        self.gen_tables()
        self.gen_activity()
        self.gen_rtotal()
        self.gen_sorption()
        self.gen_minerals()
        if c.clmcn_nrxn > 0:
            self.gen_sandbox()
        return "\n".join(o) + "\n"


class _GenW(_Gen):
    """SPEC_W warps per group of 32 cells (csrc/pfrx_specw.cuh)"""

    def __init__(self, cfg: abi.ReactionConfig, warps: int):
        super().__init__(cfg)
        self.W = warps
        ok, why = supported_multiwarp(cfg, warps)
        if not ok:
            raise ValueError("multi-warp specialisation: " + why)
        # species owner: coupled species by matrix position (the LU's row ownership),
        # the others round-robin
        self.owner = {}
        dec = 0
        for i in range(self.n):
            if i in self.cpos:
                self.owner[i] = self.cpos[i] % warps
            else:
                self.owner[i] = dec % warps
                dec += 1
        self.cx_owner = [k % warps for k in range(self.ncx)]

    def lna(self, i: int) -> str:
        return f"SW(EXS({self.cpos[i]}))"

    def ic(self, i: int) -> str:
        return f"SW(SW_OFF_IC + {self.cpos[i]})"

    def lnqk(self, logk: float, h2o: float, ptr, ids, st, k: int) -> str:
        expr = _lit(-float(logk) * LOG_TO_LN)
        if h2o != 0.0:
            expr += _term(h2o, "s.ln_act_h2o")
        for p in range(ptr[k], ptr[k + 1]):
            expr += _term(float(st[p]), self.lna(int(ids[p])))
        return expr

    def gen_maps(self) -> None:
        def chain(pairs, var, default):
            body = " : ".join(f"{var} == {a} ? {b}" for a, b in pairs)
            return (body + f" : {default}") if body else str(default)

        self.w("__host__ __device__ constexpr int spec_cmap(int i) { return " +
               chain(self.cpos.items(), "i", -1) + "; }")
        self.w("__host__ __device__ constexpr int spec_sp_of(int ci) { return " +
               chain([(ci, sp) for sp, ci in self.cpos.items()], "ci", 0) + "; }")
        self.w("__host__ __device__ constexpr int spec_owner(int i) { return " +
               chain(self.owner.items(), "i", 0) + "; }")
        self.w("__host__ __device__ constexpr int spec_cx_owner(int k) { return k % SPEC_W; }")
        z2 = [(i, _lit(float(self.a["primary_spec_Z"][i]) ** 2)) for i in range(self.naq)]
        self.w("__host__ __device__ constexpr double spec_z2(int i) { return " + chain(z2, "i", "0.0") + "; }")

    def gen_activity_w(self, w: int) -> None:
        c = self.c
        self.w(f"template <> __device__ __forceinline__ void specw_activity<{w}>(double I, CellW &s) {{")
        need = set()
        for i in range(self.naq):
            if self.owner[i] == w and self.pri_cls[i] >= 0:
                need.add(self.pri_cls[i])
        for k in range(self.ncx):
            if self.cx_owner[k] == w and self.cx_cls[k] >= 0:
                need.add(self.cx_cls[k])
        self.w("  const double sq = sqrt(I);")
        A, B, Bd = _lit(c.debyeA), _lit(c.debyeB), _lit(c.debyeBdot)
        for q in sorted(need):
            negz2, a0 = self.cls[q]
            self.w(f"  s.lgcls[{q}] = ({_lit(negz2)} * sq * {A} / (1.0 + {_lit(a0)} * {B} * sq) + {Bd} * I) * SPEC_LN;")
        for i in range(self.naq):
            if self.owner[i] == w:
                q = self.pri_cls[i]
                self.w(f"  s.lngam[{i}] = {'0.0' if q < 0 else f's.lgcls[{q}]'};")
        self.w("}")
        self.w()

    def gen_complexes_w(self, w: int) -> None:
        a = self.a
        self.w(f"template <> __device__ __forceinline__ void specw_complexes<{w}>(CellW &s, const double *W,")
        self.w("    double *sec_out, long long ld, bool store) {")
        self.w("  double Is = 0.0;")
        if self.ncx:
            ptr, ids, st = a["eqcplx_ptr"], a["eqcplx_specid"], a["eqcplx_stoich"]
            for k in range(self.ncx):
                if self.cx_owner[k] != w:
                    continue
                expr = self.lnqk(a["eqcplx_logK"][k], float(a["eqcplx_h2ostoich"][k]), ptr, ids, st, k)
                q = self.cx_cls[k]
                arg = f"({expr})" if q < 0 else f"({expr}) - s.lgcls[{q}]"
                self.w("  {")
                self.w(f"    const double sk = exp({arg});")
                self.w(f"    if (store) sec_out[{k} * ld] = sk;")
                z2 = float(a["eqcplx_Z"][k]) ** 2
                if z2 != 0.0:
                    self.w(f"    Is += sk * {_lit(z2)};")
                self.w("  }")
        self.w("  s.Is_part = Is;")
        self.w("}")
        self.w()

    def gen_rows_w(self, w: int) -> None:
        a = self.a
        own = [i for i in self.coupled if self.owner[i] == w]
        self.w(f"template <> __device__ __forceinline__ void specw_rows<{w}>(const CellW &s, double *W,")
        self.w("    const double *sec_in, long long ld, double dt, double (&tot)[SPEC_N]) {")
        self.w("  const double denL = s.den_kg * 1.e-3;")
        self.w("  const double psvd = s.por * s.sat * 1000.0 * s.vol / dt;")
        for i in own:
            self.w(f"  tot[{i}] = SW(SW_OFF_C + {self.cpos[i]});")
        hits = {}
        if self.ncx:
            ptr, ids, st = a["eqcplx_ptr"], a["eqcplx_specid"], a["eqcplx_stoich"]
            for k in range(self.ncx):
                sp = range(ptr[k], ptr[k + 1])
                for p2 in sp:
                    for p in sp:
                        if int(ids[p]) in own:
                            e = (int(ids[p]), int(ids[p2]))
                            hits[e] = hits.get(e, 0) + 1
        hot = [e for e in sorted(hits, key=lambda e: -hits[e])[:8] if hits[e] >= 4]
        for (i, j) in hot:
            self.w(f"  double jh_{i}_{j} = {'1.0' if i == j else '0.0'};")
        written = set()
        if self.ncx:
            for k in range(self.ncx):
                sp = list(range(ptr[k], ptr[k + 1]))
                mine = [p for p in sp if int(ids[p]) in own]
                if not mine:
                    continue
                self.w("  {")
                self.w(f"    const double sk = __ldcg(sec_in + {k} * ld);")
                for p in mine:
                    i, s_i = int(ids[p]), float(st[p])
                    if s_i == 1.0:
                        self.w(f"    tot[{i}] += sk;")
                    elif s_i == -1.0:
                        self.w(f"    tot[{i}] -= sk;")
                    else:
                        self.w(f"    tot[{i}] += {_lit(s_i)} * sk;")
                for p2 in sp:
                    j, s_j = int(ids[p2]), float(st[p2])
                    tj = f"(sk * {self.ic(j)})" if s_j == 1.0 else f"(({_lit(s_j)} * sk) * {self.ic(j)})"
                    self.w(f"    {{ const double t = {tj};")
                    for p in mine:
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
        for i in own:
            self.w(f"  tot[{i}] *= denL;")
        for i in own:
            for j in self.coupled:
                e = self.J(i, j)
                if (i, j) in hot:
                    self.w(f"  {e} = (jh_{i}_{j} * denL) * psvd;")
                elif (i, j) in written:
                    self.w(f"  {e} = ({e} * denL) * psvd;")
                elif i == j:
                    self.w(f"  {e} = (1.0 * denL) * psvd;")
                else:
                    self.w(f"  {e} = 0.0;")
        if own:
            self.w("  if (s.dry) {")
            for i in own:
                ci = self.cpos[i]
                self.w("#pragma unroll 1")
                self.w(f"    for (int j = 0; j < SPEC_NC; j++) W[JX({ci}, j)] = (j == {ci}) ? 1.0 : 0.0;")
            self.w("  }")
        self.w("}")
        self.w()

    def gen_eval(self) -> None:
        c, a = self.c, self.a
        self.w("__device__ __noinline__ void specw_eval(const double *W, const DevState &st, long long cell,")
        self.w("    double den_kg, double por, double vol, double spd, double temp, double ln_act_h2o, bool apply,")
        self.w("    EvalOut *out) {")
        self.w("  struct { double ln_act_h2o; } s = {ln_act_h2o};")
        for e in range(c.neqsrfcplxrxn):
            r = int(a["eqsrfcplxrxn_to_srfcplxrxn"][e])
            cx = [int(v) for v in a["srfcplxrxn_to_complex"][a["srfcplxrxn_ptr"][r]:a["srfcplxrxn_ptr"][r + 1]]]
            ty = int(a["srfcplxrxn_surf_type"][r])
            dens = _lit(float(a["srfcplxrxn_site_density"][r]))
            ptr, ids, st_ = a["srfcplx_ptr"], a["srfcplx_specid"], a["srfcplx_stoich"]
            species = sorted({int(ids[p]) for k in cx for p in range(ptr[k], ptr[k + 1])})
            self.w("  {")
            if ty == chem.MINERAL_SURFACE:
                self.w(f"    const double dens = {dens} * __ldcg(st.mnrl_volfrac + {int(a['srfcplxrxn_to_surf'][r])} * st.ld + cell);")
            elif ty == chem.ROCK_SURFACE:
                self.w(f"    const double dens = {dens} * spd * (1.0 - por);")
            else:
                self.w(f"    const double dens = {dens};")
            self.w("    if (dens < 1.e-40) {")
            self.w(f"      out->fsite[{r}] = 0.0;")
            for k in cx:
                self.w(f"      out->S[{k}] = 0.0; out->nuis[{k}] = 0.0;")
            for i in species:
                self.w(f"      out->dsx[{r} * SPEC_NC + {self.cpos[i]}] = 0.0;")
            self.w("    } else {")
            for q, k in enumerate(cx):
                expr = self.lnqk(a["srfcplx_logK"][k], float(a["srfcplx_h2ostoich"][k]), ptr, ids, st_, k)
                self.w(f"      const double e{q} = exp({expr});")
            self.w("      double esum = 0.0;")
            for q in range(len(cx)):
                self.w(f"      esum += e{q};")
            self.w("      const double fs = dens / (1.0 + esum);")
            self.w(f"      out->fsite[{r}] = fs;")
            for q, k in enumerate(cx):
                self.w(f"      const double S{q} = e{q} * fs;")
                self.w(f"      out->S[{k}] = S{q}; out->nuis[{k}] = S{q} / fs;")
            self.w("      double den = 0.0;")
            for q in range(len(cx)):
                self.w(f"      den += S{q};")
            self.w("      den = den / fs + 1.0;")
            for i in species:
                self.w(f"      double tmp{i} = 0.0;")
            for q, k in enumerate(cx):
                for p in range(ptr[k], ptr[k + 1]):
                    i, nu = int(ids[p]), float(st_[p])
                    v = f"S{q}" if nu == 1.0 else f"{_lit(nu)} * S{q}"
                    self.w(f"      tmp{i} += {v};")
            for i in species:
                self.w(f"      out->dsx[{r} * SPEC_NC + {self.cpos[i]}] = (-tmp{i} / den) * {self.ic(i)};")
            self.w("    }")
            self.w("  }")
        for m in range(c.nkinmnrl):
            ptr, ids, st_ = a["kinmnrl_ptr"], a["kinmnrl_specid"], a["kinmnrl_stoich"]
            expr = self.lnqk(a["kinmnrl_logK"][m], float(a["kinmnrl_h2ostoich"][m]), ptr, ids, st_, m)
            thr = float(a["kinmnrl_affinity_threshold"][m])
            lim = float(a["kinmnrl_rate_limiter"][m])
            eact = float(a["kinmnrl_activation_energy"][m])
            irr = int(a["kinmnrl_irreversible"][m])
            rate = _lit(float(a["kinmnrl_rate_constant"][m]))
            self.w("  {")
            self.w(f"    const double QK = exp({expr});")
            self.w("    double aff = 1.0 - QK;")
            self.w("    const double sgn = copysign(1.0, aff);")
            self.w(f"    bool active = (__ldcg(st.mnrl_volfrac + {m} * st.ld + cell) > 0.0 || sgn < 0.0);")
            if irr == 1:
                self.w("    if (sgn < 0.0) active = false;")
            if thr > 0.0:
                self.w(f"    if (sgn < 0.0 && QK < {_lit(thr)}) active = false;")
            self.w("    double rate_vol = 0.0, Imv = 0.0, dfac = 0.0;")
            self.w("    if (active) {")
            if lim > 0.0:
                self.w(f"      aff = aff / (1.0 + (1.0 - aff) / {_lit(lim)});")
            if eact > 0.0:
                self.w(f"      const double spr = {rate} * exp({_lit(eact)} / 8.31446 * "
                       "(1.0 / (25.0 + 273.15) - 1.0 / (temp + 273.15)));")
            else:
                self.w(f"      const double spr = {rate} * 1.0;")
            self.w(f"      double Im_const = -st.mnrl_area[{m} * st.ld + cell];")
            self.w("      double Im = Im_const * sgn * fabs(aff) * spr;")
            self.w("      rate_vol = Im;")
            self.w("      if (apply) {")
            self.w("        Im_const = Im_const * vol;")
            self.w("        Imv = Im * vol;")
            self.w("        const double dIm_dQK = -Im_const * spr;")
            if lim > 0.0:
                self.w(f"        const double den = 1.0 + (1.0 - aff) / {_lit(lim)};")
                self.w(f"        dfac = dIm_dQK * (1.0 + QK / {_lit(lim)} / den) * QK * (den_kg * 1.e-3) / den;")
            else:
                self.w("        dfac = dIm_dQK * QK * (den_kg * 1.e-3);")
            self.w("      }")
            self.w("    }")
            self.w(f"    out->Im[{m}] = Imv; out->dfac[{m}] = dfac; out->mrate[{m}] = rate_vol;")
            self.w("  }")
        self.w("}")
        self.w()

    def gen_apply_w(self, w: int) -> None:
        c, a = self.c, self.a
        self.w(f"template <> __device__ __forceinline__ void specw_apply<{w}>(const CellW &s, const EvalOut &e,")
        self.w("    double *W, double jscale, double (&ts)[SPEC_N], double (&res)[SPEC_N], bool minerals) {")
        self.w("  if (!minerals) {")
        for eq in range(c.neqsrfcplxrxn):
            r = int(a["eqsrfcplxrxn_to_srfcplxrxn"][eq])
            cx = [int(v) for v in a["srfcplxrxn_to_complex"][a["srfcplxrxn_ptr"][r]:a["srfcplxrxn_ptr"][r + 1]]]
            ptr, ids, st_ = a["srfcplx_ptr"], a["srfcplx_specid"], a["srfcplx_stoich"]
            for k in cx:
                sp = list(range(ptr[k], ptr[k + 1]))
                mine = [p for p in sp if self.owner[int(ids[p])] == w]
                if not mine:
                    continue
                for p in mine:
                    i, nu = int(ids[p]), float(st_[p])
                    v = f"e.S[{k}]" if nu == 1.0 else f"{_lit(nu)} * e.S[{k}]"
                    self.w(f"    ts[{i}] += {v};")
                for p2 in sp:
                    j, nu_j = int(ids[p2]), float(st_[p2])
                    a1 = (f"e.S[{k}] * {self.ic(j)}" if nu_j == 1.0 else f"{_lit(nu_j)} * e.S[{k}] * {self.ic(j)}")
                    self.w(f"    {{ const double t = {a1} + e.nuis[{k}] * e.dsx[{r} * SPEC_NC + {self.cpos[j]}];")
                    for p in mine:
                        i, nu_i = int(ids[p]), float(st_[p])
                        v = "t" if nu_i == 1.0 else f"({_lit(nu_i)} * t)"
                        self.w(f"      {self.J(i, j)} += jscale * {v};")
                    self.w("    }")
        self.w("  } else {")
        for m in range(c.nkinmnrl):
            ptr, ids, st_ = a["kinmnrl_ptr"], a["kinmnrl_specid"], a["kinmnrl_stoich"]
            sp = list(range(ptr[m], ptr[m + 1]))
            mine = [p for p in sp if self.owner[int(ids[p])] == w]
            if not mine:
                continue
            for p in mine:
                i, nu = int(ids[p]), float(st_[p])
                self.w(f"    res[{i}] +={' ' if nu == 1.0 else f' {_lit(nu)} *'} e.Im[{m}];")
            for p2 in sp:
                j, nu_j = int(ids[p2]), float(st_[p2])
                self.w(f"    {{ const double t = e.dfac[{m}] * ({_lit(nu_j)} * {self.ic(j)});")
                for p in mine:
                    i, nu_i = int(ids[p]), float(st_[p])
                    v = "t" if nu_i == 1.0 else f"{_lit(nu_i)} * t"
                    self.w(f"      {self.J(i, j)} += {v};")
                self.w("    }")
        self.w("  }")
        self.w("}")
        self.w()

    def source(self) -> str:
        c, n, W = self.c, self.n, self.W
        slots = self.nc * (self.nc + 1) + 2 * self.nc + 3 * W
        per_block = slots * 32 * 8 + 1024
        groups = max(1, min(16 // W, (227 * 1024) // per_block))  # <= 16 warps/SM keeps 128 registers per thread
        self.threads, self.minblocks, self.slots = 32 * W, groups, slots
        self.out.clear()
        self.w("// generated by pflotran_elm_interface_b200/specialize.py -- do not edit")
        for k, v in (("SPEC_N", n), ("SPEC_NAQ", self.naq), ("SPEC_NC", self.nc), ("SPEC_NCX", self.ncx),
                     ("SPEC_NCLS", len(self.cls)), ("SPEC_NKIN", c.nkinmnrl), ("SPEC_NSRFRXN", c.nsrfcplxrxn),
                     ("SPEC_NSRFCPLX", c.nsrfcplx), ("SPEC_NEQSR", c.neqsrfcplxrxn),
                     ("SPEC_USE_LOG", int(c.use_log_formulation)), ("SPEC_ACT_UPD", int(self.act_upd)),
                     ("SPEC_USE_ACT_H2O", int(c.use_activity_h2o)), ("SPEC_W", W), ("SPEC_MINBLOCKS", groups)):
            self.w(f"#define {k} {v}")
        self.w(f"#define SPEC_SIG {signature(self.cfg)}ull")
        self.gen_maps()
        self.w("__device__ __forceinline__ double spec_cx_z2(int k);")
        self.w("__device__ __forceinline__ int spec_cx_cls(int k);")
        self.w("__device__ __forceinline__ double spec_mn_vol(int m);")
        self.w('#include "pfrx_specw.cuh"')
        self.w()
        self.gen_tables()
        self.gen_eval()
        for w in range(W):
            self.gen_activity_w(w)
            self.gen_complexes_w(w)
            self.gen_rows_w(w)
            self.gen_apply_w(w)
        self.w('#include "pfrx_specw_kernel.cuh"')
        return "\n".join(self.out) + "\n"


class _GenR(_GenW):
    """rolled variant (csrc/pfrx_specr.cuh): the network as __constant__ tables, sizes
    as macros, one instruction stream for the SPEC_W warps of a group"""

    def table(self, ctype: str, name: str, vals, fmt) -> None:
        vals = list(vals)
        body = ", ".join(fmt(v) for v in vals) if vals else ("0" if ctype == "int" else "0.0")
        self.w(f"__constant__ {ctype} {name}[{max(1, len(vals))}] = {{{body}}};")

    def source(self) -> str:
        c, a, n, W = self.c, self.a, self.n, self.W
        I = lambda v: str(int(v))
        D = lambda v: _lit(float(v))
        slots = self.nc * (self.nc + 1) + 2 * self.nc + 3 * W
        per_block = slots * 32 * 8 + 1024
        groups = max(1, min(16 // W, (227 * 1024) // per_block))
        self.threads, self.minblocks, self.slots = 32 * W, groups, slots
        nr = c.nsrfcplxrxn
        maxq = 1
        if nr:
            rp = a["srfcplxrxn_ptr"]
            maxq = max(int(rp[r + 1] - rp[r]) for r in range(nr))
            cxs = [int(v) for v in a["srfcplxrxn_to_complex"]]
            if len(set(cxs)) != len(cxs):
                raise ValueError("a surface complex that belongs to two reactions")
        self.out.clear()
        self.w("// generated by pflotran_elm_interface_b200/specialize.py -- do not edit")
        for k, v in (("SPEC_N", n), ("SPEC_NAQ", self.naq), ("SPEC_NC", self.nc), ("SPEC_NCX", self.ncx),
                     ("SPEC_NCLS", len(self.cls)), ("SPEC_NKIN", c.nkinmnrl), ("SPEC_NSRFRXN", nr),
                     ("SPEC_NSRFCPLX", c.nsrfcplx), ("SPEC_NEQSR", c.neqsrfcplxrxn), ("SPEC_MAXQ", maxq),
                     ("SPEC_USE_LOG", int(c.use_log_formulation)), ("SPEC_ACT_UPD", int(self.act_upd)),
                     ("SPEC_W", W), ("SPEC_MINBLOCKS", groups), ("SPEC_DEBYE_A", _lit(c.debyeA)),
                     ("SPEC_DEBYE_B", _lit(c.debyeB)), ("SPEC_DEBYE_BDOT", _lit(c.debyeBdot))):
            self.w(f"#define {k} {v}")
        self.w(f"#define SPEC_SIG {signature(self.cfg)}ull")
        self.w("#include <cuda_runtime.h>")
        cp = self.cpos
        self.table("int", "T_cmap", [cp.get(i, -1) for i in range(n)], I)
        self.table("double", "T_z2", [float(a["primary_spec_Z"][i]) ** 2 if i < self.naq else 0.0 for i in range(n)], D)
        self.table("int", "T_pcls", [self.pri_cls[i] if i < self.naq else -1 for i in range(n)], I)
        self.table("double", "T_cls_negz2", [q[0] for q in self.cls], D)
        self.table("double", "T_cls_a0", [q[1] for q in self.cls], D)
        # secondary complexes and their transpose (species -> complexes, ascending k)
        if self.ncx:
            ptr, ids, st = a["eqcplx_ptr"], a["eqcplx_specid"], a["eqcplx_stoich"]
        else:
            ptr, ids, st = [0], [], []
        self.table("int", "T_cx_ptr", ptr, I)
        self.table("int", "T_cx_id", [cp[int(v)] for v in ids], I)
        self.table("double", "T_cx_nu", st, D)
        self.table("double", "T_cx_lnk", [-float(v) * LOG_TO_LN for v in (a["eqcplx_logK"] if self.ncx else [])], D)
        self.table("double", "T_cx_h2o", a["eqcplx_h2ostoich"] if self.ncx else [], D)
        self.table("double", "T_cx_z2", [float(v) ** 2 for v in (a["eqcplx_Z"] if self.ncx else [])], D)
        self.table("int", "T_cx_cls", self.cx_cls, I)
        sp_ptr, sp_cx, sp_nu = [0], [], []
        for ci, sp in enumerate(self.coupled):
            for k in range(self.ncx):
                for p in range(ptr[k], ptr[k + 1]):
                    if int(ids[p]) == sp:
                        sp_cx.append(k)
                        sp_nu.append(float(st[p]))
            sp_ptr.append(len(sp_cx))
        self.table("int", "T_sp_ptr", sp_ptr, I)
        self.table("int", "T_sp_cx", sp_cx, I)
        self.table("double", "T_sp_nu", sp_nu, D)
        # kinetic minerals
        nk = c.nkinmnrl
        if nk:
            mptr, mids, mst = a["kinmnrl_ptr"], a["kinmnrl_specid"], a["kinmnrl_stoich"]
        else:
            mptr, mids, mst = [0], [], []
        self.table("int", "T_mn_ptr", mptr, I)
        self.table("int", "T_mn_id", [cp[int(v)] for v in mids], I)
        self.table("int", "T_mn_sp", mids, I)
        self.table("double", "T_mn_nu", mst, D)
        self.table("double", "T_mn_lnk", [-float(v) * LOG_TO_LN for v in (a["kinmnrl_logK"] if nk else [])], D)
        for nm, key in (("T_mn_h2o", "kinmnrl_h2ostoich"), ("T_mn_vol", "kinmnrl_molar_vol"),
                        ("T_mn_rate", "kinmnrl_rate_constant"), ("T_mn_eact", "kinmnrl_activation_energy"),
                        ("T_mn_thr", "kinmnrl_affinity_threshold"), ("T_mn_lim", "kinmnrl_rate_limiter")):
            self.table("double", nm, a[key] if nk else [], D)
        self.table("int", "T_mn_irr", a["kinmnrl_irreversible"] if nk else [], I)
        # equilibrium surface complexation
        if nr:
            kinds = {chem.MINERAL_SURFACE: 1, chem.ROCK_SURFACE: 2}
            self.table("int", "T_sr_ptr", a["srfcplxrxn_ptr"], I)
            self.table("int", "T_sr_cx", a["srfcplxrxn_to_complex"], I)
            self.table("int", "T_sr_type", [kinds.get(int(v), 0) for v in a["srfcplxrxn_surf_type"]], I)
            self.table("int", "T_sr_surf", [max(0, int(v)) for v in a["srfcplxrxn_to_surf"]], I)
            self.table("double", "T_sr_dens", a["srfcplxrxn_site_density"], D)
            sptr, sids, sst = a["srfcplx_ptr"], a["srfcplx_specid"], a["srfcplx_stoich"]
            self.table("int", "T_sc_ptr", sptr, I)
            self.table("int", "T_sc_id", [cp[int(v)] for v in sids], I)
            self.table("int", "T_sc_sp", sids, I)
            self.table("double", "T_sc_nu", sst, D)
            self.table("double", "T_sc_lnk", [-float(v) * LOG_TO_LN for v in a["srfcplx_logK"]], D)
            self.table("double", "T_sc_h2o", a["srfcplx_h2ostoich"], D)
            self.table("int", "T_eq", a["eqsrfcplxrxn_to_srfcplxrxn"], I)
            dnu = [0.0] * (nr * maxq * self.nc)
            rp = a["srfcplxrxn_ptr"]
            for r in range(nr):
                for qq, k in enumerate(a["srfcplxrxn_to_complex"][rp[r]:rp[r + 1]]):
                    for p in range(sptr[k], sptr[k + 1]):
                        dnu[(r * maxq + qq) * self.nc + cp[int(sids[p])]] = float(sst[p])
            self.table("double", "T_sr_dnu", dnu, D)
        else:
            for nm in ("T_sr_ptr", "T_sr_cx", "T_sr_type", "T_sr_surf", "T_sc_ptr", "T_sc_id", "T_sc_sp", "T_eq"):
                self.table("int", nm, [0, 0] if nm.endswith("ptr") else [], I)
            for nm in ("T_sr_dens", "T_sc_nu", "T_sc_lnk", "T_sc_h2o", "T_sr_dnu"):
                self.table("double", nm, [], D)
        self.w('#include "pfrx_specr.cuh"')
        return "\n".join(self.out) + "\n"


def supported_multiwarp(cfg: abi.ReactionConfig, warps: int) -> Tuple[bool, str]:
    ok, why = supported(cfg)
    if not ok:
        return ok, why
    c = cfg.c
    used = set()
    for ids in ("eqcplx_specid", "kinmnrl_specid", "srfcplx_specid"):
        if ids in cfg.arrays:
            used.update(int(v) for v in cfg.arrays[ids])
    if warps not in (2, 4, 8):
        return False, "2, 4 or 8 warps"
    if c.clmcn_nrxn > 0:
        return False, "reaction sandbox"
    if len(used) < 2 * warps:
        return False, "too few coupled species for that many warps"
    if not c.use_log_formulation:
        return False, "linear update needs a global minimum ratio"
    if c.act_coef_update_frequency != chem.ACT_COEF_FREQUENCY_NEWTON_ITER:
        return False, "frozen activity coefficients of the complexes would need NCX more slots"
    if c.use_activity_h2o:
        return False, "activity of water"
    return True, ""


def default_variant(cfg: abi.ReactionConfig) -> Tuple[int, str]:
    """(warps per 32 cells, style).  Measured on B200 with the Hanford 15/88 network
    (profiles/r01_spec_variants.md): straight-line code with one warp per 32 cells
    135 ms, rolled tables with four warps 396 ms, straight-line with four warps
    635 ms per 4.19 M cells -- so the first is the default for every network; the
    other two stay selectable for experiments:
    PFRX_SPEC_VARIANT=<style><warps> with style s (straight), r (rolled) or l (straight-line
    assembly, dense solve as rolled loops; one warp only)."""
    env = os.environ.get("PFRX_SPEC_VARIANT")
    if env:
        return int(env[1:]), VARIANT_STYLES[env[0]]
    return 1, "straight"


def generate_source(cfg: abi.ReactionConfig, warps: Optional[int] = None, style: Optional[str] = None) -> str:
    warps, style = _variant(cfg, warps, style)
    if style == "rolled":
        return _GenR(cfg, warps).source()
    if warps > 1:
        return _GenW(cfg, warps).source()
    g = _Gen(cfg)
    g.loop_lu = style in ("looplu", "klooplu")
    g.lockstep = style in ("lockstep", "klooplu")
    return g.source()


def _stamp(src: str) -> str:
    import hashlib

    h = hashlib.sha1(src.encode())
    for d in ("pfrx_fastmath.cuh", "pfrx_spec.cuh", "pfrx_specw.cuh", "pfrx_specw_kernel.cuh", "pfrx_specr.cuh", "pfrx_types.cuh"):
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
