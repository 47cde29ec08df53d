"""Host-side chemistry setup: reaction database (.dat), the CHEMISTRY and
CONSTRAINT blocks of a PFLOTRAN input deck, and the basis set-up that turns
them into the flat tables of ``include/pfrx.h``.

This is the part of the reference that "stays host" (SURVEY.md section 2 row
16): it runs once per simulation and never touches the GPU.  It mirrors

* ``DatabaseRead``            src/pflotran/reaction_database.F90:26-446
* ``BasisInit``               src/pflotran/reaction_database.F90:812-3710
* ``ReactionReadPass1``       src/pflotran/reaction.F90:121-936
* ``SurfaceComplexationRead`` src/pflotran/reaction_surf_complex.F90:28-420
* ``CLM_CN_Read/Map``         src/pflotran/reaction_sandbox_clm_cn.F90:98-465

only as far as the decks of the hot-path configurations need (SURVEY.md
section 8): primary/secondary/immobile species, kinetic minerals, equilibrium
and multirate surface complexation, the CLM-CN sandbox.  Everything is plain
numpy; no oracle code is imported here.
"""
from __future__ import annotations

import math
import re
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

# pflotran_constants.F90:84 (truncated in the reference; keep it truncated)
LOG_TO_LN = 2.30258509299

ACT_COEF_FREQUENCY_OFF = 0
ACT_COEF_FREQUENCY_TIMESTEP = 1
ACT_COEF_FREQUENCY_NEWTON_ITER = 2
ACT_COEF_ALGORITHM_LAG = 3
ACT_COEF_ALGORITHM_NEWTON = 4

NULL_SURFACE = 0
ROCK_SURFACE = 1
MINERAL_SURFACE = 2

MAX_PREFACTORS = 10
MAX_PREFACTOR_SPECIES = 5


# --------------------------------------------------------------------------- #
# input-deck reading helpers (input_aux.F90 semantics, as far as needed)
# --------------------------------------------------------------------------- #
def _fnum(tok: str) -> float:
    """Fortran-style real: 1.d-5, 1.e-5, 1.5D0."""
    return float(tok.replace("d", "e").replace("D", "e"))


def _is_num(tok: str) -> bool:
    try:
        _fnum(tok)
        return True
    except ValueError:
        return False


def deck_lines(text: str) -> List[List[str]]:
    """Tokenised, comment-free deck lines with skip/noskip and '\\'
    continuation handled (InputReadPflotranString, input_aux.F90)."""
    out: List[List[str]] = []
    skip = 0
    pending = ""
    for raw in text.splitlines():
        line = raw
        for cc in ("#", "!"):
            k = line.find(cc)
            if k >= 0:
                line = line[:k]
        line = line.strip()
        if not line:
            continue
        low = line.lower()
        if low.startswith("noskip"):
            skip = max(0, skip - 1)
            continue
        if low.startswith("skip"):
            skip += 1
            continue
        if skip:
            continue
        if line.endswith("\\"):
            pending += line[:-1] + " "
            continue
        line = pending + line
        pending = ""
        out.append(line.replace(",", " ").split())
    return out


def _is_end(tokens: Sequence[str]) -> bool:
    t = tokens[0].upper()
    return t == "/" or t == "END" or t.startswith("END_")


class _Cursor:
    def __init__(self, lines: List[List[str]], pos: int = 0):
        self.lines = lines
        self.pos = pos

    def next(self) -> Optional[List[str]]:
        if self.pos >= len(self.lines):
            return None
        t = self.lines[self.pos]
        self.pos += 1
        return t

    def block(self):
        """iterate lines until the block terminator"""
        while True:
            t = self.next()
            if t is None or _is_end(t):
                return
            yield t

    def skip_block(self):
        """skip a block that may contain nested sub-blocks we do not know
        about: only used for OUTPUT-like leaf blocks"""
        for _ in self.block():
            pass


_TIME_UNITS = {
    "s": 1.0, "sec": 1.0, "second": 1.0,
    "min": 60.0, "h": 3600.0, "hr": 3600.0, "hour": 3600.0,
    "d": 86400.0, "day": 86400.0,
    "w": 7 * 86400.0, "week": 7 * 86400.0,
    "mo": 365.0 / 12.0 * 86400.0,
    "y": 365.0 * 86400.0, "yr": 365.0 * 86400.0, "year": 365.0 * 86400.0,
}


def time_to_sec(value: float, unit: str) -> float:
    """units.F90 UnitsConvertToInternal for times (y = 365 d)."""
    return value * _TIME_UNITS[unit.lower()]


# --------------------------------------------------------------------------- #
# reaction database
# --------------------------------------------------------------------------- #
@dataclass
class DbRxn:
    names: List[str]
    stoich: List[float]
    logK: List[float]


@dataclass
class DbAqSpecies:
    name: str
    a0: float
    Z: float
    mw: float
    rxn: Optional[DbRxn] = None


@dataclass
class DbGas:
    name: str
    molar_volume: float
    rxn: DbRxn
    mw: float


@dataclass
class DbMineral:
    name: str
    molar_volume: float  # m^3/mol
    rxn: DbRxn
    mw: float


@dataclass
class DbSrfCplx:
    name: str
    free_site_name: str
    free_site_stoich: float
    rxn: DbRxn
    Z: float


_QTOK = re.compile(r"'([^']*)'|(\S+)")


def _db_tokens(line: str) -> List[str]:
    return [a if a != "" or b == "" else b for a, b in _QTOK.findall(line)]


class Database:
    """Parsed reaction database (reaction_database.F90:125-446): line 1
    ``'temperature points' N T1..TN``, then five sections separated by
    ``'null'`` lines: primary, secondary, gas, mineral, surface complex."""

    def __init__(self, text: str):
        self.temperatures: List[float] = []
        self.primary: Dict[str, DbAqSpecies] = {}
        self.secondary: Dict[str, DbAqSpecies] = {}
        self.gas: Dict[str, DbGas] = {}
        self.mineral: Dict[str, DbMineral] = {}
        self.srfcplx: Dict[str, DbSrfCplx] = {}
        self.raw: Dict[Tuple[int, str], str] = {}  # (section, name) -> line
        self.header = ""
        self.null_lines: List[str] = []
        self._parse(text)

    @classmethod
    def from_file(cls, path: str) -> "Database":
        with open(path, "r") as f:
            return cls(f.read())

    def _parse(self, text: str) -> None:
        lines = [ln for ln in text.splitlines() if ln.strip() and not ln.lstrip().startswith(("#", "!"))]
        self.header = lines[0]
        tok = _db_tokens(lines[0])
        n = int(tok[1])
        self.temperatures = [_fnum(t) for t in tok[2:2 + n]]
        nulls = 0
        for ln in lines[1:]:
            tok = _db_tokens(ln)
            name = tok[0]
            if name == "null":
                nulls += 1
                self.null_lines.append(ln)
                if nulls >= 5:
                    break
                continue
            self.raw[(nulls, name)] = ln
            if nulls == 0:
                self.primary[name] = DbAqSpecies(name, _fnum(tok[1]), _fnum(tok[2]), _fnum(tok[3]))
            elif nulls == 1:
                ns = int(tok[1])
                p = 2
                names, st = [], []
                for _ in range(ns):
                    st.append(_fnum(tok[p]))
                    names.append(tok[p + 1])
                    p += 2
                logK = [_fnum(t) for t in tok[p:p + n]]
                p += n
                self.secondary[name] = DbAqSpecies(name, _fnum(tok[p]), _fnum(tok[p + 1]), _fnum(tok[p + 2]),
                                                   DbRxn(names, st, logK))
            elif nulls in (2, 3):
                vm = _fnum(tok[1]) * 1.0e-6  # cm^3/mol -> m^3/mol
                ns = int(tok[2])
                p = 3
                names, st = [], []
                for _ in range(ns):
                    st.append(_fnum(tok[p]))
                    names.append(tok[p + 1])
                    p += 2
                logK = [_fnum(t) for t in tok[p:p + n]]
                p += n
                mw = _fnum(tok[p])
                if nulls == 2:
                    self.gas[name] = DbGas(name, vm, DbRxn(names, st, logK), mw)
                else:
                    self.mineral[name] = DbMineral(name, vm, DbRxn(names, st, logK), mw)
            elif nulls == 4:
                ns = int(tok[1])
                p = 2
                names, st = [], []
                fs_name, fs_st = "", 0.0
                for _ in range(ns):
                    s, nm = _fnum(tok[p]), tok[p + 1]
                    p += 2
                    if nm.startswith(">"):
                        fs_name, fs_st = nm, s
                    else:
                        st.append(s)
                        names.append(nm)
                logK = [_fnum(t) for t in tok[p:p + n]]
                p += n
                self.srfcplx[name] = DbSrfCplx(name, fs_name, fs_st, DbRxn(names, st, logK), _fnum(tok[p]))

    def subset_text(self, names: Sequence[str]) -> str:
        """A database holding only ``names`` (fixture extraction: the GPU box
        has no /root/reference, so tests carry trimmed copies)."""
        want = set(names)
        out = [self.header]
        for sec in range(5):
            for (s, nm), ln in self.raw.items():
                if s == sec and nm in want:
                    out.append(ln)
            out.append(self.null_lines[sec] if sec < len(self.null_lines) else "'null' 0 0 0")
        return "\n".join(out) + "\n"


# --------------------------------------------------------------------------- #
# deck -> chemistry description
# --------------------------------------------------------------------------- #
@dataclass
class MineralKinetics:
    name: str
    rate_constant: float = 0.0            # mol/m^2/s
    activation_energy: float = 0.0
    affinity_threshold: float = 0.0
    rate_limiter: float = 0.0
    irreversible: int = 0
    affinity_power: Optional[float] = None
    temkin: Optional[float] = None
    min_scale_factor: Optional[float] = None
    prefactors: List[dict] = field(default_factory=list)


@dataclass
class SrfCplxRxn:
    itype: str = "EQUILIBRIUM"            # or MULTIRATE_KINETIC
    surface_type: int = NULL_SURFACE
    surface_name: str = ""
    free_site_name: str = ""
    site_density: float = 0.0
    complexes: List[str] = field(default_factory=list)
    rates: List[float] = field(default_factory=list)
    site_fractions: List[float] = field(default_factory=list)
    kinmr_scale_factor: float = 1.0


@dataclass
class ClmCnSandbox:
    pools: List[Tuple[str, Optional[float]]] = field(default_factory=list)  # (name, mass C:N or None)
    reactions: List[dict] = field(default_factory=list)


@dataclass
class Chemistry:
    primary: List[str] = field(default_factory=list)
    secondary: List[str] = field(default_factory=list)
    immobile: List[str] = field(default_factory=list)
    gases: List[str] = field(default_factory=list)
    decoupled: List[str] = field(default_factory=list)
    minerals: List[str] = field(default_factory=list)
    mineral_kinetics: List[MineralKinetics] = field(default_factory=list)
    srfcplx_rxns: List[SrfCplxRxn] = field(default_factory=list)
    clm_cn: Optional[ClmCnSandbox] = None
    database: str = ""
    use_log_formulation: bool = False
    act_coef_update_frequency: int = ACT_COEF_FREQUENCY_OFF
    act_coef_update_algorithm: int = ACT_COEF_ALGORITHM_LAG
    act_coef_use_bdot: bool = True
    use_activity_h2o: bool = False
    initialize_with_molality: bool = False
    use_total_as_guess: bool = False
    max_dlnC: float = 5.0
    max_dlnC_rreact: float = 5.0
    max_relative_change_tolerance: float = 1.0e-6
    max_residual_tolerance: float = 1.0e-12
    max_rel_residual_tolerance: float = 1.0e-8
    maximum_reaction_iterations: int = 20
    maximum_reaction_cuts: int = 10
    unsupported: List[str] = field(default_factory=list)


_RATE_UNITS = {  # -> mol/m^2-sec
    "mol/m^2-sec": 1.0, "mol/m^2-s": 1.0,
    "mol/cm^2-sec": 1.0e4, "mol/cm^2-s": 1.0e4,
    "mol/dm^2-sec": 1.0e2, "mol/dm^2-s": 1.0e2,
}

_AREA_UNITS = {  # -> m^2/m^3
    "m^2/m^3": 1.0, "cm^2/cm^3": 100.0, "dm^2/dm^3": 10.0, "m^2/g": None,
}


def _read_names(cur: _Cursor) -> List[str]:
    return [t[0] for t in cur.block()]


def _read_mineral_kinetics(cur: _Cursor) -> List[MineralKinetics]:
    out = []
    for t in cur.block():
        mk = MineralKinetics(t[0])
        for u in cur.block():
            key = u[0].upper()
            if key == "RATE_CONSTANT":
                r = _fnum(u[1])
                if r < 0.0:
                    r = 10.0 ** r
                if len(u) > 2:
                    r = r * _RATE_UNITS[u[2].lower()]
                mk.rate_constant = r
            elif key == "ACTIVATION_ENERGY":
                mk.activation_energy = _fnum(u[1])
            elif key == "AFFINITY_THRESHOLD":
                mk.affinity_threshold = _fnum(u[1])
            elif key == "AFFINITY_POWER":
                mk.affinity_power = _fnum(u[1])
            elif key in ("TEMKIN_CONSTANT", "TEMPKINS_CONSTANT"):
                mk.temkin = _fnum(u[1])
            elif key in ("MINERAL_SCALE_FACTOR",):
                mk.min_scale_factor = _fnum(u[1])
            elif key == "RATE_LIMITER":
                mk.rate_limiter = _fnum(u[1])
            elif key == "IRREVERSIBLE":
                mk.irreversible = 1
            elif key == "PREFACTOR":
                pf = {"rate": 0.0, "activation_energy": 0.0, "species": []}
                for v in cur.block():
                    k2 = v[0].upper()
                    if k2 == "RATE_CONSTANT":
                        r = _fnum(v[1])
                        if r < 0.0:
                            r = 10.0 ** r
                        if len(v) > 2:
                            r = r * _RATE_UNITS[v[2].lower()]
                        pf["rate"] = r
                    elif k2 == "ACTIVATION_ENERGY":
                        pf["activation_energy"] = _fnum(v[1])
                    elif k2 == "PREFACTOR_SPECIES":
                        sp = {"name": v[1], "alpha": 0.0, "beta": 0.0, "atten": 0.0}
                        for w in cur.block():
                            k3 = w[0].upper()
                            if k3 == "ALPHA":
                                sp["alpha"] = _fnum(w[1])
                            elif k3 == "BETA":
                                sp["beta"] = _fnum(w[1])
                            elif k3 == "ATTENUATION_COEF":
                                sp["atten"] = _fnum(w[1])
                        pf["species"].append(sp)
                mk.prefactors.append(pf)
            else:
                raise ValueError(f"MINERAL_KINETICS keyword {key} not supported")
        out.append(mk)
    return out


def _read_float_array(first: List[str]) -> List[float]:
    return [_fnum(x) for x in first if _is_num(x)]


def _read_srfcplx_rxn(cur: _Cursor) -> SrfCplxRxn:
    rx = SrfCplxRxn()
    for t in cur.block():
        key = t[0].upper()
        if key == "EQUILIBRIUM":
            rx.itype = "EQUILIBRIUM"
        elif key == "MULTIRATE_KINETIC":
            rx.itype = "MULTIRATE_KINETIC"
        elif key in ("RATE", "RATES"):
            rx.itype = "MULTIRATE_KINETIC"
            rx.rates = _read_float_array(t[1:])
        elif key == "SITE_FRACTION":
            rx.site_fractions = _read_float_array(t[1:])
        elif key == "MULTIRATE_SCALE_FACTOR":
            rx.kinmr_scale_factor = _fnum(t[1])
        elif key == "MINERAL":
            rx.surface_type = MINERAL_SURFACE
            rx.surface_name = t[1]
        elif key == "ROCK_DENSITY":
            rx.surface_type = ROCK_SURFACE
        elif key == "SITE":
            rx.free_site_name = t[1]
            rx.site_density = _fnum(t[2])
        elif key == "COMPLEXES":
            rx.complexes = _read_names(cur)
        else:
            raise ValueError(f"SURFACE_COMPLEXATION_RXN keyword {key} not supported")
    if rx.itype == "MULTIRATE_KINETIC":
        if not rx.site_fractions and rx.rates:
            rx.site_fractions = [1.0 / float(len(rx.rates))] * len(rx.rates)
        rx.rates = [r * rx.kinmr_scale_factor for r in rx.rates]
    return rx


def _read_clm_cn(cur: _Cursor) -> ClmCnSandbox:
    # reaction_sandbox_clm_cn.F90:98-289
    CN_ratio_mass_to_mol = 1.16616
    sb = ClmCnSandbox()
    for t in cur.block():
        key = t[0].upper()
        if key == "POOLS":
            for u in cur.block():
                if len(u) > 1 and _is_num(u[1]):
                    sb.pools.append((u[0], _fnum(u[1]) * CN_ratio_mass_to_mol))
                else:
                    sb.pools.append((u[0], None))
        elif key == "REACTION":
            rx = {"up": "", "down": "", "rate_constant": 0.0, "turnover": 0.0, "resp": -999.0, "inhib": 0.0}
            for u in cur.block():
                k2 = u[0].upper()
                if k2 == "UPSTREAM_POOL":
                    rx["up"] = u[1]
                elif k2 == "DOWNSTREAM_POOL":
                    rx["down"] = u[1]
                elif k2 == "RATE_CONSTANT":
                    rx["rate_constant"] = _fnum(u[1])
                    if len(u) > 2:
                        rx["rate_constant"] /= time_to_sec(1.0, u[2].split("/")[-1])
                elif k2 == "TURNOVER_TIME":
                    rx["turnover"] = time_to_sec(_fnum(u[1]), u[2]) if len(u) > 2 else _fnum(u[1])
                elif k2 == "RESPIRATION_FRACTION":
                    rx["resp"] = _fnum(u[1])
                elif k2 == "N_INHIBITION":
                    rx["inhib"] = _fnum(u[1])
                else:
                    raise ValueError(f"CLM-CN REACTION keyword {k2}")
            if rx["turnover"] > 0.0:
                rx["rate_constant"] = 1.0 / rx["turnover"]
            sb.reactions.append(rx)
        else:
            raise ValueError(f"CLM-CN keyword {key}")
    return sb


def read_chemistry(cur: _Cursor) -> Chemistry:
    """CHEMISTRY block (ReactionReadPass1, reaction.F90:121-936)."""
    ch = Chemistry()
    for t in cur.block():
        key = t[0].upper()
        if key == "PRIMARY_SPECIES":
            ch.primary = _read_names(cur)
        elif key == "SECONDARY_SPECIES":
            ch.secondary = _read_names(cur)
        elif key == "IMMOBILE_SPECIES":
            ch.immobile = _read_names(cur)
        elif key in ("GAS_SPECIES", "PASSIVE_GAS_SPECIES", "ACTIVE_GAS_SPECIES"):
            if key == "ACTIVE_GAS_SPECIES":
                ch.unsupported.append(key)
            ch.gases += [n for n in _read_names(cur) if n.upper() != "GAS_TRANSPORT_IS_UNVETTED"]
        elif key == "DECOUPLED_EQUILIBRIUM_REACTIONS":
            ch.decoupled = _read_names(cur)
        elif key == "MINERALS":
            ch.minerals = _read_names(cur)
        elif key == "MINERAL_KINETICS":
            ch.mineral_kinetics = _read_mineral_kinetics(cur)
        elif key == "SORPTION":
            for u in cur.block():
                k2 = u[0].upper()
                if k2 == "SURFACE_COMPLEXATION_RXN":
                    ch.srfcplx_rxns.append(_read_srfcplx_rxn(cur))
                else:
                    ch.unsupported.append("SORPTION," + k2)
                    _skip_nested(cur)
        elif key == "REACTION_SANDBOX":
            for u in cur.block():
                k2 = u[0].upper()
                if k2 == "CLM-CN":
                    ch.clm_cn = _read_clm_cn(cur)
                else:
                    ch.unsupported.append("REACTION_SANDBOX," + k2)
                    _skip_nested(cur)
        elif key == "DATABASE":
            ch.database = t[1]
        elif key == "LOG_FORMULATION":
            ch.use_log_formulation = True
        elif key == "ACTIVITY_COEFFICIENTS":
            ch.act_coef_update_algorithm = ACT_COEF_ALGORITHM_LAG
            ch.act_coef_update_frequency = ACT_COEF_FREQUENCY_TIMESTEP
            for w in t[1:]:
                w = w.upper()
                if w == "OFF":
                    ch.act_coef_update_frequency = ACT_COEF_FREQUENCY_OFF
                elif w == "LAG":
                    ch.act_coef_update_algorithm = ACT_COEF_ALGORITHM_LAG
                elif w == "NEWTON":
                    ch.act_coef_update_algorithm = ACT_COEF_ALGORITHM_NEWTON
                elif w == "TIMESTEP":
                    ch.act_coef_update_frequency = ACT_COEF_FREQUENCY_TIMESTEP
                elif w == "NEWTON_ITERATION":
                    ch.act_coef_update_frequency = ACT_COEF_FREQUENCY_NEWTON_ITER
        elif key == "NO_BDOT":
            ch.act_coef_use_bdot = False
        elif key == "ACTIVITY_H2O" or key == "ACTIVITY_WATER":
            ch.use_activity_h2o = True
        elif key == "MOLAL" or key == "MOLALITY":
            ch.initialize_with_molality = True
        elif key == "USE_TOTAL_CONCENTRATION_AS_GUESS":
            ch.use_total_as_guess = True
        elif key == "MAX_DLNC":
            ch.max_dlnC = _fnum(t[1])
        elif key == "MAX_DLNC_RREACT":
            ch.max_dlnC_rreact = _fnum(t[1])
        elif key == "MAX_RELATIVE_CHANGE_TOLERANCE":
            ch.max_relative_change_tolerance = _fnum(t[1])
        elif key == "MAX_RESIDUAL_TOLERANCE":
            ch.max_residual_tolerance = _fnum(t[1])
        elif key == "MAXIMUM_REACTION_ITERATIONS":
            ch.maximum_reaction_iterations = int(t[1])
        elif key == "MAXIMUM_REACTION_CUTS":
            ch.maximum_reaction_cuts = int(t[1])
        elif key == "OUTPUT":
            cur.skip_block()
        elif key in ("USE_FULL_GEOCHEMISTRY", "NO_CHECKPOINT_ACT_COEFS", "NO_CHECK_UPDATE",
                     "DONT_STOP_ON_RREACT_FAILURE", "UPDATE_MINERAL_SURFACE_AREA", "UPDATE_POROSITY"):
            pass
        else:
            ch.unsupported.append(key)
    return ch


def _skip_nested(cur: _Cursor) -> None:
    """skip an unknown block; sub-blocks are recognised by a lone trailing
    keyword line followed by their own terminator -- good enough for the
    SORPTION sub-blocks we do not handle"""
    depth = 1
    while depth > 0:
        t = cur.next()
        if t is None:
            return
        if _is_end(t):
            depth -= 1
        elif len(t) == 1 and t[0].isupper() and t[0] in ("CATIONS", "COMPLEXES", "ISOTHERM_REACTIONS"):
            depth += 1


@dataclass
class Constraint:
    name: str
    conc: List[Tuple[str, float, str, str]] = field(default_factory=list)  # (species, value, type, aux name)
    minerals: Dict[str, Tuple[float, float]] = field(default_factory=dict)  # name -> (vol frac, area m^2/m^3)
    immobile: Dict[str, float] = field(default_factory=dict)
    free_site_guess: Dict[str, float] = field(default_factory=dict)


def read_constraint(cur: _Cursor, name: str) -> Constraint:
    """CONSTRAINT block (transport_constraint_rt.F90)."""
    cn = Constraint(name)
    for t in cur.block():
        key = t[0].upper()
        if key in ("CONCENTRATIONS", "CONC"):
            for u in cur.block():
                typ = u[2].upper() if len(u) > 2 else "T"
                aux = u[3] if len(u) > 3 else ""
                cn.conc.append((u[0], _fnum(u[1]), typ, aux))
        elif key in ("MINERALS", "MNRL"):
            for u in cur.block():
                vf = _fnum(u[1])
                area = _fnum(u[2])
                unit = u[3].lower() if len(u) > 3 else "m^2/m^3"
                fac = _AREA_UNITS.get(unit)
                if fac is None:
                    raise ValueError(f"mineral area unit {unit} not supported")
                cn.minerals[u[0]] = (vf, area * fac)
        elif key == "IMMOBILE":
            for u in cur.block():
                cn.immobile[u[0]] = _fnum(u[1])
        elif key == "FREE_ION_GUESS":
            cur.skip_block()
        else:
            cur.skip_block()
    return cn


@dataclass
class Deck:
    chemistry: Optional[Chemistry] = None
    constraints: Dict[str, Constraint] = field(default_factory=dict)
    porosity: List[float] = field(default_factory=list)
    rock_density: List[float] = field(default_factory=list)
    reference_liquid_density: Optional[float] = None
    reference_temperature: float = 25.0
    final_time: float = 0.0
    initial_dt: float = 1.0
    maximum_dt: float = 1.0e20
    ts_acceleration: int = 5
    newton: Dict[str, float] = field(default_factory=dict)
    osrt: bool = False
    max_steps: Optional[int] = None


def read_deck(text: str) -> Deck:
    lines = deck_lines(text)
    cur = _Cursor(lines)
    dk = Deck()
    in_transport_nm = False
    while True:
        t = cur.next()
        if t is None:
            break
        key = t[0].upper()
        if key == "CHEMISTRY":
            dk.chemistry = read_chemistry(cur)
        elif key == "CONSTRAINT" and len(t) > 1:
            dk.constraints[t[1]] = read_constraint(cur, t[1])
        elif key == "POROSITY" and len(t) > 1 and _is_num(t[1]):
            dk.porosity.append(_fnum(t[1]))
        elif key == "ROCK_DENSITY" and len(t) > 1 and _is_num(t[1]):
            dk.rock_density.append(_fnum(t[1]))
        elif key == "REFERENCE_LIQUID_DENSITY":
            dk.reference_liquid_density = _fnum(t[1])
        elif key == "REFERENCE_TEMPERATURE":
            dk.reference_temperature = _fnum(t[1])
        elif key == "MODE" and len(t) > 1 and t[1].upper() == "OSRT":
            dk.osrt = True
        elif key == "FINAL_TIME":
            dk.final_time = time_to_sec(_fnum(t[1]), t[2])
        elif key == "INITIAL_TIMESTEP_SIZE":
            dk.initial_dt = time_to_sec(_fnum(t[1]), t[2])
        elif key == "MAXIMUM_TIMESTEP_SIZE" and len(t) == 3:
            dk.maximum_dt = time_to_sec(_fnum(t[1]), t[2])
        elif key == "NUMERICAL_METHODS":
            in_transport_nm = len(t) > 1 and t[1].upper() == "TRANSPORT"
        elif key == "MAX_STEPS" and in_transport_nm:
            dk.max_steps = int(t[1])
        elif key == "TS_ACCELERATION" and in_transport_nm:
            dk.ts_acceleration = int(t[1])
        elif key in ("ATOL", "RTOL", "STOL", "MAXIMUM_NUMBER_OF_ITERATIONS", "MAXIT") and in_transport_nm:
            dk.newton[key] = _fnum(t[1])
    return dk


# --------------------------------------------------------------------------- #
# basis set-up -> flat tables
# --------------------------------------------------------------------------- #
def _interpolate(x_high, x_low, x, y_high, y_low):
    """utility.F90:889-913 Interpolate, same arithmetic"""
    x_diff = x_high - x_low
    if abs(x_diff) < 1.0e-10:
        return y_low
    weight = (x - x_low) / x_diff
    return y_low + weight * (y_high - y_low)


def debye_huckel_constants(tref: float, use_bdot: bool = True):
    """reaction_database.F90:931-1023"""
    table = [  # T, A, B, Bdot
        (0.0, 0.4939, 0.3253, 0.0374), (25.0, 0.5114, 0.3288, 0.0410), (60.0, 0.5465, 0.3346, 0.0440),
        (100.0, 0.5995, 0.3421, 0.0460), (150.0, 0.6855, 0.3525, 0.0470), (200.0, 0.7994, 0.3639, 0.0470),
        (250.0, 0.9593, 0.3766, 0.0340), (300.0, 1.2180, 0.3925, 0.0000), (350.0, 1.2180, 0.3925, 0.0000),
    ]
    if tref <= 0.01:
        A, B, Bd = table[0][1:]
    elif tref > 350.0:
        A, B, Bd = table[-1][1:]
    else:
        for lo, hi in zip(table[:-1], table[1:]):
            if lo[0] < tref <= hi[0] or (lo[0] == 0.0 and 0.0 < tref <= hi[0]):
                A = _interpolate(hi[0], lo[0], tref, hi[1], lo[1])
                B = _interpolate(hi[0], lo[0], tref, hi[2], lo[2])
                Bd = _interpolate(hi[0], lo[0], tref, hi[3], lo[3])
                break
    if not use_bdot:
        Bd = 0.0
    return A, B, Bd


def fit_logK_coefs(temps: Sequence[float], logK: Sequence[float]) -> np.ndarray:
    """ReactionFitLogKCoef (reaction_aux.F90:1159-1230): least squares on the
    basis {ln T, 1, T, 1/T, 1/T^2}, skipping logK = 500 entries."""
    tk = np.asarray(temps, dtype=np.float64) + 273.15
    vec = np.stack([np.log(tk), np.ones_like(tk), tk, 1.0 / tk, 1.0 / (tk * tk)])
    lk = np.asarray(logK, dtype=np.float64)
    ok = np.abs(lk - 500.0) >= 1.0e-10
    rhs = (vec[:, ok] * lk[ok]).sum(axis=1)
    a = vec[:, ok] @ vec[:, ok].T
    return np.linalg.solve(a, rhs)


@dataclass
class Rxn:
    ids: List[int]            # primary ids, 0-based, ascending (deck order)
    stoich: List[float]
    h2o_stoich: float
    logK_T: List[float]       # per database temperature


class ReactionNetwork:
    """The flattened ``reaction_rt_type`` subset of include/pfrx.h."""

    def __init__(self, chem: Chemistry, db: Database, reference_temperature: float = 25.0,
                 use_isothermal: bool = True):
        self.chem = chem
        self.db = db
        self.tref = reference_temperature
        self.use_isothermal = use_isothermal
        self.primary_names = list(chem.primary)
        self.secondary_names = list(chem.secondary)
        self.immobile_names = list(chem.immobile)
        self.naqcomp = len(self.primary_names)
        self.nimcomp = len(self.immobile_names)
        self.ncomp = self.naqcomp + self.nimcomp
        self._basis()
        self._minerals()
        self._surface_complexation()
        self._clm_cn()

    # -- temperature handling (reaction_database.F90:1025-1050) ------------- #
    def _itemp(self):
        T = self.db.temperatures
        tr = self.tref
        if tr <= T[0]:
            return 0, 0
        if tr > T[-1]:
            return len(T) - 1, len(T) - 1
        for i in range(len(T) - 1):
            if T[i] < tr <= T[i + 1]:
                return i, i + 1
        return 0, 0

    def logK_at_tref(self, logK_T: Sequence[float]) -> float:
        lo, hi = self._itemp()
        T = self.db.temperatures
        return _interpolate(T[hi], T[lo], self.tref, logK_T[hi], logK_T[lo])

    # -- aqueous basis (reaction_database.F90:1060-1460, 1690-1800) -------- #
    def _basis(self):
        db, chem = self.db, self.chem
        pri = self.primary_names
        sec = self.secondary_names
        gas = list(chem.gases)
        self.primary_Z = np.zeros(self.naqcomp)
        self.primary_a0 = np.zeros(self.naqcomp)
        self.primary_mw = np.zeros(self.naqcomp)
        pri_rxn: List[Optional[DbRxn]] = []
        for i, nm in enumerate(pri):
            if nm in db.primary:
                s = db.primary[nm]
                pri_rxn.append(None)
            elif nm in db.secondary:
                s = db.secondary[nm]
                pri_rxn.append(None if nm in chem.decoupled else s.rxn)
            else:
                raise KeyError(f"primary species {nm} not found in database")
            self.primary_Z[i], self.primary_a0[i], self.primary_mw[i] = s.Z, s.a0, s.mw
        nT = len(db.temperatures)
        pri_names = ["H2O"] + pri                      # column 0 is water
        col = {n: i for i, n in enumerate(pri_names)}
        sec_like = sec + gas
        scol = {n: i for i, n in enumerate(sec_like)}
        rows: List[Tuple[str, DbRxn]] = []
        for nm, rx in zip(pri, pri_rxn):
            if rx is not None:
                rows.append((nm, rx))
        for nm in sec:
            if nm not in db.secondary:
                raise KeyError(f"secondary species {nm} not found in database")
            rows.append((nm, db.secondary[nm].rxn))
        for nm in gas:
            if nm not in db.gas:
                raise KeyError(f"gas species {nm} not found in database")
            rows.append((nm, db.gas[nm].rxn))
        ns = len(sec_like)
        if len(rows) != ns:
            raise ValueError("number of database reactions does not match number of secondary species + gases "
                             f"({len(rows)} vs {ns}); see reaction_database.F90:1133-1175")
        pri_matrix = np.zeros((ns, len(pri_names)))
        sec_matrix = np.zeros((ns, ns))
        logKvec = np.zeros((nT, ns))
        for r, (nm, rx) in enumerate(rows):
            logKvec[:, r] = rx.logK
            if nm in col:
                pri_matrix[r, col[nm]] = -1.0
            else:
                sec_matrix[r, scol[nm]] = -1.0
            for sn, st in zip(rx.names, rx.stoich):
                if sn in col:
                    pri_matrix[r, col[sn]] = st
                elif sn in scol:
                    sec_matrix[r, scol[sn]] = st
                else:
                    raise KeyError(f"species {sn} in reaction of {nm} is neither primary nor secondary")
        if ns:
            identity_like = np.array_equal(sec_matrix, -np.eye(ns))
            if identity_like:
                stoich_matrix = pri_matrix.copy()
                logK_sw = logKvec.copy()
            else:
                inv = np.linalg.inv(sec_matrix)
                stoich_matrix = -1.0 * (inv @ pri_matrix)
                logK_sw = -(inv @ logKvec.T).T
        self.sec_rxn: List[Rxn] = []
        self.gas_rxn: Dict[str, Rxn] = {}
        self._sec_full: Dict[str, Tuple[Dict[str, float], np.ndarray]] = {}
        for r, nm in enumerate(sec_like):
            ids, st, h2o = [], [], 0.0
            full: Dict[str, float] = {}
            for c in range(len(pri_names)):
                v = stoich_matrix[r, c]
                if abs(v) > 1.0e-40:
                    full[pri_names[c]] = v
                    if c == 0:
                        h2o = v
                    else:
                        ids.append(c - 1)
                        st.append(v)
            rx = Rxn(ids, st, h2o, list(logK_sw[:, r]))
            self._sec_full[nm] = (full, logK_sw[:, r].copy())
            if r < len(sec):
                self.sec_rxn.append(rx)
            else:
                self.gas_rxn[nm] = rx
        self.neqcplx = len(sec)
        self.eqcplx_Z = np.array([db.secondary[n].Z for n in sec], dtype=np.float64)
        self.eqcplx_a0 = np.array([db.secondary[n].a0 for n in sec], dtype=np.float64)
        self.eqcplx_mw = np.array([db.secondary[n].mw for n in sec], dtype=np.float64)
        self.debyeA, self.debyeB, self.debyeBdot = debye_huckel_constants(self.tref, chem.act_coef_use_bdot)

    def _to_basis(self, names: Sequence[str], stoich: Sequence[float], logK: Sequence[float]) -> Rxn:
        """substitute secondary/gas species, then align to the basis order
        (BasisSubSpeciesIn*Rxn + BasisAlignSpeciesInRxn,
        reaction_database_aux.F90:320-560)"""
        acc: Dict[str, float] = {}
        lk = np.asarray(logK, dtype=np.float64).copy()
        for nm, st in zip(names, stoich):
            if nm == "H2O" or nm in self.primary_names:
                acc[nm] = acc.get(nm, 0.0) + st
            elif nm in self._sec_full:
                full, slk = self._sec_full[nm]
                for k, v in full.items():
                    acc[k] = acc.get(k, 0.0) + st * v
                lk = lk + st * slk
            else:
                raise KeyError(f"species {nm} not in basis")
        ids, sts = [], []
        for i, nm in enumerate(self.primary_names):
            v = acc.get(nm, 0.0)
            if abs(v) > 1.0e-10:
                ids.append(i)
                sts.append(v)
        h2o = acc.get("H2O", 0.0)
        if abs(h2o) <= 1.0e-10:
            h2o = 0.0
        return Rxn(ids, sts, h2o, list(lk))

    # -- minerals (reaction_database.F90:1960-2400) ------------------------- #
    def _minerals(self):
        chem, db = self.chem, self.db
        self.mineral_names = list(chem.minerals)
        self.mnrl_rxn: Dict[str, Rxn] = {}
        self.mnrl_molar_vol: Dict[str, float] = {}
        for nm in self.mineral_names:
            if nm not in db.mineral:
                raise KeyError(f"mineral {nm} not found in database")
            m = db.mineral[nm]
            self.mnrl_rxn[nm] = self._to_basis(m.rxn.names, m.rxn.stoich, m.rxn.logK)
            self.mnrl_molar_vol[nm] = m.molar_volume
        self.kinmnrl_names = [mk.name for mk in chem.mineral_kinetics]
        self.nkinmnrl = len(self.kinmnrl_names)
        self.kinmnrl = chem.mineral_kinetics

    # -- surface complexation (reaction_database.F90:2640-3100) ------------- #
    def _surface_complexation(self):
        chem, db = self.chem, self.db
        self.srfcplx_names: List[str] = []
        for rx in chem.srfcplx_rxns:
            for c in rx.complexes:
                if c not in self.srfcplx_names:
                    self.srfcplx_names.append(c)
        self.srfcplx_rxn: List[Rxn] = []
        self.srfcplx_free_site_stoich: List[float] = []
        self.srfcplx_Z: List[float] = []
        for nm in self.srfcplx_names:
            if nm not in db.srfcplx:
                raise KeyError(f"surface complex {nm} not found in database")
            s = db.srfcplx[nm]
            self.srfcplx_rxn.append(self._to_basis(s.rxn.names, s.rxn.stoich, s.rxn.logK))
            self.srfcplx_free_site_stoich.append(s.free_site_stoich)
            self.srfcplx_Z.append(s.Z)
        self.srfcplxrxn = chem.srfcplx_rxns
        self.eq_rxn_ids = [i for i, r in enumerate(self.srfcplxrxn) if r.itype == "EQUILIBRIUM"]
        self.mr_rxn_ids = [i for i, r in enumerate(self.srfcplxrxn) if r.itype == "MULTIRATE_KINETIC"]

    # -- CLM-CN (reaction_sandbox_clm_cn.F90:314-465) ------------------------ #
    def _clm_cn(self):
        sb = self.chem.clm_cn
        self.clmcn = None
        if sb is None:
            return
        imm = {n: i for i, n in enumerate(self.immobile_names)}
        pools = [p[0] for p in sb.pools]
        CN = np.array([(-999.0 if p[1] is None else p[1]) for p in sb.pools], dtype=np.float64)
        nspec = np.zeros(len(pools), dtype=np.int32)
        cid = np.zeros(len(pools), dtype=np.int32)
        nid = np.full(len(pools), -1, dtype=np.int32)
        for i, (nm, ratio) in enumerate(sb.pools):
            if ratio is None:
                cid[i], nid[i], nspec[i] = imm[nm + "C"], imm[nm + "N"], 2
            else:
                cid[i], nspec[i] = imm[nm], 1
        up = np.array([pools.index(r["up"]) for r in sb.reactions], dtype=np.int32)
        down = np.array([(pools.index(r["down"]) if r["down"] else -1) for r in sb.reactions], dtype=np.int32)
        for d in down:
            if d >= 0 and CN[d] < 0.0:
                raise ValueError("CLM-CN downstream pools must have a constant C:N ratio")
        self.clmcn = dict(
            nrxn=len(sb.reactions), npool=len(pools), C_id=imm["C"], N_id=imm["N"], CN_ratio=CN,
            pool_nspec=nspec, pool_C_id=cid, pool_N_id=nid, up=up, down=down,
            rate_constant=np.array([r["rate_constant"] for r in sb.reactions], dtype=np.float64),
            resp=np.array([r["resp"] for r in sb.reactions], dtype=np.float64),
            inhib=np.array([r["inhib"] for r in sb.reactions], dtype=np.float64),
        )

    # -- helpers -------------------------------------------------------------- #
    def csr(self, rxns: Sequence[Rxn]):
        ptr = np.zeros(len(rxns) + 1, dtype=np.int32)
        ids: List[int] = []
        st: List[float] = []
        for i, r in enumerate(rxns):
            ids += r.ids
            st += r.stoich
            ptr[i + 1] = len(ids)
        return ptr, np.array(ids, dtype=np.int32), np.array(st, dtype=np.float64)

    def logKs(self, rxns: Sequence[Rxn]) -> np.ndarray:
        return np.array([self.logK_at_tref(r.logK_T) for r in rxns], dtype=np.float64)

    def logK_coefs(self, rxns: Sequence[Rxn]) -> np.ndarray:
        if not rxns:
            return np.zeros((0, 5))
        return np.stack([fit_logK_coefs(self.db.temperatures, r.logK_T) for r in rxns])


def load_network(deck_text: str, db_text: str, use_isothermal: bool = True) -> Tuple[Deck, ReactionNetwork]:
    dk = read_deck(deck_text)
    if dk.chemistry is None:
        raise ValueError("deck has no CHEMISTRY block")
    net = ReactionNetwork(dk.chemistry, Database(db_text), dk.reference_temperature, use_isothermal)
    return dk, net
