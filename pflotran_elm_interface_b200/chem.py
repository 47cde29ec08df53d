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
class SomDecSandbox:
    """SOMDECOMP block as read (SomDecRead, reaction_sandbox_somdec.F90:350-830)"""
    pools: List[Tuple[str, Optional[float]]] = field(default_factory=list)   # (name, N:C mol or None)
    reactions: List[dict] = field(default_factory=list)
    abiotic: dict = field(default_factory=dict)
    o2_species: str = ""
    co2_species: str = ""
    x0eps: float = 1.0e-20
    inhibition_nh4_no3: float = 1.0
    n2o_frac_mineralization: float = 0.02


@dataclass
class Chemistry:
    primary: List[str] = field(default_factory=list)
    secondary: List[str] = field(default_factory=list)
    immobile: List[str] = field(default_factory=list)
    gases: List[str] = field(default_factory=list)
    active_gases: List[str] = field(default_factory=list)   # ACTIVE_GAS_SPECIES (RTotalGas)
    radon: Optional[dict] = None
    decoupled: List[str] = field(default_factory=list)
    minerals: List[str] = field(default_factory=list)
    mineral_kinetics: List[MineralKinetics] = field(default_factory=list)
    srfcplx_rxns: List[SrfCplxRxn] = field(default_factory=list)
    ionx_rxns: List[dict] = field(default_factory=list)
    isotherm_rxns: List[dict] = field(default_factory=list)
    dynamic_kd_rxns: List[dict] = field(default_factory=list)
    clm_cn: Optional[ClmCnSandbox] = None
    somdec: Optional[SomDecSandbox] = None
    nitrif: Optional[dict] = None
    denitr: Optional[dict] = None
    plantn: Optional[dict] = None
    langmuir: Optional[dict] = None
    cndegas: Optional[dict] = None
    calcite_sandbox: Optional[dict] = None
    sandbox_order: List[str] = field(default_factory=list)
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
    # GENERAL_REACTION / RADIOACTIVE_DECAY_REACTION / IMMOBILE_DECAY_REACTION blocks, deck order
    general_rxns: List[dict] = field(default_factory=list)       # {reaction, kf, kr}
    radiodecay_rxns: List[dict] = field(default_factory=list)    # {reaction, k}
    immobile_decay_rxns: List[dict] = field(default_factory=list)  # {species, k}
    microbial_rxns: List[dict] = field(default_factory=list)
    microbial_units: int = 0        # 0 = not set -> MOLARITY (reaction_microbial.F90:233)
    unsupported: List[str] = field(default_factory=list)


_RATE_UNITS = {  # -> mol/m^2-sec
    "mol/m^2-sec": 1.0, "mol/m^2-s": 1.0,
    "mol/cm^2-sec": 1.0e4, "mol/cm^2-s": 1.0e4,
    "mol/dm^2-sec": 1.0e2, "mol/dm^2-s": 1.0e2,
}

_AREA_PER_MASS_UNITS = {"m^2/kg": 1.0, "m^2/g": 1.0e3, "cm^2/g": 1.0e-1, "cm^2/kg": 1.0e-4}  # -> m^2/kg
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


def _read_ionx_rxn(cur: _Cursor) -> dict:
    # reaction.F90:625-730: the REFERENCE cation (k = 1) moves to the front of the list
    rx = {"mineral": "", "CEC": None, "cations": []}
    for t in cur.block():
        key = t[0].upper()
        if key == "MINERAL":
            rx["mineral"] = t[1]
        elif key == "CEC":
            rx["CEC"] = _fnum(t[1])
        elif key == "CATIONS":
            ref = None
            for u in cur.block():
                rx["cations"].append((u[0], _fnum(u[1])))
                if len(u) > 2 and u[2].upper() == "REFERENCE":
                    ref = u[0]
            if ref is None:
                raise ValueError("Reference cation missing in Ion Exchange reaction.")
            i = [n for n, _ in rx["cations"]].index(ref)
            if abs(rx["cations"][i][1] - 1.0) > 1.0e-40:
                raise ValueError(f'Reference cation "{ref}" must have k = 1.d0.')
            rx["cations"].insert(0, rx["cations"].pop(i))
        else:
            raise ValueError(f"ION_EXCHANGE_RXN keyword {key}")
    if rx["CEC"] is None:
        raise ValueError("A CEC must be defined for all ion exchange reactions.")
    return rx


def _read_isotherm_rxns(cur: _Cursor) -> List[dict]:
    # IsothermRead, reaction_isotherm.F90:30-228
    out = []
    for t in cur.block():
        rx = {"species": t[0], "type": 1, "kd": None, "kd_units": "", "langmuir_b": 0.0, "freundlich_n": 0.0,
              "mineral": ""}
        for u in cur.block():
            key = u[0].upper()
            if key == "TYPE":
                rx["type"] = {"LINEAR": 1, "LANGMUIR": 2, "FREUNDLICH": 3}[u[1].upper()]
            elif key in ("DISTRIBUTION_COEFFICIENT", "KD"):
                rx["kd"] = _fnum(u[1])
                if len(u) > 2:
                    rx["kd_units"] = u[2]
            elif key == "LANGMUIR_B":
                rx["langmuir_b"] = _fnum(u[1])
                rx["type"] = 2
            elif key == "FREUNDLICH_N":
                rx["freundlich_n"] = _fnum(u[1])
                rx["type"] = 3
            elif key == "KD_MINERAL_NAME":
                rx["mineral"] = u[1]
            else:
                raise ValueError(f"ISOTHERM_REACTIONS keyword {key}")
        if rx["kd"] is None:
            raise ValueError("DISTRIBUTION_COEFFICIENT missing in ISOTHERM_REACTIONS")
        out.append(rx)
    return out


def _read_dynamic_kd_rxns(cur: _Cursor) -> List[dict]:
    # reaction.F90:540-612
    out = []
    for t in cur.block():
        rx = {"species": t[0], "ref": "", "ref_high": 0.0, "low": 0.0, "high": 0.0, "power": 0.0}
        keys = {"REFERENCE_SPECIES_HIGH": "ref_high", "KD_LOW": "low", "KD_HIGH": "high", "KD_POWER": "power"}
        for u in cur.block():
            key = u[0].upper()
            if key == "REFERENCE_SPECIES":
                rx["ref"] = u[1]
            elif key in keys:
                rx[keys[key]] = _fnum(u[1])
            else:
                raise ValueError(f"DYNAMIC_KD_REACTIONS keyword {key}")
        out.append(rx)
    return out


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


# elm_rspfuncs.F90:17-47
TEMPERATURE_RESPONSE = {"OFF": 0, "CLMCN": 1, "Q10": 2, "DLEM": 3, "ARRHENIUS": 4}
MOISTURE_RESPONSE = {"OFF": 0, "CLMCN": 1, "DLEM": 2, "LOGTHETA": 3}
OX_RESPONSE = {"OFF": 0, "MONOD": 1, "WFPS": 2}
INHIBITION_TYPE = {"THRESHOLD": 1, "MONOD": 3, "INVERSE_MONOD": 4}
CN_RATIO_MASS_TO_MOL = 1.166156023644992  # elm_rspfuncs.F90:36


def _default_abiotic() -> dict:
    # AbioticFactorsCreate, reaction_sandbox_somdec.F90:239-262
    return {"temperature": 0, "moisture": 0, "q10": 1.5, "ea": 51.7, "ox": 0, "ox_half_saturation": 1.0e-15,
            "depth_efolding": 0.0}


def _read_abiotic_factors(cur: _Cursor, ab: dict) -> None:
    # SomDecRead_AbioticFactors, reaction_sandbox_somdec.F90:834-983
    for t in cur.block():
        key = t[0].upper()
        if key == "TEMPERATURE_RESPONSE_FUNCTION":
            for u in cur.block():
                k2 = u[0].upper()
                ab["temperature"] = TEMPERATURE_RESPONSE.get(k2, 0)
                if k2 in ("DLEM", "Q10"):
                    ab["q10"] = _fnum(u[1])
                elif k2 == "ARRHENIUS":
                    ab["ea"] = _fnum(u[1])
        elif key == "MOISTURE_RESPONSE_FUNCTION":
            for u in cur.block():
                ab["moisture"] = MOISTURE_RESPONSE.get(u[0].upper(), 0)
        elif key == "OX_RESPONSE_FUNCTION":
            for u in cur.block():
                k2 = u[0].upper()
                ab["ox"] = OX_RESPONSE.get(k2, 0)
                if k2 == "MONOD":
                    ab["ox_half_saturation"] = _fnum(u[1])
        elif key == "DECOMP_DEPTH_EFOLDING":
            ab["depth_efolding"] = _fnum(t[1])
        else:
            raise ValueError(f"SOMDECOMP ABIOTIC_FACTORS keyword {key}")


def _rate_per_sec(value: float, unit: Optional[str]) -> float:
    """'x unitless/d', '1/s' ... -> 1/s (UnitsConvertToInternal on the time part)"""
    if not unit:
        return value
    return value / time_to_sec(1.0, unit.split("/")[-1])


def _read_somdec(cur: _Cursor) -> SomDecSandbox:
    sb = SomDecSandbox()
    sb.abiotic = _default_abiotic()
    for t in cur.block():
        key = t[0].upper()
        if key == "O2_SPECIES_NAME":
            sb.o2_species = t[1]
        elif key == "CO2_SPECIES_NAME":
            sb.co2_species = t[1]
        elif key == "ABIOTIC_FACTORS":
            _read_abiotic_factors(cur, sb.abiotic)
        elif key == "X0EPS":
            sb.x0eps = _fnum(t[1])
        elif key == "AMMONIUM_INHIBITION_NITRATE":
            sb.inhibition_nh4_no3 = _fnum(t[1])
        elif key == "N2O_FRAC_MINERALIZATION":
            sb.n2o_frac_mineralization = _fnum(t[1])
        elif key == "POOLS":
            for u in cur.block():
                if len(u) > 1 and _is_num(u[1]) and _fnum(u[1]) > 0.0:
                    sb.pools.append((u[0], 1.0 / _fnum(u[1]) / CN_RATIO_MASS_TO_MOL))
                else:
                    sb.pools.append((u[0], None))
        elif key == "REACTION":
            rx = {"up": "", "down": [], "rate_constant": -1.0, "turnover": -1.0, "rate_decomposition": -1.0,
                  "rate_ad_factor": 1.0, "monod": [], "inhibition": [], "abiotic": None, "ox_species": "",
                  "cox_species": ""}
            for u in cur.block():
                k2 = u[0].upper()
                if k2 == "UPSTREAM_POOL":
                    rx["up"] = u[1]
                elif k2 == "DOWNSTREAM_POOL":
                    rx["down"].append((u[1], _fnum(u[2])))
                elif k2 == "RATE_CONSTANT":
                    rx["rate_constant"] = _rate_per_sec(_fnum(u[1]), u[2] if len(u) > 2 else None)
                elif k2 == "TURNOVER_TIME":
                    rx["turnover"] = time_to_sec(_fnum(u[1]), u[2]) if len(u) > 2 else _fnum(u[1])
                elif k2 == "RATE_DECOMPOSITION":
                    rx["rate_decomposition"] = _rate_per_sec(_fnum(u[1]), u[2] if len(u) > 2 else None)
                elif k2 == "RATE_AD_FACTOR":
                    rx["rate_ad_factor"] = _fnum(u[1])
                elif k2 == "MONOD":
                    m = {"species": "", "half_saturation": 1.0e-15, "threshold": 0.0, "pool_normalized": 0}
                    for v in cur.block():
                        k3 = v[0].upper()
                        if k3 == "SPECIES_NAME":
                            m["species"] = v[1]
                        elif k3 == "HALF_SATURATION_CONSTANT":
                            m["half_saturation"] = _fnum(v[1])
                        elif k3 == "THRESHOLD_CONCENTRATION":
                            m["threshold"] = _fnum(v[1])
                        elif k3 == "POOL_NORMALIZED":
                            m["pool_normalized"] = 1
                        else:
                            raise ValueError(f"SOMDECOMP MONOD keyword {k3}")
                    rx["monod"].append(m)
                elif k2 == "INHIBITION":
                    ih = {"species": "", "itype": -999, "constant": 1.0e-15, "constant2": 0.0}
                    for v in cur.block():
                        k3 = v[0].upper()
                        if k3 == "SPECIES_NAME":
                            ih["species"] = v[1]
                        elif k3 == "TYPE":
                            ih["itype"] = INHIBITION_TYPE[v[1].upper()]
                            if v[1].upper() == "THRESHOLD":
                                ih["constant2"] = _fnum(v[2])
                        elif k3 == "INHIBITION_CONSTANT":
                            ih["constant"] = _fnum(v[1])
                        else:
                            raise ValueError(f"SOMDECOMP INHIBITION keyword {k3}")
                    rx["inhibition"].append(ih)
                elif k2 == "ABIOTIC_FACTORS":
                    rx["abiotic"] = _default_abiotic()
                    _read_abiotic_factors(cur, rx["abiotic"])
                elif k2 == "OX_SPECIES_NAME":
                    rx["ox_species"] = u[1]
                elif k2 == "COX_SPECIES_NAME":
                    rx["cox_species"] = u[1]
                else:
                    raise ValueError(f"SOMDECOMP REACTION keyword {k2}")
            nset = sum(1 for k in ("turnover", "rate_constant", "rate_decomposition") if rx[k] > 0.0)
            if nset != 1:
                raise ValueError("exactly one of TURNOVER_TIME / RATE_CONSTANT / RATE_DECOMPOSITION per SOMDECOMP reaction")
            if rx["turnover"] > 0.0:
                rx["rate_constant"] = 1.0 / rx["turnover"]
            sb.reactions.append(rx)
        else:
            raise ValueError(f"SOMDECOMP keyword {key}")
    # reactions without their own block copy the sandbox-wide factors as they
    # stand when the reaction is read (reaction_sandbox_somdec.F90:783-788);
    # decks put ABIOTIC_FACTORS first, so the final values are the same
    for rx in sb.reactions:
        if rx["abiotic"] is None:
            rx["abiotic"] = dict(sb.abiotic)
    return sb


def _read_nitrif(cur: _Cursor) -> dict:
    # NitrifCreate/NitrifRead, reaction_sandbox_nitrif.F90:58-150
    d = {"k_nitr_max": 1.0e-6, "k_nitr_n2o": 3.5e-8, "x0eps": 1.0e-20}
    for t in cur.block():
        key = t[0].upper()
        if key == "TEMPERATURE_RESPONSE_FUNCTION":
            cur.skip_block()   # read by the reference but not used by NitrifReact
        elif key == "X0EPS":
            d["x0eps"] = _fnum(t[1])
        elif key == "NITRIFICATION_RATE_COEF":
            d["k_nitr_max"] = _fnum(t[1])
        elif key == "N2O_RATE_COEF_NITRIFICATION":
            d["k_nitr_n2o"] = _fnum(t[1])
        elif key == "AMMONIUM_HALF_SATURATION":
            pass               # stored, never used (reaction_sandbox_nitrif.F90:234-502)
        else:
            raise ValueError(f"NITRIFICATION keyword {key}")
    return d


def _read_denitr(cur: _Cursor) -> dict:
    # DenitrCreate/DenitrRead, reaction_sandbox_denitr.F90:38-150
    d = {"half_saturation": 1.0e-15, "k_deni_max": 2.5e-6, "x0eps": 1.0e-20}
    for t in cur.block():
        key = t[0].upper()
        if key == "TEMPERATURE_RESPONSE_FUNCTION":
            cur.skip_block()   # not used by DenitrReact
        elif key == "DENITRIFICATION_RATE_COEF":
            d["k_deni_max"] = _fnum(t[1])
        elif key == "NITRATE_HALF_SATURATION":
            d["half_saturation"] = _fnum(t[1])
        elif key == "X0EPS":
            d["x0eps"] = _fnum(t[1])
        else:
            raise ValueError(f"DENITRIFICATION keyword {key}")
    return d


def _read_plantn(cur: _Cursor) -> dict:
    # PlantNCreate/PlantNRead, reaction_sandbox_plantn.F90:40-150
    d = {"half_saturation_nh4": 1.0e-15, "half_saturation_no3": 1.0e-15, "inhibition_nh4_no3": 1.0,
         "x0eps_nh4": 1.0e-20, "x0eps_no3": 1.0e-20}
    keys = {"AMMONIUM_HALF_SATURATION": "half_saturation_nh4", "NITRATE_HALF_SATURATION": "half_saturation_no3",
            "AMMONIUM_INHIBITION_NITRATE": "inhibition_nh4_no3", "X0EPS_NH4": "x0eps_nh4", "X0EPS_NO3": "x0eps_no3"}
    for t in cur.block():
        key = t[0].upper()
        if key == "RATE_PLANTNDEMAND":
            pass      # read, then overwritten by PlantNReact (reaction_sandbox_plantn.F90:399-409)
        elif key in keys:
            d[keys[key]] = _fnum(t[1])
        else:
            raise ValueError(f"PLANTN keyword {key}")
    return d


def _read_langmuir(cur: _Cursor) -> dict:
    # LangmuirCreate/LangmuirRead, reaction_sandbox_langmu.F90:44-140
    d = {"name_aq": "", "name_sorb": "", "k_kinetic": 1.0e-5, "k_equilibrium": 2.5e3, "s_max": 1.0e-3}
    for t in cur.block():
        key = t[0].upper()
        if key == "NAME_AQ":
            d["name_aq"] = t[1]
        elif key == "NAME_SORB":
            d["name_sorb"] = t[1]
        elif key == "EQUILIBRIUM_CONSTANT":
            d["k_equilibrium"] = _fnum(t[1])
        elif key == "KINETIC_CONSTANT":
            d["k_kinetic"] = _fnum(t[1])
        elif key == "S_MAX":
            d["s_max"] = _fnum(t[1])
        else:
            raise ValueError(f"LANGMUIR keyword {key}")
    return d


def _read_radon(cur: _Cursor) -> dict:
    # RadonReadInput, reaction_sandbox_radon.F90:54-100
    d: Dict[str, object] = {}
    for t in cur.block():
        key = t[0].upper()
        if key == "SPECIES_NAME":
            d["species_name"] = t[1]
        elif key == "MINERAL_NAME":
            d["mineral_name"] = t[1]
        elif key == "RADON_GENERATION_RATE":
            d["radon_generation_rate"] = _fnum(t[1])
        else:
            raise ValueError(f"RADON sandbox keyword {key}")
    if len(d) != 3:
        raise ValueError("SPECIES_NAME, MINERAL_NAME and RADON_GENERATION_RATE must be set for REACTION_SANDBOX,RADON")
    return d


def _read_calcite_sandbox(cur: _Cursor) -> dict:
    # CalciteReadInput, reaction_sandbox_calcite.F90:52-104
    d: Dict[str, float] = {}
    for t in cur.block():
        key = t[0].upper()
        if key in ("RATE_CONSTANT1", "RATE_CONSTANT2"):
            d[key.lower()] = _fnum(t[1])
        else:
            raise ValueError(f"CALCITE sandbox keyword {key}")
    if len(d) != 2:
        raise ValueError("RATE_CONSTANT1 and RATE_CONSTANT2 must be set for REACTION_SANDBOX,CALCITE")
    return d


def _read_cndegas(cur: _Cursor) -> dict:
    # CNdegasCreate / CNdegasRead, reaction_sandbox_cndegas.F90:44-130
    d = {"k_kinetic_co2": 1.0e-5, "k_kinetic_n2o": 1.0e-5, "k_kinetic_n2": 1.0e-5, "k_kinetic_h": 1.0e-5,
         "fixph": 6.5, "fixph_on": 0}
    keys = {"KINETIC_CONSTANT_CO2": "k_kinetic_co2", "KINETIC_CONSTANT_N2O": "k_kinetic_n2o",
            "KINETIC_CONSTANT_N2": "k_kinetic_n2", "KINETIC_CONSTANT_H+": "k_kinetic_h"}
    for t in cur.block():
        key = t[0].upper()
        if key in keys:
            d[keys[key]] = _fnum(t[1])
        elif key == "FIXPH":
            d["fixph"] = _fnum(t[1])
            d["fixph_on"] = 1
        else:
            raise ValueError(f"CNDEGAS keyword {key}")
    return d


def _per_time(tokens, i) -> float:
    """value at tokens[i] with an optional 1/<time> unit behind it -> 1/s"""
    v = _fnum(tokens[i])
    if len(tokens) > i + 1 and not tokens[i + 1].startswith(("!", "#")):
        u = tokens[i + 1].lower()
        if not u.startswith("1/") or u[2:] not in _TIME_UNITS:
            raise ValueError(f"rate unit {tokens[i + 1]} not supported")
        v /= _TIME_UNITS[u[2:]]
    return v


def _read_kinetic_rxn_block(cur: "_Cursor", kind: str) -> dict:
    """GENERAL_REACTION (reaction.F90:357-460), RADIOACTIVE_DECAY_REACTION (:294-356) and
    IMMOBILE_DECAY_REACTION (reaction_immobile.F90:94-160) blocks"""
    r = {"reaction": "", "kf": 0.0, "kr": 0.0, "k": None, "species": ""}
    for t in cur.block():
        key = t[0].upper()
        if key == "REACTION":
            r["reaction"] = " ".join(x for x in t[1:])
        elif key == "SPECIES_NAME":
            r["species"] = t[1]
        elif key == "FORWARD_RATE":
            r["kf"] = _per_time(t, 1)  # kg^(n-1)/mol^(n-1)-sec; only 1/<time> units are converted here
        elif key == "BACKWARD_RATE":
            r["kr"] = _per_time(t, 1)
        elif key == "RATE_CONSTANT":
            r["k"] = _per_time(t, 1)
        elif key == "HALF_LIFE":
            r["k"] = -1.0 * math.log(0.5) / time_to_sec(_fnum(t[1]), t[2] if len(t) > 2 else "s")
        else:
            raise ValueError(f"{kind}: keyword {key} not supported")
    return r


MICROBIAL_UNITS = {"MOLALITY": 1, "ACTIVITY": 2, "MOLARITY": 3}
INHIBITION_TYPES = {"THRESHOLD": 1, "MONOD": 3, "INVERSE_MONOD": 4, "SMOOTHSTEP": 5}
_ENERGY_UNITS = {"j/mol": 1.0, "kj/mol": 1.0e3, "cal/mol": 4.184, "kcal/mol": 4184.0}


def _read_microbial_rxn(cur: "_Cursor", ch: "Chemistry") -> dict:
    """MICROBIAL_REACTION block (MicrobialRead, reaction_microbial.F90:40-282)"""
    r = {"reaction": "", "rate_constant": 0.0, "activation_energy": 0.0, "monod": [], "inhibition": [],
         "biomass": None, "yield": 0.0}
    for t in cur.block():
        key = t[0].upper()
        if key == "REACTION":
            r["reaction"] = " ".join(t[1:])
        elif key == "CONCENTRATION_UNITS":
            u = MICROBIAL_UNITS[t[1].upper()]
            if ch.microbial_units and ch.microbial_units != u:
                raise ValueError("Concentration units must be consistent for all microbial reactions")
            ch.microbial_units = u
        elif key == "RATE_CONSTANT":
            if len(t) > 2 and not t[2].startswith(("!", "#")):
                raise ValueError("MICROBIAL_REACTION RATE_CONSTANT with units is not supported")
            r["rate_constant"] = _fnum(t[1])
        elif key == "ACTIVATION_ENERGY":
            u = t[2].lower() if len(t) > 2 and not t[2].startswith(("!", "#")) else "j/mol"
            r["activation_energy"] = _fnum(t[1]) * _ENERGY_UNITS[u]
        elif key == "MONOD":
            m = {"species": "", "K": 0.0, "Cth": 0.0}
            for u in cur.block():
                k2 = u[0].upper()
                if k2 == "SPECIES_NAME":
                    m["species"] = u[1]
                elif k2 == "HALF_SATURATION_CONSTANT":
                    m["K"] = _fnum(u[1])
                elif k2 == "THRESHOLD_CONCENTRATION":
                    m["Cth"] = _fnum(u[1])
                else:
                    raise ValueError(f"MICROBIAL_REACTION,MONOD: keyword {k2} not supported")
            r["monod"].append(m)
        elif key == "INHIBITION":
            h = {"species": "", "type": 0, "C": 0.0, "C2": 0.0}
            for u in cur.block():
                k2 = u[0].upper()
                if k2 == "SPECIES_NAME":
                    h["species"] = u[1]
                elif k2 == "TYPE":
                    h["type"] = INHIBITION_TYPES[u[1].upper()]
                    if h["type"] in (1, 5):
                        h["C2"] = _fnum(u[2])
                elif k2 == "INHIBITION_CONSTANT":
                    h["C"] = _fnum(u[1])
                else:
                    raise ValueError(f"MICROBIAL_REACTION,INHIBITION: keyword {k2} not supported")
            if not h["type"]:
                raise ValueError("MICROBIAL_REACTION,INHIBITION needs a TYPE")
            r["inhibition"].append(h)
        elif key == "BIOMASS":
            for u in cur.block():
                k2 = u[0].upper()
                if k2 == "SPECIES_NAME":
                    r["biomass"] = u[1]
                elif k2 == "YIELD":
                    r["yield"] = _fnum(u[1])
        else:
            raise ValueError(f"MICROBIAL_REACTION: keyword {key} not supported")
    return r


def parse_reaction_string(text: str):
    """'A(aq) + 2 B(aq) <-> C(aq)' -> [(name, stoich)], reactants negative
    (DatabaseRxnCreateFromRxnString, reaction_database_aux.F90:91-316)"""
    if "<->" not in text:
        raise ValueError(f"reaction '{text}' has no <->")
    out = []
    for side, sign in zip(text.split("<->", 1), (-1.0, 1.0)):
        coef = None
        for tok in side.split():
            if tok == "+":
                continue
            if tok.startswith(("!", "#")):
                break
            if coef is None and _is_num(tok):
                coef = _fnum(tok)
                continue
            out.append((tok, sign * (1.0 if coef is None else coef)))
            coef = None
    return out


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
            names = [n for n in _read_names(cur) if n.upper() != "GAS_TRANSPORT_IS_UNVETTED"]
            if key == "ACTIVE_GAS_SPECIES":
                ch.active_gases += names
            ch.gases += [n for n in names if n not in ch.gases]
        elif key == "DECOUPLED_EQUILIBRIUM_REACTIONS":
            ch.decoupled = _read_names(cur)
        elif key == "MINERALS":
            ch.minerals = _read_names(cur)
        elif key == "MINERAL_KINETICS":
            ch.mineral_kinetics = _read_mineral_kinetics(cur)
        elif key == "GENERAL_REACTION":
            ch.general_rxns.append(_read_kinetic_rxn_block(cur, key))
        elif key == "MICROBIAL_REACTION":
            ch.microbial_rxns.append(_read_microbial_rxn(cur, ch))
        elif key == "RADIOACTIVE_DECAY_REACTION":
            r = _read_kinetic_rxn_block(cur, key)
            if r["k"] is None:
                raise ValueError("RATE_CONSTANT or HALF_LIFE must be set in RADIOACTIVE_DECAY_REACTION")
            ch.radiodecay_rxns.append(r)
        elif key == "IMMOBILE_DECAY_REACTION":
            r = _read_kinetic_rxn_block(cur, key)
            if r["k"] is None:
                raise ValueError("RATE_CONSTANT or HALF_LIFE must be set in IMMOBILE_DECAY_REACTION")
            ch.immobile_decay_rxns.append(r)
        elif key == "SORPTION":
            for u in cur.block():
                k2 = u[0].upper()
                if k2 == "SURFACE_COMPLEXATION_RXN":
                    ch.srfcplx_rxns.append(_read_srfcplx_rxn(cur))
                elif k2 == "ION_EXCHANGE_RXN":
                    ch.ionx_rxns.append(_read_ionx_rxn(cur))
                elif k2 == "ISOTHERM_REACTIONS":
                    ch.isotherm_rxns += _read_isotherm_rxns(cur)
                elif k2 == "DYNAMIC_KD_REACTIONS":
                    ch.dynamic_kd_rxns += _read_dynamic_kd_rxns(cur)
                else:
                    ch.unsupported.append("SORPTION," + k2)
                    _skip_nested(cur)
        elif key == "REACTION_SANDBOX":
            for u in cur.block():
                k2 = u[0].upper()
                if k2 == "CLM-CN":
                    ch.clm_cn = _read_clm_cn(cur)
                    ch.sandbox_order.append(k2)
                elif k2 == "SOMDECOMP" and ch.somdec is None:
                    ch.somdec = _read_somdec(cur)
                    ch.sandbox_order.append(k2)
                elif k2 == "NITRIFICATION" and ch.nitrif is None:
                    ch.nitrif = _read_nitrif(cur)
                    ch.sandbox_order.append(k2)
                elif k2 == "DENITRIFICATION" and ch.denitr is None:
                    ch.denitr = _read_denitr(cur)
                    ch.sandbox_order.append(k2)
                elif k2 == "PLANTN" and ch.plantn is None:
                    ch.plantn = _read_plantn(cur)
                    ch.sandbox_order.append(k2)
                elif k2 == "LANGMUIR" and ch.langmuir is None:
                    ch.langmuir = _read_langmuir(cur)
                    ch.sandbox_order.append(k2)
                elif k2 == "CNDEGAS" and ch.cndegas is None:
                    ch.cndegas = _read_cndegas(cur)
                    ch.sandbox_order.append(k2)
                elif k2 == "CALCITE" and ch.calcite_sandbox is None:
                    ch.calcite_sandbox = _read_calcite_sandbox(cur)
                    ch.sandbox_order.append(k2)
                elif k2 == "RADON" and ch.radon is None:
                    ch.radon = _read_radon(cur)
                    ch.sandbox_order.append(k2)
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
    # name -> (vol frac, specific surface area per mass as written, its factor to m^2/kg): resolved
    # into `minerals` by load_network once molar weight and volume are known (MineralProcessConstraint)
    minerals_per_mass: Dict[str, Tuple[float, float, float]] = field(default_factory=dict)
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
                if unit in _AREA_PER_MASS_UNITS:
                    cn.minerals_per_mass[u[0]] = (vf, area, _AREA_PER_MASS_UNITS[unit])
                    cn.minerals[u[0]] = (vf, float("nan"))
                    continue
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
    reference_saturation: float = 1.0     # option_flow.F90:141
    final_time: float = 0.0
    initial_dt: float = 1.0
    maximum_dt: float = 1.0e20
    minimum_dt: float = 1.0e-20      # timestepper_base.F90:23 default_dt_min
    ts_acceleration: int = 5
    newton: Dict[str, float] = field(default_factory=dict)
    osrt: bool = False
    max_steps: Optional[int] = None
    numerical_jacobian: bool = False


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
        elif key == "REFERENCE_SATURATION":
            dk.reference_saturation = _fnum(t[1])
        elif key == "EOS" and len(t) > 1 and t[1].upper() == "WATER":
            # EOS WATER / DENSITY CONSTANT rho: without a flow mode the liquid density of every
            # cell is this constant (eos_water.F90 EOSWaterSetDensity('CONSTANT'))
            while True:
                u = cur.next()
                if u is None or u[0].upper() in ("END", "/"):
                    break
                if u[0].upper() == "DENSITY" and len(u) > 2 and u[1].upper() == "CONSTANT":
                    dk.reference_liquid_density = _fnum(u[2])
        elif key == "MODE" and len(t) > 1 and t[1].upper() == "OSRT":
            dk.osrt = True
        elif key == "FINAL_TIME":
            dk.final_time = time_to_sec(_fnum(t[1]), t[2])
        elif key == "INITIAL_TIMESTEP_SIZE":
            dk.initial_dt = time_to_sec(_fnum(t[1]), t[2])
        elif key == "MAXIMUM_TIMESTEP_SIZE" and len(t) == 3:
            dk.maximum_dt = time_to_sec(_fnum(t[1]), t[2])
        elif key == "MINIMUM_TIMESTEP_SIZE" and len(t) == 3:
            dk.minimum_dt = time_to_sec(_fnum(t[1]), t[2])
        elif key == "NUMERICAL_METHODS":
            in_transport_nm = len(t) > 1 and t[1].upper() == "TRANSPORT"
        elif key == "MAX_STEPS" and in_transport_nm:
            dk.max_steps = int(t[1])
        elif key == "NUMERICAL_JACOBIAN" and in_transport_nm:
            dk.numerical_jacobian = True
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
    nok = int(ok.sum())
    if nok >= 5:
        rhs = (vec[:, ok] * lk[ok]).sum(axis=1)
        a = vec[:, ok] @ vec[:, ok].T
        return np.linalg.solve(a, rhs)
    # Fewer valid temperatures than coefficients (most Hanford complexes carry one value, at 25 C):
    # the reference's normal equations are singular there and its LU only survives through the
    # 1e-20 pivot substitution, so what it returns is rounding noise.  The set-up is not the path
    # this package accelerates; give such species the lowest-order fit their data supports
    # (constant for one point) so that an anisothermal run stays meaningful.
    coefs = np.zeros(5)
    if nok == 0:
        return coefs
    order = [1, 2, 0, 3, 4][:nok]            # 1, T, ln T, 1/T, 1/T^2
    sol, *_ = np.linalg.lstsq(vec[order][:, ok].T, lk[ok], rcond=None)
    coefs[order] = sol
    return coefs


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
        self._sorption_isotherms()
        self._kinetic_rxns()
        self.sandbox_order = list(chem.sandbox_order)
        self.elm_pflotran = False
        self._somdec()
        self._nitrif_denitr()

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

    # -- general / radioactive decay / immobile decay (reaction_database.F90:2940-3145) -------- #
    def _kinetic_rxns(self):
        pri = {n: i for i, n in enumerate(self.primary_names)}
        imm = {n: i for i, n in enumerate(self.immobile_names)}
        self.general = self.radiodecay = self.immdecay = None
        if self.chem.general_rxns:
            ptr, ids, st, fptr, fids, fst, bptr, bids, bst = [0], [], [], [0], [], [], [0], [], []
            for r in self.chem.general_rxns:
                for nm, v in parse_reaction_string(r["reaction"]):
                    ids.append(pri[nm])
                    st.append(v)
                    if v < 0.0:       # forward stoichiometries are stored positive (:3119-3126)
                        fids.append(pri[nm])
                        fst.append(abs(v))
                    elif v > 0.0:
                        bids.append(pri[nm])
                        bst.append(v)
                ptr.append(len(ids)); fptr.append(len(fids)); bptr.append(len(bids))
            self.general = dict(ptr=ptr, specid=ids, stoich=st, fwd_ptr=fptr, fwd_specid=fids, fwd_stoich=fst,
                                bwd_ptr=bptr, bwd_specid=bids, bwd_stoich=bst,
                                kf=[r["kf"] for r in self.chem.general_rxns],
                                kr=[r["kr"] for r in self.chem.general_rxns])
        if self.chem.radiodecay_rxns:
            ptr, ids, st, fwd = [0], [], [], []
            for r in self.chem.radiodecay_rxns:
                parent = None
                for nm, v in parse_reaction_string(r["reaction"]):
                    ids.append(pri[nm])
                    st.append(v)
                    if v < 0.0:
                        parent = pri[nm]      # the last negative one, like :3006-3010
                if parent is None:
                    raise ValueError("RADIOACTIVE_DECAY_REACTION without a parent species")
                fwd.append(parent)
                ptr.append(len(ids))
            self.radiodecay = dict(ptr=ptr, specid=ids, stoich=st, forward_specid=fwd,
                                   kf=[r["k"] for r in self.chem.radiodecay_rxns])
        self.microbial = None
        if self.chem.microbial_rxns:
            ptr, ids, st, mptr, mid, mK, mC = [0], [], [], [0], [], [], []
            iptr, iid, ityp, iC, iC2, bio, yld = [0], [], [], [], [], [], []
            for r in self.chem.microbial_rxns:
                names = []
                for nm, v in parse_reaction_string(r["reaction"]):
                    ids.append(pri[nm]); st.append(v); names.append(nm)
                ptr.append(len(ids))
                for m in r["monod"]:
                    if m["species"] not in names:
                        raise ValueError(f"Monod species {m['species']} not found in microbial reaction")
                    mid.append(pri[m["species"]]); mK.append(m["K"]); mC.append(m["Cth"])
                mptr.append(len(mid))
                for h in r["inhibition"]:
                    iid.append(pri[h["species"]]); ityp.append(h["type"]); iC.append(h["C"]); iC2.append(h["C2"])
                iptr.append(len(iid))
                b = r["biomass"]
                if b is None:
                    bio.append(0)
                elif b in pri:
                    bio.append(pri[b] + 1)
                elif b in imm:
                    bio.append(-(imm[b] + 1))
                else:
                    raise ValueError(f"Biomass species {b} not found among the primary aqueous or immobile species")
                if b is not None and b in names:
                    raise ValueError(f"Biomass species {b} should not be included in the microbial reaction")
                yld.append(r["yield"])
            self.microbial = dict(units=self.chem.microbial_units or 3, ptr=ptr, specid=ids, stoich=st,
                                  rate_constant=[r["rate_constant"] for r in self.chem.microbial_rxns],
                                  activation_energy=[r["activation_energy"] for r in self.chem.microbial_rxns],
                                  monod_ptr=mptr, monod_specid=mid, monod_K=mK, monod_Cth=mC,
                                  inhibition_ptr=iptr, inhibition_specid=iid, inhibition_type=ityp,
                                  inhibition_C=iC, inhibition_C2=iC2, biomassid=bio, biomass_yield=yld)
        if self.chem.immobile_decay_rxns:
            self.immdecay = dict(specid=[imm[r["species"]] for r in self.chem.immobile_decay_rxns],
                                 k=[r["k"] for r in self.chem.immobile_decay_rxns])

    # -- ion exchange / KD isotherms / dynamic KD (reaction_database.F90:2800-2925) -------- #
    def _sorption_isotherms(self):
        pri = {n: i for i, n in enumerate(self.primary_names)}
        kin = {n: i for i, n in enumerate(self.kinmnrl_names)}
        self.ionx = self.kd = self.dynkd = None
        if self.chem.ionx_rxns:
            ptr, ids, ks, cec, surf, zf = [0], [], [], [], [], []
            for rx in self.chem.ionx_rxns:
                for nm, k in rx["cations"]:
                    ids.append(pri[nm])
                    ks.append(k)
                ptr.append(len(ids))
                cec.append(rx["CEC"])
                surf.append(kin[rx["mineral"]] if rx["mineral"] else -1)
                z = [self.primary_Z[pri[nm]] for nm, _ in rx["cations"]]
                zf.append(int(any(abs(a - b) > 0.1 for a in z for b in z)))
            self.ionx = dict(ptr=ptr, cationid=ids, k=ks, CEC=cec, to_surf=surf, Z_flag=zf)
        if self.chem.isotherm_rxns:
            # IsothermConvertKDUnits: kg water / m^3 bulk, or mL water / g soil (L/kg)
            units = set()
            coeff = []
            for rx in self.chem.isotherm_rxns:
                u = rx["kd_units"].lower()
                if u in ("", "kg/m^3"):
                    units.add(0)
                    coeff.append(rx["kd"])
                elif u in ("l/kg", "ml/g"):
                    units.add(1)
                    coeff.append(rx["kd"])
                else:
                    raise ValueError(f"Unrecognized kd_units: {rx['kd_units']}")
            if len(units) != 1:
                raise ValueError("all KD isotherms must use the same units")
            self.kd = dict(specid=[pri[r["species"]] for r in self.chem.isotherm_rxns],
                           type=[r["type"] for r in self.chem.isotherm_rxns],
                           mineral=[(kin[r["mineral"]] if r["mineral"] else -1) for r in self.chem.isotherm_rxns],
                           coeff=coeff, langmuir_b=[r["langmuir_b"] for r in self.chem.isotherm_rxns],
                           freundlich_n=[r["freundlich_n"] for r in self.chem.isotherm_rxns],
                           ikd_units=units.pop())
        if self.chem.dynamic_kd_rxns:
            d = self.chem.dynamic_kd_rxns
            self.dynkd = dict(specid=[pri[r["species"]] for r in d], refspecid=[pri[r["ref"]] for r in d],
                              refspechigh=[r["ref_high"] for r in d], low=[r["low"] for r in d],
                              high=[r["high"] for r in d], power=[r["power"] for r in d])

    # -- SOMDECOMP (SomDecSetup, reaction_sandbox_somdec.F90:987-1501) -------- #
    def _species(self, name: str) -> Tuple[int, int]:
        """SomDec_SpeciesID (:3744-3806): primary first, then immobile; gases are
        not supported on this path"""
        if name in self.primary_names:
            return self.primary_names.index(name), 0
        if name in self.immobile_names:
            return self.immobile_names.index(name), 2
        raise KeyError(f"species {name} is neither a primary nor an immobile species")

    def _somdec(self):
        sb = self.chem.somdec
        self.somdec = None
        if sb is None:
            return
        pri = {n: i for i, n in enumerate(self.primary_names)}
        imm = {n: i for i, n in enumerate(self.immobile_names)}
        pool = {}
        for nm, nc in sb.pools:
            d = {"nc": -999.0 if nc is None else nc, "aq": 0, "c": -1, "n": -1}
            if nc is None:
                if nm + "C" in imm:
                    d["c"], d["n"] = imm[nm + "C"], imm.get(nm + "N", -1)
                elif nm + "C" in pri:
                    d["c"], d["n"], d["aq"] = pri[nm + "C"], pri.get(nm + "N", -1), 1
                if d["c"] < 0 or d["n"] < 0:
                    raise KeyError(f"SOMDECOMP pool {nm}: species {nm}C / {nm}N not found")
            else:
                if nm in imm:
                    d["c"] = imm[nm]
                elif nm in pri:
                    d["c"], d["aq"] = pri[nm], 1
                else:
                    raise KeyError(f"SOMDECOMP pool {nm} not found")
            for tag, key in (("CHR", "hr"), ("NMIN", "nmin"), ("NIMP", "nimp"), ("NIMM", "nimm")):
                d[key] = imm.get(nm + tag, -1)
            pool[nm] = d
        nrxn = len(sb.reactions)
        I = {k: [] for k in ("upstream_c_id", "upstream_n_id", "upstream_is_aqueous", "upstream_hr_id",
                             "upstream_nmin_id", "upstream_nimp_id", "upstream_nimm_id", "downstream_ptr",
                             "downstream_c_id", "downstream_n_id", "downstream_is_aqueous",
                             "temperature_response_function", "moisture_response_function",
                             "ox_response_function", "ox_specid", "ox_specitype", "monod_ptr", "monod_specid",
                             "monod_specitype", "monod_pool_normalized", "inhib_ptr", "inhib_itype",
                             "inhib_specid", "inhib_specitype")}
        R = {k: [] for k in ("rate_constant", "rate_decomposition", "rate_ad_factor", "upstream_nc",
                             "mineral_c_stoich", "mineral_n_stoich", "downstream_stoich", "downstream_nc", "q10",
                             "ea", "ox_half_saturation", "decomp_depth_efolding", "monod_half_saturation",
                             "monod_threshold", "inhib_constant", "inhib_constant2")}
        I["downstream_ptr"].append(0)
        I["monod_ptr"].append(0)
        I["inhib_ptr"].append(0)
        for rx in sb.reactions:
            up = pool[rx["up"]]
            I["upstream_c_id"].append(up["c"])
            I["upstream_n_id"].append(up["n"])
            I["upstream_is_aqueous"].append(up["aq"])
            for key in ("hr", "nmin", "nimp", "nimm"):
                I[f"upstream_{key}_id"].append(up[key])
            R["upstream_nc"].append(up["nc"])
            if up["n"] < 0 and up["nc"] < 0.0:
                raise ValueError("SOMDECOMP upstream pool has a negative C:N ratio")
            for nm, st in rx["down"]:
                dn = pool[nm]
                I["downstream_c_id"].append(dn["c"])
                I["downstream_n_id"].append(dn["n"])
                I["downstream_is_aqueous"].append(dn["aq"])
                R["downstream_stoich"].append(st)
                R["downstream_nc"].append(dn["nc"])
            I["downstream_ptr"].append(len(I["downstream_c_id"]))
            if rx["rate_constant"] > 0.0:
                R["rate_constant"].append(rx["rate_constant"])
                R["rate_decomposition"].append(-1.0)
            else:
                R["rate_constant"].append(-1.0)
                R["rate_decomposition"].append(rx["rate_decomposition"])
            R["rate_ad_factor"].append(rx["rate_ad_factor"])
            # fixed C:N reactions: stoichiometry at set-up (:1395-1425)
            if up["n"] >= 0:
                R["mineral_c_stoich"].append(0.0)
                R["mineral_n_stoich"].append(0.0)
            else:
                sc, sn = 1.0, up["nc"]
                for nm, st in rx["down"]:
                    sc = sc - st
                    sn = sn - st * pool[nm]["nc"]
                if abs(sc) < 1.0e-15:
                    sc = 0.0
                if abs(sn) < 1.0e-15:
                    sn = 0.0
                if sc < 0.0 or sn < 0.0:
                    raise ValueError("SOMDECOMP fixed-C:N reaction with negative respiration or N mineralisation")
                R["mineral_c_stoich"].append(sc)
                R["mineral_n_stoich"].append(sn)
            ab = rx["abiotic"]
            I["temperature_response_function"].append(ab["temperature"])
            I["moisture_response_function"].append(ab["moisture"])
            I["ox_response_function"].append(ab["ox"])
            R["q10"].append(ab["q10"])
            R["ea"].append(ab["ea"])
            R["ox_half_saturation"].append(ab["ox_half_saturation"])
            R["decomp_depth_efolding"].append(ab["depth_efolding"])
            if rx["ox_species"]:
                sid, st = self._species(rx["ox_species"])
            else:
                sid, st = -1, -1
            I["ox_specid"].append(sid)
            I["ox_specitype"].append(st)
            for m in rx["monod"]:
                sid, st = self._species(m["species"])
                I["monod_specid"].append(sid)
                I["monod_specitype"].append(st)
                I["monod_pool_normalized"].append(m["pool_normalized"])
                R["monod_half_saturation"].append(m["half_saturation"])
                R["monod_threshold"].append(m["threshold"])
            I["monod_ptr"].append(len(I["monod_specid"]))
            for ih in rx["inhibition"]:
                sid, st = self._species(ih["species"])
                I["inhib_itype"].append(ih["itype"])
                I["inhib_specid"].append(sid)
                I["inhib_specitype"].append(st)
                R["inhib_constant"].append(ih["constant"])
                R["inhib_constant2"].append(ih["constant2"])
            I["inhib_ptr"].append(len(I["inhib_specid"]))
        # CO2 species (:1437-1472)
        if sb.co2_species:
            co2_id, co2_itype = self._species(sb.co2_species)
        else:
            co2_itype = 0
            for nm in ("CO2(g)*", "CO2(aq)", "HCO3-"):
                if nm in pri:
                    co2_id = pri[nm]
                    break
            else:
                raise KeyError("SOMDECOMP: none of CO2(g)*, CO2(aq), HCO3- is a primary species")
        o2_id, o2_itype = self._species(sb.o2_species) if sb.o2_species else (-1, -1)
        scal = dict(nrxn=nrxn, co2_id=co2_id, co2_itype=co2_itype, o2_id=o2_id, o2_itype=o2_itype,
                    nh4_id=pri.get("NH4+", -1), no3_id=pri.get("NO3-", -1), n2o_id=pri.get("N2O(aq)", -1),
                    proton_id=pri.get("H+", -1), hr_id=imm.get("HRimm", -1), nmin_id=imm.get("Nmin", -1),
                    nimm_id=imm.get("Nimm", -1), nimp_id=imm.get("Nimp", -1), ngasmin_id=imm.get("NGASmin", -1),
                    x0eps=sb.x0eps, n2o_frac_mineralization=sb.n2o_frac_mineralization,
                    inhibition_nh4_no3=sb.inhibition_nh4_no3)
        self.somdec = {"scalars": scal, "int_arrays": I, "real_arrays": R}

    def _nitrif_denitr(self):
        pri = {n: i for i, n in enumerate(self.primary_names)}
        imm = {n: i for i, n in enumerate(self.immobile_names)}
        self.nitrif = self.denitr = None
        if self.chem.nitrif is not None:
            # NitrifSetup, reaction_sandbox_nitrif.F90:160-232
            d = dict(self.chem.nitrif)
            for nm, key in (("H+", "proton_id"), ("NH4+", "nh4_id"), ("NO3-", "no3_id"), ("N2O(aq)", "n2o_id")):
                d[key] = pri.get(nm, -1)
            if d["nh4_id"] < 0:
                raise KeyError("NITRIFICATION needs NH4+ as a primary species")
            d["ngasnit_id"] = imm.get("NGASnitr", -1)
            self.nitrif = d
        if self.chem.denitr is not None:
            # DenitrSetup, reaction_sandbox_denitr.F90:152-210
            d = dict(self.chem.denitr)
            for nm, key in (("NO3-", "no3_id"), ("N2(aq)", "n2_id"), ("N2O(aq)", "n2o_id")):
                d[key] = pri.get(nm, -1)
            if d["no3_id"] < 0:
                raise KeyError("DENITRIFICATION needs NO3- as a primary species")
            d["ngasdeni_id"] = imm.get("NGASdeni", -1)
            self.denitr = d
        self.plantn = self.langmuir = None
        if self.chem.plantn is not None:
            # PlantNSetup, reaction_sandbox_plantn.F90:152-220
            d = dict(self.chem.plantn)
            d["nh4_id"], d["no3_id"] = pri.get("NH4+", -1), pri.get("NO3-", -1)
            if d["nh4_id"] < 0 and d["no3_id"] < 0:
                raise KeyError("PLANTN needs NH4+ or NO3- as a primary species")
            if "PlantN" not in imm:
                raise KeyError("PLANTN needs the immobile species PlantN")
            d["plantn_id"] = imm["PlantN"]
            d["plantndemand_id"] = imm.get("Plantndemand", -1)
            d["plantnh4uptake_id"] = imm.get("Plantnh4uptake", -1)
            d["plantno3uptake_id"] = imm.get("Plantno3uptake", -1)
            self.plantn = d
        if self.chem.langmuir is not None:
            # LangmuirSetup, reaction_sandbox_langmu.F90:142-181
            g = self.chem.langmuir
            self.langmuir = {"aq_id": pri[g["name_aq"]], "sorb_id": imm[g["name_sorb"]],
                             "k_kinetic": g["k_kinetic"], "k_equilibrium": g["k_equilibrium"], "s_max": g["s_max"]}

        self.cndegas = None
        if self.chem.cndegas is not None:
            # CNdegasSetup, reaction_sandbox_cndegas.F90:140-206.  The reservoir of a gas is looked up as
            # the PRIMARY species 'X(g)*' and, failing that, as the gas species 'X(g)'; either id is then
            # added to reaction%offset_immobile (:317, :367, :415), i.e. read as an immobile index.
            g = dict(self.chem.cndegas)

            def gas_id(name):
                if name + "*" in pri:
                    return pri[name + "*"]
                if name in self.chem.gases:
                    return self.chem.gases.index(name)
                return g.get("gas_ids", {}).get(name, -1)

            d = {"co2a_id": pri.get("CO2(aq)", -1), "n2oa_id": pri.get("N2O(aq)", -1), "n2a_id": pri.get("N2(aq)", -1),
                 "co2g_id": gas_id("CO2(g)"), "n2og_id": gas_id("N2O(g)"), "n2g_id": gas_id("N2(g)"),
                 "proton_id": -1, "himm_id": -1, "fixph_on": int(g["fixph_on"]),
                 "initialize_with_molality": int(g.get("initialize_with_molality", 0)),
                 "cell_state_mode": int(g.get("cell_state_mode", 0)), "pad_": 0,
                 "k_kinetic_co2": g["k_kinetic_co2"], "k_kinetic_n2o": g["k_kinetic_n2o"],
                 "k_kinetic_n2": g["k_kinetic_n2"], "k_kinetic_h": g["k_kinetic_h"], "fixph": g["fixph"],
                 "reference_temperature": float(g.get("reference_temperature", self.tref)),
                 "reference_pressure": float(g.get("reference_pressure", 101325.0))}
            if d["fixph_on"]:
                if "H+" not in pri:
                    raise KeyError("CNDEGAS: H+ is not defined even though pH needs to be fixed")
                if "Himm" not in imm:
                    raise KeyError("CNDEGAS: Himm is not defined even though pH needs to be fixed")
                d["proton_id"], d["himm_id"] = pri["H+"], imm["Himm"]
            self.cndegas = d

        self.calcite_sandbox = None
        if self.chem.calcite_sandbox is not None:
            # CalciteSetup, reaction_sandbox_calcite.F90:108-146
            g = self.chem.calcite_sandbox
            for nm in ("H+", "Ca++", "HCO3-"):
                if nm not in pri:
                    raise KeyError(f"CALCITE sandbox: primary species {nm} not found")
            if "Calcite" not in self.kinmnrl_names:
                raise KeyError("CALCITE sandbox: Calcite is not among the kinetic minerals")
            self.calcite_sandbox = {"mineral_id": self.kinmnrl_names.index("Calcite"), "h_ion_id": pri["H+"],
                                    "calcium_id": pri["Ca++"], "bicarbonate_id": pri["HCO3-"],
                                    "rate_constant1": g["rate_constant1"], "rate_constant2": g["rate_constant2"]}

        self.radon = None
        if self.chem.radon is not None:
            # RadonSetup, reaction_sandbox_radon.F90:104-146
            g = self.chem.radon
            if g["species_name"] not in pri:
                raise KeyError(f"RADON sandbox: primary species {g['species_name']} not found")
            if g["mineral_name"] not in self.kinmnrl_names:
                raise KeyError(f"RADON sandbox: {g['mineral_name']} is not among the kinetic minerals")
            self.radon = {"species_id": pri[g["species_name"]],
                          "mineral_id": self.kinmnrl_names.index(g["mineral_name"]),
                          "radon_generation_rate": g["radon_generation_rate"]}

        # active gas species: the reactions of the chosen gases in the basis of the primaries
        # (reaction_database.F90: gas%acteq*), read by RTotalGas
        self.active_gas = None
        if self.chem.active_gases:
            rx = [self.gas_rxn[n] for n in self.chem.active_gases]
            ptr, ids, st = self.csr(rx)
            self.active_gas = {"names": list(self.chem.active_gases), "ptr": ptr, "specid": ids, "stoich": st,
                               "h2ostoich": [r.h2o_stoich for r in rx], "logK": self.logKs(rx),
                               "logK_coef": self.logK_coefs(rx)}

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
    db = Database(db_text)
    net = ReactionNetwork(dk.chemistry, db, dk.reference_temperature, use_isothermal)
    for cn in dk.constraints.values():
        # specific surface area per mass of mineral (reaction_mineral.F90:588-641):
        # m^2/kg * 1e-3 kg/g * g/mol / (m^3/mol) = m^2 per m^3 mineral, times the volume fraction
        for nm, (vf, ssa, to_m2_kg) in cn.minerals_per_mass.items():
            m = db.mineral[nm]
            if not (m.mw > 0.0 and m.molar_volume > 0.0) or vf <= 0.0:
                raise ValueError(f"mineral {nm}: a mass-based surface area needs molar weight, molar volume "
                                 "and a non-zero volume fraction")
            conv = to_m2_kg * 1.0e-3 * m.mw / m.molar_volume      # same order of operations as :621-640
            conv = conv * vf
            cn.minerals[nm] = (vf, conv * ssa)
    return dk, net
