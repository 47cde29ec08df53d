#!/usr/bin/env python
"""Copy the reference's OWN regression fixtures (decks, tiny .dat databases and
.regression.gold files -- test data, not source) for the hot-path chemistry
into tests/golden/, and write trimmed extracts of the big databases.

Run in the build container (needs /root/reference); the GPU box only sees the
committed copies.  Provenance of every file is recorded in
tests/golden/PROVENANCE.txt.
"""
import os
import shutil
import sys

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from pflotran_elm_interface_b200 import chem  # noqa: E402

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(__file__), "..", "tests", "golden")

COPY = [
    # (reference path, local name)
    ("regression_tests/ascem/batch/calcite-kinetics.in", None),
    ("regression_tests/ascem/batch/calcite-kinetics.regression.gold", None),
    ("regression_tests/ascem/batch/calcite-kinetics-volume-fractions.in", None),
    ("regression_tests/ascem/batch/calcite-kinetics-volume-fractions.regression.gold", None),
    ("regression_tests/ascem/batch/calcite.dat", None),
    ("regression_tests/ascem/batch/carbonate-unit-activity.in", None),
    ("regression_tests/ascem/batch/carbonate-unit-activity.regression.gold", None),
    ("regression_tests/ascem/batch/carbonate-debye-huckel-activity.in", None),
    ("regression_tests/ascem/batch/carbonate-debye-huckel-activity.regression.gold", None),
    ("regression_tests/ascem/batch/carbonate.dat", None),
    ("regression_tests/ascem/batch/ca-carbonate-unit-activity.in", None),
    ("regression_tests/ascem/batch/ca-carbonate-unit-activity.regression.gold", None),
    ("regression_tests/ascem/batch/ca-carbonate-debye-huckel-activity.in", None),
    ("regression_tests/ascem/batch/ca-carbonate-debye-huckel-activity.regression.gold", None),
    ("regression_tests/ascem/batch/ca-carbonate.dat", None),
    ("regression_tests/ascem/batch/surface-complexation-1.in", None),
    ("regression_tests/ascem/batch/surface-complexation-1.regression.gold", None),
    ("regression_tests/ascem/batch/surface-complexation.dat", None),
    ("regression_tests/ascem/batch/ion-exchange-valocchi.in", None),
    ("regression_tests/ascem/batch/ion-exchange-valocchi.regression.gold", None),
    ("regression_tests/ngee/CLM-CN.in", None),
    ("regression_tests/ngee/CLM-CN.regression.gold", None),
    ("regression_tests/ngee/CLM-CN_database.dat", None),
    ("regression_tests/default/543/543_hanford_srfcplx_base.in", None),
    ("regression_tests/default/543/543_hanford_srfcplx_base.regression.gold", None),
    ("regression_tests/default/543/543_hanford_srfcplx_mr.in", None),
    ("regression_tests/default/batch/radon.in", None),
    ("regression_tests/default/batch/radon.regression.gold", None),
    ("regression_tests/default/column/surface_complexation_mr_os.in", None),
    ("regression_tests/default/column/tracer_os.in", None),
    ("regression_tests/default/column/tracer_os.regression.gold", None),
    ("regression_tests/default/column/tracer_os_no_geochem.regression.gold", None),
    ("shortcourse/1D_Calcite/calcite_tran_only.in", None),
    # SOMDECOMP sandbox golds (ngee/CLMCNplus, the decks that use SOMDECOMP only)
    ("regression_tests/ngee/CLMCNplus/clm_lit1.in", "clmcnplus_clm_lit1.in"),
    ("regression_tests/ngee/CLMCNplus/clm_lit1.regression.gold", "clmcnplus_clm_lit1.regression.gold"),
    ("regression_tests/ngee/CLMCNplus/clm_lit2.in", "clmcnplus_clm_lit2.in"),
    ("regression_tests/ngee/CLMCNplus/clm_lit2.regression.gold", "clmcnplus_clm_lit2.regression.gold"),
    ("regression_tests/ngee/CLMCNplus/clm_lit3.in", "clmcnplus_clm_lit3.in"),
    ("regression_tests/ngee/CLMCNplus/clm_lit3.regression.gold", "clmcnplus_clm_lit3.regression.gold"),
    ("regression_tests/ngee/CLMCNplus/clm_som1.in", "clmcnplus_clm_som1.in"),
    ("regression_tests/ngee/CLMCNplus/clm_som1.regression.gold", "clmcnplus_clm_som1.regression.gold"),
    ("regression_tests/ngee/CLMCNplus/clm_som2.in", "clmcnplus_clm_som2.in"),
    ("regression_tests/ngee/CLMCNplus/clm_som2.regression.gold", "clmcnplus_clm_som2.regression.gold"),
    ("regression_tests/ngee/CLMCNplus/clm_som3.in", "clmcnplus_clm_som3.in"),
    ("regression_tests/ngee/CLMCNplus/clm_som3.regression.gold", "clmcnplus_clm_som3.regression.gold"),
    ("regression_tests/ngee/CLMCNplus/clm_som4.in", "clmcnplus_clm_som4.in"),
    ("regression_tests/ngee/CLMCNplus/clm_som4.regression.gold", "clmcnplus_clm_som4.regression.gold"),
    ("regression_tests/ngee/CLMCNplus/clm_nmin.in", "clmcnplus_clm_nmin.in"),
    ("regression_tests/ngee/CLMCNplus/clm_nmin.regression.gold", "clmcnplus_clm_nmin.regression.gold"),
    ("regression_tests/ngee/CLMCNplus/clm_nimm1.in", "clmcnplus_clm_nimm1.in"),
    ("regression_tests/ngee/CLMCNplus/clm_nimm1.regression.gold", "clmcnplus_clm_nimm1.regression.gold"),
    ("regression_tests/ngee/CLMCNplus/clm_nimm2.in", "clmcnplus_clm_nimm2.in"),
    ("regression_tests/ngee/CLMCNplus/clm_nimm2.regression.gold", "clmcnplus_clm_nimm2.regression.gold"),
    ("regression_tests/ngee/CLMCNplus/clm_nimm3.in", "clmcnplus_clm_nimm3.in"),
    ("regression_tests/ngee/CLMCNplus/clm_nimm3.regression.gold", "clmcnplus_clm_nimm3.regression.gold"),
    ("regression_tests/ngee/CLMCNplus/clm_nimm4.in", "clmcnplus_clm_nimm4.in"),
    ("regression_tests/ngee/CLMCNplus/clm_nimm4.regression.gold", "clmcnplus_clm_nimm4.regression.gold"),
    ("regression_tests/ngee/CLMCNplus/clm_nmit.in", "clmcnplus_clm_nmit.in"),
    ("regression_tests/ngee/CLMCNplus/clm_nmit.regression.gold", "clmcnplus_clm_nmit.regression.gold"),
    ("regression_tests/ngee/CLMCNplus/clm_cn1.in", "clmcnplus_clm_cn1.in"),
    ("regression_tests/ngee/CLMCNplus/clm_cn1.regression.gold", "clmcnplus_clm_cn1.regression.gold"),
    ("regression_tests/ngee/CLMCNplus/clm_cn2.in", "clmcnplus_clm_cn2.in"),
    ("regression_tests/ngee/CLMCNplus/clm_cn2.regression.gold", "clmcnplus_clm_cn2.regression.gold"),
    ("regression_tests/ngee/CLMCNplus/clm_cn3.in", "clmcnplus_clm_cn3.in"),
    ("regression_tests/ngee/CLMCNplus/clm_cn3.regression.gold", "clmcnplus_clm_cn3.regression.gold"),
    ("regression_tests/ngee/CLMCNplus/clm_nuptake1.in", "clmcnplus_clm_nuptake1.in"),
    ("regression_tests/ngee/CLMCNplus/clm_nuptake1.regression.gold", "clmcnplus_clm_nuptake1.regression.gold"),
    ("regression_tests/ngee/CLMCNplus/clm_nuptake2.in", "clmcnplus_clm_nuptake2.in"),
    ("regression_tests/ngee/CLMCNplus/clm_nuptake2.regression.gold", "clmcnplus_clm_nuptake2.regression.gold"),
    ("regression_tests/ngee/CLMCNplus/clm_nuptake3.in", "clmcnplus_clm_nuptake3.in"),
    ("regression_tests/ngee/CLMCNplus/clm_nuptake3.regression.gold", "clmcnplus_clm_nuptake3.regression.gold"),
    ("regression_tests/ngee/CLMCNplus/clm_nh4absorption.in", "clmcnplus_clm_nh4absorption.in"),
    ("regression_tests/ngee/CLMCNplus/clm_nh4absorption.regression.gold", "clmcnplus_clm_nh4absorption.regression.gold"),
    ("regression_tests/ngee/CLMCNplus/clm_nh4desorption.in", "clmcnplus_clm_nh4desorption.in"),
    ("regression_tests/ngee/CLMCNplus/clm_nh4desorption.regression.gold", "clmcnplus_clm_nh4desorption.regression.gold"),
    ("regression_tests/ngee/CLMCNplus/CLM-CN_database.dat", "clmcnplus_CLM-CN_database.dat"),
    ("regression_tests/ascem/batch/calcite-area-per-mass.in", None),
    ("regression_tests/ascem/batch/calcite-area-per-mass.regression.gold", None),
    ("regression_tests/ascem/batch/general-reaction.in", None),
    ("regression_tests/ascem/batch/general-reaction.regression.gold", None),
    # RMicrobial (+ RImmobileDecay, RGeneral) batch golds
    ("regression_tests/default/batch/ABCD_microbial.in", None),
    ("regression_tests/default/batch/ABCD_microbial.regression.gold", None),
    ("regression_tests/default/batch/ABCD_microbial_activation_high.in", None),
    ("regression_tests/default/batch/ABCD_microbial_activation_high.regression.gold", None),
    ("regression_tests/default/batch/ABCD_microbial_activation_low.in", None),
    ("regression_tests/default/batch/ABCD_microbial_activation_low.regression.gold", None),
    ("regression_tests/default/batch/ABCD_microbial_activity.in", None),
    ("regression_tests/default/batch/ABCD_microbial_activity.regression.gold", None),
    ("regression_tests/default/batch/ABCD_microbial_aq_biomass.in", None),
    ("regression_tests/default/batch/ABCD_microbial_aq_biomass.regression.gold", None),
    ("regression_tests/default/batch/ABCD_microbial_molality.in", None),
    ("regression_tests/default/batch/ABCD_microbial_molality.regression.gold", None),
    ("regression_tests/default/batch/ABCD_microbial_molarity.in", None),
    ("regression_tests/default/batch/ABCD_microbial_molarity.regression.gold", None),
    # KD isotherms and dynamic KD (default/batch)
    ("regression_tests/default/batch/dynamic_KD.in", None),
    ("regression_tests/default/batch/dynamic_KD.regression.gold", None),
    ("regression_tests/default/batch/solute_KD_wo_mineral.in", None),
    ("regression_tests/default/batch/solute_KD_wo_mineral.regression.gold", None),
    ("regression_tests/default/batch/solute_KD_w_mineral.in", None),
    ("regression_tests/default/batch/solute_KD_w_mineral.regression.gold", None),
]


def main():
    os.makedirs(OUT, exist_ok=True)
    prov = []
    for src, name in COPY:
        p = os.path.join(REF, src)
        if not os.path.exists(p):
            print("missing", src)
            continue
        dst = os.path.join(OUT, name or os.path.basename(src))
        shutil.copyfile(p, dst)
        os.chmod(dst, 0o644)
        prov.append(f"{os.path.basename(dst)} <- {src} (verbatim copy)")
    # trimmed hanford.dat: only the species the Hanford decks and the C2/C5
    # synthetic configurations name
    with open(os.path.join(OUT, "543_hanford_srfcplx_base.in")) as f:
        dk = chem.read_deck(f.read())
    ch = dk.chemistry
    names = (["H2O"] + ch.primary + ch.secondary + ch.gases + ch.minerals
             + [c for r in ch.srfcplx_rxns for c in r.complexes]
             + ["Dolomite", "Gypsum", "Fluorite", "Schoepite", "O2(aq)", "O2(g)", "Halite", "A(aq)", "A(s)", "B(aq)", "C(aq)", "AB(aq)", "D(aq)",
                "Rn(aq)", "Rn(g)", "Quartz", "SiO2(aq)"])
    db = chem.Database.from_file(os.path.join(REF, "database/hanford.dat"))
    with open(os.path.join(OUT, "hanford_subset.dat"), "w") as f:
        f.write(db.subset_text(names))
    prov.append("hanford_subset.dat <- database/hanford.dat (lines of the species named by "
                "543_hanford_srfcplx_base.in, verbatim; made by tools/fetch_fixtures.py)")
    with open(os.path.join(OUT, "PROVENANCE.txt"), "w") as f:
        f.write("Reference fixtures (test data) copied from /root/reference by tools/fetch_fixtures.py\n")
        f.write("\n".join(prov) + "\n")
    print("\n".join(prov))


if __name__ == "__main__":
    main()
