import sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np
import oracle_lib as orc
from pflotran_elm_interface_b200 import abi, rstep, workloads as W, chem, constraint, eos

def run(deck_mod, label, n=3000, dt=3600.0):
    deck = W.C6_DECK
    for a, b in deck_mod:
        assert a in deck, a[:30]
        deck = deck.replace(a, b)
    W_C6 = W.C6_DECK
    W.C6_DECK = deck
    try:
        wl = W.by_name("c6", ncell=n, tran_dt=dt)
    finally:
        W.C6_DECK = W_C6
    ref = wl.state.copy()
    r0 = orc.rstep(wl.cfg, ref, dt, 4)
    step = rstep.ChemistryStep(wl.cfg, 0)
    dev = rstep.DeviceState.from_host(wl.state, "cuda:0")
    step.bind(dev)
    r1 = step.rstep(dt)
    got = dev.to_host()
    out = []
    for f in ("total", "pri_molal", "total_sorb_eq", "eqionx_conc", "eqionx_ref_cation_sorbed_conc"):
        a, b = ref.a[f], got.a[f]
        if a.size == 0: continue
        sc = np.maximum(np.abs(a), np.abs(b)); sc[sc == 0] = 1
        e = np.abs(a - b) / sc
        out.append((f, float(e.max()), np.unravel_index(e.argmax(), e.shape)))
    same = np.array_equal(ref.a["num_iterations"], got.a["num_iterations"])
    print(label, "its same", same, "mean its", ref.a["num_iterations"].mean(), out, flush=True)
    step.close()

ION1 = """    ION_EXCHANGE_RXN
      CEC 750. eq/m^3
      CATIONS
        Ca++  3.38638672536d0
        Na+   1.d0 REFERENCE
        Mg++  6.00240096038d0
      /
    /
"""
ION2 = """    ION_EXCHANGE_RXN
      MINERAL Halite
      CEC 5.d4
      CATIONS
        Na+   1.d0 REFERENCE
        K+    2.5d0
      /
    /
"""
ISO = W.C6_DECK[W.C6_DECK.index("    ISOTHERM_REACTIONS"):W.C6_DECK.index("    DYNAMIC_KD_REACTIONS")]
DYN = W.C6_DECK[W.C6_DECK.index("    DYNAMIC_KD_REACTIONS"):W.C6_DECK.index("  /\n  DATABASE")]
run([], "all")
run([(ION2, ""), (ISO, ""), (DYN, "")], "ionx1 only")
run([(ION1, ""), (ISO, ""), (DYN, "")], "ionx2 only")
run([(ION1, ""), (ION2, ""), (DYN, "")], "kd only")
run([(ION1, ""), (ION2, ""), (ISO, "")], "dynkd only")
