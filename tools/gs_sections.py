"""Section timing of the master-step head kernel (SM clock cycles, GPARML_A_GS_EXTRA[8..12]):
python tools/gs_sections.py [M] [Q] [D]   (needs a GPU)"""
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gparml_b200 import _lib
from gparml_b200.engine import ShardContext
from gparml_b200.synthetic import make_problem

M, Q, D = (int(v) for v in (sys.argv[1:4] + ["100", "10", "10"])[:3])
n = 20000
p = make_problem(n, M, Q, D, seed=3)
with ShardContext(M, Q, D, n) as c:
    c.upload_shard(p["Y"], p["X_mu"], p["X_S"])
    for rep in range(3):
        c.set_globals(p["Z"], p["sf2"], p["alpha"], p["beta"])
        c.statistics()
        c.enable_timing(True)
        F, g = c.global_step()
        ms = c.phase_times_ms()["global_step"]
        c.enable_timing(False)
    ex = c.download(_lib.A_GS_EXTRA)
    names = ("form A", "inversion", "C, tr", "dF/dPsi2", "pair tables")
    tot = sum(ex[8:13])
    print("M=%d Q=%d D=%d: head %.1f us (events), %.0f cycles in sections" % (M, Q, D, 1e3 * ms, tot))
    for nm, cyc in zip(names, ex[8:13]):
        print("  %-12s %9.0f cycles  %6.1f us  %4.1f %%" % (nm, cyc, cyc / 1965.0, 100.0 * cyc / tot))
