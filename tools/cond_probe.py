#!/usr/bin/env python
"""Master-step accuracy against the conditioning of Kmm (development probe, GPU box):
two inducing points are moved towards each other; the oracle's master step (LAPACK inverse / slogdet, the reference's
formulas) runs on the GPU's own statistics, so only the master step differs.   python tools/cond_probe.py [M]"""
import sys
import numpy as np
sys.path.insert(0, ".")
from gparml_b200.engine import ShardContext
from gparml_b200.synthetic import make_problem
from oracle import gparml_oracle as O

M = int(sys.argv[1]) if len(sys.argv) > 1 else 100
Q, D, n = 10, 10, 4000
p = make_problem(n, M, Q, D, seed=5, generic_hypers=True)
for delta in (1.0, 1e-1, 1e-2, 1e-3, 1e-4, 3e-5):
    Z = p["Z"].copy()
    Z[1] = Z[0] + delta * (Z[1] - Z[0]) / np.linalg.norm(Z[1] - Z[0])
    with ShardContext(M, Q, D, n) as c:
        c.upload_shard(p["Y"], p["X_mu"], p["X_S"])
        c.set_globals(Z, p["sf2"], p["alpha"], p["beta"])
        c.statistics()
        st = c.stats_named()
        try:
            F, g = c.global_step()
        except Exception as e:
            print("delta %.0e raised %s" % (delta, e))
            continue
        ref = O.global_step(st, Z, p["sf2"], p["alpha"], p["beta"], n)
        K = O.kmm(Z, p["sf2"], p["alpha"])
        rel = lambda a, b: float(np.max(np.abs(np.asarray(a) - np.asarray(b))) / np.max(np.abs(np.asarray(b))))
        print("delta %.0e  log10 cond(Kmm) %5.2f  F %.2e  gZ %.2e  galpha %.2e  gsf2 %.2e  gbeta %.2e  jitter %d" % (
            delta, np.log10(np.linalg.cond(K)), abs(F - ref["F"]) / abs(ref["F"]), rel(g["Z"], ref["grad_Z"]),
            rel(g["alpha"], ref["grad_alpha"]), rel(g["sf2"], ref["grad_sf2"]), rel(g["beta"], ref["grad_beta"]), c.jitter_events))
