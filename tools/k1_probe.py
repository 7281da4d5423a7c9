"""K1 (psi1_stats) alone at a BASELINE shape, for timing / ncu: python tools/k1_probe.py [cfg] [n]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gparml_b200.engine import ShardContext
from gparml_b200.synthetic import CONFIGS, make_problem

cfg = sys.argv[1] if len(sys.argv) > 1 else "c4"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
k = CONFIGS[cfg]
p = make_problem(n, k["M"], k["Q"], k["D"], seed=4)
with ShardContext(k["M"], k["Q"], k["D"], n) as c:
    c.upload_shard(p["Y"], p["X_mu"], p["X_S"])
    c.enable_timing(True)
    for rep in range(3):
        c.set_globals(p["Z"], p["sf2"], p["alpha"], p["beta"])
        c.statistics()
        t = c.phase_times_ms()
    print(cfg, "n", n, {kk: round(v, 3) for kk, v in t.items()})
