"""One evaluation at M=500 (multi-kernel master step) for a launch-list capture:
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python tools/gs_large_launches.py"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gparml_b200.engine import ShardContext
from gparml_b200.synthetic import make_problem

M, Q, D, n = int(os.environ.get("GS_M", "500")), 10, int(os.environ.get("GS_D", "50")), 4000
p = make_problem(n, M, Q, D, seed=4)
with ShardContext(M, Q, D, n) as c:
    c.upload_shard(p["Y"], p["X_mu"], p["X_S"])
    for rep in range(2):
        c.set_globals(p["Z"], p["sf2"], p["alpha"], p["beta"])
        c.statistics()
        c.enable_timing(True)
        F, g = c.global_step()
        c.embedding_grads()
        print("rep", rep, "F", F, "global_step ms", c.phase_times_ms()["global_step"])
        c.enable_timing(False)
