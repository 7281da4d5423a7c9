import sys, time, numpy as np
sys.path.insert(0, '.')
from gparml_b200.engine import ShardContext
from gparml_b200.synthetic import make_problem
for M in (100, 50, 116):
    p = make_problem(20000, M, 10, 10, seed=3)
    c = ShardContext(M, 10, 10, 20000)
    c.upload_shard(p["Y"], p["X_mu"], p["X_S"])
    c.set_globals(p["Z"], p["sf2"], p["alpha"], p["beta"])
    c.statistics()
    for name, fn in (("kmm_only", c.update_global_statistics), ("full", c.global_step)):
        fn(); c.synchronize()
        t = time.perf_counter()
        for _ in range(20):
            fn()
        c.synchronize()
        print(M, name, "%.1f us" % ((time.perf_counter() - t) / 20 * 1e6))
    c.close()
