#!/usr/bin/env python
"""Where the end-to-end evaluation loses time against the device-resident one (development probe, one GPU):
the same loop with the upload and / or the download switched on, host wall-clock per step after warm-up."""
import sys
import time
import numpy as np
import torch
sys.path.insert(0, ".")
from gparml_b200.engine import ShardContext
from gparml_b200.synthetic import CONFIGS, block_problem_globals, block_problem_rows

cfg = sys.argv[1] if len(sys.argv) > 1 else "c3"
chunks = int(sys.argv[2]) if len(sys.argv) > 2 else 3
k = CONFIGS[cfg]
N, M, Q, D = k["N"], k["M"], k["Q"], k["D"]
g = block_problem_globals(cfg)
p = block_problem_rows(cfg, 0, N, with_direction=False)
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
Yp, MUp, Sp = pin(p["Y"]), pin(p["X_mu"]), pin(p["X_S"])
GLp = torch.empty((2, N, Q), dtype=torch.float64).pin_memory()
c = ShardContext(M, Q, D, N)
c.upload_shard_ptrs(Yp.data_ptr(), MUp.data_ptr(), Sp.data_ptr(), N)


def step(up, down):
    if up:
        c.upload_shard_ptrs(Yp.data_ptr(), MUp.data_ptr(), Sp.data_ptr(), N)
    c.set_globals(g["Z"], g["sf2"], g["alpha"], g["beta"])
    c.set_step(0.0)
    c.statistics_launch()
    c.global_step_begin()
    if down:
        c.embedding_grads_into(GLp.data_ptr(), chunks=chunks)
    else:
        c.embedding_grads()
    return c.global_step_end()


for up, down in ((0, 0), (1, 0), (0, 1), (1, 1)):
    for _ in range(4):
        step(up, down)
    c.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        step(up, down)
    c.synchronize()
    print("upload %d download %d: %.3f ms per evaluation  (dl ratio %s)" % (up, down, (time.perf_counter() - t0) * 100.0, ""))
