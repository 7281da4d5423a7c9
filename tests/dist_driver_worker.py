"""Worker for tests/test_gpu_distributed.py::test_driver_*: one rank of the replayed reference driver
(`gparml_b200.parallel_GPLVM`) on the `b200_MapReduce` backend under torch.distributed.run.
All ranks may share cuda:0 (backend gloo) or own one GPU each (backend nccl).

argv: backend, work dir (input/ embeddings/ statistics/ tmp/ exist, input shards written), M, Q, D, iterations
"""
import os
import shutil
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    backend, work = sys.argv[1], sys.argv[2]
    M, Q, D, iters = (int(v) for v in sys.argv[3:7])
    import torch.distributed as dist
    from gparml_b200 import b200_MapReduce, parallel_GPLVM as drv
    from gparml_b200.scg_adapted import SCG_adapted
    dirs = {d: os.path.join(work, d) for d in ("input", "embeddings", "statistics", "tmp")}
    np.random.seed(0)
    opts = drv.default_options(M=M, Q=Q, D=D, iterations=iters, init="PCA", display=False, b200_backend=backend,
                               b200_write_files=False, **dirs)
    opts = b200_MapReduce.init(opts)
    rank, world, _ = b200_MapReduce.dist_info(opts)
    opts, gs = drv.init_statistics(b200_MapReduce, opts)
    x0 = drv.flatten_global_statistics(opts, gs)
    x0 = np.array([drv.sp.transform_back(b, x) for b, x in zip(opts["flat_global_statistics_bounds"], x0)])
    dist.barrier()
    if rank == 0:       # the initial embeddings (flush overwrites them): what the oracle run starts from
        shutil.copytree(dirs["embeddings"], os.path.join(work, "embeddings_init"))
    dist.barrier()
    drv.options, drv.map_reduce = opts, b200_MapReduce
    x, flog, nev, status, tacc = SCG_adapted(drv.likelihood_and_gradient, x0.copy(), opts["embeddings"], False,
                                             display=False, maxiters=iters, xtol=0, ftol=0, gtol=0)
    drv.likelihood_and_gradient(x, "f")              # writes the checkpoint files (rank 0)
    b200_MapReduce.flush(opts)
    n_ctx = len(b200_MapReduce.session_contexts(opts["embeddings"]))
    np.savez(os.path.join(work, "rank%d.npz" % rank), x0=x0, x=x, flog=np.array(flog), N=opts["N"], n_ctx=n_ctx,
             world=world)
    dist.barrier()
    b200_MapReduce.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
