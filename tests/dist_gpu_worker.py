"""Worker for tests/test_gpu_distributed.py: one rank of a one-process-per-shard evaluation.
All ranks may share cuda:0 (backend gloo) or own one GPU each (backend nccl)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from gparml_b200 import _lib  # noqa: E402
from gparml_b200 import distributed as gd  # noqa: E402
from gparml_b200.engine import ShardContext, evaluate  # noqa: E402
from gparml_b200.synthetic import make_problem  # noqa: E402


def main():
    backend = sys.argv[1]
    rank, world, local_rank = gd.env_rank_world()
    dev = local_rank if backend == "nccl" else 0
    torch.cuda.set_device(dev)
    if backend == "nccl":
        dist.init_process_group("nccl", device_id=torch.device("cuda", dev))
    else:
        dist.init_process_group("gloo")
    N, M, Q, D = 3001, 30, 4, 3
    p = make_problem(N, M, Q, D, seed=77, generic_hypers=True, with_direction=True)
    lo, hi = gd.shard_range(N, world, rank)
    ctx = ShardContext(M, Q, D, N, device=dev)
    ctx.use_torch_stream()
    ctx.upload_shard(p["Y"][lo:hi], p["X_mu"][lo:hi], p["X_S"][lo:hi])
    ctx.upload(_lib.A_GRAD_D, p["d"][:, lo:hi])
    Fs = []
    for rep in range(5):            # repeated: a stream-ordering bug shows up as run-to-run differences
        F, g = evaluate([ctx], p["Z"], p["sf2"], p["alpha"], p["beta"], step_size=1e-3, reduce_fn=gd.packed_reduce_fn())
        Fs.append(F)
    assert len(set(Fs)) == 1, Fs
    gl = ctx.grad_latest()
    # optimiser inner products across ranks
    ctx.scg_set_grads()
    ops = gd.DistributedLocalOps(ctx, device=torch.device("cuda", dev) if backend == "nccl" else None)
    kappa = ops.embeddings_get_grads_kappa("unused")
    # device-side initialisation across ranks: PCA + Lloyd iterations from a common guess
    from gparml_b200 import init_device
    tdev = torch.device("cuda", dev) if backend == "nccl" else None
    red = lambda a: gd.allreduce_numpy(a, device=tdev)   # noqa: E731
    Y2 = make_problem(N, M, Q, 6, seed=78)["Y"]
    ctx2 = ShardContext(M, Q, 6, N, device=dev)
    ctx2.upload_outputs(Y2[lo:hi])
    init_device.pca([ctx2], reduce_fn=red)
    X0 = ctx2.download(_lib.A_X_MU, (hi - lo, Q))
    book, distortion = init_device.kmeans_from_guess([ctx2], p["Z"][:8] * 0.3, reduce_fn=red)
    ctx2.close()
    np.savez(os.path.join(sys.argv[2], "rank%d.npz" % rank), F=F, flat=g["flat"], gl=gl, lo=lo, hi=hi, kappa=kappa,
             X0=X0, book=book, distortion=distortion)
    dist.barrier()
    dist.destroy_process_group()
    ctx.close()


if __name__ == "__main__":
    main()
