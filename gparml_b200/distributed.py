"""One-process-per-GPU plumbing (torch.distributed): the reference's file-based reducer
(``local_MapReduce.py:250-277``, one ``load`` + ``+=`` per statistic and shard) becomes ONE
sum all-reduce of the packed statistics buffer (NCCL over NVLink on B200; gloo on CPU for the
host-logic tests), and the optimiser's per-shard partial inner products
(``scg_adapted_local_MapReduce.py:60-156``) become one small all-reduce.

Nothing here touches the data path: the points never leave their GPU.
"""
import os

from .synthetic import split_rows


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def shard_range(n_total, world, rank):
    """Contiguous rows [lo, hi) of rank's shard (uneven sizes allowed: the statistics are sums)."""
    return split_rows(n_total, world)[rank]


def init_process_group(backend=None):
    import torch
    import torch.distributed as dist
    rank, world, local_rank = env_rank_world()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            kw["device_id"] = torch.device("cuda", local_rank)
        dist.init_process_group(backend, rank=rank, world_size=world, **kw)
    return rank, world, local_rank


def allreduce_sum_(tensor):
    """In-place sum over ranks (no-op for a single rank)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(tensor, op=dist.ReduceOp.SUM)
    return tensor


def allreduce_scalars(values, op="sum", device=None):
    """Reduce a short list of Python floats over ranks; returns Python floats."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
        return [float(v) for v in values]
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM if op == "sum" else dist.ReduceOp.MAX)
    return [float(v) for v in t.cpu()]


def allreduce_numpy(array, device=None):
    """Sum a small float64 numpy array over ranks (the ``reduce_fn`` of ``init_device``: column
    sums, the D x D scatter matrix, k-means cluster sums).  ``device``: where the temporary tensor
    lives (a CUDA device for NCCL, None for gloo)."""
    import numpy
    import torch
    import torch.distributed as dist
    array = numpy.ascontiguousarray(array, dtype=numpy.float64)
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
        return array
    t = torch.from_numpy(array.copy())
    if device is not None:
        t = t.to(device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()


def packed_reduce_fn():
    """reduce_fn for ``engine.evaluate``: all-reduces the context's packed statistics in place."""
    def fn(ctx):
        allreduce_sum_(ctx.stats_torch_view())
    return fn


class DistributedLocalOps(object):
    """The 12 local-state operations for a rank that owns ONE shard context, with the partial
    inner products summed (max for max_d) over ranks.  Same function names as
    ``scg_adapted_local_MapReduce`` so that ``SCG_adapted(..., local_ops=this)`` runs unchanged on
    every rank with identical scalars."""

    def __init__(self, ctx, device=None):
        self.ctx, self.device = ctx, device
        self.time_acc = {}

    def _red(self, v, op="sum"):
        return allreduce_scalars([v], op, self.device)[0]

    def embeddings_set_grads(self, folder): self.ctx.scg_set_grads()
    def embeddings_get_grads_mu(self, folder): return self._red(self.ctx.scg_get_mu())
    def embeddings_get_grads_kappa(self, folder): return self._red(self.ctx.scg_get_kappa())
    def embeddings_get_grads_theta(self, folder): return self._red(self.ctx.scg_get_theta())
    def embeddings_get_grads_current_grad(self, folder): return self._red(self.ctx.scg_get_current_grad())
    def embeddings_get_grads_gamma(self, folder): return self._red(self.ctx.scg_get_gamma())
    def embeddings_get_grads_max_d(self, folder, alpha): return self._red(self.ctx.scg_get_max_d(alpha), "max")
    def embeddings_set_grads_reset_d(self, folder): self.ctx.scg_reset_d()
    def embeddings_set_grads_update_d(self, folder, gamma): self.ctx.scg_update_d(gamma)
    def embeddings_set_grads_update_X(self, folder, alpha): self.ctx.scg_update_X(alpha)
    def embeddings_set_grads_update_grad_old(self, folder): self.ctx.scg_update_grad_old()
    def embeddings_set_grads_update_grad_new(self, folder): self.ctx.scg_update_grad_new()
