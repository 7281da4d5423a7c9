"""Optimiser local-state operations, interchangeable with the reference's
``scg_adapted_local_MapReduce`` module (``scg_adapted_local_MapReduce.py:29-243``): the same
12 functions taking the embeddings ``folder``, returning Python floats, and the same
``time_acc`` dictionary (returned by ``SCG_adapted``, scg_adapted.py:336).

The reference re-reads ``(2, n, Q)`` ``.npy`` files for every inner product; here the four
gradient vectors and the embeddings are device-resident in the shard contexts of the session
registered for ``folder`` (``b200_MapReduce.session_contexts``), each operation is one fused
streaming kernel per shard and the per-shard partial results are added in shard order.
"""
import time

from . import b200_MapReduce

time_acc = {
    'embeddings_set_grads': [], 'embeddings_get_grads_mu': [], 'embeddings_get_grads_kappa': [],
    'embeddings_get_grads_theta': [], 'embeddings_get_grads_current_grad': [], 'embeddings_get_grads_gamma': [],
    'embeddings_get_grads_max_d': [], 'embeddings_set_grads_reset_d': [], 'embeddings_set_grads_update_d': [],
    'embeddings_set_grads_update_X': [], 'embeddings_set_grads_update_grad_old': [],
    'embeddings_set_grads_update_grad_new': [],
}


def _timed(name, fn):
    start = time.time()
    out = fn()
    time_acc[name] += [time.time() - start]
    return out


def _ctx(folder):
    return b200_MapReduce.session_contexts(folder)


def _sum(folder, method):
    """Partial inner products of this process's shards, added in shard order and -- under
    torch.distributed.run -- summed over the ranks, so that every rank's optimiser sees the same scalar."""
    total = 0
    for c in _ctx(folder):
        total += getattr(c, method)()
    return _over_ranks(folder, total, "sum")


def _over_ranks(folder, value, op):
    s = b200_MapReduce.session_of(folder)
    if s.world == 1:
        return value
    from . import distributed as gd
    return gd.allreduce_scalars([value], op, s.tdev)[0]


def embeddings_set_grads(folder):                       # :29-55
    _timed('embeddings_set_grads', lambda: [c.scg_set_grads() for c in _ctx(folder)])


def embeddings_get_grads_mu(folder):                    # :60-75
    return _timed('embeddings_get_grads_mu', lambda: _sum(folder, 'scg_get_mu'))


def embeddings_get_grads_kappa(folder):                 # :77-90
    return _timed('embeddings_get_grads_kappa', lambda: _sum(folder, 'scg_get_kappa'))


def embeddings_get_grads_theta(folder):                 # :92-109
    return _timed('embeddings_get_grads_theta', lambda: _sum(folder, 'scg_get_theta'))


def embeddings_get_grads_current_grad(folder):          # :111-124
    return _timed('embeddings_get_grads_current_grad', lambda: _sum(folder, 'scg_get_current_grad'))


def embeddings_get_grads_gamma(folder):                 # :126-141
    return _timed('embeddings_get_grads_gamma', lambda: _sum(folder, 'scg_get_gamma'))


def embeddings_get_grads_max_d(folder, alpha):          # :143-156
    return _timed('embeddings_get_grads_max_d', lambda: _over_ranks(folder, max([0] + [c.scg_get_max_d(alpha) for c in _ctx(folder)]), "max"))


def embeddings_set_grads_reset_d(folder):               # :161-174
    _timed('embeddings_set_grads_reset_d', lambda: [c.scg_reset_d() for c in _ctx(folder)])


def embeddings_set_grads_update_d(folder, gamma):       # :176-191
    _timed('embeddings_set_grads_update_d', lambda: [c.scg_update_d(gamma) for c in _ctx(folder)])


def embeddings_set_grads_update_X(folder, alpha):       # :193-216
    _timed('embeddings_set_grads_update_X', lambda: [c.scg_update_X(alpha) for c in _ctx(folder)])


def embeddings_set_grads_update_grad_old(folder):       # :218-230
    _timed('embeddings_set_grads_update_grad_old', lambda: [c.scg_update_grad_old() for c in _ctx(folder)])


def embeddings_set_grads_update_grad_new(folder):       # :232-243
    _timed('embeddings_set_grads_update_grad_new', lambda: [c.scg_update_grad_new() for c in _ctx(folder)])
