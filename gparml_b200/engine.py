"""Thin object wrapper over the C ABI: one :class:`ShardContext` = one shard of the
data resident on one B200 (the reference's "node": one input file handled by one
mapper process, ``local_MapReduce.py:132-137``).

Everything numeric happens in ``libgparml_b200.so``; this module only moves numpy
arrays across the boundary and sequences calls.
"""
import ctypes

import numpy as np

from . import _lib
from ._lib import STAT_FIELDS

_SCALAR_STATS = ("sum_YYT", "sum_exp_K_ii", "sum_KL", "sum_d_exp_K_ii_d_sf2")


def stat_shapes(M, Q, D):
    """Shapes of the 12 accumulated statistics (parallel_GPLVM.py:142-151)."""
    return {
        "sum_YYT": (), "sum_exp_K_ii": (), "sum_KL": (), "sum_d_exp_K_ii_d_sf2": (),
        "sum_exp_K_mi_K_im": (M, M), "sum_exp_K_miY": (M, D),
        "sum_d_exp_K_miY_d_Z": (M, Q, D), "sum_d_exp_K_mi_K_im_d_Z": (M, Q, M),
        "sum_d_exp_K_miY_d_alpha": (Q, M, D), "sum_d_exp_K_mi_K_im_d_alpha": (Q, M, M),
        "sum_d_exp_K_miY_d_sf2": (M, D), "sum_d_exp_K_mi_K_im_d_sf2": (M, M),
    }


class ShardContext(object):
    def __init__(self, M, Q, D, n_total, device=0, fixed_embeddings=False, fixed_beta=False, fp32_map=False):
        self._lib = _lib.load()
        self.M, self.Q, self.D = int(M), int(Q), int(D)
        self.n_total = int(n_total)
        self.device = int(device)
        self.fixed_embeddings = bool(fixed_embeddings)
        flags = 0
        if fp32_map:
            flags |= _lib.FLAG_FP32_MAP
        if fixed_embeddings:
            flags |= _lib.FLAG_FIXED_EMBEDDINGS
        if fixed_beta:
            flags |= _lib.FLAG_FIXED_BETA
        h = ctypes.c_void_p()
        _lib.check(self._lib.gparml_create(ctypes.byref(h), self.device, self.M, self.Q, self.D, self.n_total, flags))
        self._h = h
        self._torch_view = None
        self.last_F = None           # bound of the last finished master step

    # -- lifetime -----------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self._lib.gparml_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- plumbing -----------------------------------------------------------------
    def set_stream(self, cuda_stream_ptr):
        """Order the context's work on the given cudaStream_t (0 = CUDA's legacy default stream)."""
        _lib.check(self._lib.gparml_set_stream(self._h, ctypes.c_void_p(int(cuda_stream_ptr or 0))))

    def use_own_stream(self):
        _lib.check(self._lib.gparml_use_own_stream(self._h))

    def use_torch_stream(self):
        """Order this context's work on torch's current stream (needed when the packed
        statistics are all-reduced with torch.distributed / NCCL)."""
        import torch
        self.set_stream(torch.cuda.current_stream(self.device).cuda_stream)

    def synchronize(self):
        _lib.check(self._lib.gparml_synchronize(self._h))

    def set_n_total(self, n_total):
        self.n_total = int(n_total)
        _lib.check(self._lib.gparml_set_n_total(self._h, self.n_total))

    @property
    def n_local(self):
        return int(self._lib.gparml_n_local(self._h))

    @property
    def jitter_events(self):
        """Evaluations whose Kmm / Kmm + beta Psi2 needed the reference's 1e-7 jitter (partial_terms.py:453-457)."""
        return int(self._lib.gparml_jitter_events(self._h))

    @property
    def launch_count(self):
        return int(self._lib.gparml_launch_count(self._h))

    # -- shard + globals -------------------------------------------------------------
    def upload_shard(self, Y, X_mu, X_S, positive_variance=False):
        X_mu = _lib.as_f64(X_mu)
        n = X_mu.shape[0]
        if X_mu.ndim != 2 or X_mu.shape[1] != self.Q:
            raise ValueError("X_mu must be (n, Q=%d), got %r" % (self.Q, X_mu.shape))
        Y = _lib.as_f64(Y)
        if Y.ndim == 1:                       # local_MapReduce.py:198-199
            Y = np.ascontiguousarray(Y.reshape(-1, 1))
        Y = _lib.as_f64(Y, (n, self.D))
        X_S = _lib.as_f64(X_S, (n, self.Q))
        dom = _lib.VARIANCE_POSITIVE if positive_variance else _lib.VARIANCE_UNCONSTRAINED
        _lib.check(self._lib.gparml_upload_shard(self._h, _lib.ptr(Y), _lib.ptr(X_mu), _lib.ptr(X_S), n, dom))

    def upload_shard_ptrs(self, y_ptr, mu_ptr, s_ptr, n, positive_variance=False):
        """Same as upload_shard but from raw host pointers (e.g. pinned torch tensors)."""
        dom = _lib.VARIANCE_POSITIVE if positive_variance else _lib.VARIANCE_UNCONSTRAINED
        _lib.check(self._lib.gparml_upload_shard(self._h, ctypes.c_void_p(y_ptr), ctypes.c_void_p(mu_ptr),
                                                 ctypes.c_void_p(s_ptr), int(n), dom))

    def set_globals(self, Z, sf2, alpha, beta):
        Z = _lib.as_f64(Z, (self.M, self.Q))
        alpha = _lib.as_f64(np.atleast_1d(np.squeeze(alpha)), (self.Q,))
        _lib.check(self._lib.gparml_set_globals(self._h, _lib.ptr(Z), float(sf2), _lib.ptr(alpha), float(beta)))

    def set_step(self, step_size):
        _lib.check(self._lib.gparml_set_step(self._h, float(step_size)))

    # -- generic arrays ---------------------------------------------------------------
    def download(self, array_id, shape=None):
        n = int(self._lib.gparml_array_count(self._h, array_id))
        if n < 0:
            raise ValueError(_lib.last_error())
        out = np.empty(n, dtype=np.float64)
        _lib.check(self._lib.gparml_download(self._h, array_id, _lib.ptr(out), n))
        return out.reshape(shape) if shape is not None else out

    def download_into_ptr(self, array_id, host_ptr, count):
        _lib.check(self._lib.gparml_download(self._h, array_id, ctypes.c_void_p(host_ptr), int(count)))

    def upload(self, array_id, arr):
        arr = _lib.as_f64(arr)
        _lib.check(self._lib.gparml_upload(self._h, array_id, _lib.ptr(arr), arr.size))

    def device_ptr(self, array_id):
        p = ctypes.c_void_p()
        _lib.check(self._lib.gparml_array_device_ptr(self._h, array_id, ctypes.byref(p)))
        return p.value

    # -- map 1 ---------------------------------------------------------------------------
    def statistics(self):
        """The statistics map; blocks until the device-side input check (variances in range) is known."""
        _lib.check(self._lib.gparml_statistics(self._h))

    def statistics_launch(self):
        """Queue the statistics map and return at once (no host synchronisation); a failed input check is
        raised by the next :meth:`status` or :meth:`global_step_end`."""
        _lib.check(self._lib.gparml_statistics_launch(self._h))

    def status(self):
        """Wait for the context's stream and raise what the device status word holds (AssertionError for a
        variance out of range, LinAlgError for a failed pivot)."""
        _lib.check(self._lib.gparml_status(self._h))

    @property
    def stats_count(self):
        return int(self._lib.gparml_stats_count(self._h))

    def stats_packed(self):
        return self.download(_lib.A_STATS)

    def set_stats_packed(self, arr):
        self.upload(_lib.A_STATS, arr)

    def stats_device_ptr(self):
        p = ctypes.c_void_p()
        _lib.check(self._lib.gparml_stats_device_ptr(self._h, ctypes.byref(p)))
        return p.value

    def stats_add(self, other, scale=1.0):
        """stats = (stats + other.stats) * scale, on device (same GPU)."""
        if other.device != self.device:
            raise ValueError("stats_add needs both contexts on one device; use the NCCL all-reduce across devices")
        other.synchronize()
        _lib.check(self._lib.gparml_stats_add(self._h, ctypes.c_void_p(other.stats_device_ptr()), float(scale)))

    def stats_add_any(self, other, scale=1.0):
        """stats_add that also works when ``other`` lives on a different GPU of this process
        (stream-ordered, the host does not wait)."""
        _lib.check(self._lib.gparml_stats_add_peer(self._h, other._h, float(scale)))

    def stats_copy_from(self, other):
        """stats = other's packed buffer (any device of this process; stream-ordered, asynchronous)."""
        _lib.check(self._lib.gparml_stats_copy_peer(self._h, other._h))

    def kmm_derivative(self, which):
        shape = {0: (self.M, self.Q, self.M), 1: (self.Q, self.M, self.M), 2: (self.M, self.M)}[which]
        out = np.empty(shape, dtype=np.float64)
        _lib.check(self._lib.gparml_kmm_derivative(self._h, which, _lib.ptr(out)))
        return out

    def grad_contract(self, which, dF_dKmm, dKmm_dx, dF_dP1Y, dP1Y_dx, dF_dP2, dP2_dx):
        M, Q, D = self.M, self.Q, self.D
        lead = {0: (M, Q), 1: (Q, M), 2: (M,)}[which]
        shp_k = {0: (M, Q, M), 1: (Q, M, M), 2: (M, M)}[which]
        shp_1 = {0: (M, Q, D), 1: (Q, M, D), 2: (M, D)}[which]
        a = _lib.as_f64(dF_dKmm, (M, M)); b = _lib.as_f64(dKmm_dx, shp_k)
        c = _lib.as_f64(dF_dP1Y, (M, D)); e = _lib.as_f64(dP1Y_dx, shp_1)
        g = _lib.as_f64(dF_dP2, (M, M)); h = _lib.as_f64(dP2_dx, shp_k)
        out = np.empty({0: (M, Q), 1: (Q,), 2: (1,)}[which], dtype=np.float64)
        del lead
        _lib.check(self._lib.gparml_grad_contract(self._h, which, _lib.ptr(a), _lib.ptr(b), _lib.ptr(c), _lib.ptr(e),
                                                  _lib.ptr(g), _lib.ptr(h), _lib.ptr(out)))
        return out

    def stats_torch_view(self):
        """Zero-copy torch view (float64, cuda) of the packed device buffer, for
        ``torch.distributed.all_reduce`` over NCCL."""
        if self._torch_view is None:
            import torch

            class _Holder(object):
                pass
            h = _Holder()
            h.__cuda_array_interface__ = {
                "shape": (self.stats_count,), "typestr": "<f8", "data": (self.stats_device_ptr(), False),
                "version": 2, "strides": None,
            }
            self._torch_view = torch.as_tensor(h, device=torch.device("cuda", self.device))
        return self._torch_view

    def stats_named(self, names=STAT_FIELDS):
        shapes = stat_shapes(self.M, self.Q, self.D)
        ns = _lib.NamedStats()
        out = {}
        for k in names:
            out[k] = np.zeros(shapes[k] if shapes[k] else (1,), dtype=np.float64)
            setattr(ns, k, out[k].ctypes.data_as(ctypes.POINTER(ctypes.c_double)))
        _lib.check(self._lib.gparml_stats_expand(self._h, ctypes.byref(ns)))
        for k in names:
            if k in _SCALAR_STATS:
                out[k] = float(out[k][0])
        return out

    def set_stats_named(self, stats):
        shapes = stat_shapes(self.M, self.Q, self.D)
        ns = _lib.NamedStats()
        keep = []
        for k, v in stats.items():
            if k not in shapes:
                raise ValueError("unknown statistic %r" % k)
            if k in ("sum_d_exp_K_miY_d_sf2", "sum_d_exp_K_mi_K_im_d_sf2"):
                continue                    # pure rescalings of other statistics (partial_terms.py:310-316)
            a = _lib.as_f64(np.atleast_1d(v)).reshape(shapes[k] if shapes[k] else (1,))
            a = np.ascontiguousarray(a)
            keep.append(a)
            setattr(ns, k, a.ctypes.data_as(ctypes.POINTER(ctypes.c_double)))
        _lib.check(self._lib.gparml_stats_set_named(self._h, ctypes.byref(ns)))

    # -- master ---------------------------------------------------------------------------
    def update_global_statistics(self):
        _lib.check(self._lib.gparml_update_global_statistics(self._h))

    def global_step(self):
        """Returns (F, grad) with grad = dict(Z (M,Q), sf2, alpha (Q,), beta), positive domain."""
        F = np.zeros(1)
        g = np.zeros(self.M * self.Q + self.Q + 2)
        _lib.check(self._lib.gparml_global_step(self._h, _lib.ptr(F), _lib.ptr(g)))
        mq = self.M * self.Q
        self.last_F = float(F[0])
        return float(F[0]), {"Z": g[:mq].reshape(self.M, self.Q).copy(), "sf2": float(g[mq]),
                             "alpha": g[mq + 1:mq + 1 + self.Q].copy(), "beta": float(g[mq + 1 + self.Q]),
                             "flat": g}

    def global_step_begin(self):
        """Asynchronous first half of :meth:`global_step`: afterwards :meth:`embedding_grads` may be
        launched; the bound and the global gradients finish on a side stream next to it."""
        _lib.check(self._lib.gparml_global_step_begin(self._h))

    def global_step_end(self):
        """Second half: blocks until F and the global gradients are on the host; same return value as
        :meth:`global_step`."""
        F = np.zeros(1)
        g = np.zeros(self.M * self.Q + self.Q + 2)
        _lib.check(self._lib.gparml_global_step_end(self._h, _lib.ptr(F), _lib.ptr(g)))
        mq = self.M * self.Q
        self.last_F = float(F[0])
        return float(F[0]), {"Z": g[:mq].reshape(self.M, self.Q).copy(), "sf2": float(g[mq]),
                             "alpha": g[mq + 1:mq + 1 + self.Q].copy(), "beta": float(g[mq + 1 + self.Q]),
                             "flat": g}

    # -- map 2 ----------------------------------------------------------------------------
    def embedding_grads(self):
        _lib.check(self._lib.gparml_embedding_grads(self._h))

    def embedding_grads_into(self, host_ptr, chunks=4):
        """embedding_grads + download of grad_latest (2, n, Q) into host memory at ``host_ptr``
        (e.g. a pinned torch tensor); the copy of one point range overlaps the next range's kernels."""
        _lib.check(self._lib.gparml_embedding_grads_download(self._h, ctypes.c_void_p(host_ptr), int(chunks)))

    def embedding_grads_numpy(self, chunks=4):
        out = np.empty((2, self.n_local, self.Q), dtype=np.float64)
        self.embedding_grads_into(out.ctypes.data, chunks)
        return out

    def grad_latest(self):
        return self.download(_lib.A_GRAD_LATEST, (2, self.n_local, self.Q))

    # -- optimiser local state -----------------------------------------------------------------
    def _scalar(self, fn, *args):
        out = ctypes.c_double()
        _lib.check(fn(self._h, *args, ctypes.byref(out)))
        return float(out.value)

    def scg_set_grads(self): _lib.check(self._lib.gparml_scg_set_grads(self._h))
    def scg_get_mu(self): return self._scalar(self._lib.gparml_scg_get_mu)
    def scg_get_kappa(self): return self._scalar(self._lib.gparml_scg_get_kappa)
    def scg_get_theta(self): return self._scalar(self._lib.gparml_scg_get_theta)
    def scg_get_current_grad(self): return self._scalar(self._lib.gparml_scg_get_current_grad)
    def scg_get_gamma(self): return self._scalar(self._lib.gparml_scg_get_gamma)
    def scg_get_max_d(self, alpha): return self._scalar(self._lib.gparml_scg_get_max_d, float(alpha))
    def scg_reset_d(self): _lib.check(self._lib.gparml_scg_reset_d(self._h))
    def scg_update_d(self, gamma): _lib.check(self._lib.gparml_scg_update_d(self._h, float(gamma)))
    def scg_update_X(self, alpha): _lib.check(self._lib.gparml_scg_update_X(self._h, float(alpha)))
    def scg_update_grad_old(self): _lib.check(self._lib.gparml_scg_update_grad_old(self._h))
    def scg_update_grad_new(self): _lib.check(self._lib.gparml_scg_update_grad_new(self._h))

    # -- probes ----------------------------------------------------------------------------------
    def enable_timing(self, on=True):
        _lib.check(self._lib.gparml_enable_timing(self._h, 1 if on else 0))

    def phase_times_ms(self):
        out = (ctypes.c_double * 5)()
        _lib.check(self._lib.gparml_phase_times(self._h, out))
        return dict(zip(("prep_points", "psi1_stats", "psi2_stats", "global_step", "embed_grads"), list(out)))

    def measure_dfma_peak(self):
        out = ctypes.c_double()
        _lib.check(self._lib.gparml_measure_dfma_peak(self._h, ctypes.byref(out)))
        return float(out.value)

    # -- one-off initialisation on the device (see init_device.py) --------
    def upload_outputs(self, Y):
        Y = _lib.as_f64(Y)
        if Y.ndim == 1:
            Y = np.ascontiguousarray(Y.reshape(-1, 1))
        Y = _lib.as_f64(Y, (Y.shape[0], self.D))
        _lib.check(self._lib.gparml_upload_outputs(self._h, _lib.ptr(Y), Y.shape[0]))

    def init_column_sums(self):
        out = np.empty(self.D)
        _lib.check(self._lib.gparml_init_column_sums(self._h, _lib.ptr(out)))
        return out

    def init_scatter(self, mean):
        mean = _lib.as_f64(mean, (self.D,))
        out = np.empty((self.D, self.D))
        _lib.check(self._lib.gparml_init_scatter(self._h, _lib.ptr(mean), _lib.ptr(out)))
        return out

    def init_project(self, mean, W):
        mean = _lib.as_f64(mean, (self.D,))
        W = _lib.as_f64(W, (self.D, self.Q))
        _lib.check(self._lib.gparml_init_project(self._h, _lib.ptr(mean), _lib.ptr(W)))

    def init_random_variances(self, seed, row_offset=0):
        _lib.check(self._lib.gparml_init_random(self._h, 0, int(seed) & (2 ** 64 - 1), int(row_offset)))

    def init_random_means(self, seed, row_offset=0):
        _lib.check(self._lib.gparml_init_random(self._h, 1, int(seed) & (2 ** 64 - 1), int(row_offset)))

    def kmeans_step(self, centroids):
        """-> (counts (k,), sums (k, Q), summed distance to the nearest centroid)."""
        centroids = _lib.as_f64(centroids)
        k = centroids.shape[0]
        centroids = _lib.as_f64(centroids, (k, self.Q))
        out = np.empty(k * (1 + self.Q) + 1)
        _lib.check(self._lib.gparml_kmeans_step(self._h, _lib.ptr(centroids), k, _lib.ptr(out)))
        t = out[:-1].reshape(k, 1 + self.Q)
        return t[:, 0].copy(), t[:, 1:].copy(), float(out[-1])


def allreduce_contexts(contexts, scale=1.0):
    """The reducer for shard contexts driven by this host thread (any mix of GPUs): afterwards EVERY
    context's packed buffer holds ``scale`` * the sum over all of them (``gparml_stats_allreduce_peers``:
    one kernel over NVLink peer memory, no host synchronisation)."""
    lib = _lib.load()
    n = len(contexts)
    if n > _lib.MAX_PEERS:               # two levels: groups of MAX_PEERS, then the group heads, then hand the sum back
        groups = [contexts[i:i + _lib.MAX_PEERS] for i in range(0, n, _lib.MAX_PEERS)]
        for g in groups:
            allreduce_contexts(g, 1.0)
        allreduce_contexts([g[0] for g in groups], scale)
        for g in groups:
            for c in g[1:]:
                c.stats_copy_from(g[0])
        return
    arr = (ctypes.c_void_p * n)(*[c._h for c in contexts])
    _lib.check(lib.gparml_stats_allreduce_peers(arr, n, float(scale)))


def evaluate(contexts, Z, sf2, alpha, beta, step_size=0.0, reduce_fn=None):
    """One ELBO + gradient evaluation over shard contexts living in THIS process
    (SURVEY.md 3.2 steps 4-9).  ``contexts`` share one device unless ``reduce_fn`` is
    given; ``reduce_fn(ctx)`` must sum the packed buffer across ranks in place (e.g. an
    NCCL all-reduce of ``ctx.stats_torch_view()``), in which case ``contexts`` holds this
    rank's single shard.

    Returns (F, grad dict).  Per-point gradients stay on device (``ctx.grad_latest()``).
    """
    for c in contexts:                      # every shard's map is queued before the host waits for any
        c.set_globals(Z, sf2, alpha, beta)
        c.set_step(step_size)
        c.statistics_launch()
    root = contexts[0]
    if len(contexts) > 1:
        allreduce_contexts(contexts)        # every context now holds the sum (local_MapReduce.py:250-277)
    if reduce_fn is not None:
        reduce_fn(root)
        for c in contexts[1:]:
            c.stats_copy_from(root)
    if root.fixed_embeddings:
        return root.global_step()
    # the embeddings map only waits for the partial derivatives; F and the global gradients are finished
    # and downloaded concurrently with it.  The master step is replicated on every context (identical
    # inputs, identical outputs: local_MapReduce.py:315-321,350-354 read the same files in every mapper)
    for c in contexts:
        c.global_step_begin()
        c.embedding_grads()
    out = root.global_step_end()
    for c in contexts[1:]:
        c.global_step_end()
    return out
