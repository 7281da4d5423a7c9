"""One-off initialisation on the device.

The reference initialises on the master: it concatenates ALL outputs and takes an SVD
(``local_MapReduce.py:52-65`` -> ``supporting_functions.PCA`` ``:102-121``), draws the initial
variances (``local_MapReduce.py:88-93``) and runs ``scipy.cluster.vq.kmeans`` on the first
shard's embeddings for the inducing inputs (``parallel_GPLVM.py:170-186``).  At N = 10M that is
the wall-clock bottleneck of a run.  Here the shards stay on their GPUs; each returns small
partial sums (D column sums, a D x D scatter matrix, k (1+Q) cluster sums) which are added over
shards -- and over ranks through ``reduce_fn`` -- and only a D x D eigen-problem runs on the host.

``reduce_fn(array) -> array`` sums a small float64 numpy array over ranks (identity when all
shards live in this process); see ``distributed.allreduce_numpy``.
"""
import collections

import numpy

from . import _lib


def _sum(parts, reduce_fn):
    tot = parts[0].copy()
    for p in parts[1:]:
        tot += p
    return reduce_fn(tot) if reduce_fn is not None else tot


def pca(contexts, reduce_fn=None):
    """PCA initialisation of X_mu on every context (outputs already uploaded).

    Same result as ``supporting_functions.PCA`` (:102-121) up to the sign of each latent column
    (an SVD's signs are arbitrary): with Yc = Y - mean = U S V^T the reference returns
    ``U[:, :Q] / std`` = ``Yc v_q sqrt(N / lambda_q)``, lambda_q the q-th largest eigenvalue of
    the scatter matrix Yc^T Yc.  The sign is fixed by making the largest-magnitude entry of each
    v_q positive.  Returns ``(mean, W)`` with ``X_mu = (Y - mean) W``.
    """
    Q, D = contexts[0].Q, contexts[0].D
    if D < Q:
        raise ValueError("PCA initialisation needs D >= Q (D=%d, Q=%d)" % (D, Q))
    head = _sum([numpy.concatenate((c.init_column_sums(), [float(c.n_local)])) for c in contexts], reduce_fn)
    N = head[-1]
    mean = head[:-1] / N
    scatter = _sum([c.init_scatter(mean) for c in contexts], reduce_fn)
    lam, V = numpy.linalg.eigh(scatter)
    order = numpy.argsort(lam)[::-1][:Q]
    lam, V = lam[order], V[:, order]
    if not numpy.all(lam > 0):
        raise numpy.linalg.LinAlgError("PCA initialisation: the outputs span fewer than Q=%d directions" % Q)
    sign = numpy.sign(V[numpy.abs(V).argmax(axis=0), numpy.arange(Q)])
    W = V * sign * numpy.sqrt(N / lam)
    for c in contexts:
        c.init_project(mean, W)
    return mean, W


def random_variances(contexts, seed, row_offsets=None):
    """X_S (unconstrained domain) = transformVar_back(clip(0.5 + 0.01 N(0,1), 0.001, 1))
    (local_MapReduce.py:90-93).  ``row_offsets``: global index of each context's first row
    (default: contexts are consecutive from 0) -- the draw is a function of the global row."""
    for c, off in zip(contexts, _offsets(contexts, row_offsets)):
        c.init_random_variances(seed, off)


def random_means(contexts, seed, row_offsets=None):
    """X_mu = N(0,1) (``--init random``, local_MapReduce.py:86-87)."""
    for c, off in zip(contexts, _offsets(contexts, row_offsets)):
        c.init_random_means(seed, off)


def _offsets(contexts, row_offsets):
    if row_offsets is not None:
        return list(row_offsets)
    out, lo = [], 0
    for c in contexts:
        out.append(lo)
        lo += c.n_local
    return out


def kmeans_from_guess(contexts, guess, thresh=1e-5, reduce_fn=None, max_passes=10000):
    """scipy.cluster.vq._kmeans: Lloyd iterations from ``guess`` until the mean distance to the
    nearest centroid changes by at most ``thresh``; empty clusters are dropped.  One device pass
    over the embeddings per iteration.  Returns ``(code_book, mean_distance)``."""
    book = numpy.array(guess, dtype=numpy.float64, copy=True)
    Q = book.shape[1]
    prev = collections.deque([numpy.inf], maxlen=2)
    diff = numpy.inf
    passes = 0
    def one_pass(book):
        k = book.shape[0]
        parts = []
        for c in contexts:
            counts, sums, dist = c.kmeans_step(book)
            parts.append(numpy.concatenate((counts, sums.ravel(), [dist, float(c.n_local)])))
        tot = _sum(parts, reduce_fn)
        return tot[:k], tot[k:k + k * Q].reshape(k, Q), tot[-2] / tot[-1]

    while diff > thresh and passes < max_passes:
        counts, sums, avg = one_pass(book)
        prev.append(avg)
        has = counts > 0
        book = sums[has] / counts[has, None]
        diff = abs(prev[0] - prev[1])
        passes += 1
    return book, one_pass(book)[2]          # scipy reports the distortion of the final code book


def kmeans(contexts, k, iter=20, thresh=1e-5, rng=None, reduce_fn=None, candidates=None):
    """scipy.cluster.vq.kmeans(obs, k): ``iter`` runs from random distinct observations, the
    code book with the lowest distortion wins (parallel_GPLVM.py:181).  ``candidates``: the rows
    the initial guesses are drawn from (default: the first context's embeddings; with several
    ranks pass the same array everywhere so that every rank follows the same sequence)."""
    rng = numpy.random if rng is None else rng
    if candidates is None:
        c0 = contexts[0]
        candidates = c0.download(_lib.A_X_MU, (c0.n_local, c0.Q))
    if candidates.shape[0] < k:
        raise ValueError("kmeans: %d candidate rows for k=%d" % (candidates.shape[0], k))
    best_book, best = None, numpy.inf
    for _ in range(iter):
        guess = candidates[rng.choice(candidates.shape[0], size=k, replace=False)]
        book, dist = kmeans_from_guess(contexts, guess, thresh=thresh, reduce_fn=reduce_fn)
        if dist < best:
            best_book, best = book, dist
    return best_book, best
