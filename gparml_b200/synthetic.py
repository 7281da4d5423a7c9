"""Seeded synthetic GPLVM / sparse-GP problems.

The data generator generalises the reference's toy generator
(``tools/easy_dataset.py:12-23``: per output column a random linear map of the
latent point plus half a sine of a random projection plus 0.05 noise) to Q latent
dimensions; the variational initialisation mirrors ``local_MapReduce.py:90-93``
(variance 0.5 +- 0.01 clipped to [0.001, 1], stored in the softplus^-1 domain) and
the inducing inputs mirror ``parallel_GPLVM.py:180-186`` without the k-means step
(M distinct rows of the means plus 0.05 noise).

Pure numpy host code; used by tests, ``bench.py`` and the golden generator.
"""
import numpy as np

BASE_SEED = 20141208

# BASELINE.json configs (c1..c5); N is the *total* number of points.
CONFIGS = {
    "c1": dict(N=1000, M=2, Q=2, D=4, fixed_embeddings=False),
    "c2": dict(N=100000, M=50, Q=4, D=1, fixed_embeddings=True),
    "c3": dict(N=1000000, M=100, Q=10, D=10, fixed_embeddings=False),
    "c4": dict(N=1000000, M=500, Q=10, D=50, fixed_embeddings=False),
    "c5": dict(N=10000000, M=100, Q=10, D=10, fixed_embeddings=False),
}


def softplus_inv(y):
    return np.log(np.expm1(y))


def make_problem(N, M, Q, D, seed=0, fixed_embeddings=False, generic_hypers=False,
                 with_direction=False, dtype=np.float64):
    """Return a dict with Y (N,D), X_mu (N,Q), X_S (N,Q; unconstrained unless
    fixed_embeddings, then zeros), Z (M,Q), sf2, alpha (Q,), beta and optionally a
    local search direction d (2,N,Q).

    ``generic_hypers=False`` gives the reference's initial values sf2=alpha=beta=1
    (``parallel_GPLVM.py:189-194``); ``True`` gives sf2=1.7, alpha~U(0.5,1.5),
    beta=2.5 so that no term hides behind a unit factor in the parity runs.
    """
    rng = np.random.default_rng(BASE_SEED + seed)
    Xs = rng.standard_normal((N, Q))
    Y = np.empty((N, D))
    for d in range(D):
        lin = rng.standard_normal(Q)
        proj = rng.standard_normal(Q)
        phase = 5.0 * rng.standard_normal()
        Y[:, d] = 1.03 * (Xs @ lin) / np.sqrt(Q) + 0.5 * np.sin(2.0 * (Xs @ proj) / np.sqrt(Q) + phase)
    Y += 0.05 * rng.standard_normal((N, D))
    Y = (Y - Y.mean(axis=0)) / Y.std(axis=0)
    if fixed_embeddings:
        X_mu = Xs.copy()
        X_S = np.zeros((N, Q))
    else:
        X_mu = Xs + 0.05 * rng.standard_normal((N, Q))
        X_S = softplus_inv(np.clip(0.5 + 0.01 * rng.standard_normal((N, Q)), 0.001, 1.0))
    idx = rng.choice(N, size=M, replace=False) if N >= M else rng.integers(0, N, size=M)
    Z = X_mu[idx] + 0.05 * rng.standard_normal((M, Q))
    if generic_hypers:
        sf2, alpha, beta = 1.7, rng.uniform(0.5, 1.5, size=Q), 2.5
    else:
        sf2, alpha, beta = 1.0, np.ones(Q), 1.0
    out = dict(Y=np.ascontiguousarray(Y, dtype=dtype), X_mu=np.ascontiguousarray(X_mu, dtype=dtype),
               X_S=np.ascontiguousarray(X_S, dtype=dtype), Z=np.ascontiguousarray(Z, dtype=dtype),
               sf2=float(sf2), alpha=np.ascontiguousarray(alpha, dtype=dtype), beta=float(beta),
               N=N, M=M, Q=Q, D=D, fixed_embeddings=fixed_embeddings)
    if with_direction:
        out["d"] = rng.standard_normal((2, N, Q))
    return out


def split_rows(n, parts):
    """Contiguous row ranges [lo, hi) of an n-row problem split into ``parts`` shards
    (sizes differ by at most one)."""
    base, rem = divmod(n, parts)
    out, lo = [], 0
    for r in range(parts):
        hi = lo + base + (1 if r < rem else 0)
        out.append((lo, hi))
        lo = hi
    return out
