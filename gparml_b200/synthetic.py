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


# ------------------------------------------------------------------------------------------------
# the same kind of problem generated in independent row blocks (bench.py): a rank builds only the
# rows it owns, and the data -- hence the bound F -- is identical for 1, 2, 4 or 8 ranks
# ------------------------------------------------------------------------------------------------
ROW_BLOCKS = 8


def block_problem_globals(cfg_name):
    """Everything that does not depend on a row: hyper-parameters (the reference's initial values sf2 = alpha =
    beta = 1, parallel_GPLVM.py:189-194), the inducing inputs (M draws from the distribution of the means; the
    reference takes k-means centres of the embeddings plus 0.05 noise, :180-186) and the output maps."""
    k = CONFIGS[cfg_name]
    M, Q, D = k["M"], k["Q"], k["D"]
    rng = np.random.default_rng([BASE_SEED, int(cfg_name[1]), 0])
    lin = rng.standard_normal((D, Q))
    proj = rng.standard_normal((D, Q))
    phase = 5.0 * rng.standard_normal(D)
    Z = rng.standard_normal((M, Q)) * np.sqrt(1.0 + 0.05 ** 2) + 0.05 * rng.standard_normal((M, Q))
    return dict(lin=lin, proj=proj, phase=phase, Z=np.ascontiguousarray(Z), sf2=1.0, alpha=np.ones(Q), beta=1.0,
                N=k["N"], M=M, Q=Q, D=D, fixed_embeddings=k["fixed_embeddings"])


def block_problem_rows(cfg_name, lo, hi, with_direction=True):
    """Rows [lo, hi) of the block-generated problem of BASELINE config ``cfg_name``: dict with Y, X_mu, X_S
    (unconstrained; zeros with fixed embeddings) and optionally d (2, n, Q).  [lo, hi) must be a union of whole
    blocks (N / ROW_BLOCKS rows each), which every split over 1, 2, 4 or 8 ranks is."""
    g = block_problem_globals(cfg_name)
    N, Q, D = g["N"], g["Q"], g["D"]
    assert N % ROW_BLOCKS == 0
    bs = N // ROW_BLOCKS
    assert lo % bs == 0 and hi % bs == 0 and 0 <= lo <= hi <= N, (lo, hi, bs)
    scale = 1.0 / np.sqrt(1.03 ** 2 * np.sum(g["lin"] ** 2, axis=1) / Q + 0.125 + 0.05 ** 2)     # unit variance columns
    Ys, MUs, Ss, Ds = [], [], [], []
    for b in range(lo // bs, hi // bs):
        rng = np.random.default_rng([BASE_SEED, int(cfg_name[1]), 1 + b])
        Xs = rng.standard_normal((bs, Q))
        Y = 1.03 * (Xs @ g["lin"].T) / np.sqrt(Q) + 0.5 * np.sin(2.0 * (Xs @ g["proj"].T) / np.sqrt(Q) + g["phase"])
        Y += 0.05 * rng.standard_normal((bs, D))
        Ys.append(Y * scale)
        if g["fixed_embeddings"]:
            MUs.append(Xs)
            Ss.append(np.zeros((bs, Q)))
        else:
            MUs.append(Xs + 0.05 * rng.standard_normal((bs, Q)))
            Ss.append(softplus_inv(np.clip(0.5 + 0.01 * rng.standard_normal((bs, Q)), 0.001, 1.0)))
            if with_direction:
                Ds.append(rng.standard_normal((2, bs, Q)))
    cat = lambda parts, axis=0: np.ascontiguousarray(np.concatenate(parts, axis=axis)) if parts else None   # noqa: E731
    out = dict(Y=cat(Ys), X_mu=cat(MUs), X_S=cat(Ss))
    if Ds:
        out["d"] = cat(Ds, 1)
    return out
