"""Device-side initialisation against the reference's host algorithms:
PCA by SVD (supporting_functions.py:102-121), the variance draw (local_MapReduce.py:88-93) and
scipy's k-means (parallel_GPLVM.py:181)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def relerr(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def ref_pca(Y, Q):
    """supporting_functions.py:116-121, verbatim algorithm."""
    Z = np.linalg.svd(Y - Y.mean(axis=0), full_matrices=False)
    X = Z[0][:, 0:Q]
    return X / X.std(axis=0)


def _contexts(Y, Q, M, splits):
    from gparml_b200.engine import ShardContext
    ctxs = []
    for lo, hi in splits:
        c = ShardContext(M, Q, Y.shape[1], Y.shape[0])
        c.upload_outputs(Y[lo:hi])
        ctxs.append(c)
    return ctxs


@pytest.mark.parametrize("N,D,Q,splits", [
    (5000, 7, 3, [(0, 1700), (1700, 1701), (1701, 5000)]),
    (3000, 50, 10, [(0, 3000)]),
    (777, 300, 4, [(0, 400), (400, 777)]),           # D > 256: the wide-column path
    (640, 4, 4, [(0, 640)]),
])
def test_pca_matches_svd(N, D, Q, splits):
    from gparml_b200 import _lib, init_device
    rng = np.random.default_rng(5)
    Y = rng.standard_normal((N, min(D, 12))) @ rng.standard_normal((min(D, 12), D)) + 0.1 * rng.standard_normal((N, D)) + 3.0
    ctxs = _contexts(Y, Q, 5, splits)
    try:
        mean, W = init_device.pca(ctxs)
        X = np.concatenate([c.download(_lib.A_X_MU, (c.n_local, Q)) for c in ctxs])
    finally:
        for c in ctxs:
            c.close()
    ref = ref_pca(Y, Q)
    sign = np.sign(np.sum(X * ref, axis=0))          # an SVD's column signs are arbitrary
    assert relerr(X * sign, ref) < 1e-9
    assert relerr(mean, Y.mean(axis=0)) < 1e-13
    assert np.allclose(X.std(axis=0), 1.0, rtol=1e-10)


def test_partial_sums_match_numpy():
    rng = np.random.default_rng(6)
    Y = rng.standard_normal((4099, 10)) * 2 + 1
    ctxs = _contexts(Y, 2, 3, [(0, 4099)])
    try:
        c = ctxs[0]
        assert relerr(c.init_column_sums(), Y.sum(axis=0)) < 1e-13
        m = Y.mean(axis=0)
        S = c.init_scatter(m)
        assert relerr(S, (Y - m).T @ (Y - m)) < 1e-13
        assert np.array_equal(S, S.T)
    finally:
        c.close()


def test_random_draws_are_sharding_independent_and_distributed_like_the_reference():
    from gparml_b200 import _lib, init_device, transforms as sp
    N, Q = 200000, 5
    Y = np.zeros((N, 1))
    one = _contexts(Y, Q, 3, [(0, N)])
    two = _contexts(Y, Q, 3, [(0, 70001), (70001, N)])
    try:
        init_device.random_variances(one, 1234)
        init_device.random_variances(two, 1234)
        init_device.random_means(one, 99)
        init_device.random_means(two, 99)
        S1 = one[0].download(_lib.A_X_S, (N, Q))
        S2 = np.concatenate([c.download(_lib.A_X_S, (c.n_local, Q)) for c in two])
        M1 = one[0].download(_lib.A_X_MU, (N, Q))
        M2 = np.concatenate([c.download(_lib.A_X_MU, (c.n_local, Q)) for c in two])
        init_device.random_variances(one, 1235)
        S3 = one[0].download(_lib.A_X_S, (N, Q))
    finally:
        for c in one + two:
            c.close()
    assert np.array_equal(S1, S2) and np.array_equal(M1, M2)
    assert not np.array_equal(S1, S3)
    pos = sp.transformVar(S1)                         # local_MapReduce.py:90-93: clip(0.5 + 0.01 randn, 0.001, 1)
    assert pos.min() >= 0.001 - 1e-12 and pos.max() <= 1 + 1e-12
    assert abs(pos.mean() - 0.5) < 1e-4 and abs(pos.std() - 0.01) < 1e-4
    z = (pos - 0.5) / 0.01
    assert abs(np.mean(z ** 3)) < 0.02 and abs(np.mean(z ** 4) - 3.0) < 0.05
    assert abs(M1.mean()) < 5e-3 and abs(M1.std() - 1.0) < 5e-3 and abs(np.mean(M1 ** 4) - 3.0) < 0.05
    assert abs(np.corrcoef(M1[:-1].ravel(), M1[1:].ravel())[0, 1]) < 5e-3


def test_kmeans_matches_scipy():
    import scipy.cluster.vq as cl
    from gparml_b200 import init_device
    from gparml_b200.engine import ShardContext
    rng = np.random.default_rng(7)
    Q, k, N = 3, 12, 6000
    centres = rng.standard_normal((k, Q)) * 4
    X = centres[rng.integers(0, k, N)] + 0.3 * rng.standard_normal((N, Q))
    ctxs = []
    try:
        for lo, hi in [(0, 2500), (2500, N)]:
            c = ShardContext(k, Q, 1, N)
            c.upload_shard(np.zeros((hi - lo, 1)), X[lo:hi], np.zeros((hi - lo, Q)))
            ctxs.append(c)
        guess = X[rng.choice(N, k, replace=False)]
        # one assignment pass against numpy
        d = np.sqrt(((X[:, None, :] - guess[None]) ** 2).sum(-1))
        code = d.argmin(axis=1)
        counts = sum(c.kmeans_step(guess)[0] for c in ctxs)
        sums = sum(c.kmeans_step(guess)[1] for c in ctxs)
        dist = sum(c.kmeans_step(guess)[2] for c in ctxs)
        assert np.array_equal(counts, np.bincount(code, minlength=k))
        ref_sums = np.zeros((k, Q))
        np.add.at(ref_sums, code, X)
        assert relerr(sums, ref_sums) < 1e-12
        assert abs(dist - d.min(axis=1).sum()) < 1e-10 * dist
        # Lloyd iterations from the same guess == scipy.cluster.vq.kmeans(obs, guess)
        book, distortion = init_device.kmeans_from_guess(ctxs, guess)
        ref_book, ref_dist = cl.kmeans(X, guess)
        assert book.shape == ref_book.shape and relerr(book, ref_book) < 1e-9
        assert abs(distortion - ref_dist) < 1e-9 * ref_dist
        # the full routine: 20 random starts, a code book with low distortion and no empty cluster
        book2, dist2 = init_device.kmeans(ctxs, k, rng=np.random.RandomState(3))
        assert book2.shape[1] == Q and book2.shape[0] <= k and dist2 <= ref_dist * 1.5
    finally:
        for c in ctxs:
            c.close()


def test_init_writes_reference_files_and_keeps_the_session(tmp_path):
    """b200_MapReduce.init with the device initialisation: the files of local_MapReduce.py:75-93
    exist with the reference's shapes / domains, and the evaluation runs on the session init built."""
    from gparml_b200 import b200_MapReduce, parallel_GPLVM as drv, transforms as sp
    from gparml_b200.synthetic import make_problem, split_rows
    p = make_problem(900, 6, 3, 5, seed=4)
    dirs = {}
    for k in ("input", "embeddings", "statistics", "tmp"):
        dirs[k] = str(tmp_path / k)
        os.makedirs(dirs[k])
    for i, (lo, hi) in enumerate(split_rows(900, 3)):
        np.savetxt(os.path.join(dirs["input"], "part_%d" % i), p["Y"][lo:hi], delimiter=",")
    np.random.seed(2)
    opts = drv.default_options(M=6, Q=3, D=5, iterations=1, init="PCA", display=False, **dirs)
    try:
        opts = b200_MapReduce.init(opts)
        assert opts["N"] == 900
        X = np.concatenate([np.load(os.path.join(dirs["embeddings"], "part_%d.embedding.npy" % i)) for i in range(3)])
        ref = ref_pca(p["Y"], 3)
        assert relerr(X * np.sign(np.sum(X * ref, axis=0)), ref) < 1e-9
        V = np.concatenate([np.load(os.path.join(dirs["embeddings"], "part_%d.variance.npy" % i)) for i in range(3)])
        pos = sp.transformVar(V)
        assert V.shape == (900, 3) and pos.min() >= 0.001 - 1e-12 and pos.max() <= 1 + 1e-12 and abs(pos.mean() - 0.5) < 2e-3
        ctx_before = list(b200_MapReduce.session_contexts(dirs["embeddings"]))
        opts, gs = drv.init_statistics(b200_MapReduce, opts)
        assert gs["Z"].shape == (6, 3) and np.all(np.isfinite(gs["Z"]))
        x0 = drv.flatten_global_statistics(opts, gs)
        x0 = np.array([drv.sp.transform_back(b, x) for b, x in zip(opts["flat_global_statistics_bounds"], x0)])
        drv.options, drv.map_reduce = opts, b200_MapReduce
        f, g = drv.likelihood_and_gradient(x0, 0, 0)
        assert np.isfinite(f) and np.all(np.isfinite(g))
        assert b200_MapReduce.session_contexts(dirs["embeddings"]) == ctx_before     # no second CSV parse / upload
        # host initialisation stays available
        opts2 = dict(opts, b200_device_init=False)
        b200_MapReduce.init(opts2)
        X2 = np.concatenate([np.load(os.path.join(dirs["embeddings"], n + ".embedding.npy")) for n in sorted(os.listdir(dirs["input"]))])
        assert X2.shape == (900, 3)
    finally:
        b200_MapReduce.close()
