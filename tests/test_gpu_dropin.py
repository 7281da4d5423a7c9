"""GPU tests of the drop-in boundary: the ``partial_terms`` mirror driven the way the
reference's own tests drive the original (test.py), the ``b200_MapReduce`` module driven by
the replayed ``parallel_GPLVM`` protocol, and the device-resident optimiser state."""
import os

import numpy as np
import pytest

from conftest import load_golden, relerr

pytestmark = pytest.mark.gpu


def _mirror_from_golden(name):
    from gparml_b200.partial_terms import partial_terms
    from oracle import gparml_oracle as O
    g = load_golden(name)
    sh = g["shards"][0]
    N, D = sh["Y"].shape
    M, Q = g["Z"].shape
    pt = partial_terms(g["Z"].copy(), g["sf2"], g["alpha"].copy(), g["beta"], M, Q, N, D)
    S = O.softplus(sh["X_S"])
    pt.set_data(sh["Y"], sh["X_mu"], S, is_set_statistics=True)
    return g, sh, S, pt


def test_partial_terms_mirror_matches_reference_golden():
    """Same call sequence as test.py:setUp + scg_adapted-example.py:140-211 on the fixture
    D=7, Q=2, N=5, M=10."""
    g, sh, S, pt = _mirror_from_golden("t5")
    M, Q = g["Z"].shape
    assert relerr(pt.Kmm, g["glob_Kmm"]) < 1e-12 and relerr(pt.Kmm_inv, g["glob_Kmm_inv"]) < 1e-9
    assert relerr(pt.sum_exp_K_mi_K_im, g["stat_sum_exp_K_mi_K_im"]) < 1e-12
    assert relerr(pt.exp_K_miY, g["stat_sum_exp_K_miY"]) < 1e-12
    assert abs(pt.KL - float(g["stat_sum_KL"])) < 1e-12 * abs(float(g["stat_sum_KL"]))
    assert abs(pt.logmarglik() - float(g["glob_F"])) < 1e-9 * abs(float(g["glob_F"]))
    dF_dKmm, dF_d1, dF_d2, dF_d0 = pt.dF_dKmm(), pt.dF_dexp_K_miY(), pt.dF_dexp_K_mi_K_im(), pt.dF_dexp_K_ii()
    assert relerr(dF_dKmm, g["glob_dF_dKmm"]) < 1e-9
    assert relerr(dF_d1, g["glob_dF_dsum_exp_K_miY"]) < 1e-9
    assert relerr(dF_d2, g["glob_dF_dsum_exp_K_mi_K_im"]) < 1e-9
    gZ = pt.grad_Z(dF_dKmm, pt.dKmm_dZ(), dF_d1, pt.dexp_K_miY_dZ(), dF_d2, pt.dexp_K_mi_K_im_dZ())
    ga = pt.grad_alpha(dF_dKmm, pt.dKmm_dalpha(), dF_d1, pt.dexp_K_miY_dalpha(), dF_d2, pt.dexp_K_mi_K_im_dalpha())
    gs = pt.grad_sf2(dF_dKmm, pt.dKmm_dsf2(), dF_d0, pt.dexp_K_ii_dsf2(), dF_d1, pt.dexp_K_miY_dsf2(), dF_d2,
                     pt.dexp_K_mi_K_im_dsf2())
    assert relerr(gZ, g["glob_grad_Z"]) < 1e-9
    assert relerr(ga, g["glob_grad_alpha"]) < 1e-9
    assert relerr(gs, g["glob_grad_sf2"]) < 1e-9
    assert relerr(pt.grad_beta(), g["glob_grad_beta"]) < 1e-9
    from oracle import gparml_oracle as O
    gl = -np.array([pt.grad_X_mu(), pt.grad_X_S() * O.softplus_grad(sh["X_S"])])
    assert relerr(gl, g["grad_latest_0"]) < 1e-9
    assert relerr(pt.exp_K_mi, O.psi1(g["Z"], g["sf2"], g["alpha"], sh["X_mu"], S)) < 1e-12
    assert pt.hyp.sf == pytest.approx(g["sf2"] ** 0.5) and np.allclose(pt.hyp.ard, g["alpha"] ** -0.5)
    pt.close()


def test_partial_terms_mirror_finite_differences_like_reference_tests():
    """test.py:62-93 (Z), :150-184 (sf via hyp.sf += d), :186-201 (beta), :270-296 (mu, S):
    attribute pokes + set_data + update_global_statistics, forward differences, 1 % bar."""
    g, sh, S, pt = _mirror_from_golden("t5")
    M, Q = g["Z"].shape
    F0 = pt.logmarglik()
    dF_dKmm, dF_d1, dF_d2 = pt.dF_dKmm(), pt.dF_dexp_K_miY(), pt.dF_dexp_K_mi_K_im()
    gZ = pt.grad_Z(dF_dKmm, pt.dKmm_dZ(), dF_d1, pt.dexp_K_miY_dZ(), dF_d2, pt.dexp_K_mi_K_im_dZ())
    g_beta = pt.grad_beta()
    g_mu, g_S = pt.grad_X_mu(), pt.grad_X_S()
    h = 1e-6
    Zp = g["Z"].copy(); Zp[4, 1] += h
    pt.Z = Zp
    pt.set_data(sh["Y"], sh["X_mu"], S, is_set_statistics=True)
    pt.update_global_statistics()
    assert abs((pt.logmarglik() - F0) / h - gZ[4, 1]) < 0.01 * abs(gZ[4, 1])
    pt.Z = g["Z"].copy()
    pt.beta += h
    pt.set_data(sh["Y"], sh["X_mu"], S, is_set_statistics=True)
    pt.update_global_statistics()
    assert abs((pt.logmarglik() - F0) / h - g_beta) < 0.01 * abs(g_beta)
    pt.beta -= h
    mu = sh["X_mu"].copy(); mu[1, 0] += h
    pt.set_data(sh["Y"], mu, S, is_set_statistics=True)
    assert abs((pt.logmarglik() - F0) / h - g_mu[1, 0]) < 0.01 * abs(g_mu[1, 0])
    S2 = S.copy(); S2[2, 1] += h
    pt.set_data(sh["Y"], sh["X_mu"], S2, is_set_statistics=True)
    assert abs((pt.logmarglik() - F0) / h - g_S[2, 1]) < 0.01 * abs(g_S[2, 1])
    pt.close()


def _write_problem(tmp_path, p, parts):
    from gparml_b200.synthetic import split_rows
    dirs = {}
    for d in ("input", "embeddings", "statistics", "tmp"):
        (tmp_path / d).mkdir()
        dirs[d] = str(tmp_path / d)
    for i, (lo, hi) in enumerate(split_rows(p["N"], parts)):
        np.savetxt(os.path.join(dirs["input"], "easy_%d" % i), p["Y"][lo:hi], delimiter=",", fmt="%.17g")
    return dirs


def test_driver_protocol_c1_matches_oracle_scg(tmp_path):
    """BASELINE config 1 (README minimal run): local MapReduce protocol, 5 SCG iterations,
    M=2 Q=2 D=4, N=1k in 4 shards -- the replayed driver on the b200 backend against the same
    driver on the oracle backend (objective trace and final parameters)."""
    from gparml_b200 import b200_MapReduce, parallel_GPLVM as drv
    from gparml_b200.scg_adapted import SCG_adapted
    from gparml_b200.synthetic import make_problem
    from oracle import c_oracle
    from oracle_backend import OracleBackend
    p = make_problem(1000, 2, 2, 4, seed=1)
    dirs = _write_problem(tmp_path, p, 4)
    np.random.seed(0)
    opts = drv.default_options(M=2, Q=2, D=4, iterations=5, init="PCA", display=False, **dirs)
    opts = b200_MapReduce.init(opts)
    assert opts["N"] == 1000
    opts, gs = drv.init_statistics(b200_MapReduce, opts)
    x0 = drv.flatten_global_statistics(opts, gs)
    x0 = np.array([drv.sp.transform_back(b, x) for b, x in zip(opts["flat_global_statistics_bounds"], x0)])
    names = sorted(os.listdir(dirs["input"]))
    shards = [dict(Y=np.genfromtxt(os.path.join(dirs["input"], n), delimiter=","),
                   X_mu=np.load(os.path.join(dirs["embeddings"], n + ".embedding.npy")),
                   X_S=np.load(os.path.join(dirs["embeddings"], n + ".variance.npy"))) for n in names]
    drv.options, drv.map_reduce = opts, b200_MapReduce
    try:
        xg, flog_g, _, _, tacc = SCG_adapted(drv.likelihood_and_gradient, x0.copy(), opts["embeddings"], False,
                                             display=False, maxiters=5, xtol=0, ftol=0, gtol=0)
        # file protocol artefacts of the last evaluation exist where the reference writes them
        assert os.path.exists(os.path.join(dirs["statistics"], "accumulated_statistics_sum_exp_K_mi_K_im_%d.npy" % opts["i"]))
        assert os.path.exists(os.path.join(dirs["statistics"], "cache_Kmm_inv_%d.npy" % opts["i"]))
        assert len(tacc["embeddings_get_grads_mu"]) > 0
        be = OracleBackend(shards, 2, 2, evaluate=c_oracle.evaluate)
        xo, flog_o, _, _, _ = SCG_adapted(be.f_and_gradf, x0.copy(), "unused", False, display=False, maxiters=5,
                                          xtol=0, ftol=0, gtol=0, local_ops=be)
        assert relerr(np.array(flog_g), np.array(flog_o)) < 1e-8
        assert relerr(xg, xo) < 1e-6
        assert flog_g[-1] < flog_g[0]
        # device-resident embeddings after the run == oracle's files-in-memory
        b200_MapReduce.flush(opts)
        for n, s in zip(names, be.st):
            assert relerr(np.load(os.path.join(dirs["embeddings"], n + ".embedding.npy")), s["X_mu"]) < 1e-6
            assert relerr(np.load(os.path.join(dirs["embeddings"], n + ".grad_d.npy")), s["d"]) < 1e-5
        # the file-less fast path gives the same evaluation
        f1, g1 = drv.likelihood_and_gradient(xg, 7, 0)
        opts["b200_write_files"] = False
        f2, g2 = drv.likelihood_and_gradient(xg, 8, 0)
        assert abs(f1 - f2) <= 1e-12 * abs(f1) and relerr(g2, g1) < 1e-10
    finally:
        b200_MapReduce.close()


def test_main_runs_fixed_embeddings_sparse_gp(tmp_path):
    """--fixed_embeddings (BASELINE config 2 shape, small N): the whole main() sequence incl. the
    final 'f' checkpoint evaluation (parallel_GPLVM.py:120)."""
    from gparml_b200 import b200_MapReduce, parallel_GPLVM as drv
    from gparml_b200.synthetic import make_problem, split_rows
    p = make_problem(600, 20, 4, 1, seed=2, fixed_embeddings=True)
    dirs = _write_problem(tmp_path, p, 2)
    for i, (lo, hi) in enumerate(split_rows(600, 2)):
        np.save(os.path.join(dirs["embeddings"], "easy_%d.embedding.npy" % i), p["X_mu"][lo:hi])
    np.random.seed(1)
    opts = drv.default_options(M=20, Q=4, D=1, iterations=3, fixed_embeddings=True, display=False, **dirs)
    try:
        x_opt = drv.main(opts)
        flog = x_opt[1]
        assert len(flog) == 4 and flog[-1] <= flog[0]
        assert os.path.exists(os.path.join(dirs["statistics"], "partial_derivatives_F_f.npy"))
        assert os.path.exists(os.path.join(dirs["statistics"], "global_statistics_Z_f.npy"))
        assert os.path.exists(os.path.join(dirs["statistics"], "nlml_acc.obj"))
    finally:
        b200_MapReduce.close()


def test_dropout_rescales_like_reference(tmp_path):
    """--drop_out_fraction (local_MapReduce.py:121-129,263-264): every accumulated statistic is the sum over the
    kept shards divided by kept / (kept + dropped) -- checked on all 12 tensors against the oracle's per-shard
    sums -- including the reference's fallback when every node is dropped (one node drawn with random.randint,
    the dropped list left untouched, so the factor is (1 + n) / 1)."""
    import random
    from gparml_b200 import b200_MapReduce, parallel_GPLVM as drv
    from gparml_b200.synthetic import make_problem
    from oracle import gparml_oracle as O
    p = make_problem(403, 5, 2, 3, seed=3)
    dirs = _write_problem(tmp_path, p, 4)
    np.random.seed(5)
    opts = drv.default_options(M=5, Q=2, D=3, iterations=1, init="random", display=False, **dirs)
    opts = b200_MapReduce.init(opts)
    opts, gs = drv.init_statistics(b200_MapReduce, opts)
    names = sorted(os.listdir(dirs["input"]))
    Z, sf2 = gs["Z"], float(np.squeeze(gs["sf2"]))
    alpha = np.atleast_1d(np.squeeze(gs["alpha"]))
    per_shard = []
    for n in names:
        Y = np.genfromtxt(os.path.join(dirs["input"], n), delimiter=",")
        mu = np.load(os.path.join(dirs["embeddings"], n + ".embedding.npy"))
        S = O.softplus(np.load(os.path.join(dirs["embeddings"], n + ".variance.npy")))
        per_shard.append(O.shard_statistics(Y, mu, S, Z, sf2, alpha))
    try:
        opts["i"], opts["step_size"] = 0, 0
        b200_MapReduce.cache(opts, gs)
        for frac, seed in ((0.5, 11), (0.5, 12), (1.0, 13)):
            opts["drop_out_fraction"] = frac
            np.random.seed(seed)
            random.seed(seed)
            files, _, _ = b200_MapReduce.statistics_MR(opts)
            kept = list(b200_MapReduce.non_dropped_out_nodes)
            dropped = list(b200_MapReduce.dropped_out_nodes)
            if frac == 1.0:
                assert len(kept) == 1 and len(dropped) == 4          # the reference's fallback branch
            else:
                assert sorted(kept + dropped) == [0, 1, 2, 3] and 1 <= len(kept) <= 4
            scale = float(len(kept) + len(dropped)) / len(kept)
            got = {k: np.load(f) for k, f in files}
            want = O.reduce_statistics([per_shard[i] for i in kept])
            for k in got:
                assert relerr(got[k], np.asarray(want[k]) * scale) < 1e-11, (frac, seed, k)
            # the embeddings map still runs on every shard, dropped or not (local_MapReduce.py:292-294)
            F, _ = b200_MapReduce.session_contexts(opts["embeddings"])[kept[0]].global_step()
            assert np.isfinite(F)
            b200_MapReduce.embeddings_MR(opts)
            for c in b200_MapReduce.session_contexts(opts["embeddings"]):
                assert np.all(np.isfinite(c.grad_latest()))
    finally:
        b200_MapReduce.close()


def test_load_resumes_from_checkpoint_and_keep_keeps_files(tmp_path):
    """--load (parallel_GPLVM.py:195-200, local_MapReduce.py:49): a second main() started from the 'f' checkpoint
    and the flushed embeddings continues where the first one stopped; --keep (parallel_GPLVM.py:373-404) leaves the
    per-iteration files in place, without it they are cleaned."""
    from gparml_b200 import b200_MapReduce, parallel_GPLVM as drv
    from gparml_b200.synthetic import make_problem
    p = make_problem(600, 6, 2, 3, seed=12)
    dirs = _write_problem(tmp_path, p, 3)
    st = dirs["statistics"]
    np.random.seed(2)
    try:
        x1 = drv.main(drv.default_options(M=6, Q=2, D=3, iterations=3, init="PCA", display=False, **dirs))
    finally:
        b200_MapReduce.close()
    flog1 = x1[1]
    assert not [f for f in os.listdir(st) if f.startswith("global_statistics_Z_") and not f.endswith("_f.npy")
                and f not in ("global_statistics_Z_%d.npy" % k for k in (len(flog1) - 1, len(flog1) - 2))]
    Z_f = np.load(os.path.join(st, "global_statistics_Z_f.npy"))
    F_f = float(np.load(os.path.join(st, "partial_derivatives_F_f.npy")))
    assert -F_f == pytest.approx(flog1[-1], rel=1e-12)
    emb1 = np.load(os.path.join(dirs["embeddings"], "easy_0.embedding.npy"))
    try:
        x2 = drv.main(drv.default_options(M=6, Q=2, D=3, iterations=2, load=True, keep=True, display=False, **dirs))
    finally:
        b200_MapReduce.close()
    flog2 = x2[1]
    # the resumed run starts at the checkpoint: same parameters, same embeddings, hence the same objective
    assert flog2[0] == pytest.approx(flog1[-1], rel=1e-11)
    assert flog2[-1] < flog2[0]
    assert not np.array_equal(np.load(os.path.join(st, "global_statistics_Z_f.npy")), Z_f)       # a new checkpoint
    assert not np.array_equal(np.load(os.path.join(dirs["embeddings"], "easy_0.embedding.npy")), emb1)
    # keep=True: the per-iteration files of the second run are all still there
    for k in range(0, 3):
        assert os.path.exists(os.path.join(st, "global_statistics_Z_%d.npy" % k)), k
        assert os.path.exists(os.path.join(st, "accumulated_statistics_sum_exp_K_mi_K_im_%d.npy" % k)), k


def test_checkpoint_written_without_per_evaluation_files(tmp_path):
    """b200_write_files=False skips the per-evaluation file transport but the final 'f' evaluation still writes the
    checkpoint predict.py / --load need (global statistics, accumulated statistics, cache, partial derivatives)."""
    from gparml_b200 import b200_MapReduce, parallel_GPLVM as drv
    from gparml_b200.synthetic import make_problem
    p = make_problem(300, 4, 2, 3, seed=14)
    dirs = _write_problem(tmp_path, p, 2)
    st = dirs["statistics"]
    np.random.seed(6)
    try:
        x1 = drv.main(drv.default_options(M=4, Q=2, D=3, iterations=2, init="PCA", display=False, b200_write_files=False, **dirs))
        files = sorted(os.listdir(st))
        assert not [f for f in files if f.endswith(".npy") and not f.endswith("_f.npy")], files
        for f in ("global_statistics_Z_f.npy", "global_statistics_beta_f.npy", "accumulated_statistics_sum_exp_K_mi_K_im_f.npy",
                  "accumulated_statistics_sum_KL_f.npy", "cache_Kmm_f.npy", "cache_Kmm_inv_f.npy",
                  "partial_derivatives_F_f.npy", "partial_derivatives_dF_dKmm_f.npy"):
            assert f in files, f
        assert -float(np.load(os.path.join(st, "partial_derivatives_F_f.npy"))) == pytest.approx(x1[1][-1], rel=1e-12)
    finally:
        b200_MapReduce.close()
    try:        # and that checkpoint is loadable
        x2 = drv.main(drv.default_options(M=4, Q=2, D=3, iterations=1, load=True, display=False, b200_write_files=False, **dirs))
        assert x2[1][0] == pytest.approx(x1[1][-1], rel=1e-11)
    finally:
        b200_MapReduce.close()


def test_predict_path_matches_oracle_and_improves(tmp_path):
    """predict.py:116-144 replay: the test points' statistics are added to the stored 'f' statistics;
    one evaluation against the oracle (1e-9), then a short optimisation must increase the bound."""
    from gparml_b200 import b200_MapReduce, parallel_GPLVM as drv, predict
    from gparml_b200.synthetic import make_problem
    from oracle import gparml_oracle as O
    p = make_problem(500, 8, 2, 3, seed=9)
    dirs = _write_problem(tmp_path, p, 2)
    np.random.seed(3)
    opts = drv.default_options(M=8, Q=2, D=3, iterations=2, init="PCA", display=False, **dirs)
    try:
        drv.main(opts)                      # writes the *_f.npy checkpoint the prediction path reads
    finally:
        b200_MapReduce.close()
    rng = np.random.default_rng(5)
    Y_test = p["Y"][:7] + 0.01 * rng.standard_normal((7, 3))
    try:
        s = predict.setup(opts, Y_test)
        gs, acc = s["global_statistics"], s["accumulated_statistics"]
        mu = rng.standard_normal((7, 2)); S = rng.uniform(0.2, 0.8, (7, 2))
        x = np.concatenate([mu.ravel(), O.softplus_inv(S).ravel()])
        f, g = predict.likelihood_and_gradient(x)
        # oracle: same composition of statistics
        Z, sf2 = gs["Z"], float(np.squeeze(gs["sf2"]))
        alpha, beta = np.atleast_1d(np.squeeze(gs["alpha"])), float(np.squeeze(gs["beta"]))
        new = O.shard_statistics(Y_test, mu, S, Z, sf2, alpha)
        tot = dict(new)
        for k in ("sum_YYT", "sum_exp_K_mi_K_im", "sum_exp_K_miY", "sum_exp_K_ii", "sum_KL"):
            tot[k] = acc[k] + new[k]
        G = O.global_step(tot, Z, sf2, alpha, beta, opts["N"])
        gm, gS = O.embedding_grads(Y_test, mu, S, Z, sf2, alpha, G["dF_dsum_exp_K_miY"], G["dF_dsum_exp_K_mi_K_im"])
        g_ref = -np.concatenate([gm.ravel(), (gS * O.softplus_grad(O.softplus_inv(S))).ravel()])
        assert abs(f + G["F"]) <= 1e-9 * abs(G["F"])
        assert relerr(g, g_ref) < 1e-9
        np.random.seed(4)
        best = predict.test(opts, Y_test, random_iterations=15)
        f0, _ = predict.likelihood_and_gradient(np.concatenate([best[0].ravel(), O.softplus_inv(best[1]).ravel()]))
        assert best[0].shape == (7, 2) and np.all(best[1] > 0)
        assert abs(-f0 - best[2]) <= 1e-8 * abs(best[2])
    finally:
        predict.close()
