"""GPU parity tests (``-m gpu``): the CUDA path, called through the C ABI, against
(a) the golden vectors generated from the live reference and (b) the C / numpy oracles on
seeded inputs at the BASELINE shapes.  Bar: 1e-9 max-norm relative error in fp64
(BASELINE.json north_star) on the 12 reduced statistics, F, the four global gradients and
the per-point gradients, single- and multi-shard, step_size in {0, 1e-3}."""
import numpy as np
import pytest

from conftest import GOLDEN_CASES, check_against_golden, load_golden, relerr

pytestmark = pytest.mark.gpu

TOL = 1e-9


def _gpu_evaluate(shards, Z, sf2, alpha, beta, step_size=0.0, fixed_embeddings=False, want_stats=True):
    from gparml_b200 import _lib
    from gparml_b200.engine import ShardContext, evaluate
    M, Q = Z.shape
    D = shards[0]["Y"].shape[1] if shards[0]["Y"].ndim == 2 else 1
    N = sum(s["X_mu"].shape[0] for s in shards)
    ctxs = []
    try:
        for s in shards:
            c = ShardContext(M, Q, D, N, fixed_embeddings=fixed_embeddings)
            c.upload_shard(s["Y"], s["X_mu"], s["X_S"])
            if s.get("d") is not None:
                c.upload(_lib.A_GRAD_D, s["d"])
            ctxs.append(c)
        F, grad = evaluate(ctxs, Z, sf2, alpha, beta, step_size=step_size)
        root = ctxs[0]
        out = {"global": {"F": F, "grad_Z": grad["Z"], "grad_sf2": grad["sf2"], "grad_alpha": grad["alpha"],
                          "grad_beta": grad["beta"],
                          "dF_dKmm": root.download(_lib.A_DF_DKMM, (M, M)),
                          "dF_dsum_exp_K_miY": root.download(_lib.A_DF_DPSI1Y, (M, D)),
                          "dF_dsum_exp_K_mi_K_im": root.download(_lib.A_DF_DPSI2, (M, M)),
                          "Kmm": root.download(_lib.A_KMM, (M, M)),
                          "Kmm_inv": root.download(_lib.A_KMM_INV, (M, M))},
               "grad_latest": [] if fixed_embeddings else [c.grad_latest() for c in ctxs]}
        if want_stats:
            out["stats"] = root.stats_named()
        return out
    finally:
        for c in ctxs:
            c.close()


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_cuda_matches_reference_golden(name):
    g = load_golden(name)
    res = _gpu_evaluate(g["shards"], g["Z"], g["sf2"], g["alpha"], g["beta"], step_size=g["step_size"],
                        fixed_embeddings=g["fixed_embeddings"])
    errs = check_against_golden(res, g, TOL, "CUDA vs reference golden " + name)
    print(name, "max rel err %.2e (cond Kmm %.1f)" % (max(errs.values()), float(g["glob_cond_Kmm"])))


def _shards_of(p, parts):
    from gparml_b200.synthetic import split_rows
    out = []
    for lo, hi in split_rows(p["N"], parts):
        sh = dict(Y=p["Y"][lo:hi], X_mu=p["X_mu"][lo:hi], X_S=p["X_S"][lo:hi])
        if "d" in p:
            sh["d"] = p["d"][:, lo:hi]
        out.append(sh)
    return out


# (config, N_parity, shards, step)  -- SURVEY.md 8d parity protocol, BASELINE shapes
FULL_SHAPE_CASES = [
    ("c1", 1000, 4, 1e-3),
    ("c2", 8192, 1, 0.0),
    ("c3", 8192, 1, 0.0),
    ("c3", 4099, 8, 1e-3),      # ragged 8-shard split, prime N
    ("c4", 384, 2, 0.0),
]


@pytest.mark.parametrize("cfg,n,parts,step", FULL_SHAPE_CASES)
def test_cuda_matches_c_oracle_full_shapes(cfg, n, parts, step):
    from gparml_b200.synthetic import CONFIGS, make_problem
    from oracle import c_oracle
    k = CONFIGS[cfg]
    p = make_problem(n, k["M"], k["Q"], k["D"], seed=int(cfg[1]), fixed_embeddings=k["fixed_embeddings"],
                     generic_hypers=True, with_direction=not k["fixed_embeddings"])
    shards = _shards_of(p, parts)
    ref = c_oracle.evaluate(shards, p["Z"], p["sf2"], p["alpha"], p["beta"], step_size=step,
                            fixed_embeddings=k["fixed_embeddings"])
    res = _gpu_evaluate(shards, p["Z"], p["sf2"], p["alpha"], p["beta"], step_size=step,
                        fixed_embeddings=k["fixed_embeddings"])
    errs = {}
    for key, v in ref["stats"].items():
        errs[key] = relerr(res["stats"][key], v)
    for key in ("F", "grad_Z", "grad_alpha", "grad_sf2", "grad_beta", "dF_dKmm", "dF_dsum_exp_K_miY",
                "dF_dsum_exp_K_mi_K_im"):
        errs[key] = relerr(res["global"][key], ref["global"][key])
    for i, (a, b) in enumerate(zip(res["grad_latest"], ref["grad_latest"])):
        errs["grad_latest_%d" % i] = relerr(a, b)
    print(cfg, n, parts, "log10 cond(Kmm) = %.2f" % np.log10(ref["global"]["cond_Kmm"]),
          "max rel err %.2e at %s" % (max(errs.values()), max(errs, key=errs.get)))
    bad = {k2: v for k2, v in errs.items() if not v <= TOL}
    assert not bad, bad


def test_statistics_are_additive_over_shards():
    """Size-independent property: 1-shard and 5-shard evaluations agree (the sums are
    exact up to fp64 reassociation)."""
    from gparml_b200.synthetic import make_problem
    p = make_problem(3001, 20, 3, 2, seed=7, generic_hypers=True)
    a = _gpu_evaluate(_shards_of(p, 1), p["Z"], p["sf2"], p["alpha"], p["beta"])
    b = _gpu_evaluate(_shards_of(p, 5), p["Z"], p["sf2"], p["alpha"], p["beta"])
    for k in a["stats"]:
        assert relerr(b["stats"][k], a["stats"][k]) < 1e-12, k
    assert abs(a["global"]["F"] - b["global"]["F"]) < 1e-9 * abs(a["global"]["F"])
    assert relerr(np.concatenate(b["grad_latest"], axis=1), a["grad_latest"][0]) < 1e-10


def test_evaluation_is_deterministic():
    from gparml_b200.synthetic import make_problem
    p = make_problem(2000, 30, 4, 3, seed=8, generic_hypers=True)
    a = _gpu_evaluate(_shards_of(p, 1), p["Z"], p["sf2"], p["alpha"], p["beta"])
    b = _gpu_evaluate(_shards_of(p, 1), p["Z"], p["sf2"], p["alpha"], p["beta"])
    assert a["global"]["F"] == b["global"]["F"]
    assert np.array_equal(a["global"]["grad_Z"], b["global"]["grad_Z"])
    assert np.array_equal(a["grad_latest"][0], b["grad_latest"][0])


def test_finite_differences_through_cuda_path():
    """The reference's own test strategy (test.py:62-93,150-201,270-296): analytic gradients
    against finite differences of the bound -- here central differences through the CUDA path
    on the fixture shape D=7, Q=2, N=5, M=10."""
    g = load_golden("t5")
    sh = g["shards"]

    def F_of(**kw):
        a = dict(Z=g["Z"], sf2=g["sf2"], alpha=g["alpha"], beta=g["beta"])
        shards = sh
        for k, v in kw.items():
            if k in a:
                a[k] = v
            else:
                shards = [dict(sh[0], **{k: v})]
        return _gpu_evaluate(shards, a["Z"], a["sf2"], a["alpha"], a["beta"], want_stats=False)

    base = F_of()
    h = 1e-6
    for (j, k) in [(0, 0), (7, 1)]:
        Zp, Zm = g["Z"].copy(), g["Z"].copy(); Zp[j, k] += h; Zm[j, k] -= h
        fd = (F_of(Z=Zp)["global"]["F"] - F_of(Z=Zm)["global"]["F"]) / (2 * h)
        assert abs(fd - base["global"]["grad_Z"][j, k]) <= 2e-6 * max(1.0, abs(fd))
    fd = (F_of(sf2=g["sf2"] + h)["global"]["F"] - F_of(sf2=g["sf2"] - h)["global"]["F"]) / (2 * h)
    assert abs(fd - base["global"]["grad_sf2"]) <= 2e-6 * max(1.0, abs(fd))
    fd = (F_of(beta=g["beta"] + h)["global"]["F"] - F_of(beta=g["beta"] - h)["global"]["F"]) / (2 * h)
    assert abs(fd - base["global"]["grad_beta"]) <= 2e-6 * max(1.0, abs(fd))
    ap, am = g["alpha"].copy(), g["alpha"].copy(); ap[1] += h; am[1] -= h
    fd = (F_of(alpha=ap)["global"]["F"] - F_of(alpha=am)["global"]["F"]) / (2 * h)
    assert abs(fd - base["global"]["grad_alpha"][1]) <= 2e-6 * max(1.0, abs(fd))
    mp, mm = sh[0]["X_mu"].copy(), sh[0]["X_mu"].copy(); mp[2, 1] += h; mm[2, 1] -= h
    fd = (F_of(X_mu=mp)["global"]["F"] - F_of(X_mu=mm)["global"]["F"]) / (2 * h)
    assert abs(fd + base["grad_latest"][0][0, 2, 1]) <= 2e-6 * max(1.0, abs(fd))
    sp, sm_ = sh[0]["X_S"].copy(), sh[0]["X_S"].copy(); sp[3, 0] += h; sm_[3, 0] -= h
    fd = (F_of(X_S=sp)["global"]["F"] - F_of(X_S=sm_)["global"]["F"]) / (2 * h)
    assert abs(fd + base["grad_latest"][0][1, 3, 0]) <= 2e-6 * max(1.0, abs(fd))


@pytest.mark.parametrize("M", [6, 130])
def test_error_mapping_not_pd_and_range(M):
    """Device-side numerical failure surfaces as the exception types the reference's
    optimiser wrapper survives (scg_adapted.py:55); M = 130 takes the multi-kernel master step."""
    from gparml_b200.engine import ShardContext
    rng = np.random.default_rng(0)
    Q, D, n = 2, 2, 40
    Z = rng.standard_normal((M, Q)) * (1.0 if M < 50 else 8.0)     # 130 points in the unit square would make Kmm singular to 1e-15
    with ShardContext(M, Q, D, n) as c:
        c.upload_shard(rng.standard_normal((n, D)), rng.standard_normal((n, Q)), rng.standard_normal((n, Q)) - 1.0)
        c.set_globals(Z, 1.0, np.ones(Q), 1.0)
        c.statistics()
        F, _ = c.global_step()
        assert np.isfinite(F)
        # a negative-definite "Psi2" makes Kmm + beta*Psi2 indefinite -> failed Cholesky pivot
        c.set_stats_named({"sum_exp_K_mi_K_im": -10.0 * np.eye(M)})
        with pytest.raises(np.linalg.LinAlgError):
            c.global_step()
        bad = rng.standard_normal((n, Q)); bad[5, 1] = 40.0   # supporting_functions.py:154 assert
        c.upload_shard(rng.standard_normal((n, D)), rng.standard_normal((n, Q)), bad)
        c.set_globals(Z + 0.1, 1.0, np.ones(Q), 1.0)
        with pytest.raises(AssertionError):
            c.statistics()


@pytest.mark.parametrize("M", [12, 130])
def test_jitter_retry_follows_reference_branch(M):
    """partial_terms.py:453-457: when a factorisation says "not positive definite" the reference retries with 1e-7
    on the diagonal and only gives up (assert) if that fails too.  The device does the same (M = 12: single-CTA
    master step; M = 130: the multi-kernel block sweep, whose retry pass is decided on the device): a Kmm + beta Psi2 with one eigenvalue of -1e-9 evaluates to the bound of the reference's jitter branch
    (the oracle restates it: oracle/gparml_oracle.py global_step), the event is counted, and a matrix that is
    indefinite beyond the jitter still raises LinAlgError (test_error_mapping_not_pd_and_range)."""
    from gparml_b200.engine import ShardContext
    from gparml_b200.synthetic import make_problem
    from oracle import gparml_oracle as O
    Q, D, n = 3, 2, 300
    p = make_problem(n, M, Q, D, seed=91, generic_hypers=True)
    rng = np.random.default_rng(91)
    V, _ = np.linalg.qr(rng.standard_normal((M, M)))
    lam = np.concatenate([[-1.0e-9], rng.uniform(0.5, 2.0, M - 1)])
    E = (V * lam) @ V.T
    E = 0.5 * (E + E.T)
    K = O.kmm(p["Z"], p["sf2"], p["alpha"])
    P2 = (E - K) / p["beta"]                      # Kmm + beta Psi2 = E: indefinite by 1e-9
    with ShardContext(M, Q, D, n) as c:
        c.upload_shard(p["Y"], p["X_mu"], p["X_S"])
        c.set_globals(p["Z"], p["sf2"], p["alpha"], p["beta"])
        c.statistics()
        F0, _ = c.global_step()
        assert c.jitter_events == 0 and np.isfinite(F0)
        c.set_stats_named({"sum_exp_K_mi_K_im": P2})
        stats = c.stats_named()
        F, g = c.global_step()
        assert c.jitter_events == 1
        ref = O.global_step(stats, p["Z"], p["sf2"], p["alpha"], p["beta"], n)
        assert np.linalg.slogdet(K + p["beta"] * stats["sum_exp_K_mi_K_im"])[0] < 0      # the reference's trigger (:455)
        print("jitter branch: F %.12g  oracle %.12g  rel %.2e" % (F, ref["F"], abs(F - ref["F"]) / abs(ref["F"])))
        assert abs(F - ref["F"]) <= 1e-7 * abs(ref["F"])
        assert np.all(np.isfinite(g["flat"]))
        # the next evaluation with healthy statistics is not affected
        c.statistics()
        F1, _ = c.global_step()
        assert F1 == F0 and c.jitter_events == 1


def test_scg_local_ops_match_numpy():
    from gparml_b200 import _lib
    from gparml_b200.engine import ShardContext
    from oracle import gparml_oracle as O
    rng = np.random.default_rng(3)
    n, Q = 777, 5
    with ShardContext(4, Q, 2, n) as c:
        mu0, s0 = rng.standard_normal((n, Q)), rng.standard_normal((n, Q))
        c.upload_shard(rng.standard_normal((n, 2)), mu0, s0)
        st = [dict(latest=rng.standard_normal((2, n, Q)), X_mu=mu0.copy(), X_S=s0.copy())]
        c.upload(_lib.A_GRAD_LATEST, st[0]["latest"])
        c.scg_set_grads(); O.scg_set_grads(st)
        st[0]["latest"] = rng.standard_normal((2, n, Q)); c.upload(_lib.A_GRAD_LATEST, st[0]["latest"])
        assert c.scg_get_mu() == pytest.approx(O.scg_get_mu(st), rel=1e-12)
        assert c.scg_get_kappa() == pytest.approx(O.scg_get_kappa(st), rel=1e-12)
        assert c.scg_get_theta() == pytest.approx(O.scg_get_theta(st), rel=1e-11)
        assert c.scg_get_current_grad() == pytest.approx(O.scg_get_current_grad(st), rel=1e-12)
        assert c.scg_get_max_d(0.37) == O.scg_get_max_d(st, 0.37)
        c.scg_update_X(0.37); O.scg_update_X(st, 0.37)
        assert np.array_equal(c.download(_lib.A_X_MU, (n, Q)), st[0]["X_mu"])
        assert np.array_equal(c.download(_lib.A_X_S, (n, Q)), st[0]["X_S"])
        c.scg_update_grad_old(); O.scg_update_grad_old(st)
        c.scg_update_grad_new(); O.scg_update_grad_new(st)
        assert c.scg_get_gamma() == pytest.approx(O.scg_get_gamma(st), rel=1e-12)
        c.scg_update_d(0.81); O.scg_update_d(st, 0.81)
        assert np.allclose(c.download(_lib.A_GRAD_D, (2, n, Q)), st[0]["d"], rtol=1e-15, atol=0)
        c.scg_reset_d(); O.scg_reset_d(st)
        assert np.array_equal(c.download(_lib.A_GRAD_D, (2, n, Q)), st[0]["d"])


@pytest.mark.parametrize("Q,D,M", [(1, 1, 5), (3, 3, 17), (7, 13, 40), (12, 2, 33), (16, 5, 20), (5, 21, 130)])
def test_template_paths_odd_shapes(Q, D, M):
    """Every template path: odd/even Q (record padding), Q > 10 (one pair per thread), Q = 16,
    D that needs several column chunks, M > 116 (multi-kernel master step), ragged 3-shard split."""
    from gparml_b200.synthetic import make_problem
    from oracle import c_oracle
    n = 700 + 13 * Q
    p = make_problem(n, M, Q, D, seed=100 + Q, generic_hypers=True, with_direction=True)
    shards = _shards_of(p, 3)
    ref = c_oracle.evaluate(shards, p["Z"], p["sf2"], p["alpha"], p["beta"], step_size=1e-3)
    res = _gpu_evaluate(shards, p["Z"], p["sf2"], p["alpha"], p["beta"], step_size=1e-3)
    errs = {k: relerr(res["stats"][k], v) for k, v in ref["stats"].items()}
    for key in ("F", "grad_Z", "grad_alpha", "grad_sf2", "grad_beta", "dF_dKmm", "dF_dsum_exp_K_miY", "dF_dsum_exp_K_mi_K_im"):
        errs[key] = relerr(res["global"][key], ref["global"][key])
    for i, (a, b) in enumerate(zip(res["grad_latest"], ref["grad_latest"])):
        errs["grad_latest_%d" % i] = relerr(a, b)
    print(Q, D, M, "log10 cond %.2f" % np.log10(ref["global"]["cond_Kmm"]), "max rel err %.2e at %s" % (max(errs.values()), max(errs, key=errs.get)))
    bad = {k2: v for k2, v in errs.items() if not v <= TOL}
    assert not bad, bad


def test_random_shapes_against_c_oracle():
    """Twenty seeded random problem shapes (M, Q, D, shard sizes, step on / off, fixed embeddings on / off) through the
    C ABI against the C oracle: every instantiation boundary gets hit by something nobody chose by hand -- Q on both
    sides of the psi2x_stats limit (10), D on both sides of the wide-D Psi1 kernel (17), M on both sides of the
    single-CTA master step (116), ragged shards including one-point shards."""
    from gparml_b200.synthetic import make_problem
    from oracle import c_oracle
    rng = np.random.default_rng(20141208 + 77)
    worst = {}
    for case in range(20):
        Q = int(rng.integers(1, 17))
        D = int(rng.choice([1, 2, 3, 5, 8, 10, 16, 17, 19, 33]))
        M = int(rng.choice([2, 3, 7, 16, 31, 50, 64, 100, 116, 117, 140]))
        parts = int(rng.integers(1, 5))
        sizes = [int(rng.choice([1, 2, 33, 127, 128, 129, 400, 777])) for _ in range(parts)]
        if sum(sizes) < 5:
            sizes.append(33)
        n = sum(sizes)
        fixed = bool(rng.random() < 0.25)
        step = 0.0 if (fixed or rng.random() < 0.5) else 1e-3
        p = make_problem(n, M, Q, D, seed=1000 + case, generic_hypers=True, with_direction=step != 0.0, fixed_embeddings=fixed)
        cuts = np.concatenate([[0], np.cumsum(sizes)])
        shards = [dict(Y=p["Y"][a:b], X_mu=p["X_mu"][a:b], X_S=p["X_S"][a:b],
                       d=(np.stack([p["d"][0][a:b], p["d"][1][a:b]]) if step != 0.0 else None)) for a, b in zip(cuts[:-1], cuts[1:])]
        ref = c_oracle.evaluate(shards, p["Z"], p["sf2"], p["alpha"], p["beta"], step_size=step, fixed_embeddings=fixed)
        res = _gpu_evaluate(shards, p["Z"], p["sf2"], p["alpha"], p["beta"], step_size=step, fixed_embeddings=fixed)
        errs = _compare_all(res, ref)
        cond = float(ref["global"]["cond_Kmm"])
        # two backward-stable inversions differ by ~cond eps (profiles/cond_probe_r02.txt): the statistics and per-point
        # gradients are held to TOL always, the master step's outputs where Kmm is not worse than 1e6
        tol = {k: (TOL if (k in ref["stats"] or k.startswith("grad_latest") or cond < 1e6) else max(TOL, 100 * cond * 2.2e-16)) for k in errs}
        bad = {k: v for k, v in errs.items() if not v <= tol[k]}
        print("case %2d Q %2d D %2d M %3d shards %s step %g fixed %d cond %.1e: max rel err %.2e at %s" % (
            case, Q, D, M, sizes, step, fixed, cond, max(errs.values()), max(errs, key=errs.get)))
        assert not bad, (case, bad)
        worst[case] = max(errs.values())


@pytest.mark.parametrize("Q,D", [(10, 8), (10, 9), (10, 11), (10, 16), (10, 17), (4, 24), (12, 9), (16, 10)])
def test_psi1_contraction_column_shapes(Q, D):
    """psi1_stats column chunking: 8-wide tensor-core tiles, the DFMA remainder columns and several
    chunks, for the wide (Q <= 10) and narrow (Q > 10) accumulator budgets; M not a multiple of 8."""
    from gparml_b200.synthetic import make_problem
    from oracle import c_oracle
    M, n = 27, 531
    p = make_problem(n, M, Q, D, seed=300 + Q + D, generic_hypers=True, with_direction=True)
    shards = _shards_of(p, 2)
    ref = c_oracle.evaluate(shards, p["Z"], p["sf2"], p["alpha"], p["beta"], step_size=1e-3)
    res = _gpu_evaluate(shards, p["Z"], p["sf2"], p["alpha"], p["beta"], step_size=1e-3)
    errs = {k: relerr(res["stats"][k], v) for k, v in ref["stats"].items()}
    errs["F"] = relerr(res["global"]["F"], ref["global"]["F"])
    for i, (a, b) in enumerate(zip(res["grad_latest"], ref["grad_latest"])):
        errs["grad_latest_%d" % i] = relerr(a, b)
    print(Q, D, "max rel err %.2e at %s" % (max(errs.values()), max(errs, key=errs.get)))
    bad = {k2: v for k2, v in errs.items() if not v <= TOL}
    assert not bad, bad


def _compare_all(res, ref, keys=("F", "grad_Z", "grad_alpha", "grad_sf2", "grad_beta")):
    errs = {k: relerr(res["stats"][k], v) for k, v in ref["stats"].items()}
    for key in keys:
        errs[key] = relerr(res["global"][key], ref["global"][key])
    for i, (a, b) in enumerate(zip(res["grad_latest"], ref["grad_latest"])):
        errs["grad_latest_%d" % i] = relerr(a, b)
    return errs


@pytest.mark.parametrize("offset", [30.0, 1000.0, 1.0e4])
def test_latent_space_far_from_origin(offset):
    """embed_grads evaluates the Psi2 exponent in an expanded (dot-product) form centred on the column
    means of Z; a common translation of X_mu and Z far from the origin must not cost accuracy.  Generic
    (unquantised) inputs: here the comparison is limited by the REFERENCE's own rounding of
    mu - 0.5 z_m - 0.5 z_m' (kernel_exp.py:145), eps * offset in every distance."""
    from gparml_b200.synthetic import make_problem, split_rows
    from oracle import c_oracle
    M, Q, D, n = 30, 5, 3, 900
    p = make_problem(n, M, Q, D, seed=77, generic_hypers=True, with_direction=True)
    shift = offset * np.array([1.0, -0.5, 0.25, 2.0, -1.0])
    X_mu = p["X_mu"] + shift
    Z = p["Z"] + shift
    shards = []
    for lo, hi in split_rows(n, 2):
        shards.append(dict(Y=p["Y"][lo:hi], X_mu=X_mu[lo:hi], X_S=p["X_S"][lo:hi], d=p["d"][:, lo:hi]))
    ref = c_oracle.evaluate(shards, Z, p["sf2"], p["alpha"], p["beta"], step_size=1e-3)
    res = _gpu_evaluate(shards, Z, p["sf2"], p["alpha"], p["beta"], step_size=1e-3)
    errs = _compare_all(res, ref)
    print("offset %g: max rel err %.2e at %s" % (offset, max(errs.values()), max(errs, key=errs.get)))
    bad = {k2: v for k2, v in errs.items() if not v <= TOL}
    assert not bad, bad


@pytest.mark.parametrize("log2_offset", [13, 17])
def test_expanded_basis_is_centred_exact_translation(log2_offset):
    """Pins the centring of the expanded basis (capi.cu gparml_set_globals: center = column means of Z).
    X_mu and Z are quantised to multiples of 2^-30 and translated by multiples of 2^log2_offset (8192 and
    131072 units), so the translated inputs are exact doubles and every difference the reference forms
    (mu - 0.5 z - 0.5 z', z - z', mu - z) is exact: the oracle in the translated frame is as accurate as
    at the origin and the comparison isolates the CUDA path's own error.  With center = 0 (the round-1
    defect) the expanded exponent loses eps * w * offset^2 = 1e-8 .. 4e-6 here; centred it must stay
    <= 1e-11 on everything, per-point gradients included."""
    from gparml_b200.synthetic import make_problem, split_rows
    from oracle import c_oracle
    M, Q, D, n = 30, 5, 3, 900
    p = make_problem(n, M, Q, D, seed=78, generic_hypers=True)
    quant = lambda a: np.round(a * 2.0 ** 30) / 2.0 ** 30
    shift = 2.0 ** log2_offset * np.array([1.0, -0.5, 0.25, 2.0, -1.0])
    X_mu = quant(p["X_mu"]) + shift
    Z = quant(p["Z"]) + shift
    assert np.array_equal(X_mu - shift, quant(p["X_mu"])) and np.array_equal(Z - shift, quant(p["Z"]))
    shards = [dict(Y=p["Y"][lo:hi], X_mu=X_mu[lo:hi], X_S=p["X_S"][lo:hi]) for lo, hi in split_rows(n, 2)]
    ref = c_oracle.evaluate(shards, Z, p["sf2"], p["alpha"], p["beta"])
    res = _gpu_evaluate(shards, Z, p["sf2"], p["alpha"], p["beta"])
    errs = _compare_all(res, ref)
    worst_gl = max(v for k, v in errs.items() if k.startswith("grad_latest"))
    print("offset 2^%d: grad_latest rel err %.2e, overall max %.2e at %s"
          % (log2_offset, worst_gl, max(errs.values()), max(errs, key=errs.get)))
    bad = {k2: v for k2, v in errs.items() if not v <= 1e-11}
    assert not bad, bad


def test_unnormalised_regression_inputs_huge_negative_exponents():
    """BASELINE config-2 style sparse GP regression on un-normalised inputs: S = 0, alpha = 1 (the reference's
    initial value, parallel_GPLVM.py:189-194), inputs spanning 1e4 per dimension, so the Psi exponents
    (kernel_exp.py:80,143-146) reach -1e8, beyond where k = round(x 32/ln2) fits the 32-bit word the table
    exp reads it from (x < -4.65e7).  numpy's exp gives 0 there; the device exp clamps its argument
    (gp_exp.cuh gp_exp_clamp) and must agree on every statistic, the bound and the gradients."""
    from oracle import c_oracle
    rng = np.random.default_rng(20141208 + 202)
    n, M, Q, D = 4096, 50, 4, 1
    X = rng.uniform(0.0, 1.0e4, (n, Q))
    # half of the inducing points sit on data points (so Psi1 / Psi2 are not all zero), a few are near each
    # other (non-trivial Kmm), the rest are far from everything
    Z = np.concatenate([X[rng.choice(n, 25, replace=False)] + 0.3 * rng.standard_normal((25, Q)),
                        rng.uniform(0.0, 1.0e4, (25, Q))])
    Z[1] = Z[0] + 0.7
    Z[30] = Z[29] + np.array([1.0, -0.5, 0.2, 0.1])
    Y = np.sin(X[:, :1] / 500.0) + 0.1 * rng.standard_normal((n, 1))
    S = np.zeros((n, Q))
    alpha = np.ones(Q)
    lk_min = -0.25 * np.max(np.sum((Z[:, None, :] - Z[None, :, :]) ** 2, axis=2))
    assert lk_min < -4.65e7, lk_min                       # the exponent range this test is about
    shards = [dict(Y=Y[:2000], X_mu=X[:2000], X_S=S[:2000]), dict(Y=Y[2000:], X_mu=X[2000:], X_S=S[2000:])]
    ref = c_oracle.evaluate(shards, Z, 1.0, alpha, 1.0, fixed_embeddings=True)
    res = _gpu_evaluate(shards, Z, 1.0, alpha, 1.0, fixed_embeddings=True)
    assert np.max(np.abs(ref["stats"]["sum_exp_K_mi_K_im"])) > 0.1      # not a vacuous comparison
    for k, v in res["stats"].items():
        assert np.all(np.isfinite(np.asarray(v))), k
    errs = _compare_all(res, ref)
    print("min pair exponent %.3g: max rel err %.2e at %s" % (lk_min, max(errs.values()), max(errs, key=errs.get)))
    bad = {k2: v for k2, v in errs.items() if not v <= TOL}
    assert not bad, bad


def test_huge_negative_exponents_gplvm_path():
    """The same exponent range through the GPLVM kernels (psi2_stats with S > 0, embed_psi1 / embed_psi2x):
    latent means spanning 2e4 with a few inducing points on data points."""
    from gparml_b200.synthetic import softplus_inv
    from oracle import c_oracle
    rng = np.random.default_rng(20141208 + 203)
    n, M, Q, D = 1200, 24, 3, 2
    X = rng.uniform(-1.0e4, 1.0e4, (n, Q))
    Z = np.concatenate([X[rng.choice(n, 12, replace=False)] + 0.2 * rng.standard_normal((12, Q)),
                        rng.uniform(-1.0e4, 1.0e4, (12, Q))])
    Z[1] = Z[0] + 0.5
    Y = rng.standard_normal((n, D))
    S_raw = softplus_inv(np.clip(0.5 + 0.01 * rng.standard_normal((n, Q)), 0.001, 1.0))
    alpha = np.array([1.0, 0.7, 1.3])
    shards = [dict(Y=Y, X_mu=X, X_S=S_raw)]
    ref = c_oracle.evaluate(shards, Z, 1.3, alpha, 2.0)
    res = _gpu_evaluate(shards, Z, 1.3, alpha, 2.0)
    assert np.all(np.isfinite(res["grad_latest"][0]))
    errs = _compare_all(res, ref)
    print("max rel err %.2e at %s" % (max(errs.values()), max(errs, key=errs.get)))
    bad = {k2: v for k2, v in errs.items() if not v <= TOL}
    assert not bad, bad


@pytest.mark.parametrize("alpha_scale,expect_robust", [(1.0, False), (40.0, False), (2.0e3, True), (3.0e5, True)])
def test_variance_large_against_length_scale(alpha_scale, expect_robust):
    """psi2x_stats builds the Psi2 exponent from the u = t^2 + v it accumulates anyway (sum_q u_q (-1/w_q) started
    at lc2 + sum_q alpha_q S_q), which cancels sum_q alpha_q S_q in rounding.  Posterior variances far above the
    squared length-scale (alpha S up to 1e3 on this path) must stay inside the tolerance; beyond alpha S = 1024
    prep_points switches the evaluation to the cancellation-free instantiation (one more FMA per q), which has to
    agree with the reference formulas (kernel_exp.py:126-148, partial_terms.py:190-205,273-284) for any alpha S."""
    from gparml_b200.synthetic import softplus_inv
    from oracle import c_oracle
    rng = np.random.default_rng(20141208 + 240)
    n, M, Q, D = 1500, 30, 10, 3
    X = rng.standard_normal((n, Q))
    Z = X[rng.choice(n, M, replace=False)] + 0.05 * rng.standard_normal((M, Q))
    Y = rng.standard_normal((n, D))
    S = np.clip(0.5 + 0.2 * rng.standard_normal((n, Q)), 0.05, 1.5)
    S[rng.random((n, Q)) < 0.02] = 14.0                                  # a few very uncertain coordinates
    alpha = rng.uniform(0.5, 1.5, Q)
    alpha[[1, 6]] *= alpha_scale                                          # two short length-scales next to ordinary ones (all ten
    Z[:, [1, 6]] = X[:M, [1, 6]] + 0.02 / np.sqrt(alpha_scale) * rng.standard_normal((M, 2))   # would make every Psi vanish)
    assert (np.max(alpha[None, :] * S) > 1024.0) == expect_robust
    shards = [dict(Y=Y[:700], X_mu=X[:700], X_S=softplus_inv(S[:700])), dict(Y=Y[700:], X_mu=X[700:], X_S=softplus_inv(S[700:]))]
    ref = c_oracle.evaluate(shards, Z, 1.3, alpha, 2.0)
    res = _gpu_evaluate(shards, Z, 1.3, alpha, 2.0)
    # not a vacuous comparison
    assert np.max(np.abs(ref["stats"]["sum_exp_K_mi_K_im"])) > 1e-6
    errs = _compare_all(res, ref)
    print("max alpha S %.3g: max rel err %.2e at %s" % (np.max(alpha[None, :] * S), max(errs.values()), max(errs, key=errs.get)))
    bad = {k2: v for k2, v in errs.items() if not v <= TOL}
    assert not bad, bad


@pytest.mark.parametrize("zero_dims", [(1,), (0, 3)])
def test_alpha_zero_switches_dimension_off_with_finite_gradient(zero_dims):
    """alpha_q = 0 (infinite length-scale) is legal in the reference (kernel_exp.py:30,130 assert >= 0;
    kernels.py self-test) and its alpha-derivatives are finite there (partial_terms.py:256-284 never divide by
    alpha).  The packed buffer keeps those sums scaled by alpha_q^2, so the device substitutes 2^-400 for 0
    (capi.cu gparml_set_globals); statistics, bound and ALL gradients -- grad_alpha[q] of the switched-off
    dimensions included -- must match the oracle evaluated at alpha_q = 0 exactly."""
    from gparml_b200.synthetic import make_problem
    from oracle import c_oracle
    M, Q, D, n = 12, 5, 4, 700           # log10 cond(Kmm) = 0.9 / 1.1 with the dimensions switched off
    p = make_problem(n, M, Q, D, seed=55, generic_hypers=True, with_direction=True)
    alpha = p["alpha"].copy()
    for q in zero_dims:
        alpha[q] = 0.0
    shards = _shards_of(p, 2)
    ref = c_oracle.evaluate(shards, p["Z"], p["sf2"], alpha, p["beta"], step_size=1e-3)
    res = _gpu_evaluate(shards, p["Z"], p["sf2"], alpha, p["beta"], step_size=1e-3)
    assert np.all(np.isfinite(res["global"]["grad_alpha"]))
    assert all(abs(ref["global"]["grad_alpha"][q]) > 1e-3 for q in zero_dims)      # a real number to match
    errs = _compare_all(res, ref, keys=("F", "grad_Z", "grad_alpha", "grad_sf2", "grad_beta", "dF_dKmm",
                                        "dF_dsum_exp_K_miY", "dF_dsum_exp_K_mi_K_im"))
    errs["grad_alpha_elementwise"] = float(np.max(np.abs(res["global"]["grad_alpha"] - ref["global"]["grad_alpha"])
                                                  / np.abs(ref["global"]["grad_alpha"])))
    print("alpha = 0 at", zero_dims, "grad_alpha", res["global"]["grad_alpha"],
          "max rel err %.2e at %s" % (max(errs.values()), max(errs, key=errs.get)))
    bad = {k2: v for k2, v in errs.items() if not v <= TOL}
    assert not bad, bad


def test_split_master_step_matches_blocking_call():
    """gparml_global_step_begin / embedding_grads / gparml_global_step_end (tail of the master step
    concurrent with the embeddings map) gives bit-identical results to the blocking sequence, also when the
    next evaluation is started while a begin is still pending."""
    from gparml_b200 import _lib
    from gparml_b200.engine import ShardContext
    from gparml_b200.synthetic import make_problem
    M, Q, D, n = 40, 6, 5, 3000
    p = make_problem(n, M, Q, D, seed=11, generic_hypers=True, with_direction=True)
    with ShardContext(M, Q, D, n) as c:
        c.upload_shard(p["Y"], p["X_mu"], p["X_S"])
        c.upload(_lib.A_GRAD_D, p["d"])
        c.set_globals(p["Z"], p["sf2"], p["alpha"], p["beta"])
        c.set_step(1e-3)
        c.statistics()
        F0, g0 = c.global_step()
        c.embedding_grads()
        gl0 = c.grad_latest()
        for _ in range(3):
            c.set_globals(p["Z"], p["sf2"], p["alpha"], p["beta"])
            c.statistics()
            c.global_step_begin()
            c.embedding_grads()
            F1, g1 = c.global_step_end()
            gl1 = c.grad_latest()
            assert F1 == F0 and np.array_equal(g1["flat"], g0["flat"]) and np.array_equal(gl1, gl0)
        # a pending begin is finished by the next call that overwrites its inputs
        c.global_step_begin()
        c.set_globals(p["Z"] * 1.01, p["sf2"], p["alpha"], p["beta"])
        c.statistics()
        F2, _ = c.global_step()
        assert np.isfinite(F2) and F2 != F0
        with pytest.raises(ValueError):
            c.global_step_end()


def test_kmm_like_reference_kernel_selftest():
    """kernels.py:131-157 (`python kernels.py -t`): the closed form of one entry, and an infinite
    length-scale (alpha = 0) switches its input dimension off."""
    from gparml_b200 import _lib
    from gparml_b200.engine import ShardContext
    rng = np.random.default_rng(4)
    M = 10
    Z3 = rng.uniform(-5.0, 5.0, (M, 3))
    with ShardContext(M, 3, 1, 0) as c3, ShardContext(M, 2, 1, 0) as c2:
        c3.set_globals(Z3, 16.0, np.array([0.5, 0.5, 0.0]), 1.0)
        c3.update_global_statistics()
        K3 = c3.download(_lib.A_KMM, (M, M))
        c2.set_globals(np.ascontiguousarray(Z3[:, :2]), 16.0, np.array([0.5, 0.5]), 1.0)
        c2.update_global_statistics()
        K2 = c2.download(_lib.A_KMM, (M, M))
    a, b = 3, 5
    expect = 16.0 * np.exp(-np.sum((Z3[a, :2] - Z3[b, :2]) ** 2) / 4.0)
    assert abs(K3[a, b] - expect) <= 1e-12 * expect
    assert np.array_equal(K3, K3.T)
    assert np.allclose(K3, K2, rtol=1e-14, atol=0.0)


def test_statistics_pipelined_with_upload_match_resident_path():
    """gparml_upload_shard sends X_mu / X_S in row ranges and the first gparml_statistics consumes them range
    by range (prep_points + psi2_stats per range); a second call on the resident data takes the one-launch
    path.  Same statistics, bound and gradients (summation order differs: 1e-12)."""
    from gparml_b200 import _lib
    from gparml_b200.engine import ShardContext
    from gparml_b200.synthetic import make_problem
    M, Q, D, n = 24, 3, 2, 200000            # 3 row ranges
    p = make_problem(n, M, Q, D, seed=21, generic_hypers=True, with_direction=True)
    with ShardContext(M, Q, D, n) as c:
        out = []
        for rep in range(2):
            if rep == 0:
                c.upload_shard(p["Y"], p["X_mu"], p["X_S"])
                c.set_globals(p["Z"], p["sf2"], p["alpha"], p["beta"])
            c.statistics()
            st = c.stats_named()
            F, g = c.global_step()
            c.embedding_grads()
            out.append((st, F, g["flat"], c.grad_latest()))
        for key, v in out[0][0].items():
            assert relerr(out[1][0][key], v) <= 1e-12, key
        assert relerr(out[1][1], out[0][1]) <= 1e-12
        assert relerr(out[1][2], out[0][2]) <= 1e-11
        assert relerr(out[1][3], out[0][3]) <= 1e-11
        # a consumer of X_mu between upload and statistics orders itself behind the ranges
        c.upload_shard(p["Y"], p["X_mu"] + 1.0, p["X_S"])
        assert np.array_equal(c.download(_lib.A_X_MU, (n, Q)), p["X_mu"] + 1.0)


def test_chunked_gradient_download_and_reupload():
    """embedding_grads_download (copy of one point range overlapping the next range's kernels) gives
    the same array as embedding_grads + download for any chunk count; re-uploading a different shard
    into the same context (Y on the copy stream) cannot leak the old Y into the new evaluation."""
    from gparml_b200 import _lib
    from gparml_b200.engine import ShardContext
    from gparml_b200.synthetic import make_problem
    from oracle import c_oracle
    pa = make_problem(1537, 24, 5, 6, seed=61, generic_hypers=True)
    pb = make_problem(1537, 24, 5, 6, seed=62, generic_hypers=True)
    with ShardContext(24, 5, 6, 1537) as c:
        for p in (pa, pb, pa):
            c.upload_shard(p["Y"], p["X_mu"], p["X_S"])
            c.set_globals(pa["Z"], pa["sf2"], pa["alpha"], pa["beta"])
            c.statistics()
            F, g = c.global_step()
            c.embedding_grads()
            base = c.grad_latest()
            assert np.array_equal(c.embedding_grads_numpy(1), base)
            for chunks in (3, 8):       # other ranges may pick another m-split count: same sums, other order
                assert relerr(c.embedding_grads_numpy(chunks), base) < 1e-12
            ref = c_oracle.evaluate([dict(Y=p["Y"], X_mu=p["X_mu"], X_S=p["X_S"])], pa["Z"], pa["sf2"], pa["alpha"], pa["beta"])
            assert abs(F - ref["global"]["F"]) <= 1e-9 * abs(ref["global"]["F"])
            assert relerr(base, ref["grad_latest"][0]) < 1e-9
            assert relerr(c.stats_named()["sum_YYT"], ref["stats"]["sum_YYT"]) < 1e-13


def test_empty_and_tiny_shards():
    """Edge cases of the shard sizes: an empty shard (a rank with no points), a 1-point shard and a
    2-point shard next to a normal one; M = 1 and D = 1."""
    from gparml_b200.synthetic import make_problem
    from oracle import c_oracle
    p = make_problem(203, 1, 2, 1, seed=71, generic_hypers=True, with_direction=True)
    cuts = [(0, 0), (0, 1), (1, 3), (3, 203)]
    shards = [dict(Y=p["Y"][lo:hi], X_mu=p["X_mu"][lo:hi], X_S=p["X_S"][lo:hi], d=p["d"][:, lo:hi]) for lo, hi in cuts]
    ref = c_oracle.evaluate([s for s in shards if len(s["Y"])], p["Z"], p["sf2"], p["alpha"], p["beta"], step_size=1e-3)
    res = _gpu_evaluate(shards, p["Z"], p["sf2"], p["alpha"], p["beta"], step_size=1e-3)
    for k, v in ref["stats"].items():
        assert relerr(res["stats"][k], v) < TOL, k
    assert abs(res["global"]["F"] - ref["global"]["F"]) <= TOL * abs(ref["global"]["F"])
    assert relerr(res["global"]["grad_Z"], ref["global"]["grad_Z"]) < TOL
    assert res["grad_latest"][0].shape == (2, 0, 2)
    got = np.concatenate(res["grad_latest"][1:], axis=1)
    assert relerr(got, np.concatenate(ref["grad_latest"], axis=1)) < TOL
