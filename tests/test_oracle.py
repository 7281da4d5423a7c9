"""CPU tests: the numpy and C oracles against the golden vectors generated from the
live reference, against each other, and (where /root/reference exists) against the
live reference itself.  Mirrors the reference's own test strategy (test.py: finite
difference checks on the fixture D=7, Q=2, N=5, M=10)."""
import numpy as np
import pytest

from conftest import GOLDEN_CASES, check_against_golden, load_golden, relerr
from gparml_b200.synthetic import make_problem, split_rows
from oracle import c_oracle, gparml_oracle as O, ref_shim

TOL = 1e-9   # north_star fp64 parity bar


@pytest.mark.parametrize("name", [c for c in GOLDEN_CASES if c != "c4s"])
def test_numpy_oracle_matches_golden(name):
    g = load_golden(name)
    res = O.evaluate(g["shards"], g["Z"], g["sf2"], g["alpha"], g["beta"], step_size=g["step_size"],
                     fixed_embeddings=g["fixed_embeddings"], chunk=16)
    check_against_golden(res, g, TOL, "numpy oracle " + name)


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_c_oracle_matches_golden(name):
    g = load_golden(name)
    res = c_oracle.evaluate(g["shards"], g["Z"], g["sf2"], g["alpha"], g["beta"], step_size=g["step_size"],
                            fixed_embeddings=g["fixed_embeddings"])
    check_against_golden(res, g, TOL, "C oracle " + name)


def test_c_oracle_matches_numpy_oracle_random():
    p = make_problem(70, 9, 3, 5, seed=3, generic_hypers=True, with_direction=True)
    shards = [dict(Y=p["Y"][lo:hi], X_mu=p["X_mu"][lo:hi], X_S=p["X_S"][lo:hi], d=p["d"][:, lo:hi])
              for lo, hi in split_rows(70, 3)]
    a = O.evaluate(shards, p["Z"], p["sf2"], p["alpha"], p["beta"], step_size=2e-3)
    b = c_oracle.evaluate(shards, p["Z"], p["sf2"], p["alpha"], p["beta"], step_size=2e-3)
    for k in O.STAT_NAMES:
        assert relerr(b["stats"][k], a["stats"][k]) < 1e-12, k
    for k in ("F", "grad_Z", "grad_alpha", "grad_sf2", "grad_beta"):
        assert relerr(b["global"][k], a["global"][k]) < 1e-10, k
    for x, y in zip(b["grad_latest"], a["grad_latest"]):
        assert relerr(x, y) < 1e-11


def test_psi1_c_vs_numpy():
    p = make_problem(33, 7, 4, 2, seed=5, generic_hypers=True)
    S = O.softplus(p["X_S"])
    assert relerr(c_oracle.psi1(p["Z"], p["sf2"], p["alpha"], p["X_mu"], S),
                  O.psi1(p["Z"], p["sf2"], p["alpha"], p["X_mu"], S)) < 1e-13


def _fd_problem():
    # reference fixture shape, test.py:24-60
    g = load_golden("t5")
    sh = g["shards"][0]
    return g, sh


def _F_of(g, sh, Z=None, sf2=None, alpha=None, beta=None, mu=None, S_raw=None):
    shard = dict(Y=sh["Y"], X_mu=sh["X_mu"] if mu is None else mu, X_S=sh["X_S"] if S_raw is None else S_raw)
    r = O.evaluate([shard], g["Z"] if Z is None else Z, g["sf2"] if sf2 is None else sf2,
                   g["alpha"] if alpha is None else alpha, g["beta"] if beta is None else beta)
    return r


def test_finite_differences_global_and_local():
    """Central differences of the oracle's own bound reproduce every gradient block
    (the reference asserts 1 % with forward differences, test.py:92,184,201,282,296)."""
    g, sh = _fd_problem()
    base = _F_of(g, sh)
    h = 1e-6
    # Z
    Z = g["Z"]
    for (j, k) in [(0, 0), (3, 1), (9, 0)]:
        Zp, Zm = Z.copy(), Z.copy(); Zp[j, k] += h; Zm[j, k] -= h
        fd = (_F_of(g, sh, Z=Zp)["global"]["F"] - _F_of(g, sh, Z=Zm)["global"]["F"]) / (2 * h)
        assert abs(fd - base["global"]["grad_Z"][j, k]) <= 1e-6 * max(1.0, abs(fd))
    # sf2, beta, alpha
    fd = (_F_of(g, sh, sf2=g["sf2"] + h)["global"]["F"] - _F_of(g, sh, sf2=g["sf2"] - h)["global"]["F"]) / (2 * h)
    assert abs(fd - base["global"]["grad_sf2"]) <= 1e-6 * max(1.0, abs(fd))
    fd = (_F_of(g, sh, beta=g["beta"] + h)["global"]["F"] - _F_of(g, sh, beta=g["beta"] - h)["global"]["F"]) / (2 * h)
    assert abs(fd - base["global"]["grad_beta"]) <= 1e-6 * max(1.0, abs(fd))
    for q in range(2):
        ap, am = g["alpha"].copy(), g["alpha"].copy(); ap[q] += h; am[q] -= h
        fd = (_F_of(g, sh, alpha=ap)["global"]["F"] - _F_of(g, sh, alpha=am)["global"]["F"]) / (2 * h)
        assert abs(fd - base["global"]["grad_alpha"][q]) <= 1e-6 * max(1.0, abs(fd))
    # local mean / unconstrained variance: grad_latest = -dF/d(.)
    for (i, q) in [(0, 0), (4, 1)]:
        mp, mm = sh["X_mu"].copy(), sh["X_mu"].copy(); mp[i, q] += h; mm[i, q] -= h
        fd = (_F_of(g, sh, mu=mp)["global"]["F"] - _F_of(g, sh, mu=mm)["global"]["F"]) / (2 * h)
        assert abs(fd + base["grad_latest"][0][0, i, q]) <= 1e-6 * max(1.0, abs(fd))
        sp, sm = sh["X_S"].copy(), sh["X_S"].copy(); sp[i, q] += h; sm[i, q] -= h
        fd = (_F_of(g, sh, S_raw=sp)["global"]["F"] - _F_of(g, sh, S_raw=sm)["global"]["F"]) / (2 * h)
        assert abs(fd + base["grad_latest"][0][1, i, q]) <= 1e-6 * max(1.0, abs(fd))


def test_transforms_roundtrip_and_limits():
    x = np.array([-30.0, -1.0, 0.0, 2.5, 30.0])
    assert np.allclose(O.softplus_inv(O.softplus(x[1:])), x[1:], atol=1e-9)
    assert np.allclose(O.softplus_grad(x), 1.0 / (1.0 + np.exp(-x)))
    with pytest.raises(AssertionError):           # supporting_functions.py:154 |x| < 36.04
        O.softplus(np.array([40.0]))


def test_fixed_embeddings_KL_zero_and_no_local_grads():
    g = load_golden("c2s")
    res = O.evaluate(g["shards"], g["Z"], g["sf2"], g["alpha"], g["beta"], fixed_embeddings=True, chunk=32)
    assert res["stats"]["sum_KL"] == 0.0
    assert res["grad_latest"] == []


def test_scg_local_ops_numpy():
    rng = np.random.default_rng(0)
    st = [dict(latest=rng.standard_normal((2, 5, 3)), X_mu=rng.standard_normal((5, 3)),
               X_S=rng.standard_normal((5, 3))) for _ in range(2)]
    O.scg_set_grads(st)
    assert O.scg_get_mu(st) == pytest.approx(-O.scg_get_current_grad(st))
    assert O.scg_get_kappa(st) == pytest.approx(O.scg_get_current_grad(st))
    assert O.scg_get_theta(st) == 0.0
    x0 = st[0]["X_mu"].copy()
    O.scg_update_X(st, 0.5)
    assert np.allclose(st[0]["X_mu"], x0 + 0.5 * st[0]["d"][0])
    O.scg_update_d(st, 0.25)
    assert np.allclose(st[1]["d"], -0.25 * st[1]["latest"] - st[1]["new"])
    assert O.scg_get_max_d(st, 2.0) == pytest.approx(max(np.max(np.abs(2.0 * s["d"])) for s in st))


@pytest.mark.skipif(not ref_shim.reference_available(), reason="live reference tree not present")
@pytest.mark.parametrize("shape", [(40, 6, 3, 4, False), (30, 5, 2, 1, True), (24, 10, 2, 7, False)])
def test_numpy_oracle_matches_live_reference(shape):
    from oracle import ref_harness as R
    N, M, Q, D, fe = shape
    p = make_problem(N, M, Q, D, seed=1, fixed_embeddings=fe, generic_hypers=True, with_direction=True)
    shards = [dict(Y=p["Y"][lo:hi], X_mu=p["X_mu"][lo:hi], X_S=p["X_S"][lo:hi], d=p["d"][:, lo:hi])
              for lo, hi in split_rows(N, 3)]
    for step in (0.0, 1e-3):
        a = O.evaluate(shards, p["Z"], p["sf2"], p["alpha"], p["beta"], step_size=step, fixed_embeddings=fe, chunk=7)
        b = R.reference_evaluate(shards, p["Z"], p["sf2"], p["alpha"], p["beta"], step_size=step, fixed_embeddings=fe)
        for k in O.STAT_NAMES:
            assert relerr(a["stats"][k], b["stats"][k]) < 1e-12, k
        for k in ("F", "grad_Z", "grad_alpha", "grad_sf2", "grad_beta", "dF_dKmm",
                  "dF_dsum_exp_K_miY", "dF_dsum_exp_K_mi_K_im"):
            assert relerr(a["global"][k], b["global"][k]) < 1e-10, k
        for x, y in zip(a["grad_latest"], b["grad_latest"]):
            assert relerr(x, y) < 1e-11


def test_bench_expected_F_c2():
    """bench.py asserts F against tests/golden/bench_expected_F.json at every rank count; the c2 value (N = 100k,
    M = 50, Q = 4, D = 1, fixed embeddings) is small enough for the C oracle at its full size."""
    import json
    import os
    from gparml_b200.synthetic import block_problem_globals, block_problem_rows
    from oracle import c_oracle
    want = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "bench_expected_F.json")))["c2"]
    g = block_problem_globals("c2")
    rows = block_problem_rows("c2", 0, g["N"], with_direction=False)
    ref = c_oracle.evaluate([rows], g["Z"], g["sf2"], g["alpha"], g["beta"], fixed_embeddings=True)
    assert abs(ref["global"]["F"] - want) <= 1e-9 * abs(want), (ref["global"]["F"], want)
