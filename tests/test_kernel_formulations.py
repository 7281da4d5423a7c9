"""CPU restatement of the two reformulated Psi2 kernels -- the arithmetic they execute, in numpy, from the same derived
quantities prep_points / pair_table_kernel hand them -- against the oracle (which follows the reference's formulas,
kernel_exp.py:126-148, partial_terms.py:190-205,273-284,367-431).  The GPU parity tests check the kernels; this pins
the algebra they rest on where no GPU is needed:
  psi2x_stats (psi2.cu):   t = fma(-w, zc, w mc), u = t^2 + v, exponent = lk + lc2 + sum alpha S + sum_q u_q (-1/w_q)
  embed_psi2m (embed_m.cu): E = kn + X R^T, h = exp(E), sums = h (G R), then gp_embed_finish_one (embed.cuh)."""
import numpy as np

from gparml_b200.synthetic import make_problem
from oracle import gparml_oracle as O


def _problem(seed, n=60, M=9, Q=4, D=3):
    p = make_problem(n, M, Q, D, seed=seed, generic_hypers=True)
    S = O.softplus(p["X_S"])
    return p, S


def _pairs(M):
    a, b = np.triu_indices(M)
    return a, b


def test_psi2x_formulation_equals_reference_statistics():
    p, S = _problem(11)
    Z, alpha, sf2, mu = p["Z"], p["alpha"], p["sf2"], p["X_mu"]
    n, Q = mu.shape
    M = Z.shape[0]
    ref = O.shard_statistics(p["Y"], mu, S, Z, sf2, alpha)
    # what prep_points writes into rec2x (prep.cu) and pair_table_kernel into pair_lk / pair_zc
    c = Z.mean(axis=0)
    w = alpha / (2.0 * alpha * S + 1.0)
    nw, wmc, v, nwinv = -w, w * (mu - c), alpha * S * w, -(2.0 * alpha * S + 1.0) / alpha
    lc2 = 2.0 * np.log(sf2) - 0.5 * np.sum(np.log(2.0 * alpha * S + 1.0), axis=1)
    kn = lc2 + np.sum(alpha * S, axis=1)
    a, b = _pairs(M)
    zc = 0.5 * (Z[a] + Z[b]) - c                                      # (P, Q)
    lk = -0.25 * np.sum(alpha * (Z[a] - Z[b]) ** 2, axis=1)
    # the point step of psi2x_stats for all (pair, point) at once
    t = nw[None, :, :] * zc[:, None, :] + wmc[None, :, :]             # (P, n, Q)
    u = t * t + v[None, :, :]
    e = lk[:, None] + kn[None, :] + np.sum(u * nwinv[None, :, :], axis=2)
    psi = np.exp(e)
    S0 = psi.sum(axis=1)
    TZ = np.einsum("pn,pnq->qp", psi, t)
    TA = np.einsum("pn,pnq->qp", psi, u)
    # the cancellation-free instantiation accumulates the same exponent as lk + lc2 + sum_q (t (-1/w)) t
    e_rob = lk[:, None] + lc2[None, :] + np.sum((t * nwinv[None, :, :]) * t, axis=2)
    assert np.max(np.abs(e_rob - e)) < 1e-12
    # expansion to the reference layouts (misc.cu expand_kernel)
    P2 = np.zeros((M, M)); P2[a, b] = S0; P2[b, a] = S0
    dZ = np.zeros((M, Q, M)); dA = np.zeros((Q, M, M))
    for q in range(Q):
        tzq = np.zeros((M, M)); tzq[a, b] = TZ[q]; tzq[b, a] = TZ[q]
        taq = np.zeros((M, M)); taq[a, b] = TA[q]; taq[b, a] = TA[q]
        dz = Z[:, None, q] - Z[None, :, q]
        dZ[:, q, :] = -0.5 * alpha[q] * dz * P2 + tzq
        dA[q] = -0.25 * dz * dz * P2 - taq / alpha[q] ** 2
    rel = lambda x, y: float(np.max(np.abs(x - y)) / np.max(np.abs(y)))
    assert rel(P2, ref["sum_exp_K_mi_K_im"]) < 1e-12
    assert rel(dZ, ref["sum_d_exp_K_mi_K_im_d_Z"]) < 1e-12
    assert rel(dA, ref["sum_d_exp_K_mi_K_im_d_alpha"]) < 1e-12


def test_embed_psi2m_formulation_equals_reference_gradients():
    p, S = _problem(12)
    Z, alpha, sf2, mu, Y = p["Z"], p["alpha"], p["sf2"], p["X_mu"], p["Y"]
    n, Q = mu.shape
    M, D = Z.shape[0], Y.shape[1]
    rng = np.random.default_rng(5)
    G2 = rng.standard_normal((M, M))                                  # any dF/dPsi2 (not symmetric on purpose)
    G2[2, 5] = -G2[5, 2]                                              # one pair whose symmetrised weight is exactly zero
    g_mu, g_S = O.embedding_grads(Y, mu, S, Z, sf2, alpha, np.zeros((M, D)), G2)
    # per-point features X and kn (prologue of embed_psi2m), per-pair features R (pair_table_kernel), pair factor G
    c = Z.mean(axis=0)
    w = alpha / (2.0 * alpha * S + 1.0)
    mc = mu - c
    X = np.concatenate([2.0 * w * mc, -w], axis=1)                    # (n, 2Q)
    kn = 2.0 * np.log(sf2) - 0.5 * np.sum(np.log(2.0 * alpha * S + 1.0), axis=1) - np.sum(w * mc * mc, axis=1)
    a, b = _pairs(M)
    zc = 0.5 * (Z[a] + Z[b]) - c
    R = np.concatenate([zc, zc * zc, np.ones((len(a), 1))], axis=1)   # (P, 2Q + 1)
    lk = -0.25 * np.sum(alpha * (Z[a] - Z[b]) ** 2, axis=1)
    Gs = np.where(a == b, G2[a, b], G2[a, b] + G2[b, a])              # gs_common.cuh: pair weight of the upper triangle
    with np.errstate(divide="ignore"):
        lg = np.log(np.abs(Gs))
    lkg = lk + np.where(lg > -700.0, lg, -700.0)                      # pair_h
    G = np.where(Gs < 0.0, -1.0, 1.0) * np.exp(lkg)                   # pair_ra_kernel
    E = kn[:, None] + X @ R[:, :2 * Q].T                              # first product
    h = np.exp(E)
    sums = h @ (G[:, None] * R)                                       # second product: (n, 2Q + 1)
    am, as_, ah = sums[:, :Q], sums[:, Q:2 * Q], sums[:, 2 * Q:2 * Q + 1]
    # gp_embed_finish_one with a zero Psi1 part
    t1 = w * (mc * ah - am)
    t2 = w * (w * (mc * (mc * ah - 2.0 * am) + as_))
    gm = -mu - 2.0 * t1
    gs = -0.5 * (1.0 - 1.0 / S) + (2.0 * t2 - w * ah)
    rel = lambda x, y: float(np.max(np.abs(x - y)) / np.max(np.abs(y)))
    assert rel(gm, g_mu) < 1e-11
    assert rel(gs, g_S) < 1e-11
