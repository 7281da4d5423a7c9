"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy) of GParML's map-reduce
variational-bound path.  It is the parity checker for the CUDA path and the
``cpu_baseline`` ("port") leg of ``bench.py``; the product (``gparml_b200``)
never imports it.

Pinning: this restatement is checked against the *live* reference
(``oracle/ref_shim.py`` imports ``/root/reference/partial_terms.py`` unmodified)
in ``tests/test_oracle.py`` and against the committed golden
vectors ``tests/golden/*.npz`` that ``oracle/gen_golden.py`` produced from the
live reference.  The reference itself ships no golden vectors or seeded tests
(SURVEY.md section 4), so the reference-run fixtures are the pin.

Every function cites the reference lines it restates (paths relative to
``/root/reference``).  The loop structure deliberately follows the reference
(one Python iteration per data point, an (M, M, Q) broadcast + ``exp`` per
point, derivative tensors built from a stored per-point Psi2 tensor) so that
timing this module is a fair stand-in for timing the reference's numpy path.

Notation: mu = X_mu, S = X_S (already positive), a_q = alpha_q (inverse squared
length-scale), Psi1 = <K_mi>, Psi2_n = <K_mi K_im> for point n.
"""
import math
import sys

import numpy as np

LIM_VAL = -math.log(sys.float_info.epsilon)  # supporting_functions.py:125


# --------------------------------------------------------------------------
# positive-parameter transforms (supporting_functions.py:125-168)
# --------------------------------------------------------------------------
def softplus(x):
    """supporting_functions.py:127-131,153-156 (transform / transformVar)."""
    x = np.asarray(x, dtype=np.float64)
    assert np.all(np.abs(x) < LIM_VAL)
    return np.log(1.0 + np.exp(x))


def softplus_inv(y):
    """supporting_functions.py:134-139,159-162 (transform_back / transformVar_back)."""
    y = np.asarray(y, dtype=np.float64)
    assert np.all(y > sys.float_info.epsilon) and np.all(y < LIM_VAL)
    return np.log(np.exp(y) - 1.0)


def softplus_grad(x):
    """supporting_functions.py:142-148,165-168 (transform_grad / transformVar_grad)."""
    x = np.asarray(x, dtype=np.float64)
    assert np.all(np.abs(x) < LIM_VAL)
    return 1.0 / (np.exp(-x) + 1.0)


# --------------------------------------------------------------------------
# kernel matrix and Psi statistics (kernels.py, kernel_exp.py)
# --------------------------------------------------------------------------
def kmm(Z, sf2, alpha):
    """kernels.py:72-113 with ard = alpha**-0.5 (partial_terms.py:32):
    sf^2 exp(-sum_q (z_mq - z_m'q)^2 / (2 ard_q^2))."""
    diff = Z[:, None, :] - Z[None, :, :]
    return sf2 * np.exp(-0.5 * np.einsum("abq,q->ab", diff * diff, alpha))


def psi1(Z, sf2, alpha, mu, S):
    """kernel_exp.py:51-82 -> (N, M)."""
    den = alpha[None, :] * S + 1.0                                    # (N, Q)
    pref = sf2 / np.sqrt(np.prod(den, axis=1))                        # (N,)
    diff = Z[None, :, :] - mu[:, None, :]                             # (N, M, Q)
    arg = np.sum(diff * diff * (alpha[None, :] / den)[:, None, :], axis=2)
    return pref[:, None] * np.exp(-0.5 * arg)


def psi1_y(P1, Y):
    """kernel_exp.py:13-49: sum over points of outer(Psi1[n], Y[n]) -> (M, D)."""
    out = np.zeros((P1.shape[1], Y.shape[1]))
    for row, y in zip(P1, Y):
        out += row[:, None] * y[None, :]
    return out


def psi2_point(Z, sf2, alpha, mu_n, S_n):
    """kernel_exp.py:126-148 for a single point -> (M, M)."""
    den = 2.0 * alpha * S_n + 1.0                                     # (Q,)
    pref = sf2 * sf2 / math.sqrt(np.prod(den))
    dz = Z[:, None, :] - Z[None, :, :]
    zbar = 0.5 * (Z[:, None, :] + Z[None, :, :])
    e_z = -0.25 * np.sum(alpha * dz * dz, axis=2)
    dm = mu_n - zbar
    e_x = -np.sum(alpha * dm * dm / den, axis=2)
    return pref * np.exp(e_z + e_x)


# --------------------------------------------------------------------------
# per-shard map: statistics (local_MapReduce.py:183-248, partial_terms.py)
# --------------------------------------------------------------------------
STAT_NAMES = (  # parallel_GPLVM.py:142-151
    "sum_YYT", "sum_exp_K_mi_K_im", "sum_exp_K_miY", "sum_exp_K_ii", "sum_KL",
    "sum_d_exp_K_miY_d_Z", "sum_d_exp_K_mi_K_im_d_Z",
    "sum_d_exp_K_miY_d_alpha", "sum_d_exp_K_mi_K_im_d_alpha",
    "sum_d_exp_K_ii_d_sf2", "sum_d_exp_K_miY_d_sf2", "sum_d_exp_K_mi_K_im_d_sf2",
)


def kl_term(mu, S):
    """partial_terms.py:83-87."""
    if np.all(S == 0):
        return 0.0
    return 0.5 * float(np.sum(np.sum(S - np.log(S), axis=1) + np.sum(mu * mu, axis=1) - mu.shape[1]))


def shard_statistics(Y, mu, S, Z, sf2, alpha):
    """The 12 partial sums one mapper emits for one shard
    (local_MapReduce.py:224-240 calling partial_terms.py:38-52,74-87,162-205,
    256-284,306-320).  One Python iteration per point, Psi2 tensor stored."""
    n, Q = mu.shape
    M = Z.shape[0]
    D = Y.shape[1]
    # set_data: partial_terms.py:40,45-49
    yyt = float(sum(float(y.dot(y)) for y in Y))
    P2 = np.zeros((n, M, M))
    for i in range(n):
        P2[i] = psi2_point(Z, sf2, alpha, mu[i], S[i])
    P1 = psi1(Z, sf2, alpha, mu, S)
    # update_local_statistics: partial_terms.py:79-87
    s_p2 = P2.sum(axis=0)
    s_p1y = psi1_y(P1, Y)
    s_kii = sf2 * n
    kl = kl_term(mu, S)

    # dexp_K_miY_dZ: partial_terms.py:162-188
    d1_dZ = np.zeros((M, Q, D))
    for j in range(M):
        for k in range(Q):
            wgt = P1[:, j] * alpha[k] * (mu[:, k] - Z[j, k]) / (alpha[k] * S[:, k] + 1.0)
            d1_dZ[j, k, :] = wgt.dot(Y)

    # dexp_K_mi_K_im_dZ: partial_terms.py:190-205
    d2_dZ = np.zeros((M, Q, M))
    dzT = Z[:, :, None] - Z.T[None, :, :]                              # (M, Q, M)
    szT = Z[:, :, None] + Z.T[None, :, :]
    aQ = alpha[None, :, None]
    for i in range(n):
        inner = (-0.5 * aQ * dzT
                 + 0.5 * aQ * (2.0 * mu[i][None, :, None] - szT) / (2.0 * aQ * S[i][None, :, None] + 1.0))
        d2_dZ += P2[i][:, None, :] * inner

    # dexp_K_miY_dalpha: partial_terms.py:256-271
    d1_da = np.zeros((Q, M, D))
    for q in range(Q):
        for i in range(n):
            den = alpha[q] * S[i, q] + 1.0
            v = -0.5 * P1[i] * (((mu[i, q] - Z[:, q]) / den) ** 2 + S[i, q] / den)
            d1_da[q] += v[:, None] * Y[i][None, :]

    # dexp_K_mi_K_im_dalpha: partial_terms.py:273-284
    d2_da = np.zeros((Q, M, M))
    dz2 = np.moveaxis((Z[:, None, :] - Z[None, :, :]) ** 2, 2, 0)      # (Q, M, M)
    sz = np.moveaxis(Z[:, None, :] + Z[None, :, :], 2, 0)              # (Q, M, M)
    for i in range(n):
        den = (2.0 * alpha * S[i] + 1.0)[:, None, None]
        inner = (-0.25 * dz2
                 - 0.25 * ((2.0 * mu[i][:, None, None] - sz) / den) ** 2
                 - S[i][:, None, None] / den)
        d2_da += P2[i][None, :, :] * inner

    return {
        "sum_YYT": yyt,
        "sum_exp_K_ii": s_kii,
        "sum_exp_K_mi_K_im": s_p2,
        "sum_exp_K_miY": s_p1y,
        "sum_KL": kl,
        "sum_d_exp_K_miY_d_Z": d1_dZ,
        "sum_d_exp_K_mi_K_im_d_Z": d2_dZ,
        "sum_d_exp_K_miY_d_alpha": d1_da,
        "sum_d_exp_K_mi_K_im_d_alpha": d2_da,
        # partial_terms.py:306-320
        "sum_d_exp_K_ii_d_sf2": float(n),
        "sum_d_exp_K_miY_d_sf2": s_p1y / sf2,
        "sum_d_exp_K_mi_K_im_d_sf2": 2.0 * s_p2 / sf2,
    }


def shard_statistics_chunked(Y, mu, S, Z, sf2, alpha, chunk=512):
    """Same sums as :func:`shard_statistics`, accumulated over row chunks so the
    (n, M, M) tensor stays bounded (all 12 statistics are additive over points;
    SURVEY.md section 8c).  KL's all-zero test is applied to the whole shard."""
    n = mu.shape[0]
    total = None
    for lo in range(0, n, chunk):
        part = shard_statistics(Y[lo:lo + chunk], mu[lo:lo + chunk], S[lo:lo + chunk], Z, sf2, alpha)
        total = part if total is None else reduce_statistics([total, part])
    if np.all(S == 0):
        total["sum_KL"] = 0.0
    return total


def reduce_statistics(parts):
    """statistics_reducer: local_MapReduce.py:250-277 (sum per key; no dropout)."""
    out = {}
    for k in STAT_NAMES:
        acc = parts[0][k]
        acc = acc.copy() if isinstance(acc, np.ndarray) else acc
        for p in parts[1:]:
            acc = acc + p[k]
        out[k] = acc
    return out


# --------------------------------------------------------------------------
# master: global step (parallel_GPLVM.py:302-369, partial_terms.py:54-61,
# 102-160, 207-360, 436-473)
# --------------------------------------------------------------------------
def global_step(stats, Z, sf2, alpha, beta, N, fixed_beta=False):
    """Returns dict with F, the partial derivatives of F w.r.t. the summed
    statistics and Kmm, and the gradients w.r.t. Z, sf2, alpha, beta."""
    M, Q = Z.shape
    P2 = stats["sum_exp_K_mi_K_im"]
    P1Y = stats["sum_exp_K_miY"]
    D = P1Y.shape[1]
    K = kmm(Z, sf2, alpha)                                  # local_MapReduce.py:386-387
    Kinv = np.linalg.inv(K)                                 # local_MapReduce.py:392
    A = K + beta * P2
    Ainv = np.linalg.inv(A)                                 # partial_terms.py:60

    # logmarglik: partial_terms.py:436-473
    s1, ldK = np.linalg.slogdet(K)
    s2, ldA = np.linalg.slogdet(A)
    A_for_quad = A
    if s1 < 0:
        s1, ldK = np.linalg.slogdet(K + np.eye(M) * 1e-7)
    if s2 < 0:
        A_for_quad = A + np.eye(M) * 1e-7
        s2, ldA = np.linalg.slogdet(A_for_quad)
    assert s1 >= 0 and s2 >= 0
    F = (-0.5 * N * D * math.log(2.0 * math.pi)
         + 0.5 * D * N * math.log(beta)
         + 0.5 * D * ldK
         - 0.5 * D * ldA
         - 0.5 * beta * stats["sum_YYT"]
         - 0.5 * beta * D * stats["sum_exp_K_ii"]
         + 0.5 * beta * D * np.trace(Kinv.dot(P2))
         + 0.5 * beta ** 2 * np.trace(P1Y.T.dot(np.linalg.inv(A_for_quad).dot(P1Y)))
         - stats["sum_KL"])

    # partial_terms.py:102-138
    C = Ainv.dot(P1Y)
    E = C.dot(C.T)
    dF_dKmm = 0.5 * D * Kinv - 0.5 * D * Ainv - 0.5 * beta * D * Kinv.dot(P2.dot(Kinv)) - 0.5 * beta ** 2 * E
    dF_dP1Y = beta ** 2 * C
    dF_dP2 = -0.5 * beta * D * Ainv + 0.5 * beta * D * Kinv - 0.5 * beta ** 3 * E
    dF_dKii = -0.5 * beta * D

    # grad_Z: partial_terms.py:146-160,207-240
    dK_dZ = K[:, None, :] * (-alpha[None, :, None]) * (Z[:, :, None] - Z.T[None, :, :])
    # the reference masks row j and column j of an (M, M) zero matrix (:226-228);
    # the doubly-hit [j, j] entry of dK_dZ is exactly 0, so row + column = sym.
    sym = dF_dKmm + dF_dKmm.T
    gZ = (np.einsum("jm,jkm->jk", sym, dK_dZ)
          + np.einsum("jd,jkd->jk", dF_dP1Y, stats["sum_d_exp_K_miY_d_Z"])
          + 2.0 * np.einsum("jm,jkm->jk", dF_dP2, stats["sum_d_exp_K_mi_K_im_d_Z"]))

    # grad_alpha: partial_terms.py:247-254,286-299
    dK_da = -0.5 * K[None, :, :] * np.moveaxis((Z[:, None, :] - Z[None, :, :]) ** 2, 2, 0)
    ga = (np.einsum("ab,qab->q", dF_dKmm, dK_da)
          + np.einsum("md,qmd->q", dF_dP1Y, stats["sum_d_exp_K_miY_d_alpha"])
          + np.einsum("ab,qab->q", dF_dP2, stats["sum_d_exp_K_mi_K_im_d_alpha"]))

    # grad_sf2: partial_terms.py:306-333
    gsf2 = (np.sum(dF_dKmm * K / sf2)
            + dF_dKii * stats["sum_d_exp_K_ii_d_sf2"]
            + np.sum(dF_dP1Y * stats["sum_d_exp_K_miY_d_sf2"])
            + np.sum(dF_dP2 * stats["sum_d_exp_K_mi_K_im_d_sf2"]))

    # grad_beta: partial_terms.py:340-360
    if fixed_beta:                                          # parallel_GPLVM.py:363-366
        gbeta = 0.0
    else:
        gbeta = (0.5 * N * D / beta
                 - 0.5 * D * np.trace(Ainv.dot(P2))
                 - 0.5 * stats["sum_YYT"]
                 - 0.5 * D * stats["sum_exp_K_ii"]
                 + 0.5 * D * np.trace(Kinv.dot(P2))
                 + beta * np.trace(P1Y.T.dot(C))
                 - 0.5 * beta ** 2 * np.trace(C.T.dot(P2).dot(C)))
    return {
        "F": float(F), "Kmm": K, "Kmm_inv": Kinv, "Kmm_plus_op_inv": Ainv,
        "dF_dKmm": dF_dKmm, "dF_dsum_exp_K_miY": dF_dP1Y,
        "dF_dsum_exp_K_mi_K_im": dF_dP2, "dF_dsum_exp_K_ii": dF_dKii,
        "grad_Z": gZ, "grad_sf2": float(gsf2), "grad_alpha": ga, "grad_beta": float(gbeta),
        "cond_Kmm": float(np.linalg.cond(K)),
    }


# --------------------------------------------------------------------------
# per-shard map: embedding gradients (partial_terms.py:367-431)
# --------------------------------------------------------------------------
def embedding_grads(Y, mu, S, Z, sf2, alpha, dF_dP1Y, dF_dP2):
    """grad_X_mu (partial_terms.py:367-398) and grad_X_S (:400-431), gradient of
    F w.r.t. the (positive-domain) variational mean / variance of each point.
    One Python iteration per point and per latent dimension like the reference."""
    n, Q = mu.shape
    P1 = psi1(Z, sf2, alpha, mu, S)
    g_mu = np.zeros((n, Q))
    g_S = np.zeros((n, Q))
    for i in range(n):
        P2i = psi2_point(Z, sf2, alpha, mu[i], S[i])
        g_mu[i] = -mu[i]                                               # :385
        g_S[i] = -0.5 * (1.0 - 1.0 / S[i])                             # :418
        for q in range(Q):
            d1 = alpha[q] * S[i, q] + 1.0
            d2 = 2.0 * alpha[q] * S[i, q] + 1.0
            dm = mu[i, q] - Z[:, q]
            sm = 2.0 * mu[i, q] - Z[:, None, q] - Z[None, :, q]
            # :388-393
            t1 = np.outer(P1[i] * (-alpha[q] * dm / d1), Y[i])
            t2 = P2i * (-alpha[q]) * sm / d2
            g_mu[i, q] += np.sum(dF_dP1Y * t1) + np.sum(dF_dP2 * t2)
            # :421-427
            u1 = np.outer(P1[i] * (0.5 * (alpha[q] * dm / d1) ** 2 - 0.5 * alpha[q] / d1), Y[i])
            u2 = P2i * (2.0 * (alpha[q] * sm / (2.0 * d2)) ** 2 - alpha[q] / d2)
            g_S[i, q] += np.sum(dF_dP1Y * u1) + np.sum(dF_dP2 * u2)
    return g_mu, g_S


# --------------------------------------------------------------------------
# mapper glue + one full evaluation (local_MapReduce.py:183-248,310-363;
# parallel_GPLVM.py:222-279)
# --------------------------------------------------------------------------
def effective_embedding(mu0, s_raw0, direction, step_size, fixed_embeddings):
    """local_MapReduce.py:203-214 / :331-341: apply the local step in memory and
    map the variance to the positive domain.  Returns (mu, S, s_raw_effective)."""
    mu = np.array(mu0, dtype=np.float64, copy=True)
    s_raw = np.array(s_raw0, dtype=np.float64, copy=True)
    if fixed_embeddings:
        return mu, s_raw, s_raw
    if direction is not None and step_size != 0:
        mu += direction[0] * step_size
        s_raw += direction[1] * step_size
    return mu, softplus(s_raw), s_raw


def flatten_globals(Z, sf2, alpha, beta):
    """parallel_GPLVM.py:286-290 with key order Z, sf2, alpha, beta (:139-141)."""
    return np.concatenate([np.ravel(Z), [sf2], np.ravel(alpha), [beta]])


def evaluate(shards, Z, sf2, alpha, beta, step_size=0.0, fixed_embeddings=False,
             fixed_beta=False, chunk=512):
    """One ELBO + gradient evaluation over a list of shards.

    ``shards`` is a list of dicts with keys Y, X_mu, X_S (X_S in the
    *unconstrained* domain unless fixed_embeddings, local_MapReduce.py:90-93,103)
    and optionally ``d`` (the (2, n, Q) local search direction).
    Returns dict with F, the positive-domain global gradients, the reduced
    statistics and per-shard ``grad_latest`` = -[g_mu, g_S * sigmoid(S_raw)]
    (local_MapReduce.py:357-360)."""
    N = sum(s["Y"].shape[0] for s in shards)
    eff = [effective_embedding(s["X_mu"], s["X_S"], s.get("d"), step_size, fixed_embeddings) for s in shards]
    parts = [shard_statistics_chunked(s["Y"], e[0], e[1], Z, sf2, alpha, chunk) for s, e in zip(shards, eff)]
    stats = reduce_statistics(parts)
    g = global_step(stats, Z, sf2, alpha, beta, N, fixed_beta=fixed_beta)
    out = {"stats": stats, "global": g, "grad_latest": []}
    if not fixed_embeddings:
        for s, e in zip(shards, eff):
            gm, gs = embedding_grads(s["Y"], e[0], e[1], Z, sf2, alpha,
                                     g["dF_dsum_exp_K_miY"], g["dF_dsum_exp_K_mi_K_im"])
            out["grad_latest"].append(-np.array([gm, gs * softplus_grad(e[2])]))
    return out


def objective_and_flat_gradient(res, x_unconstrained, M, Q):
    """parallel_GPLVM.py:266-279: (-F, -grad * transform_grad) on the flat vector
    [Z, sf2, alpha, beta]; Z entries are unbounded, the rest softplus-mapped."""
    g = res["global"]
    flat = flatten_globals(g["grad_Z"], g["grad_sf2"], g["grad_alpha"], g["grad_beta"])
    chain = np.ones_like(flat)
    chain[M * Q:] = softplus_grad(x_unconstrained[M * Q:])
    return -g["F"], -flat * chain


# --------------------------------------------------------------------------
# SCG local-state operations (scg_adapted_local_MapReduce.py:29-243) on
# in-memory per-shard dicts {latest,new,old,d: (2,n,Q); X_mu, X_S}
# --------------------------------------------------------------------------
def scg_set_grads(st):                     # :29-55
    for s in st:
        s["new"] = s["latest"].copy(); s["old"] = s["latest"].copy(); s["d"] = -s["latest"]


def scg_get_mu(st):                        # :60-75
    return float(sum((s["new"] * s["d"]).sum() for s in st))


def scg_get_kappa(st):                     # :77-90
    return float(sum((s["d"] * s["d"]).sum() for s in st))


def scg_get_theta(st):                     # :92-109
    return float(sum((s["d"] * (s["latest"] - s["new"])).sum() for s in st))


def scg_get_current_grad(st):              # :111-124
    return float(sum((s["new"] * s["new"]).sum() for s in st))


def scg_get_gamma(st):                     # :126-141
    return float(sum((s["new"] * s["old"]).sum() for s in st))


def scg_get_max_d(st, alpha):              # :143-156
    return float(max(np.max(np.abs(alpha * s["d"])) for s in st))


def scg_reset_d(st):                       # :161-174
    for s in st:
        s["d"] = -s["new"]


def scg_update_d(st, gamma):               # :176-191
    for s in st:
        s["d"] = gamma * s["d"] - s["new"]


def scg_update_X(st, alpha):               # :193-216
    for s in st:
        s["X_mu"] = s["X_mu"] + alpha * s["d"][0]
        s["X_S"] = s["X_S"] + alpha * s["d"][1]


def scg_update_grad_old(st):               # :218-230
    for s in st:
        s["old"] = s["new"].copy()


def scg_update_grad_new(st):               # :232-243
    for s in st:
        s["new"] = s["latest"].copy()
