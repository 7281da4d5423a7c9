"""TEST INFRASTRUCTURE ONLY -- ctypes front-end of the plain-C oracle
(``oracle/gparml_oracle.c``).  Same outputs as the per-shard maps of
``oracle/gparml_oracle.py`` but fast enough (OpenMP over inducing points /
data points) for the full BASELINE shapes at a few thousand points.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "gparml_oracle.c")
LIB = os.path.join(HERE, "libgparml_oracle.so")

_lib = None
_dp = ctypes.POINTER(ctypes.c_double)


def build(force=False):
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        subprocess.check_call(["gcc", "-O2", "-fopenmp", "-shared", "-fPIC", "-o", LIB, SRC, "-lm"])
    return LIB


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        _lib = ctypes.CDLL(LIB)
        _lib.oracle_shard_stats.restype = None
        _lib.oracle_embed_grads.restype = None
        _lib.oracle_psi1.restype = None
    return _lib


def _p(a):
    return a.ctypes.data_as(_dp)


def _c(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def shard_statistics(Y, mu, S, Z, sf2, alpha):
    """The 12 named partial sums for one shard (same dict as the numpy oracle)."""
    lib = _load()
    Y, mu, S, Z, alpha = _c(Y), _c(mu), _c(S), _c(Z), _c(alpha)
    n, Q = mu.shape
    M, D = Z.shape[0], Y.shape[1]
    psi2 = np.empty((M, M)); psi1y = np.empty((M, D))
    d1z = np.empty((M, Q, D)); d2z = np.empty((M, Q, M))
    d1a = np.empty((Q, M, D)); d2a = np.empty((Q, M, M)); scal = np.empty(2)
    lib.oracle_shard_stats(ctypes.c_long(n), M, Q, D, _p(Y), _p(mu), _p(S), _p(Z), ctypes.c_double(sf2),
                           _p(alpha), _p(psi2), _p(psi1y), _p(d1z), _p(d2z), _p(d1a), _p(d2a), _p(scal))
    return {
        "sum_YYT": float(scal[0]), "sum_exp_K_ii": sf2 * n, "sum_exp_K_mi_K_im": psi2,
        "sum_exp_K_miY": psi1y, "sum_KL": float(scal[1]),
        "sum_d_exp_K_miY_d_Z": d1z, "sum_d_exp_K_mi_K_im_d_Z": d2z,
        "sum_d_exp_K_miY_d_alpha": d1a, "sum_d_exp_K_mi_K_im_d_alpha": d2a,
        "sum_d_exp_K_ii_d_sf2": float(n), "sum_d_exp_K_miY_d_sf2": psi1y / sf2,
        "sum_d_exp_K_mi_K_im_d_sf2": 2.0 * psi2 / sf2,
    }


def embedding_grads(Y, mu, S, Z, sf2, alpha, G1, G2):
    lib = _load()
    Y, mu, S, Z, alpha, G1, G2 = _c(Y), _c(mu), _c(S), _c(Z), _c(alpha), _c(G1), _c(G2)
    n, Q = mu.shape
    M, D = Z.shape[0], Y.shape[1]
    gm = np.empty((n, Q)); gs = np.empty((n, Q))
    lib.oracle_embed_grads(ctypes.c_long(n), M, Q, D, _p(Y), _p(mu), _p(S), _p(Z), ctypes.c_double(sf2),
                           _p(alpha), _p(G1), _p(G2), _p(gm), _p(gs))
    return gm, gs


def psi1(Z, sf2, alpha, mu, S):
    lib = _load()
    mu, S, Z, alpha = _c(mu), _c(S), _c(Z), _c(alpha)
    n, Q = mu.shape
    M = Z.shape[0]
    out = np.empty((n, M))
    lib.oracle_psi1(ctypes.c_long(n), M, Q, _p(mu), _p(S), _p(Z), ctypes.c_double(sf2), _p(alpha), _p(out))
    return out


def evaluate(shards, Z, sf2, alpha, beta, step_size=0.0, fixed_embeddings=False, fixed_beta=False):
    """Full evaluation with the C maps and the numpy master step (global_step is
    O(M^3) and identical to the numpy oracle's)."""
    from . import gparml_oracle as O
    N = sum(s["Y"].shape[0] for s in shards)
    eff = [O.effective_embedding(s["X_mu"], s["X_S"], s.get("d"), step_size, fixed_embeddings) for s in shards]
    parts = [shard_statistics(s["Y"], e[0], e[1], Z, sf2, alpha) for s, e in zip(shards, eff)]
    stats = O.reduce_statistics(parts)
    if fixed_embeddings:
        stats["sum_KL"] = 0.0
    g = O.global_step(stats, Z, sf2, alpha, beta, N, fixed_beta=fixed_beta)
    out = {"stats": stats, "global": g, "grad_latest": []}
    if not fixed_embeddings:
        for s, e in zip(shards, eff):
            gm, gs = embedding_grads(s["Y"], e[0], e[1], Z, sf2, alpha,
                                     g["dF_dsum_exp_K_miY"], g["dF_dsum_exp_K_mi_K_im"])
            out["grad_latest"].append(-np.array([gm, gs * O.softplus_grad(e[2])]))
    return out
