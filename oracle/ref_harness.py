"""TEST INFRASTRUCTURE ONLY -- replays one GParML evaluation around the *live,
unmodified* reference ``partial_terms`` class (via ``oracle/ref_shim.py``).

``local_MapReduce.py`` / ``parallel_GPLVM.py`` are Python-2-only and cannot be
imported, so the bodies of ``statistics_mapper`` (local_MapReduce.py:183-248),
``statistics_reducer`` (:250-277), ``calculate_global_statistics`` /
``calculate_global_derivatives`` (parallel_GPLVM.py:302-369) and
``embeddings_mapper`` (local_MapReduce.py:310-363) are replayed here on
in-memory arrays, every arithmetic call going to the reference object.
Used by ``oracle/gen_golden.py``, ``tests/test_oracle.py`` and -- as the CPU arm of ``bench.py``
(``--impl reference`` / ``cpu_baseline``) -- timed on the GPU box's host cores, where the three reference
files are present as the git-ignored staging copy ``oracle/_ref/`` (``oracle/stage_ref.py``).
"""
import numpy as np

from . import ref_shim

STAT_NAMES = (
    "sum_YYT", "sum_exp_K_mi_K_im", "sum_exp_K_miY", "sum_exp_K_ii", "sum_KL",
    "sum_d_exp_K_miY_d_Z", "sum_d_exp_K_mi_K_im_d_Z",
    "sum_d_exp_K_miY_d_alpha", "sum_d_exp_K_mi_K_im_d_alpha",
    "sum_d_exp_K_ii_d_sf2", "sum_d_exp_K_miY_d_sf2", "sum_d_exp_K_mi_K_im_d_sf2",
)


def _softplus(x):
    return np.log(1.0 + np.exp(x))


def _sigmoid(x):
    return 1.0 / (np.exp(-x) + 1.0)


def _new_pt(Z, sf2, alpha, beta, N, D):
    pt, _, kernels = ref_shim.load_reference()
    M, Q = Z.shape
    obj = pt.partial_terms(Z.copy(), float(sf2), np.array(alpha, dtype=np.float64), float(beta),
                           M, Q, N, D, update_global_statistics=False)
    # local_MapReduce.cache (:383-394) + load_cache (:396-401)
    kern = kernels.rbf(Q, sf=float(sf2) ** 0.5, ard=np.array(alpha, dtype=np.float64) ** -0.5)
    Kmm = kern.K(Z)
    obj.set_global_statistics(Kmm, np.linalg.inv(Kmm))
    return obj


def _effective(shard, step_size, fixed_embeddings):
    mu = shard["X_mu"].copy()
    s_raw = shard["X_S"].copy()
    if not fixed_embeddings:
        if shard.get("d") is not None and step_size != 0:
            mu += shard["d"][0] * step_size
            s_raw += shard["d"][1] * step_size
        return mu, _softplus(s_raw), s_raw
    return mu, s_raw, s_raw


def map_statistics(s, Z, sf2, alpha, beta, N, D, step_size=0.0, fixed_embeddings=False):
    """Body of statistics_mapper (local_MapReduce.py:183-248) for one shard, on in-memory arrays: the 12 partial
    sums, every arithmetic call going to the reference's own partial_terms object."""
    mu, S, _ = _effective(s, step_size, fixed_embeddings)
    o = _new_pt(Z, sf2, alpha, beta, N, D)
    o.set_data(s["Y"], mu, S, is_set_statistics=True)
    t = o.get_local_statistics()
    return {
        "sum_YYT": t["sum_YYT"], "sum_exp_K_ii": t["sum_exp_K_ii"],
        "sum_exp_K_mi_K_im": t["sum_exp_K_mi_K_im"], "sum_exp_K_miY": t["exp_K_miY"],
        "sum_KL": t["KL"],
        "sum_d_exp_K_miY_d_Z": o.dexp_K_miY_dZ(),
        "sum_d_exp_K_mi_K_im_d_Z": o.dexp_K_mi_K_im_dZ(),
        "sum_d_exp_K_miY_d_alpha": o.dexp_K_miY_dalpha(),
        "sum_d_exp_K_mi_K_im_d_alpha": o.dexp_K_mi_K_im_dalpha(),
        "sum_d_exp_K_ii_d_sf2": o.dexp_K_ii_dsf2(),
        "sum_d_exp_K_miY_d_sf2": o.dexp_K_miY_dsf2(),
        "sum_d_exp_K_mi_K_im_d_sf2": o.dexp_K_mi_K_im_dsf2(),
    }


def reduce_statistics(parts):
    """statistics_reducer (local_MapReduce.py:250-277): per-key sum over the shards."""
    stats = {}
    for k in STAT_NAMES:
        acc = parts[0][k]
        for p in parts[1:]:
            acc = acc + p[k]
        stats[k] = acc
    return stats


def map_embeddings(s, stats, Z, sf2, alpha, beta, N, D, step_size=0.0):
    """Body of embeddings_mapper (local_MapReduce.py:310-363) for one shard: grad_latest (2, n, Q)."""
    mu, S, s_raw = _effective(s, step_size, False)
    o = _new_pt(Z, sf2, alpha, beta, N, D)
    o.set_data(s["Y"], mu, S, is_set_statistics=False)
    o.set_local_statistics(stats["sum_YYT"], stats["sum_exp_K_mi_K_im"], stats["sum_exp_K_miY"],
                           stats["sum_exp_K_ii"], stats["sum_KL"])
    gm = o.grad_X_mu()
    gs = o.grad_X_S() * _sigmoid(s_raw)
    return -1 * np.array([gm, gs])


def reference_evaluate(shards, Z, sf2, alpha, beta, step_size=0.0, fixed_embeddings=False,
                       fixed_beta=False):
    N = sum(s["Y"].shape[0] for s in shards)
    D = shards[0]["Y"].shape[1]
    parts = [map_statistics(s, Z, sf2, alpha, beta, N, D, step_size, fixed_embeddings) for s in shards]
    stats = reduce_statistics(parts)
    return master_and_embeddings(shards, stats, Z, sf2, alpha, beta, N, D, step_size, fixed_embeddings, fixed_beta)


def master_step(stats, Z, sf2, alpha, beta, N, D, fixed_beta=False):
    """calculate_global_statistics / calculate_global_derivatives (parallel_GPLVM.py:302-369)."""
    g = _new_pt(Z, sf2, alpha, beta, N, D)
    g.set_local_statistics(stats["sum_YYT"], stats["sum_exp_K_mi_K_im"], stats["sum_exp_K_miY"],
                           stats["sum_exp_K_ii"], stats["sum_KL"])
    pd = {
        "F": g.logmarglik(), "dF_dsum_exp_K_ii": g.dF_dexp_K_ii(),
        "dF_dsum_exp_K_miY": g.dF_dexp_K_miY(), "dF_dsum_exp_K_mi_K_im": g.dF_dexp_K_mi_K_im(),
        "dF_dKmm": g.dF_dKmm(),
    }
    grad_Z = g.grad_Z(pd["dF_dKmm"], g.dKmm_dZ(), pd["dF_dsum_exp_K_miY"], stats["sum_d_exp_K_miY_d_Z"],
                      pd["dF_dsum_exp_K_mi_K_im"], stats["sum_d_exp_K_mi_K_im_d_Z"])
    grad_alpha = g.grad_alpha(pd["dF_dKmm"], g.dKmm_dalpha(), pd["dF_dsum_exp_K_miY"],
                              stats["sum_d_exp_K_miY_d_alpha"], pd["dF_dsum_exp_K_mi_K_im"],
                              stats["sum_d_exp_K_mi_K_im_d_alpha"])
    grad_sf2 = g.grad_sf2(pd["dF_dKmm"], g.dKmm_dsf2(), pd["dF_dsum_exp_K_ii"], stats["sum_d_exp_K_ii_d_sf2"],
                          pd["dF_dsum_exp_K_miY"], stats["sum_d_exp_K_miY_d_sf2"],
                          pd["dF_dsum_exp_K_mi_K_im"], stats["sum_d_exp_K_mi_K_im_d_sf2"])
    grad_beta = 0.0 if fixed_beta else g.grad_beta()
    glob = dict(pd)
    glob.update({"Kmm": g.Kmm, "Kmm_inv": g.Kmm_inv, "Kmm_plus_op_inv": g.Kmm_plus_op_inv,
                 "grad_Z": grad_Z, "grad_alpha": grad_alpha, "grad_sf2": float(grad_sf2),
                 "grad_beta": float(grad_beta), "F": float(pd["F"]),
                 "cond_Kmm": float(np.linalg.cond(g.Kmm))})
    return glob


def master_and_embeddings(shards, stats, Z, sf2, alpha, beta, N, D, step_size, fixed_embeddings, fixed_beta):
    glob = master_step(stats, Z, sf2, alpha, beta, N, D, fixed_beta)
    out = {"stats": stats, "global": glob, "grad_latest": []}
    if not fixed_embeddings:
        for s in shards:
            out["grad_latest"].append(map_embeddings(s, stats, Z, sf2, alpha, beta, N, D, step_size))
    return out
