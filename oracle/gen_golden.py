"""TEST INFRASTRUCTURE ONLY -- generate ``tests/golden/*.npz`` by running the
*live* reference (``/root/reference/partial_terms.py`` through
``oracle/ref_shim.py``) on seeded inputs.  Run once in the build container:

    python -m oracle.gen_golden

The reference ships no golden vectors (its tests are unseeded, SURVEY.md
section 4); these fixtures are therefore the pin for both the numpy/C oracles
and the CUDA path on the GPU box, where ``/root/reference`` does not exist.

Each file stores the inputs (so that nothing depends on the RNG stream of a
particular numpy version) and the reference outputs: the 12 reduced statistics,
F, the partial derivatives, the four global gradients, and per-shard
``grad_latest``.  For the M=500 case only strided samples of the two
O(Q M^2) tensors are stored to keep the fixture small.
"""
import os
import sys
import time

import numpy as np

from gparml_b200.synthetic import make_problem
from . import ref_harness

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(os.path.dirname(HERE), "tests", "golden")

# name -> (problem kwargs, shard sizes, step_size)
CASES = {
    # reference test fixture shape (test.py:24-60): D=7, Q=2, N=5, M=10, S=0.2
    "t5": (dict(N=5, M=10, Q=2, D=7, seed=11, generic_hypers=True), [5], 0.0),
    # c1 shape (README minimal run): M=2 Q=2 D=4
    "c1s": (dict(N=50, M=2, Q=2, D=4, seed=12, generic_hypers=True, with_direction=True), [30, 20], 1e-3),
    # c2 shape: sparse GP regression, fixed embeddings, D=1
    "c2s": (dict(N=96, M=50, Q=4, D=1, seed=13, generic_hypers=True, fixed_embeddings=True), [40, 56], 0.0),
    # c3/c5 shape: M=100 Q=10 D=10, ragged shards incl. a single-point shard
    "c3s": (dict(N=48, M=100, Q=10, D=10, seed=14, generic_hypers=True, with_direction=True), [20, 1, 27], 1e-3),
    # c3 shape at the reference's initial hyper-parameters (sf2=alpha=beta=1)
    "c3u": (dict(N=32, M=100, Q=10, D=10, seed=15, generic_hypers=False), [32], 0.0),
    # c4 shape: M=500 Q=10 D=50 (strided samples of the big tensors only)
    "c4s": (dict(N=12, M=500, Q=10, D=50, seed=16, generic_hypers=True), [7, 5], 0.0),
}
BIG_SAMPLE = (slice(None, None, 7), slice(None), slice(None, None, 11))


def build_case(name):
    kw, sizes, step = CASES[name]
    p = make_problem(**kw)
    if name == "t5":
        p["X_S"] = np.log(np.expm1(np.full_like(p["X_S"], 0.2)))
    if kw["N"] < kw["M"] or name in ("t5",):
        # too few points to pick distinct inducing inputs from: draw Z ~ N(0, I) (test.py:56)
        rng = np.random.default_rng(1000 + kw["seed"])
        p["Z"] = rng.standard_normal((kw["M"], kw["Q"]))
    shards, lo = [], 0
    for n in sizes:
        sh = dict(Y=p["Y"][lo:lo + n], X_mu=p["X_mu"][lo:lo + n], X_S=p["X_S"][lo:lo + n])
        if "d" in p:
            sh["d"] = p["d"][:, lo:lo + n]
        shards.append(sh)
        lo += n
    return p, shards, step


def main(names=None):
    os.makedirs(GOLDEN, exist_ok=True)
    for name in (names or CASES):
        p, shards, step = build_case(name)
        t0 = time.time()
        ref = ref_harness.reference_evaluate(shards, p["Z"], p["sf2"], p["alpha"], p["beta"], step_size=step,
                                             fixed_embeddings=p["fixed_embeddings"])
        out = dict(Y=p["Y"], X_mu=p["X_mu"], X_S=p["X_S"], Z=p["Z"], sf2=p["sf2"], alpha=p["alpha"],
                   beta=p["beta"], step_size=step, fixed_embeddings=p["fixed_embeddings"],
                   shard_sizes=np.array(CASES[name][1]))
        if "d" in p:
            out["d"] = p["d"]
        big = p["M"] >= 400
        for k, v in ref["stats"].items():
            v = np.asarray(v, dtype=np.float64)
            if big and k == "sum_d_exp_K_mi_K_im_d_sf2":
                continue                      # = 2 * sum_exp_K_mi_K_im / sf2, stored in full below
            if big and v.ndim == 3 and v.size > 100_000:
                out["stat_sample_" + k] = v[BIG_SAMPLE]
            else:
                out["stat_" + k] = v
        for k in ("F", "grad_Z", "grad_alpha", "grad_sf2", "grad_beta", "dF_dsum_exp_K_miY", "cond_Kmm"):
            out["glob_" + k] = np.asarray(ref["global"][k], dtype=np.float64)
        if not big:
            for k in ("dF_dKmm", "dF_dsum_exp_K_mi_K_im", "Kmm", "Kmm_inv"):
                out["glob_" + k] = np.asarray(ref["global"][k], dtype=np.float64)
        for i, g in enumerate(ref["grad_latest"]):
            out["grad_latest_%d" % i] = g
        path = os.path.join(GOLDEN, name + ".npz")
        np.savez_compressed(path, **out)
        print("%s: %.1fs  cond(Kmm)=%.3g  F=%.12g  -> %s (%.0f kB)" % (
            name, time.time() - t0, ref["global"]["cond_Kmm"], ref["global"]["F"], path,
            os.path.getsize(path) / 1e3))


if __name__ == "__main__":
    main(sys.argv[1:] or None)
